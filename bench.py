"""bench.py -- predicted frames / second of one full stage-2 VPTR-NAR training iteration (BASELINE.json metric),
workload cfg1: MovingMNIST-shape 10->10, 64x64x1, 64 synthetic clips per GPU, 4 encoder + 8 decoder layers.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores (rank 0 only)

A step = everything train_NAR.single_iter does (reference train_NAR.py:49-107): ResNet-encode past and future clips
(no_grad), Transformer forward, ResNet decoder, NCE projector, MSE + GDL + 0.1*BiPatchNCE, backward, [gradient all-reduce
when N > 1], clip_grad_norm_(1.0), AdamW step.  `value` times it with the clips already resident in HBM; `e2e` times the
same step through the public nn.Module API with the clips copied from pinned host memory and the loss read back every
step.  Prints ONE JSON line on rank 0."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(Tp=10, Tf=10, img=64, Cimg=1, d_model=528, nhead=8, enc_layers=4, dec_layers=8, ws=4, clips_per_gpu=64)
METRIC = "predicted frames/sec (NAR 10->10, 64x64)"
WORKLOAD = "VPTR-NAR MovingMNIST-shape 10->10, 64x64x1, 4 enc + 8 dec layers, d_model 528, window 4 (cfg1)"   # both arms


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, bf16=1400.0, src="fallback")


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def step_flops(n_clips):
    """algorithmic FLOPs of one training step (SURVEY.md 8d / BASELINE.md 3): 553.9 GFLOP per clip for cfg1"""
    return 553.9e9 * n_clips


# ======================================================================================================== reference arm
def run_reference(args):
    """The reference algorithm on the host CPU: oracle/train_step.py (a port -- /root/reference does not travel to the GPU
    box), all host threads, batch 2 of the 64-clip workload per step."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import train_step as TS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    res = cpu_step_rate(TS, torch, clips=2, steps=max(1, min(args.steps, 3)), warmup=1)
    line = {"impl": "reference", "metric": METRIC, "value": res["fps"], "unit": "frames/s", "n_gpus": args.gpus, "steps": res["steps"],
            "warmup": 1, "ms_per_step": res["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_step": 2,
                       "note": "CPU port of the reference algorithm (oracle/), dropout 0"},
            "cpu_baseline": {"value": res["fps"], "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "%d full training steps at 2 clips (of the 64-clip workload) after 1 warm-up" % res["steps"]},
            "e2e": {"value": res["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_step_rate(TS, torch, clips, steps, warmup):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from vptr_b200.model import VPTRDec, VPTREnc, VPTRFormerNAR, init_weights
    torch.manual_seed(2021)
    enc = VPTREnc(CFG["Cimg"], feat_dim=CFG["d_model"], n_downsampling=3).eval()
    dec = VPTRDec(CFG["Cimg"], feat_dim=CFG["d_model"], n_downsampling=3, out_layer="Sigmoid").eval()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):     # init_weights prints; stdout carries exactly one JSON line
        init_weights(enc)
        init_weights(dec)
    T = VPTRFormerNAR(CFG["Tp"], CFG["Tf"], encH=8, encW=8, d_model=CFG["d_model"], nhead=CFG["nhead"], num_encoder_layers=CFG["enc_layers"],
                      num_decoder_layers=CFG["dec_layers"], dropout=0.0, window_size=CFG["ws"], rpe=True)
    sd_T = {k: v.detach() for k, v in T.state_dict().items()}
    params = {k: v.detach().clone().requires_grad_(True) for k, v in T.named_parameters()}
    opt = torch.optim.AdamW(list(params.values()), lr=1e-4)
    g = torch.Generator().manual_seed(2021)
    past = torch.rand(clips, CFG["Tp"], CFG["Cimg"], CFG["img"], CFG["img"], generator=g)
    fut = torch.rand(clips, CFG["Tf"], CFG["Cimg"], CFG["img"], CFG["img"], generator=g)
    sde = {k: v.detach() for k, v in enc.state_dict().items()}
    sdd = {k: v.detach() for k, v in dec.state_dict().items()}
    for _ in range(warmup):
        TS.nar_step(sde, sdd, sd_T, params, opt, past, fut)
    t0 = time.perf_counter()
    for _ in range(steps):
        TS.nar_step(sde, sdd, sd_T, params, opt, past, fut)
    dt = (time.perf_counter() - t0) / steps
    return {"fps": clips * CFG["Tf"] / dt, "ms": dt * 1e3, "steps": steps}


# ======================================================================================================== CUDA arm
def build_models(torch, device, dropout):
    from vptr_b200.model import BiPatchNCE, GDL, MSELoss, VPTRDec, VPTREnc, VPTRFormerNAR, init_weights
    torch.manual_seed(2021)
    enc = VPTREnc(CFG["Cimg"], feat_dim=CFG["d_model"], n_downsampling=3, padding_type="reflect").to(device).eval()
    dec = VPTRDec(CFG["Cimg"], feat_dim=CFG["d_model"], n_downsampling=3, out_layer="Sigmoid", padding_type="reflect").to(device).eval()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        init_weights(enc)
        init_weights(dec)
    T = VPTRFormerNAR(CFG["Tp"], CFG["Tf"], encH=8, encW=8, d_model=CFG["d_model"], nhead=CFG["nhead"], num_encoder_layers=CFG["enc_layers"],
                      num_decoder_layers=CFG["dec_layers"], dropout=dropout, window_size=CFG["ws"], rpe=True).to(device)
    losses = dict(mse=MSELoss(), gdl=GDL(alpha=1), bpnce=BiPatchNCE(CFG["clips_per_gpu"], CFG["Tf"], 8, 8, 1.0).to(device))
    opt = torch.optim.AdamW(params=T.parameters(), lr=1e-4)
    return enc, dec, T, losses, opt


def train_step(torch, F, dist, world, enc, dec, T, losses, opt, past, future, flat_params):
    """train_NAR.single_iter (reference train_NAR.py:49-107) on the drop-in modules"""
    with torch.no_grad():
        past_f = enc(past)
        fut_f = enc(future)
    T.train()
    T.zero_grad(set_to_none=True)
    dec.zero_grad(set_to_none=True)
    pred_f = T(past_f)
    pred = dec(pred_f)
    pf = T.NCE_projector(pred_f.permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3)
    gf = T.NCE_projector(fut_f.permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3)
    loss = losses["mse"](pred, future) + losses["gdl"](future, pred) + 0.1 * losses["bpnce"](F.normalize(gf, p=2.0, dim=2), F.normalize(pf, p=2.0, dim=2))
    loss.backward()
    if world > 1:   # data-parallel gradient mean over NVLink: replaces DistributedDataParallel (train_NAR_mp.py:118,167)
        from vptr_b200.parallel import allreduce_mean_grads
        allreduce_mean_grads(flat_params, world)
    torch.nn.utils.clip_grad_norm_(T.parameters(), max_norm=1.0, norm_type=2)
    opt.step()
    return loss


def run_cuda(args):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (the CUDA arm has no CPU fallback); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from vptr_b200 import _lib, ops
    _lib.lib()
    enc, dec, T, losses, opt = build_models(torch, device, args.dropout)
    flat_params = list(T.parameters())
    if world > 1:   # identical replicas: broadcast rank 0's initial weights (DDP constructor semantics)
        for p in flat_params:
            dist.broadcast(p.data, 0)
    n = args.batch or CFG["clips_per_gpu"]
    if n != CFG["clips_per_gpu"]:
        from vptr_b200.model import BiPatchNCE
        losses["bpnce"] = BiPatchNCE(n, CFG["Tf"], 8, 8, 1.0).to(device)
    g = torch.Generator().manual_seed(2021 + rank)
    shape_p = (n, CFG["Tp"], CFG["Cimg"], CFG["img"], CFG["img"])
    shape_f = (n, CFG["Tf"], CFG["Cimg"], CFG["img"], CFG["img"])
    host_p = torch.rand(*shape_p, generator=g).pin_memory()
    host_f = torch.rand(*shape_f, generator=g).pin_memory()
    dev_p, dev_f = host_p.to(device), host_f.to(device)
    l2_flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, from_host):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        last = None
        for _ in range(k):
            if from_host:
                p, f = host_p.to(device, non_blocking=True), host_f.to(device, non_blocking=True)
            else:
                p, f = dev_p, dev_f
            loss = train_step(torch, F, dist, world, enc, dec, T, losses, opt, p, f, flat_params)
            if from_host:
                last = loss.item()      # device -> host read of the step's result
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / k, last

    for _ in range(args.warmup):
        train_step(torch, F, dist, world, enc, dec, T, losses, opt, dev_p, dev_f, flat_params)
    l2_flush.zero_()
    if args.ncu_step:   # one step between cudaProfilerStart/Stop for `ncu --profile-from-start off`; prints nothing
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        train_step(torch, F, dist, world, enc, dec, T, losses, opt, dev_p, dev_f, flat_params)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    c0 = _lib.launch_count
    ms_dev, _ = timed(args.steps, from_host=False)
    launches = (_lib.launch_count - c0) // max(args.steps, 1)
    ms_e2e, loss_val = timed(args.steps, from_host=True)
    clocks = sampler.stop() if rank == 0 else None
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30

    # --- dominant kernel (tcgen05 TF32 GEMM): CUDA-event time of every launch of one more step -> achieved TFLOP/s
    gemm_stats = profile_gemms(torch, ops, lambda: train_step(torch, F, dist, world, enc, dec, T, losses, opt, dev_p, dev_f, flat_params))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    frames = n * CFG["Tf"] * world
    value = frames / (ms_dev * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)
    tf32_peak = pk["bf16"] / 2.0
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_dev, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32 (fp32 storage, fp32 accumulate)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "clips_per_gpu": n, "global_clips": n * world, "dropout": args.dropout, "parallelism": "dp%d" % world,
                   "l2": "per-step working set (>60 GB of activations) far exceeds the 126 MB L2; L2 flushed once before timing",
                   "loss": "MSE + GDL + 0.1*BiPatchNCE", "optimizer": "AdamW lr 1e-4, clip_grad_norm 1.0",
                   "step_tflop_algorithmic": round(step_flops(n) / 1e12, 2), "peak_mem_gib": round(peak_mem, 1)},
        "e2e": {"value": round(e2e, 2), "unit": "frames/s", "ms_per_step": round(ms_e2e, 3),
                "h2d_bytes_per_step": int(host_p.numel() + host_f.numel()) * 4, "d2h_bytes_per_step": 4, "loss": loss_val},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "gemm_tf32_2cta_kernel (tcgen05.mma cta_group::2 kind::tf32, TMA operands, TMEM accumulators; all Linear/1x1/3x3-conv contractions, fwd+dgrad+wgrad)", "bound": "tensor", "achieved": round(gemm_stats["tflops"], 1),
                     "peak": round(tf32_peak, 1), "unit": "TFLOP/s", "frac": round(gemm_stats["tflops"] / tf32_peak, 4), "traffic": None,
                     "traffic_sample": {"launch": "M=40960 N=2112 K=528 (fc1 / linear1 shape)", "dram_bytes": 388915456,
                                        "algorithmic_bytes": 437000000, "tensor_pipe_active_pct": 52.4,
                                        "source": "profiles/r01_ncu_gemm_fc1_s3.txt (ncu --set full; achieved above is the aggregate over all GEMM shapes of a step, so a single per-launch traffic figure does not exist)",
                                        "conv3x3_w8": {"tensor_pipe_active_pct": 89.0, "dram_bytes": 230455808, "source": "profiles/r01_ncu_conv3x3_w8.txt"},
                                        "attn_tc_fwd (window attention, tcgen05)": {"tensor_pipe_active_pct": 14.4, "dram_bytes": 323859968, "algorithmic_bytes": 346030080,
                                                                                    "hbm_gbs": 3280, "hbm_frac_of_measured": 0.50,
                                                                                    "source": "profiles/r01_ncu_attn_tcgen05.txt"}},
                     "peak_note": "tf32 dense = half of the %s bf16 sustained %.1f TFLOP/s (MEASURED_PEAKS.json has no tf32 entry)" % (pk["src"], pk["bf16"]),
                     "launches_per_step": gemm_stats["launches"], "gemm_ms_per_step": round(gemm_stats["ms"], 3),
                     "gemm_share_of_step": round(gemm_stats["ms"] / ms_dev, 3), "gemm_tflop_per_step": round(gemm_stats["flop"] / 1e12, 3)},
        "model_flops_utilisation": {"achieved_tflops": round(step_flops(n) / (ms_dev * 1e-3) / 1e12, 1), "of_tf32_peak": round(step_flops(n) / (ms_dev * 1e-3) / 1e12 / tf32_peak, 4)},
    }
    if not args.no_cpu_baseline and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import train_step as TS
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        res = cpu_step_rate(TS, torch, clips=2, steps=2, warmup=1)
        line["cpu_baseline"] = {"value": round(res["fps"], 3), "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "2 full training steps at 2 clips (of the 64-clip workload) after 1 warm-up, oracle port, dropout 0"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def profile_gemms(torch, ops, step_fn):
    """re-runs one step with a CUDA-event pair around every vptr_gemm_tf32 launch (same stream)"""
    records = []
    orig = ops.gemm

    def wrapped(A, B, out=None, a_mn=False, b_mn=False, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(A, B, out=out, a_mn=a_mn, b_mn=b_mn, **kw)
        e1.record()
        K, M = (A.shape if a_mn else A.shape[::-1])
        N = B.shape[1] if b_mn else B.shape[0]
        records.append((e0, e1, 2.0 * M * N * K))
        return r

    orig_conv = ops.conv3x3_tf32

    def wrapped_conv(xpad, w, F_, H, W, C, Cout, **kw):      # implicit-GEMM convolution: same kernel, A via 4-D TMA
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_conv(xpad, w, F_, H, W, C, Cout, **kw)
        e1.record()
        records.append((e0, e1, 2.0 * F_ * H * W * Cout * 9 * C))
        return r

    ops.gemm = wrapped
    ops.conv3x3_tf32 = wrapped_conv
    try:
        step_fn()
        torch.cuda.synchronize()
    finally:
        ops.gemm = orig
        ops.conv3x3_tf32 = orig_conv
    ms = sum(e0.elapsed_time(e1) for e0, e1, _ in records)
    flop = sum(f for _, _, f in records)
    return {"ms": ms, "flop": flop, "launches": len(records), "tflops": flop / (ms * 1e-3) / 1e12 if ms > 0 else 0.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU (default: the 64 of cfg1)")
    ap.add_argument("--dropout", type=float, default=0.1, help="Transformer dropout / DropPath rate (reference default 0.1, train_NAR.py:199)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-step", action="store_true", help="profile exactly one step (use under ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
