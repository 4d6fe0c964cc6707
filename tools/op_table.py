"""Per-entry-point timing table of one training step (any bench.py --config): wraps the C-ABI call site with CUDA events (same stream) and
aggregates by (entry point, shape signature).  Usage: python tools/op_table.py [--batch N] [--out gpurun_out/op_table.txt]"""
import argparse
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import bench  # noqa: E402
from vptr_b200 import _lib, ops  # noqa: E402


def signature(name, a):
    if name in ("vptr_gemm_tf32", "vptr_gemm_simt"):
        return "M=%d N=%d K=%d a_mn=%d b_mn=%d res=%d flags=%d act=%d ks=%d" % (a[8], a[9], a[10], a[2], a[5], int(a[12] != 0), a[16], a[15], a[17])
    ints = [x for x in a[:-1] if isinstance(x, int) and 0 <= x < (1 << 24)]
    return " ".join(str(x) for x in ints[:10])


def gemm_flops(name, a):
    if name in ("vptr_gemm_tf32", "vptr_gemm_simt"):
        return 2.0 * a[8] * a[9] * a[10]
    return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--config", default="cfg1")
    ap.add_argument("--torch-tail", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "op_table.txt"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    from vptr_b200 import model as M
    from vptr_b200.trainer import Stage2Trainer
    c = bench.CONFIGS[args.config]
    enc, dec, T = bench.build_modules(M, torch, c, dev, args.dropout)
    n = args.batch or c["clips_per_gpu"]
    trainer = Stage2Trainer(c["kind"], enc, dec, T, use_bpnce=c["bpnce"], fused_tail=not args.torch_tail)
    past, fut = (t.to(dev) for t in bench.clip_tensors(torch, c, n, 0))

    def step():
        return trainer.step(past, fut)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    rec = []
    orig = ops._call

    def wrapped(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(name, *a)
        e1.record()
        rec.append((name, signature(name, a), gemm_flops(name, a), e0, e1))
        return r

    from vptr_b200 import tail
    ops._call = wrapped
    tail._call = wrapped
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    step()
    t1.record()
    torch.cuda.synchronize()
    ops._call = orig
    tail._call = orig
    total = t0.elapsed_time(t1)
    by_name, by_sig = defaultdict(lambda: [0.0, 0, 0.0]), defaultdict(lambda: [0.0, 0, 0.0])
    for name, sig, fl, e0, e1 in rec:
        ms = e0.elapsed_time(e1)
        for d, k in ((by_name, name), (by_sig, name + " | " + sig)):
            d[k][0] += ms
            d[k][1] += 1
            d[k][2] += fl
    lines = ["step %.2f ms (instrumented), %d C-ABI launches, sum of launches %.2f ms" % (total, len(rec), sum(v[0] for v in by_name.values()))]
    lines.append("---- by entry point")
    for k, v in sorted(by_name.items(), key=lambda kv: -kv[1][0]):
        lines.append("%9.3f ms %5.1f%% %5d x  %s%s" % (v[0], 100 * v[0] / total, v[1], k, ("  %.0f TFLOP/s" % (v[2] / v[0] / 1e9)) if v[2] else ""))
    lines.append("---- by entry point and shape")
    for k, v in sorted(by_sig.items(), key=lambda kv: -kv[1][0])[:120]:
        lines.append("%9.3f ms %5d x %8.1f us  %s%s" % (v[0], v[1], 1e3 * v[0] / v[1], k, ("  %.0f TFLOP/s" % (v[2] / v[0] / 1e9)) if v[2] else ""))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:45]))


if __name__ == "__main__":
    main()
