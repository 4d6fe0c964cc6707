"""Multi-GPU check of vptr_b200.parallel.GradReducer (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/test_dp.py
Every rank runs the same weights on its own clips; the overlapped, library-owned NCCL reduction (vptr_allreduce_grads on a side
stream, slices announced by the engine while the backward runs) must equal the plain mean of the per-rank gradients, and the
squared norm it accumulates behind the all-reduces must equal the norm of that mean."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from vptr_b200.model import VPTRFormerNAR
    from vptr_b200.parallel import GradReducer
    from vptr_b200.tail import grad_sqnorm
    torch.manual_seed(7)
    T = VPTRFormerNAR(4, 4, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=2, num_decoder_layers=2, dropout=0.0, window_size=4, rpe=True).to(dev)
    for p in T.parameters():
        dist.broadcast(p.data, 0)
    x = torch.rand(2, 4, 528, 8, 8, generator=torch.Generator().manual_seed(100 + rank)).to(dev)

    def fwd_bwd():
        T.zero_grad(set_to_none=True)
        y = T(x)
        pf = T.NCE_projector(y.permute(0, 1, 3, 4, 2))
        (y.square().mean() + pf.square().mean()).backward()

    T.train()
    fwd_bwd()
    ref = {}
    for k, p in T.named_parameters():
        g = p.grad.detach().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        ref[k] = g / world
    ref_sq = sum(float(g.double().square().sum()) for g in ref.values())
    # run-to-run noise floor of the local gradients (split-K atomics reorder fp32 sums): measured, and used as the comparison scale
    loc1 = {k: p.grad.detach().clone() for k, p in T.named_parameters()}
    fwd_bwd()
    gm = max(float(g.norm()) for g in loc1.values())
    nz = sorted(((float((p.grad - loc1[k]).norm()) / max(float(loc1[k].norm()), 1e-3 * gm), k) for k, p in T.named_parameters()), reverse=True)
    noise = nz[0][0]
    if rank == 0:
        print("noisiest:", nz[:4], flush=True)
    inside = sum(int(p.grad.untyped_storage().data_ptr() == T._vptr_gflat.untyped_storage().data_ptr()) for p in T.parameters())
    if rank == 0:
        print("run-to-run deviation of local gradients (no reduction): %.3e; %d of %d .grad tensors are views of the flat buffer"
              % (noise, inside, len(list(T.parameters()))), flush=True)
    red = GradReducer(list(T.parameters()), world, min_chunk=1 << 16)
    assert red.native is not None, "library-owned NCCL communicator was not created"
    for it in range(2):
        red.arm()
        fwd_bwd()
        red.finish()
        gmax = max(float(g.norm()) for g in ref.values())
        worst, errs = 0.0, []
        for k, p in T.named_parameters():
            e = float((p.grad - ref[k]).norm()) / max(float(ref[k].norm()), 1e-3 * gmax)
            errs.append((e, k, float(ref[k].norm())))
            worst = max(worst, e)
        tol = max(2e-5, 3.0 * noise)     # the run-to-run noise of the local gradients is the yardstick
        if worst >= tol and rank == 0:
            for e, k, n in sorted(errs, reverse=True)[:12]:
                print("  %-70s %.3e  |ref| %.3e" % (k, e, n), flush=True)
        assert worst < tol, (rank, it, worst, noise)
        assert red.sqnorm is not None and len(red.done) == 0
        sq = float(grad_sqnorm(red.rest, red.sqnorm, accumulate=True))
        assert abs(sq - ref_sq) <= 1e-5 * ref_sq, (rank, it, sq, ref_sq)
    torch.cuda.synchronize()
    if rank == 0:
        print("DP OK: world %d, worst gradient deviation %.2e, fused squared norm %.6e vs %.6e" % (world, worst, sq, ref_sq))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
