"""Data-parallel glue: ONE flat all-reduce (mean) of the Transformer gradient per step, replacing the bucketed
DistributedDataParallel reduction of the reference (train_NAR_mp.py:118,167; train_FAR_mp.py:132,178).
Works with any torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def broadcast_parameters(module, src=0):
    """identical replicas at start (what DDP's constructor does): rank `src`'s parameters and buffers to every rank"""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def allreduce_mean_grads(params, world_size=None):
    """In-place mean over ranks of every existing .grad, through one flat buffer.  Parameters whose grad is None
    contribute nothing locally; if ranks disagree on which grads exist the call raises instead of hanging."""
    world_size = dist.get_world_size() if world_size is None else world_size
    grads = [p.grad for p in params if p.grad is not None]
    n_local = torch.tensor([len(grads), sum(g.numel() for g in grads)], dtype=torch.int64,
                           device=grads[0].device if grads else "cpu")
    n_max = n_local.clone()
    dist.all_reduce(n_max, op=dist.ReduceOp.MAX)
    if not torch.equal(n_max, n_local):
        raise RuntimeError("allreduce_mean_grads: ranks hold different gradient sets (%s vs max %s)" % (n_local.tolist(), n_max.tolist()))
    if not grads:
        return 0
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world_size)
    for g, s in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(s)
    return flat.numel()
