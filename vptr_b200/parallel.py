"""Data-parallel glue: ONE flat all-reduce (mean) of the Transformer gradient per step, replacing the bucketed
DistributedDataParallel reduction of the reference (train_NAR_mp.py:118,167; train_FAR_mp.py:132,178).
Works with any torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def broadcast_parameters(module, src=0):
    """identical replicas at start (what DDP's constructor does): rank `src`'s parameters and buffers to every rank"""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


_CHECKED = set()


def _covering_view(grads):
    """If the gradients are consecutive views of ONE contiguous buffer (the engine hands out parameter gradients as views of a
    flat fp32 buffer, vptr_b200.engine.Params), return a 1-D view covering them all; else None."""
    if not grads or any(not g.is_contiguous() for g in grads):
        return None
    base = grads[0].untyped_storage().data_ptr()
    esz = grads[0].element_size()
    pos = grads[0].data_ptr()
    for g in grads:
        if g.untyped_storage().data_ptr() != base or g.dtype != grads[0].dtype or g.data_ptr() != pos:
            return None
        pos += g.numel() * esz
    total = (pos - grads[0].data_ptr()) // esz
    return torch.as_strided(grads[0], (total,), (1,), grads[0].storage_offset())


def allreduce_mean_grads(params, world_size=None):
    """In-place mean over ranks of every existing .grad with ONE collective.  Gradients that already are consecutive views of a
    flat buffer (the Transformer's, straight out of the engine's backward) are reduced in place -- no flatten / unflatten copies;
    anything else goes through a temporary flat buffer.  NCCL averages inside the collective (ReduceOp.AVG); other backends sum
    and divide.  Parameters whose grad is None contribute nothing locally; the first call for a given gradient set checks that
    all ranks hold the same set and raises instead of hanging later."""
    world_size = dist.get_world_size() if world_size is None else world_size
    grads = [p.grad for p in params if p.grad is not None]
    sig = (len(grads), sum(g.numel() for g in grads))
    if sig not in _CHECKED:
        n_local = torch.tensor(list(sig), dtype=torch.int64, device=grads[0].device if grads else "cpu")
        n_max = n_local.clone()
        dist.all_reduce(n_max, op=dist.ReduceOp.MAX)
        n_min = n_local.clone()
        dist.all_reduce(n_min, op=dist.ReduceOp.MIN)
        if not (torch.equal(n_max, n_local) and torch.equal(n_min, n_local)):
            raise RuntimeError("allreduce_mean_grads: ranks hold different gradient sets (%s vs min %s max %s)" %
                               (n_local.tolist(), n_min.tolist(), n_max.tolist()))
        _CHECKED.add(sig)
    if not grads:
        return 0
    avg = dist.get_backend() == "nccl"

    def reduce_(flat):
        if avg:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(world_size)

    # split into maximal runs that are already flat in memory (one linear pass: a gradient extends the current run iff it starts
    # where the previous one ended, in the same storage); the longest run (the Transformer) is reduced in place
    runs, cur, nxt = [], [], None
    for g in grads:
        ok = g.is_contiguous()
        if cur and ok and g.data_ptr() == nxt and g.dtype == cur[0].dtype and \
                g.untyped_storage().data_ptr() == cur[0].untyped_storage().data_ptr():
            cur.append(g)
        else:
            if cur:
                runs.append(cur)
            cur = [g]
        nxt = g.data_ptr() + g.numel() * g.element_size() if ok else None
    runs.append(cur)
    loose = []
    for r in runs:
        v = _covering_view(r)
        if v is not None and v.numel() >= (1 << 20):
            reduce_(v)
        else:
            loose.extend(r)
    if loose:
        flat = torch._utils._flatten_dense_tensors(loose)
        reduce_(flat)
        for g, s_ in zip(loose, torch._utils._unflatten_dense_tensors(flat, loose)):
            g.copy_(s_)
    return sig[1]


class NativeComm:
    """NCCL communicator owned by libvptr_b200.so (vptr_nccl_comm_init): the 128-byte unique id is made on rank 0 and shipped with
    torch.distributed (whatever backend is up); after that the gradient reduction needs no torch collective."""

    def __init__(self, device):
        import ctypes
        from . import _lib
        self.lib = _lib.lib()
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            rc = self.lib.vptr_nccl_unique_id(uid.data_ptr())
            if rc != 0:
                raise RuntimeError("vptr_nccl_unique_id failed (%d): %s" % (rc, self.lib.vptr_last_error().decode("utf-8", "replace")))
        if dist.get_backend() == "nccl":
            u = uid.to(device)
            dist.broadcast(u, 0)
            uid = u.cpu()
        else:
            dist.broadcast(uid, 0)
        comm = ctypes.c_void_p()
        with torch.cuda.device(device):
            rc = self.lib.vptr_nccl_comm_init(ctypes.byref(comm), world, rank, uid.data_ptr())
        if rc != 0:
            raise RuntimeError("vptr_nccl_comm_init failed (%d): %s" % (rc, self.lib.vptr_last_error().decode("utf-8", "replace")))
        self.comm, self.world = comm, world

    def allreduce_mean_(self, flat, sqnorm, stream):
        from . import _lib
        _lib.call("vptr_allreduce_grads", self.comm, flat.data_ptr(), flat.numel(), 0 if sqnorm is None else sqnorm.data_ptr(), stream.cuda_stream)


class GradReducer:
    """Data-parallel gradient mean overlapped with the backward pass (the role of DistributedDataParallel's bucketed hooks,
    train_NAR_mp.py:118,167): the engine finishes the Transformer's layers in reverse order and each layer's parameter gradients
    are one contiguous slice of the flat gradient buffer, so every finished slice is handed to NCCL at once (`arm()` installs the
    engine hook) while the remaining layers' backward keeps the SMs busy.  `finish()` reduces whatever is left (frame queries, final
    norms, gradients outside the flat buffer) and makes the compute stream wait for all of it.

    On CUDA with the NCCL backend the collectives go through the library's own communicator (vptr_allreduce_grads) on a side
    stream, each followed by the squared-norm accumulation of the reduced slice -- so the norm clip_grad_norm_ needs (`sqnorm`)
    is ready when the reduction is, without a pass of its own.  Other backends (gloo in the CPU tests) use torch.distributed."""

    def __init__(self, params, world_size=None, min_chunk=1 << 20, native=None):
        self.params = list(params)
        self.world = dist.get_world_size() if world_size is None else world_size
        self.min_chunk = min_chunk
        self.works, self.done, self.flat = [], [], None
        self.avg = dist.get_backend() == "nccl"
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.native = None
        if native is None:
            native = self.avg and dev.type == "cuda"
        if native:
            self.native = NativeComm(dev)
            self.side = torch.cuda.Stream(device=dev)
            self.sq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.sqnorm = None          # after finish(): (1,) float64 device tensor = sum of squares of ALL reduced gradients, or None
        self.rest = None

    def arm(self):
        from . import engine
        self.works, self.done, self.flat, self.sqnorm = [], [], None, None
        if self.native is not None:
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                self.sq.zero_()
        engine.GRAD_READY = self._ready

    def _ready(self, flat, lo, hi):
        if hi - lo < self.min_chunk:
            return
        self.flat = flat
        chunk = flat[lo:hi]
        if self.native is not None:
            self.side.wait_stream(torch.cuda.current_stream())      # the slice's producers have been enqueued on the compute stream
            self.native.allreduce_mean_(chunk, self.sq, self.side)
        elif self.avg:
            self.works.append((dist.all_reduce(chunk, op=dist.ReduceOp.AVG, async_op=True), chunk))
        else:
            self.works.append((dist.all_reduce(chunk, op=dist.ReduceOp.SUM, async_op=True), chunk))
        self.done.append((lo, hi))

    def finish(self):
        from . import engine
        engine.GRAD_READY = None
        for w, chunk in self.works:
            w.wait()
            if not self.avg:
                chunk.div_(self.world)
        rest = []
        if self.flat is not None and self.done:
            base, esz, n = self.flat.data_ptr(), self.flat.element_size(), self.flat.numel()
            spans = sorted(self.done)
            for p in self.params:
                g = p.grad
                if g is None:
                    continue
                off = (g.data_ptr() - base) // esz
                inside = 0 <= off < n and any(lo <= off and off + g.numel() <= hi for lo, hi in spans)
                if not inside:
                    rest.append(p)
        else:
            rest = self.params
        if self.native is not None:
            # the remainder (frame queries, final norms, gradients outside the flat buffer: a few MB) goes through the library's
            # communicator too, packed into one temporary -- no torch collective inside the step, so the step stays graph-capturable
            cur = torch.cuda.current_stream()
            cur.wait_stream(self.side)
            gs = [p.grad for p in rest if p.grad is not None]
            if gs:
                flat = torch.cat([g.reshape(-1) for g in gs])
                self.native.allreduce_mean_(flat, self.sq, cur)
                off = 0
                for g in gs:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
            self.sqnorm, self.rest = self.sq, []
        else:
            allreduce_mean_grads(rest, self.world)
            self.rest = rest
        self.works, self.done = [], []
