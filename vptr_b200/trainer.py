"""Host-side mirror of the reference's stage-2 training iteration -- `train_NAR.single_iter` (reference train_NAR.py:49-107,
loss `cal_lossT` :33-47) and `train_FAR.single_iter` (train_FAR.py:48-101, `cal_lossT` :32-46) -- on the drop-in modules of
vptr_b200.model, with the data-parallel gradient mean of train_{NAR,FAR}_mp.py:118,167 (DistributedDataParallel there) done by
vptr_b200.parallel.  bench.py times exactly this; tests/test_gpu_dropin.py checks it against the reference's own, unmodified
single_iter driving the same modules.

The tail of the iteration (losses, global-norm clip, AdamW) runs on the library's fused kernels
(vptr_b200.tail) unless `fused_tail=False`, in which case it is the reference's literal PyTorch sequence."""
import torch
import torch.nn.functional as F

from . import parallel


class Stage2Trainer:
    def __init__(self, kind, enc, dec, transformer, lr=1e-4, max_grad_norm=1.0, lam_pc=0.1, use_bpnce=True, world=1,
                 fused_tail=True, optimizer=None):
        """kind: 'nar' | 'far'.  use_bpnce: the single-GPU NAR script adds 0.1 * BiPatchNCE (train_NAR.py:42-45); the multi-GPU
        script disables it (train_NAR_mp.py:68-69) but still encodes the future frames and runs the projector."""
        from .model import GDL, MSELoss
        self.kind, self.enc, self.dec, self.T = kind, enc, dec, transformer
        self.max_grad_norm, self.lam_pc, self.use_bpnce, self.world = max_grad_norm, lam_pc, use_bpnce and kind == "nar", world
        self.mse, self.gdl = MSELoss(), GDL(alpha=1)
        self.bpnce = None
        self.params = list(transformer.parameters())
        self.fused_tail = fused_tail
        self.tail = None
        if fused_tail:
            from . import tail
            self.tail = tail.FusedTail(self.params, lr=lr, max_grad_norm=max_grad_norm)
            self.opt = None
        else:
            self.opt = optimizer if optimizer is not None else torch.optim.AdamW(params=self.params, lr=lr)
        self.reducer = parallel.GradReducer(self.params, world) if world > 1 else None
        self.graph_mode = False

    # ------------------------------------------------------------------------------------------------ losses
    def _bpnce(self, n, t, h, w, device):
        from .model import BiPatchNCE
        key = (n, t, h, w)
        if self.bpnce is None or self.bpnce[0] != key:
            self.bpnce = (key, BiPatchNCE(n, t, h, w, 1.0).to(device))
        return self.bpnce[1]

    def _pixel_loss(self, pred, target):
        """MSE + GDL (cal_lossT).  Fused: one kernel for the loss value and one for d(loss)/d(pred) (vptr_b200.tail)."""
        if self.fused_tail:
            from . import tail
            return tail.mse_gdl_loss(pred, target)
        return self.mse(pred, target) + self.gdl(target, pred)

    # ------------------------------------------------------------------------------------------------ one iteration
    def step(self, past, future):
        T, enc, dec = self.T, self.enc, self.dec
        if self.graph_mode:      # first node of a captured step: new dropout / DropPath epoch for this replay (common.cuh)
            from . import _lib, ops
            _lib.call("vptr_rng_advance", -1, ops._s())
        if self.kind == "far":
            with torch.no_grad():
                feats = enc(torch.cat([past, future[:, :-1]], dim=1))          # train_FAR.py:53-55
            T.train()
            T.zero_grad(set_to_none=True)
            dec.zero_grad(set_to_none=True)
            if self.reducer is not None:
                self.reducer.arm()
            pred = dec(T(feats))
            loss = self._pixel_loss(pred, torch.cat([past[:, 1:], future], dim=1))   # :80
        else:
            with torch.no_grad():                                               # train_NAR.py:54-56: Enc(past), Enc(future) -- frames
                n, tp, tf = past.shape[0], past.shape[1], future.shape[1]       # are independent, so both go through ONE encoder pass
                f = enc(torch.cat([past.flatten(0, 1), future.flatten(0, 1)], dim=0).unsqueeze(0))[0]
                past_f = f[:n * tp].unflatten(0, (n, tp))
                fut_f = f[n * tp:].unflatten(0, (n, tf))
            T.train()
            T.zero_grad(set_to_none=True)
            dec.zero_grad(set_to_none=True)
            if self.reducer is not None:
                self.reducer.arm()
            pred_f = T(past_f)
            pred = dec(pred_f)
            pf = T.NCE_projector(pred_f.permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3)     # :81-82
            gf = T.NCE_projector(fut_f.permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3)
            loss = self._pixel_loss(pred, future)
            if self.use_bpnce:
                n, t, _, h, w = pf.shape
                if self.fused_tail and h * w <= 64:
                    from . import tail
                    loss = loss + self.lam_pc * tail.bipatch_nce_normalized(gf, pf, 1.0)
                else:
                    loss = loss + self.lam_pc * self._bpnce(n, t, h, w, pf.device)(F.normalize(gf, p=2.0, dim=2), F.normalize(pf, p=2.0, dim=2))
        loss.backward()
        if self.reducer is not None:   # data-parallel gradient mean over NVLink: replaces DistributedDataParallel
            self.reducer.finish()
        if self.fused_tail:
            if self.reducer is not None and self.reducer.sqnorm is not None:
                self.tail.clip_and_step(self.reducer.sqnorm, self.reducer.rest)
            else:
                self.tail.clip_and_step()
        else:
            torch.nn.utils.clip_grad_norm_(self.params, max_norm=self.max_grad_norm, norm_type=2)   # :85
            self.opt.step()                                                                          # :86
        return loss


class GraphedStep:
    """One training iteration captured as a CUDA graph and replayed (stage-2 shapes are static: same clip shape every step).

    Every kernel of the step is a C-ABI call on borrowed pointers with an explicit stream, the optimizer's step count and the dropout
    epoch live on the device (vptr_counter_add / vptr_rng_advance), and -- with more than one rank -- the gradient reduction is the
    library's own NCCL communicator on a forked side stream, so the whole of `Stage2Trainer.step` (ResNet encoder, Transformer forward
    and backward, decoder, losses, all-reduce, clip, AdamW) becomes ONE graph launch: ~1 500 kernel launches and their host-side
    Python disappear from the critical path.  Requires the fused tail.  The replayed step reads its clips from static device buffers
    (`step(past, future)` copies into them) and returns the static loss tensor."""

    def __init__(self, trainer, past, future, warmup=3):
        if not trainer.fused_tail:
            raise RuntimeError("vptr_b200.trainer.GraphedStep needs fused_tail=True (torch.optim.AdamW keeps its step count on the host)")
        self.trainer = trainer
        self.past, self.future = past.clone(), future.clone()
        trainer.graph_mode = True
        trainer.tail.opt.device_step(True)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):               # eager warm-up on a side stream (allocator, lazy tables, optimizer state)
            for _ in range(warmup):
                trainer.step(self.past, self.future)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()                    # the warm-up's cached blocks would otherwise double the footprint next to the graph's pool
        self.graph = torch.cuda.CUDAGraph()
        try:     # (the warm-up ran on another stream than the capture; the mismatch warning is about exactly that and harmless here)
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except Exception:
            pass
        with torch.cuda.graph(self.graph, capture_error_mode="relaxed"):
            self.loss = trainer.step(self.past, self.future)
        trainer.tail.opt.note_replayed(-1)          # capturing ran the host code of one step but executed nothing on the device
        self.replays = 0

    def step(self, past, future):
        self.past.copy_(past, non_blocking=True)
        self.future.copy_(future, non_blocking=True)
        self.graph.replay()
        self.trainer.tail.opt.note_replayed()
        self.replays += 1
        return self.loss
