"""Fused tail of the stage-2 training iteration on the kernels of csrc/tail.cu:

  * `mse_gdl_loss(pred, target)`  = MSELoss()(pred, target) + GDL(alpha=1)(target, pred)   (reference model/criterion.py:105-204,
    cal_lossT train_NAR.py:33-36 / train_FAR.py:32-34): one pass for the value, one pass for d/d(pred);
  * `bipatch_nce_normalized(gt_f, pred_f)` = BiPatchNCE(...)(F.normalize(gt_f), F.normalize(pred_f)) (criterion.py:206-259,
    train_NAR.py:36): one forward and one backward kernel, one CTA per frame;
  * `FusedAdamW` -- a torch.optim.AdamW (same constructor, param_groups, state / state_dict layout, so the reference's checkpoint
    code `optimizer_T.state_dict()` / `load_state_dict`, utils/train_summary.py:22-31,139, keeps working) whose `step()` is ONE
    multi-tensor launch, optionally with the `clip_grad_norm_` coefficient (train_NAR.py:85) folded in so clipping costs no pass;
  * `grad_sqnorm(params)` -- the squared global gradient norm with one launch.

No CPU path: CUDA float32 tensors only."""
import torch

from . import _lib, ops

_call = _lib.call


# ----------------------------------------------------------------------------------------------------- MSE + GDL
class _MseGdl(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        if pred.shape != target.shape or pred.dim() < 3:
            raise RuntimeError("vptr_b200.tail.mse_gdl_loss: pred %s / target %s must share a (..., H, W) shape" % (tuple(pred.shape), tuple(target.shape)))
        ops._chk(pred, "pred"); ops._chk(target, "target")
        p, t = pred.contiguous(), target.contiguous()
        H, W = p.shape[-2:]
        planes = p.numel() // (H * W)
        sums = torch.zeros(3, dtype=torch.float64, device=p.device)
        loss3 = torch.empty(3, dtype=torch.float32, device=p.device)
        _call("vptr_mse_gdl_fwd", p.data_ptr(), t.data_ptr(), planes, H, W, sums.data_ptr(), loss3.data_ptr(), ops._s())
        ctx.save_for_backward(p, t)
        ctx.geom = (planes, H, W)
        ctx.mark_non_differentiable(loss3)
        return loss3[0].clone(), loss3

    @staticmethod
    def backward(ctx, dloss, _d3):
        p, t = ctx.saved_tensors
        planes, H, W = ctx.geom
        dl = dloss.contiguous().to(torch.float32)
        dp = torch.empty_like(p)
        _call("vptr_mse_gdl_bwd", p.data_ptr(), t.data_ptr(), dl.data_ptr(), dp.data_ptr(), planes, H, W, ops._s())
        return dp, None


def mse_gdl_loss(pred, target, parts=False):
    """MSE + GDL(alpha=1), both un-weighted means as cal_lossT uses them.  parts=True also returns the (3,) tensor {total, mse, gdl}."""
    loss, loss3 = _MseGdl.apply(pred, target)
    return (loss, loss3) if parts else loss


# ----------------------------------------------------------------------------------------------------- BiPatchNCE
class _BiPatchNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gt_rows, pred_rows, temperature):
        F_, L, C = gt_rows.shape
        ops._chk(gt_rows, "gt features"); ops._chk(pred_rows, "predicted features")
        g, p = gt_rows.contiguous(), pred_rows.contiguous()
        S = torch.empty(F_ * 64 * 64, dtype=torch.float32, device=g.device)
        stats = torch.empty(F_ * 256, dtype=torch.float32, device=g.device)
        acc = torch.zeros(1, dtype=torch.float64, device=g.device)
        loss = torch.empty(1, dtype=torch.float32, device=g.device)
        _call("vptr_bipatch_nce_fwd", g.data_ptr(), p.data_ptr(), F_, L, C, float(temperature), S.data_ptr(), stats.data_ptr(), acc.data_ptr(),
              loss.data_ptr(), ops._s())
        ctx.save_for_backward(g, p, S, stats)
        ctx.temperature = float(temperature)
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        g, p, S, stats = ctx.saved_tensors
        F_, L, C = g.shape
        dl = dloss.contiguous().to(torch.float32)
        dg, dp = torch.empty_like(g), torch.empty_like(p)
        _call("vptr_bipatch_nce_bwd", g.data_ptr(), p.data_ptr(), S.data_ptr(), stats.data_ptr(), dl.data_ptr(), F_, L, C, ctx.temperature,
              dg.data_ptr(), dp.data_ptr(), ops._s())
        return dg, dp, None


def bipatch_nce_normalized(gt_f, pred_f, temperature=1.0):
    """BiPatchNCE(N, T, h, w, temperature)(F.normalize(gt_f, p=2, dim=2), F.normalize(pred_f, p=2, dim=2)) -- the contrastive term of
    cal_lossT exactly as train_NAR.py:36 composes it -- with the channel normalisation, both score matrices, both cross-entropies
    and the reference's stop-gradients in ONE forward and ONE backward kernel.  gt_f / pred_f: (N, T, C, h, w), h*w <= 64."""
    N, T, C, h, w = gt_f.shape
    rows = lambda t: t.permute(0, 1, 3, 4, 2).reshape(N * T, h * w, C)     # free when t is a channel-last view (as the projector emits)
    return _BiPatchNCE.apply(rows(gt_f), rows(pred_f), temperature)


# ----------------------------------------------------------------------------------------------------- multi-tensor tables
class _Table:
    """device int64 table [cols x n pointers][n cumulative unit ends] for the multi-tensor kernels, rebuilt only when a pointer changes"""

    def __init__(self):
        self.key, self.dev, self.n, self.total, self.vec = None, None, 0, 0, 1

    def get(self, cols):
        """cols: list of equal-length tensor lists (same numel per row across columns)"""
        key = tuple(t.data_ptr() for c in cols for t in c)
        if key != self.key:
            first = cols[0]
            vec = all(t.numel() % 4 == 0 for t in first) and all(p % 16 == 0 for p in key)
            unit = 4 if vec else 1
            ends, acc = [], 0
            for t in first:
                acc += t.numel() // unit
                ends.append(acc)
            # (pinned source kept alive until the copy has run -- for good when the copy is captured into a CUDA graph)
            self.dev, self._host = ops.pinned_table(list(key) + ends, first[0].device)
            self.key, self.n, self.total, self.vec = key, len(first), acc, int(vec)
        return self.dev, self.n, self.total, self.vec


def _check(ts, what):
    for t in ts:
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("vptr_b200.tail: %s must be contiguous CUDA float32 tensors (got %s %s); there is no CPU path" % (what, t.device, t.dtype))


_SQ_TABLES = {}


def grad_sqnorm(params, out=None, accumulate=False):
    """(1,) float64 device tensor: sum of squares of every existing .grad -- the square of clip_grad_norm_'s total_norm.
    accumulate: add to `out` instead of overwriting it (part of the norm came from the gradient reducer)"""
    grads = [p.grad for p in params if p.grad is not None]
    dev = grads[0].device if grads else torch.device("cuda")
    if out is None:
        out = torch.zeros(1, dtype=torch.float64, device=dev)
    elif not accumulate:
        out.zero_()
    if not grads:
        return out
    _check(grads, "gradients")
    tab = _SQ_TABLES.setdefault(len(grads), _Table())
    t, n, total, vec = tab.get([grads])
    _call("vptr_sqnorm_multi", t.data_ptr(), n, total, vec, out.data_ptr(), ops._s())
    return out


class FusedAdamW(torch.optim.AdamW):
    """torch.optim.AdamW with a single-launch step.  `step(grad_sqnorm=..., max_norm=...)` additionally applies the
    clip_grad_norm_ coefficient to the gradients inside the same pass (the .grad tensors themselves are left unscaled)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, foreach=False, fused=False)
        self._tables = {}
        self._dev_step = False      # True: the update count lives in device memory (CUDA-graph replay; see device_step())
        self._ctrs = {}

    def device_step(self, on=True):
        """keep the step count on the device (vptr_counter_add + vptr_adamw_multi_dev) so that a captured graph of step() stays correct
        when replayed; call note_replayed() after each replay to keep the host-side state['step'] (checkpoints) in sync"""
        self._dev_step = on
        self._ctrs = {}

    def note_replayed(self, n=1):
        for st in self.state.values():
            if "step" in st:
                st["step"] += n

    @torch.no_grad()
    def step(self, closure=None, grad_sqnorm=None, max_norm=0.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if group.get("amsgrad") or group.get("maximize"):
                raise NotImplementedError("vptr_b200.tail.FusedAdamW: amsgrad / maximize are not implemented")
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            fresh = [p for p in ps if len(self.state[p]) == 0]
            if fresh:   # moments of all new parameters live in two flat buffers (views per parameter)
                tot = sum(p.numel() for p in fresh)
                m = torch.zeros(tot, dtype=torch.float32, device=fresh[0].device)
                v = torch.zeros(tot, dtype=torch.float32, device=fresh[0].device)
                off = 0
                for p in fresh:
                    n = p.numel()
                    self.state[p]["step"] = torch.tensor(0.0, dtype=torch.float32)
                    self.state[p]["exp_avg"] = m[off:off + n].view_as(p)
                    self.state[p]["exp_avg_sq"] = v[off:off + n].view_as(p)
                    off += n
            # parameters may have been stepped a different number of times (e.g. the NCE projector when BiPatchNCE is switched off)
            by_step = {}
            for p in ps:
                st = self.state[p]
                st["step"] += 1
                by_step.setdefault(int(st["step"]), []).append(p)
            b1, b2 = group["betas"]
            for k, (step, sub) in enumerate(sorted(by_step.items())):
                grads = [p.grad for p in sub]
                ms = [self.state[p]["exp_avg"] for p in sub]
                vs = [self.state[p]["exp_avg_sq"] for p in sub]
                _check(sub, "parameters"); _check(grads, "gradients"); _check(ms, "exp_avg"); _check(vs, "exp_avg_sq")
                tab = self._tables.setdefault((gi, k, len(sub)), _Table())
                t, n, total, vec = tab.get([[p.data for p in sub], grads, ms, vs])
                ctr = 0
                if self._dev_step:
                    c = self._ctrs.get((gi, k))
                    if c is None:
                        c = self._ctrs[(gi, k)] = torch.tensor([step - 1], dtype=torch.int64, device=sub[0].device)
                    _call("vptr_counter_add", c.data_ptr(), 1, ops._s())
                    ctr = c.data_ptr()
                _call("vptr_adamw_multi_dev", t.data_ptr(), n, total, vec, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                      float(group["weight_decay"]), step, ctr, 0 if grad_sqnorm is None else grad_sqnorm.data_ptr(), float(max_norm), ops._s())
        return loss


class FusedTail:
    """clip_grad_norm_(params, max_norm) + AdamW step (train_NAR.py:85-86) as two launches: squared norm, fused clip + update"""

    def __init__(self, params, lr=1e-4, max_grad_norm=1.0, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.params = list(params)
        self.max_grad_norm = max_grad_norm
        self.opt = FusedAdamW(self.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self._sq = None

    def clip_and_step(self, reduced_sqnorm=None, rest=None):
        """reduced_sqnorm / rest: from parallel.GradReducer -- the squared norm of the slices it already reduced (accumulated behind
        each all-reduce) and the parameters it did not cover; only those still need a norm pass"""
        if reduced_sqnorm is not None and rest is not None:
            self._sq = grad_sqnorm(rest, reduced_sqnorm, accumulate=True)
        else:
            self._sq = grad_sqnorm(self.params, self._sq if self._sq is not reduced_sqnorm else None)
        self.opt.step(grad_sqnorm=self._sq, max_norm=self.max_grad_norm)
        return self._sq

    def total_norm(self):
        return self._sq.sqrt().to(torch.float32)
