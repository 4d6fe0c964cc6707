"""Forward / backward schedule of the VidHRFormer blocks on token-major fp32 activations, expressed as calls into
libvptr_b200.so (vptr_b200.ops).  No autograd graph is built: every sub-block saves what its hand-written backward
needs, and the whole Transformer is exposed to torch.autograd as ONE Function (vptr_b200.model).

Restates (reference paths under /root/reference):
  VidHRFormerBlockEnc.forward     model/VidHRFormer_modules.py:60-93
  VidHRFormerBlockDecNAR.forward  model/VidHRFormer_modules.py:164-211
  SpatialLocalMultiheadAttention  model/VidHRFormer_modules.py:321-357  (+ MultiheadAttentionRPE, MultiHeadAttentionRPE.py:527-697)
  MlpDWBN.forward                 model/VidHRFormer_modules.py:424-442
  VidHRFormerNAR/FAR.forward      model/VidHRFormer.py:28-53,71-88
Dropout / DropPath (train mode, p > 0) are fused into the producing kernels' epilogues with a counter-based RNG (class Drop);
exact parity with the reference's Philox stream is impossible, so p > 0 is checked statistically and parity is gated at p = 0.
"""
import math

import torch

from . import ops

# Every operand of the tensor-core GEMM is rounded to nearest tf32 where it is produced (or by ops.round_copy):
# tcgen05 kind::tf32 truncates the low 13 mantissa bits, which on raw fp32 data is a biased (towards zero) error that
# compounds through the backward chain; on pre-rounded data it is exact.
ROUND_TF32 = True
RT = ROUND_TF32


# Activation-memory mode.  "fast" keeps every GEMM operand of the forward for the backward (LayerNorm outputs, GELU outputs);
# "lean" keeps only the pre-activation tensors and re-derives those operands in the backward with one extra elementwise pass each
# (LayerNorm / norm+GELU / GELU forward: HBM-bound, ~6 % of a cfg1 step) -- 58 % of the "fast" bytes per decoder block, which is
# what lets cfg4 (16 clips of 10 -> 30 frames at 128x128 per GPU; the reference needs 353 GiB there) fit 180 GB.
# "auto" (default) picks lean when the fast-mode estimate would not fit the device.
MEMORY_MODE = "auto"
FAST_FLOATS_PER_TOKEN = {"enc": 22176, "dec": 35376}      # saved fp32 values per token and block in fast mode


def set_memory_mode(mode):
    global MEMORY_MODE
    if mode not in ("auto", "fast", "lean"):
        raise ValueError("vptr_b200.engine.set_memory_mode: mode must be 'auto', 'fast' or 'lean'")
    MEMORY_MODE = mode


def lean_for(enc_tokens, n_enc, dec_tokens, n_dec, device):
    """decide the mode of one forward pass"""
    if MEMORY_MODE != "auto":
        return MEMORY_MODE == "lean"
    need = 4.0 * (enc_tokens * n_enc * FAST_FLOATS_PER_TOKEN["enc"] + dec_tokens * n_dec * FAST_FLOATS_PER_TOKEN["dec"])
    free, total = torch.cuda.mem_get_info(device)
    reusable = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    return need > 0.85 * (free + reusable) - 8e9          # 8 GB of head room for transients (decoder, loss, workspaces)


class Params:
    """name -> parameter tensor / gradient buffer (views into one flat fp32 buffer when grads are wanted)."""

    GEMM_WEIGHTS = ("proj.weight", "in_proj_weight", "fc1.weight", "fc2.weight", "linear1.weight", "linear2.weight")

    def __init__(self, named, want_grads, rounded=None, holder=None):
        """holder: the owning nn.Module.  Its flat gradient buffer is kept across steps (`holder._vptr_gflat`) and re-zeroed
        instead of re-allocated, so gradient addresses are stable from step to step (the multi-tensor optimizer keeps its pointer
        table, the all-reduce its chunk views) -- but only when no parameter's current .grad still aliases it (i.e. the caller
        did zero_grad(set_to_none=True), as train_NAR.py:60 does); gradient accumulation gets a fresh buffer."""
        self.t = {k: v for k, v in named}
        self.gflat = None
        self.gv = {}
        self.goff = {}
        # tf32-rounded GEMM weights: one multi-tensor launch per step (the backward reuses the forward's copies -- the
        # optimizer only steps after it)
        self.rounded = rounded if rounded is not None else {}
        if ROUND_TF32 and rounded is None:
            names = [k for k, v in self.t.items() if k.endswith(self.GEMM_WEIGHTS) and v.is_cuda and v.numel() % 4 == 0
                     and v.data_ptr() % 16 == 0 and v.is_contiguous()]
            names = _rounding_order(names)
            if names:
                for k, r in zip(names, ops.round_copy_multi([self.t[k] for k in names])):
                    self.rounded[k] = r
        if want_grads:
            names = [k for k, v in self.t.items() if v.requires_grad]
            total = sum(self.t[k].numel() for k in names)
            dev = next(iter(self.t.values())).device
            cached = getattr(holder, "_vptr_gflat", None) if holder is not None else None
            if cached is not None and cached.numel() == total and cached.device == dev:
                base = cached.untyped_storage().data_ptr()
                if all(self.t[k].grad is None or self.t[k].grad.untyped_storage().data_ptr() != base for k in names):
                    self.gflat = cached.zero_()
            if self.gflat is None:
                self.gflat = torch.zeros(total, dtype=torch.float32, device=dev)
                if holder is not None:
                    holder._vptr_gflat = self.gflat
            off = 0
            for k in names:
                n = self.t[k].numel()
                self.gv[k] = self.gflat[off:off + n].view(self.t[k].shape)
                self.goff[k] = (off, off + n)
                off += n

    def w(self, name):
        return self.t[name]

    def g(self, name):
        return self.gv.get(name)

    def wr(self, name):
        """tf32-rounded copy of a weight matrix (made once per forward / backward, shared by all its GEMMs)"""
        if not ROUND_TF32:
            return self.t[name]
        r = self.rounded.get(name)
        if r is None:
            r = self.rounded[name] = ops.round_copy(self.t[name])
        return r


def _rounding_order(names):
    """Order of the weights inside the flat tf32-rounded buffer: parameter order, except that q comes before k before v
    inside each window-attention module (registration order is k, v, q), so that [Wq ; Wk] and [Wk ; Wv] are contiguous row
    blocks of the buffer (_stacked)."""
    rank = {"q_proj.weight": 0, "k_proj.weight": 1, "v_proj.weight": 2}
    pos = {k: i for i, k in enumerate(names)}
    first = {}
    for k in names:
        mod, _, leaf = k.rpartition(".attn.")
        if leaf in rank:
            first[mod] = min(first.get(mod, pos[k]), pos[k])

    def key(k):
        mod, _, leaf = k.rpartition(".attn.")
        return (first[mod], rank[leaf]) if leaf in rank else (pos[k], 0)

    return sorted(names, key=key)


def _stacked(*ws):
    """[sum of rows][cols] view over weight matrices that sit back to back in memory (the tf32-rounded copies of one attention
    module's q / k / v projections do: Params rounds all GEMM weights into one flat buffer in parameter order), else None."""
    w0 = ws[0]
    if any(w.dim() != 2 or not w.is_contiguous() or w.shape[1] != w0.shape[1] or w.dtype != w0.dtype for w in ws):
        return None
    ptr = w0.data_ptr()
    for w in ws:
        if w.data_ptr() != ptr or w.untyped_storage().data_ptr() != w0.untyped_storage().data_ptr():
            return None
        ptr += w.numel() * w.element_size()
    return torch.as_strided(w0, (sum(w.shape[0] for w in ws), w0.shape[1]), (w0.shape[1], 1), w0.storage_offset())


# Bias gradients (column sums of a dY) are produced by the kernel that writes dY when dY is wide (the 2112-channel tensors of the two
# FFNs: gelu_bwd, norm+GELU backward) -- measured 34-66 us cheaper per site than a separate pass.  For the 528-wide tensors and inside
# the attention backward the fused forms LOSE (a 132-float4 row leaves the column-slab kernels with 160-thread blocks; shared-memory
# atomics in the attention tile store cost +40 %): profiles/r02_op_table_cfg1_colsum_fusion.txt.  Those sites keep vptr_colsum.
FUSE_COLSUM_MIN_WIDTH = 1024


def _rc(x, rowscale=None, group_elems=0, seed=0, p=0.0, colsum=None):
    """GEMM-operand copy of x: tf32-rounded and, for the backward of a regularised branch, masked / DropPath-scaled.
    colsum: bias-gradient accumulator that receives the column sums of the copy"""
    if not ROUND_TF32 and rowscale is None and p <= 0.0:
        if colsum is not None:
            ops.colsum(x, colsum)
        return x
    if colsum is not None and x.shape[-1] < FUSE_COLSUM_MIN_WIDTH:
        y = ops.round_copy(x, ROUND_TF32, rowscale, group_elems, seed, p)
        ops.colsum(y, colsum)
        return y
    return ops.round_copy(x, ROUND_TF32, rowscale, group_elems, seed, p, colsum=colsum)


class Drop:
    """Dropout / DropPath state of one forward call (train mode, p > 0).  Every site draws its own 64-bit seed; masks are
    regenerated from (seed, element index) in the backward.  drop_path rate == dropout rate (VPTR_modules.py:114,170).
    DropPath granularity follows the reference site by site: per clip where it is applied to (N,T,H,W,C) tensors (window attention,
    conv FFNs), per future-frame index where it is applied to the seq-first (T2, N*H*W, C) tensor (encoder-decoder attention)."""

    def __init__(self, p, n_clips, device):
        self.p, self.n_clips, self.device = float(p), n_clips, device
        self.p_path = float(p)
        self.base = int(torch.randint(0, 2 ** 62, (1,)).item())
        self.count = 0

    def seed(self):
        self.count += 1
        return (self.base + self.count * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF

    def path(self):
        """(n_clips,) DropPath keep-scales: one Bernoulli per clip, shape (N,1,1,1,1) in the reference"""
        if self.p_path <= 0.0:
            self.seed()
            return None
        return ops.droppath_scales(self.n_clips, self.seed(), self.p_path, self.device)

    def path_per_t(self, T):
        """(n_clips*T,) per-frame keep-scales for the encoder-decoder attention: the reference applies drop_path1 to the
        (T2, N*H*W, C) seq-first tensor there (VidHRFormer_modules.py:200-205), and drop_path draws along dim 0
        (:563-575) -- i.e. ONE Bernoulli per future-frame index t, shared by every clip and pixel.  Reproduced as such."""
        if self.p_path <= 0.0:
            self.seed()
            return None
        return ops.droppath_scales(T, self.seed(), self.p_path, self.device).repeat(self.n_clips)


class _NoDrop:
    p = 0.0

    def seed(self):
        return 0

    def path(self):
        return None

    def path_per_t(self, T):
        return None


NO_DROP = _NoDrop()


class exact_fp32:
    """Verification mode (tests only): every contraction runs on the fp32 FFMA kernel (vptr_gemm_simt) and nothing is
    rounded to tf32, so the forward/backward schedule can be checked against the reference at fp32 accuracy -- the tf32
    tensor-core path cannot be, because a ~5e-4 forward perturbation flips ReLU masks and moves gradients by percents."""

    def __enter__(self):
        global ROUND_TF32, RT
        self.prev = (ROUND_TF32, ops.FORCE_SIMT)
        ROUND_TF32 = RT = False
        ops.FORCE_SIMT = True
        return self

    def __exit__(self, *a):
        global ROUND_TF32, RT
        ROUND_TF32 = RT = self.prev[0]
        ops.FORCE_SIMT = self.prev[1]
        return False


class precise_3xtf32:
    """Verification mode (tests only): every contraction still runs on the product's tcgen05 GEMM kernel, but with 3xTF32 operand
    precision (ops._gemm_3xtf32: hi/lo tf32 planes concatenated along K), nothing is pre-rounded to tf32, and the attention cores
    run on the 3xTF32 mma.sync kernels.  Isolates operand rounding from everything else the product path does."""

    def __enter__(self):
        global ROUND_TF32, RT
        self.prev = (ROUND_TF32, ops.GEMM_3XTF32, ops.ATTN_TC)
        ROUND_TF32 = RT = False
        ops.GEMM_3XTF32 = True
        ops.ATTN_TC = False
        return self

    def __exit__(self, *a):
        global ROUND_TF32, RT
        ROUND_TF32 = RT = self.prev[0]
        ops.GEMM_3XTF32, ops.ATTN_TC = self.prev[1], self.prev[2]
        return False


class Geom:
    def __init__(self, N, T, H, W, C, nhead, ws, lean=False):
        self.N, self.T, self.H, self.W, self.C, self.nhead, self.ws = N, T, H, W, C, nhead, ws
        self.lean = lean
        self.d = C // nhead
        self.HW = H * W
        self.F = N * T
        self.R = N * T * H * W
        self.scale = float(self.d) ** -0.5
        pad_h = math.ceil(H / ws) * ws - H
        pad_w = math.ceil(W / ws) * ws - W
        self.ph0, self.pw0 = pad_h // 2, pad_w // 2      # PadBlock: pad//2 before, the rest after (VidHRFormer_modules.py:538-550)
        self.Hp, self.Wp = H + pad_h, W + pad_w
        self.padded = pad_h + pad_w > 0


# Weight-gradient GEMMs are off the backward's critical path (nothing downstream reads dW before the optimizer), so they are issued on
# a SIDE stream: while the tensor cores work through dY^T X, the main stream's HBM-bound kernels (norm + GELU backward, LayerNorm
# backward, attention backward, depthwise conv) run on the same SMs next to the GEMM's one CTA per SM.  Two tcgen05 GEMMs cannot
# share an SM (shared memory / TMEM), so a side GEMM simply alternates with the main stream's input-gradient GEMMs.  The main stream
# joins the side work of sub-block i at the end of sub-block i+1 (lagged: the trailing weight gradients of a sub-block overlap with
# the head of the next one); operands stay referenced until then so the caching allocator cannot hand their memory out early.
# Inside a captured CUDA graph this is an ordinary fork / join of the capture.
# MEASURED (profiles/r02_bench_cfg1_wgrad_overlap.json.log): no gain -- cfg1 166.5 ms with the fork vs 164.8 ms without, cfg2 85.2 vs
# 85.3 ms.  A side GEMM holds one CTA (210 KB of shared memory) on every SM, so the main stream's next input-gradient GEMM queues
# behind it and the critical path stretches by what the overlap saved; the HBM-bound kernels next to it slow down by their share of
# the bandwidth.  Kept as an opt-in switch, OFF by default.
OVERLAP_WGRAD = False


_SIDE_STREAMS = {}      # one side stream per device, created once (never inside a graph capture: the warm-up steps come first)


def _side_stream(P):
    st = getattr(P, "side", None)
    if st is None:
        dev = torch.cuda.current_device()
        st = _SIDE_STREAMS.get(dev)
        if st is None:
            st = _SIDE_STREAMS[dev] = torch.cuda.Stream()
        P.side = st
        P.side_refs, P.side_prev = [], None
    return st


def _wgrad_raw(P, dY, X, out):
    """out += dY^T X on the side stream (see OVERLAP_WGRAD)"""
    if not (OVERLAP_WGRAD and dY.is_cuda):
        ops.gemm(dY, X, out=out, a_mn=True, b_mn=True, accumulate=True)
        return
    st = _side_stream(P)
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        ops.gemm(dY, X, out=out, a_mn=True, b_mn=True, accumulate=True)
    P.side_refs.append((dY, X))


def _entry_done(P):
    """end of one sub-block's backward: lagged join of the side stream (waits for the PREVIOUS sub-block's weight gradients)"""
    if getattr(P, "side", None) is None:
        return
    ev = torch.cuda.Event()
    ev.record(P.side)
    if P.side_prev is not None:
        torch.cuda.current_stream().wait_event(P.side_prev[0])
        P.side_prev[1].clear()
    P.side_prev, P.side_refs = (ev, P.side_refs), []


def join_side(P):
    """full join: every weight gradient issued so far is ordered before whatever the main stream does next"""
    if getattr(P, "side", None) is None:
        return
    torch.cuda.current_stream().wait_stream(P.side)
    P.side_refs, P.side_prev = [], None


def _wgrad(P, name, dY, X):
    """dW[name] += dY^T X (contraction over tokens; both operands read as stored)."""
    g = P.g(name)
    if g is not None:
        _wgrad_raw(P, dY, X, g.view(g.shape[0], -1))


def _bgrad(P, name, dY):
    g = P.g(name)
    if g is not None:
        ops.colsum(dY, g)


# =================================================================================================== window attention
def _win_inputs(P, ln, x, g, rpe, qpos, lw_tab):
    """(v source, q/k source, mean, rstd): LN(x) [+ query_pos] [padded] [+ lw_pos]; also re-run by the lean-mode backward"""
    lnw, lnb = P.w(ln + ".weight"), P.w(ln + ".bias")
    if qpos is not None:
        a, aq, mean, rstd = ops.layernorm_fwd(x, lnw, lnb, add=qpos, add_div=1, add_mod=qpos.shape[0], round_tf32=RT)
    else:
        a, aq, mean, rstd = ops.layernorm_fwd(x, lnw, lnb, round_tf32=RT)
        aq = a
    Fr = g.F
    if g.padded:
        a_in = ops.pad_hw(a, Fr, g.H, g.W, g.Hp, g.Wp, g.ph0, g.pw0)
        aq_in = a_in if aq is a else ops.pad_hw(aq, Fr, g.H, g.W, g.Hp, g.Wp, g.ph0, g.pw0)
    else:
        a_in, aq_in = a, aq
    if not rpe:   # VidHRFormer_modules.py:341: q = k = x + lw_pos (per in-window position), v = x
        aq_in = ops.add_rows(aq_in, lw_tab, 1, lw_tab.shape[0], round_tf32=RT)
    return a_in, aq_in, mean, rstd


def window_attn_fwd(P, pre, ln, x, g, rpe, qpos, lw_tab, save, D=NO_DROP):
    """x (R,C) -> x + SLMHSA(LN(x)).  qpos (T*H*W, C) or None: q/k source = LN(x)+qpos (decoder)."""
    C = g.C
    Fr = g.F
    a_in, aq_in, mean, rstd = _win_inputs(P, ln, x, g, rpe, qpos, lw_tab)
    Rp = a_in.shape[0]
    qkv = ops.empty(Rp, 3 * C, like=x)
    o = ops.empty(Rp, C, like=x)
    at = pre + ".attn."
    # tcgen05 / TMA / TMEM forward for the path's window shape: its operands enter the tensor core as TF32, so q / k / v are
    # rounded to nearest tf32 by the projections' epilogues (then the hardware's mantissa truncation is exact and unbiased)
    use_tc = RT and ops.attn_tc_window_ok(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, g.Hp, g.Wp, g.ws, g.nhead, g.d)
    if rpe:
        ops.gemm(aq_in, P.wr(at + "q_proj.weight"), out=qkv[:, :C], bias=P.w(at + "q_proj.bias"), round_tf32=use_tc)
        ops.gemm(aq_in, P.wr(at + "k_proj.weight"), out=qkv[:, C:2 * C], bias=P.w(at + "k_proj.bias"), round_tf32=use_tc)
        ops.gemm(a_in, P.wr(at + "v_proj.weight"), out=qkv[:, 2 * C:], bias=P.w(at + "v_proj.bias"), round_tf32=use_tc)
        table = P.w(at + "relative_position_bias_table")
        wo, bo = P.wr(at + "out_proj.weight"), P.w(at + "out_proj.bias")
    else:
        Wi, bi = P.wr(at + "in_proj_weight"), P.w(at + "in_proj_bias")
        ops.gemm(aq_in, Wi[:2 * C], out=qkv[:, :2 * C], bias=bi[:2 * C], round_tf32=use_tc)
        ops.gemm(a_in, Wi[2 * C:], out=qkv[:, 2 * C:], bias=bi[2 * C:], round_tf32=use_tc)
        table = None
        wo, bo = P.wr(at + "out_proj.weight"), P.w(at + "out_proj.bias")
    s_attn, dp = D.seed(), D.path()
    rpg = g.T * g.Hp * g.Wp                      # rows per clip (DropPath is per clip)
    (ops.attn_fwd_tcgen05 if use_tc else ops.attn_fwd)(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, table, 0, Fr, g.Hp, g.Wp, g.ws, 0, 0,
                                                       g.nhead, g.d, False, g.scale, round_tf32=RT, drop_seed=s_attn, drop_p=D.p)
    if g.padded:
        yp = ops.gemm(o, wo, bias=bo, rowscale=dp, rows_per_group=rpg)
        y = ops.crop_hw(yp, Fr, g.H, g.W, g.Hp, g.Wp, g.ph0, g.pw0)
        out = ops.axpby(x, y)
    else:
        out = ops.gemm(o, wo, bias=bo, residual=x, rowscale=dp, rows_per_group=rpg)
    if save is not None:
        keep = not g.lean
        save.append(("win", dict(pre=pre, ln=ln, x=x, mean=mean, rstd=rstd, a_in=a_in if keep else None, aq_in=aq_in if keep else None,
                                 qkv=qkv, o=o, rpe=rpe, qpos=qpos, lw_tab=lw_tab, has_qpos=qpos is not None, g=g, s_attn=s_attn, dp=dp,
                                 rpg=rpg, p=D.p)))
    return out


def window_attn_bwd(P, s, dout, dqpos):
    g, C, pre, ln = s["g"], s["g"].C, s["pre"], s["ln"]
    at = pre + ".attn."
    Fr = g.F
    dy = _rc(ops.pad_hw(dout, Fr, g.H, g.W, g.Hp, g.Wp, g.ph0, g.pw0) if g.padded else dout, s["dp"], s["rpg"] * C,
             colsum=P.g(at + "out_proj.bias"))
    qkv, o = s["qkv"], s["o"]
    _wgrad(P, at + "out_proj.weight", dy, o)
    do = ops.gemm(dy, P.wr(at + "out_proj.weight"), b_mn=True)
    dqkv = torch.empty_like(qkv)
    if s["rpe"]:
        table = P.w(at + "relative_position_bias_table")
        dtable = P.g(at + "relative_position_bias_table")
    else:
        table = dtable = None
    ops.attn_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], do, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], table, dtable, 0, Fr,
                 g.Hp, g.Wp, g.ws, 0, 0, g.nhead, g.d, False, g.scale, round_tf32=RT, drop_seed=s["s_attn"], drop_p=s["p"])
    a_in, aq_in = s["a_in"], s["aq_in"]
    if a_in is None:      # lean mode: the projections' operands are re-derived from the block input
        a_in, aq_in, _, _ = _win_inputs(P, ln, s["x"], g, s["rpe"], s["qpos"], s["lw_tab"])
    if s["rpe"]:
        for i, nm in enumerate(("q_proj", "k_proj", "v_proj")):
            src = a_in if nm == "v_proj" else aq_in
            _wgrad(P, at + nm + ".weight", dqkv[:, i * C:(i + 1) * C], src)
            _bgrad(P, at + nm + ".bias", dqkv[:, i * C:(i + 1) * C])
        wq, wk, wv = P.wr(at + "q_proj.weight"), P.wr(at + "k_proj.weight"), P.wr(at + "v_proj.weight")
        wqk = _stacked(wq, wk)
        if wqk is not None:      # d(aq) = [dq | dk] [Wq ; Wk]: one contraction over 2C instead of two GEMMs chained through a residual
            d_aq = ops.gemm(dqkv[:, :2 * C], wqk, b_mn=True)
        else:
            d_aq = ops.gemm(dqkv[:, :C], wq, b_mn=True)
            ops.gemm(dqkv[:, C:2 * C], wk, b_mn=True, out=d_aq, residual=d_aq)
        d_a = ops.gemm(dqkv[:, 2 * C:], wv, b_mn=True)
    else:
        Wi = P.wr(at + "in_proj_weight")
        gW, gb = P.g(at + "in_proj_weight"), P.g(at + "in_proj_bias")
        if gW is not None:
            _wgrad_raw(P, dqkv[:, :2 * C], aq_in, gW[:2 * C])
            _wgrad_raw(P, dqkv[:, 2 * C:], a_in, gW[2 * C:])
            ops.colsum(dqkv, gb)
        d_aq = ops.gemm(dqkv[:, :2 * C], Wi[:2 * C], b_mn=True)
        d_a = ops.gemm(dqkv[:, 2 * C:], Wi[2 * C:], b_mn=True)
    if g.padded:
        d_aq = ops.crop_hw(d_aq, Fr, g.H, g.W, g.Hp, g.Wp, g.ph0, g.pw0)
        d_a = ops.crop_hw(d_a, Fr, g.H, g.W, g.Hp, g.Wp, g.ph0, g.pw0)
    if s["has_qpos"] and dqpos is not None:
        ops.rowgroup_sum(d_aq, dqpos, g.N)
    return ops.layernorm_bwd(d_a, d_aq, s["x"], P.w(ln + ".weight"), P.w(ln + ".bias"), s["mean"], s["rstd"], dout,
                             P.g(ln + ".weight"), P.g(ln + ".bias"))


# =================================================================================================== conv FFN (MlpDWBN)
_FFN_NORMS = ("norm1", "norm2", "norm3")


def _ffn_layout(P, pre, hw, layer_norm):
    """[(parameter name, rows, cols)] of the MlpDWBN parameters the kernels read in another layout than nn.Module stores them:
    the depthwise 3x3 weights (Ch,1,3,3) -> [9][Ch], and -- LayerNorm((ch,H,W)) flavour -- the three norms' affine (ch,H,W) -> [hw][ch]"""
    items = []
    if layer_norm:
        for n in _FFN_NORMS:
            ch = P.w(pre + "." + n + ".weight").shape[0]
            items += [(pre + "." + n + ".weight", ch, hw), (pre + "." + n + ".bias", ch, hw)]
    items.append((pre + ".dw3x3.weight", P.w(pre + ".dw3x3.weight").shape[0], 9))
    return items


def _ffn_params_cl(P, pre, hw, layer_norm):
    """engine-layout copies of those parameters: ONE batched transpose launch per block (was seven)"""
    items = _ffn_layout(P, pre, hw, layer_norm)
    flat = ops.empty(sum(r * c for _, r, c in items), like=P.w(items[0][0]))
    out, off, entries = {}, 0, []
    for name, r, c in items:
        out[name] = flat[off:off + r * c]
        entries.append((P.w(name), out[name], r, c))
        off += r * c
    ops.transpose_multi((pre, "fwd", flat.device.index), entries)
    aff = tuple((out[pre + "." + n + ".weight"], out[pre + "." + n + ".bias"]) if layer_norm
                else (P.w(pre + "." + n + ".weight"), P.w(pre + "." + n + ".bias")) for n in _FFN_NORMS)
    return aff, out[pre + ".dw3x3.weight"]


def _ffn_norm_stats(P, pre, name, h, g, layer_norm, training, bufs):
    if layer_norm:
        return ops.group_stats(h, g.F)
    rm, rv = bufs[pre + "." + name + ".running_mean"], bufs[pre + "." + name + ".running_var"]
    if training:
        bufs[pre + "." + name + ".num_batches_tracked"].add_(1)
        return ops.bn_stats(h, rm, rv)
    return ops.bn_eval_stats(rm, rv)


def conv_ffn_fwd(P, bufs, pre, ln, x, g, layer_norm, training, save, D=NO_DROP):
    """x (R,C) -> x + MlpDWBN(LN(x)).  layer_norm: LayerNorm((ch,H,W)) flavour (FAR, NAR decoder) else BatchNorm2d."""
    mode = 1 if layer_norm else 0
    b_, _, mean, rstd = ops.layernorm_fwd(x, P.w(ln + ".weight"), P.w(ln + ".bias"), round_tf32=RT)
    w1 = P.wr(pre + ".fc1.weight")
    Ch = w1.shape[0]
    h1 = ops.gemm(b_, w1.view(Ch, g.C), bias=P.w(pre + ".fc1.bias"))
    st1 = _ffn_norm_stats(P, pre, "norm1", h1, g, layer_norm, training, bufs)
    ((g1, b1), (g2, b2), (g3, b3)), w9 = _ffn_params_cl(P, pre, g.HW, layer_norm)
    u1 = ops.norm_act_fwd(h1, st1[0], st1[1], g1, b1, g.HW, mode)
    if layer_norm:   # frame-LayerNorm statistics of the conv output come out of the conv kernel itself
        h2, st2 = ops.dwconv3x3_stats(u1, w9, P.w(pre + ".dw3x3.bias"), g.F, g.H, g.W)
    else:
        h2 = ops.dwconv3x3(u1, w9, P.w(pre + ".dw3x3.bias"), g.F, g.H, g.W)
        st2 = _ffn_norm_stats(P, pre, "norm2", h2, g, layer_norm, training, bufs)
    s2, s3, dp = D.seed(), D.seed(), D.path()
    rpg = g.T * g.HW
    u2 = ops.norm_act_fwd(h2, st2[0], st2[1], g2, b2, g.HW, mode, round_tf32=ROUND_TF32, drop_seed=s2, drop_p=D.p)
    w2 = P.wr(pre + ".fc2.weight")
    h3 = ops.gemm(u2, w2.view(g.C, Ch), bias=P.w(pre + ".fc2.bias"))
    st3 = _ffn_norm_stats(P, pre, "norm3", h3, g, layer_norm, training, bufs)
    out = ops.norm_act_fwd(h3, st3[0], st3[1], g3, b3, g.HW, mode, res=x, rowscale=dp, rows_per_group=rpg, drop_seed=s3, drop_p=D.p)
    if save is not None:
        bwd_mode = mode if (layer_norm or training) else 2
        if g.lean:
            b_ = u1 = u2 = None
        save.append(("ffn", dict(pre=pre, ln=ln, x=x, mean=mean, rstd=rstd, b=b_, h1=h1, st1=st1, u1=u1, h2=h2, st2=st2, u2=u2, h3=h3,
                                 st3=st3, g=g, layer_norm=layer_norm, mode=bwd_mode, w9=w9, aff=((g1, b1), (g2, b2), (g3, b3)),
                                 s2=s2, s3=s3, dp=dp, rpg=rpg, p=D.p)))
    return out


def _ffn_grads_cl(P, pre, hw, layer_norm, like):
    """zeroed engine-layout accumulators for the gradients of _ffn_layout's parameters (one buffer, one fill) and the entries
    that fold them back into the flat gradient buffer (one batched transpose-accumulate at the end of the block's backward)"""
    items = _ffn_layout(P, pre, hw, layer_norm)
    flat = ops.zeros(sum(r * c for _, r, c in items), like=like)
    acc, off, entries = {}, 0, []
    for name, r, c in items:
        acc[name] = flat[off:off + r * c]
        if P.g(name) is not None:
            entries.append((acc[name], P.g(name), c, r))
        off += r * c
    return acc, entries


def _ffn_norm_bwd(P, pre, name, dy, h, st, aff, g, layer_norm, mode, acc, rnd=False, drop=None, inplace=False, colsum=None):
    """colsum: bias-gradient accumulator of the 1x1 conv in front of this norm (its output gradient is the dx computed here);
    acc: the block's engine-layout gradient accumulators (_ffn_grads_cl)"""
    ch = h.shape[1]
    if colsum is not None and ch < FUSE_COLSUM_MIN_WIDTH:      # narrow tensor: separate column-sum pass (see FUSE_COLSUM_MIN_WIDTH)
        dx = _ffn_norm_bwd(P, pre, name, dy, h, st, aff, g, layer_norm, mode, acc, rnd, drop, inplace, None)
        ops.colsum(dx, colsum)
        return dx
    drop = dict(drop or {}, colsum=colsum)
    if layer_norm:      # frame-LayerNorm affine gradients accumulate [hw][ch]; conv_ffn_bwd folds them back in one launch
        return ops.norm_act_bwd(dy, h, st[0], st[1], aff[0], aff[1], acc[pre + "." + name + ".weight"], acc[pre + "." + name + ".bias"],
                                g.HW, mode, round_tf32=rnd, inplace=inplace, **drop)
    gw, gb = P.g(pre + "." + name + ".weight"), P.g(pre + "." + name + ".bias")
    if gw is None:
        gw, gb = ops.zeros(ch, like=h), ops.zeros(ch, like=h)
    return ops.norm_act_bwd(dy, h, st[0], st[1], aff[0], aff[1], gw, gb, g.HW, mode, round_tf32=rnd, inplace=inplace, **drop)


def conv_ffn_bwd(P, s, dout):
    g, pre, ln, lnm, mode = s["g"], s["pre"], s["ln"], s["layer_norm"], s["mode"]
    Ch = s["h1"].shape[1]
    acc, fold = _ffn_grads_cl(P, pre, g.HW, lnm, dout)
    dh3 = _ffn_norm_bwd(P, pre, "norm3", dout, s["h3"], s["st3"], s["aff"][2], g, lnm, mode, acc, rnd=RT,
                        drop=dict(rowscale=s["dp"], rows_per_group=s["rpg"], drop_seed=s["s3"], drop_p=s["p"]), colsum=P.g(pre + ".fc2.bias"))
    u2 = s["u2"]
    if u2 is None:        # lean mode: u2 = drop(GELU(norm2(h2))) again, same dropout seed
        u2 = ops.norm_act_fwd(s["h2"], s["st2"][0], s["st2"][1], s["aff"][1][0], s["aff"][1][1], g.HW, 1 if lnm else 0,
                              round_tf32=ROUND_TF32, drop_seed=s["s2"], drop_p=s["p"])
    _wgrad(P, pre + ".fc2.weight", dh3, u2)
    del u2
    du2 = ops.gemm(dh3, P.wr(pre + ".fc2.weight").view(g.C, Ch), b_mn=True)
    dh2 = _ffn_norm_bwd(P, pre, "norm2", du2, s["h2"], s["st2"], s["aff"][1], g, lnm, mode, acc, drop=dict(drop_seed=s["s2"], drop_p=s["p"]),
                        inplace=True)
    gdw, gdb = P.g(pre + ".dw3x3.weight"), P.g(pre + ".dw3x3.bias")
    if gdw is not None:
        dw9 = acc[pre + ".dw3x3.weight"]
        u1 = s["u1"]
        if u1 is None:    # lean mode
            u1 = ops.norm_act_fwd(s["h1"], s["st1"][0], s["st1"][1], s["aff"][0][0], s["aff"][0][1], g.HW, 1 if lnm else 0)
        ops.dwconv3x3_wgrad(u1, dh2, dw9, gdb, g.F, g.H, g.W)
        del u1
    du1 = ops.dwconv3x3(dh2, s["w9"], None, g.F, g.H, g.W, flip=True)
    dh1 = _ffn_norm_bwd(P, pre, "norm1", du1, s["h1"], s["st1"], s["aff"][0], g, lnm, mode, acc, rnd=RT, inplace=True, colsum=P.g(pre + ".fc1.bias"))
    if fold:            # engine-layout gradients -> (ch,H,W) / (Ch,1,3,3) slices of the flat gradient buffer, one launch
        ops.transpose_multi((pre, "bwd", dout.device.index), fold, accumulate=True)
    b_ = s["b"]
    if b_ is None:        # lean mode
        b_ = ops.layernorm_fwd(s["x"], P.w(ln + ".weight"), P.w(ln + ".bias"), round_tf32=RT, save_stats=False)[0]
    _wgrad(P, pre + ".fc1.weight", dh1, b_)
    del b_
    db = ops.gemm(dh1, P.wr(pre + ".fc1.weight").view(Ch, g.C), b_mn=True)
    return ops.layernorm_bwd(db, None, s["x"], P.w(ln + ".weight"), P.w(ln + ".bias"), s["mean"], s["rstd"], dout,
                             P.g(ln + ".weight"), P.g(ln + ".bias"))


# =================================================================================================== temporal self-attention
def temporal_attn_fwd(P, pre, ln, x, g, pos, causal, save, D=NO_DROP):
    """x + MHA_t(q=k=LN(x)+pos_t, v=LN(x)) per pixel (VidHRFormer_modules.py:74-84,183-187)."""
    C = g.C
    z, zp, mean, rstd = ops.layernorm_fwd(x, P.w(ln + ".weight"), P.w(ln + ".bias"), add=pos, add_div=g.HW, add_mod=g.T, round_tf32=RT)
    Wi, bi = P.wr(pre + ".in_proj_weight"), P.w(pre + ".in_proj_bias")
    qkv = ops.empty(g.R, 3 * C, like=x)
    o = ops.empty(g.R, C, like=x)
    use_tc = RT and ops.attn_tc_temporal_ok(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, g.T, g.T, g.nhead, g.d)   # tcgen05 forward
    ops.gemm(zp, Wi[:2 * C], out=qkv[:, :2 * C], bias=bi[:2 * C], round_tf32=use_tc)
    ops.gemm(z, Wi[2 * C:], out=qkv[:, 2 * C:], bias=bi[2 * C:], round_tf32=use_tc)
    s_attn, s1 = D.seed(), D.seed()
    (ops.attn_fwd_tcgen05 if use_tc else ops.attn_fwd)(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, None, 1, g.N, g.H, g.W, 0, g.T, g.T,
                                                       g.nhead, g.d, causal, g.scale, round_tf32=RT, drop_seed=s_attn, drop_p=D.p)
    out = ops.gemm(o, P.wr(pre + ".out_proj.weight"), bias=P.w(pre + ".out_proj.bias"), residual=x, drop_seed=s1, drop_p=D.p)
    if save is not None:
        if g.lean:
            z = zp = None
        save.append(("tattn", dict(pre=pre, ln=ln, x=x, mean=mean, rstd=rstd, z=z, zp=zp, pos=pos, qkv=qkv, o=o, causal=causal, g=g,
                                   s_attn=s_attn, s1=s1, p=D.p)))
    return out


def temporal_attn_bwd(P, s, dout):
    g, C, pre, ln = s["g"], s["g"].C, s["pre"], s["ln"]
    qkv, o = s["qkv"], s["o"]
    dres, dout = dout, _rc(dout, seed=s["s1"], p=s["p"], colsum=P.g(pre + ".out_proj.bias"))
    _wgrad(P, pre + ".out_proj.weight", dout, o)
    do = ops.gemm(dout, P.wr(pre + ".out_proj.weight"), b_mn=True)
    dqkv = torch.empty_like(qkv)
    gW, gb = P.g(pre + ".in_proj_weight"), P.g(pre + ".in_proj_bias")
    ops.attn_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], do, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], None, None, 1, g.N,
                 g.H, g.W, 0, g.T, g.T, g.nhead, g.d, s["causal"], g.scale, round_tf32=RT, drop_seed=s["s_attn"], drop_p=s["p"])
    Wi = P.wr(pre + ".in_proj_weight")
    if gW is not None:
        z, zp = s["z"], s["zp"]
        if z is None:     # lean mode
            z, zp, _, _ = ops.layernorm_fwd(s["x"], P.w(ln + ".weight"), P.w(ln + ".bias"), add=s["pos"], add_div=g.HW, add_mod=g.T,
                                            round_tf32=RT, save_stats=False)
        _wgrad_raw(P, dqkv[:, :2 * C], zp, gW[:2 * C])
        _wgrad_raw(P, dqkv[:, 2 * C:], z, gW[2 * C:])
        del z, zp
        ops.colsum(dqkv, gb)
    dzp = ops.gemm(dqkv[:, :2 * C], Wi[:2 * C], b_mn=True)
    dz = ops.gemm(dqkv[:, 2 * C:], Wi[2 * C:], b_mn=True)
    return ops.layernorm_bwd(dz, dzp, s["x"], P.w(ln + ".weight"), P.w(ln + ".bias"), s["mean"], s["rstd"], dres,
                             P.g(ln + ".weight"), P.g(ln + ".bias"))


# =================================================================================================== MLP FFN
def mlp_fwd(P, pre, ln, x, g, save, D=NO_DROP):
    """x + linear2(GELU(linear1(LN(x)))) (VidHRFormer_modules.py:87-89,190-192); pre = block prefix."""
    y, _, mean, rstd = ops.layernorm_fwd(x, P.w(ln + ".weight"), P.w(ln + ".bias"), round_tf32=RT)
    h = ops.gemm(y, P.wr(pre + ".linear1.weight"), bias=P.w(pre + ".linear1.bias"))
    s2, s3 = D.seed(), D.seed()
    u = ops.gelu_fwd(h, round_tf32=ROUND_TF32, drop_seed=s2, drop_p=D.p)
    out = ops.gemm(u, P.wr(pre + ".linear2.weight"), bias=P.w(pre + ".linear2.bias"), residual=x, drop_seed=s3, drop_p=D.p)
    if save is not None:
        if g.lean:
            y = u = None
        save.append(("mlp", dict(pre=pre, ln=ln, x=x, mean=mean, rstd=rstd, y=y, h=h, u=u, g=g, s2=s2, s3=s3, p=D.p)))
    return out


def mlp_bwd(P, s, dout):
    pre, ln = s["pre"], s["ln"]
    dres, dout = dout, _rc(dout, seed=s["s3"], p=s["p"], colsum=P.g(pre + ".linear2.bias"))
    u = s["u"]
    if u is None:         # lean mode
        u = ops.gelu_fwd(s["h"], round_tf32=ROUND_TF32, drop_seed=s["s2"], drop_p=s["p"])
    _wgrad(P, pre + ".linear2.weight", dout, u)
    del u
    du = ops.gemm(dout, P.wr(pre + ".linear2.weight"), b_mn=True)
    dh = ops.gelu_bwd(du, s["h"], out=du, round_tf32=RT, drop_seed=s["s2"], drop_p=s["p"], colsum=P.g(pre + ".linear1.bias"))
    y = s["y"]
    if y is None:         # lean mode
        y = ops.layernorm_fwd(s["x"], P.w(ln + ".weight"), P.w(ln + ".bias"), round_tf32=RT, save_stats=False)[0]
    _wgrad(P, pre + ".linear1.weight", dh, y)
    del y
    dy = ops.gemm(dh, P.wr(pre + ".linear1.weight"), b_mn=True)
    return ops.layernorm_bwd(dy, None, s["x"], P.w(ln + ".weight"), P.w(ln + ".bias"), s["mean"], s["rstd"], dres,
                             P.g(ln + ".weight"), P.g(ln + ".bias"))


# =================================================================================================== encoder-decoder attention
def cross_attn_fwd(P, pre, ln, x, g, gm, qadd, mem, mem_k, save, D=NO_DROP, tslma=False):
    """x + MHA(q = LN(x)+query_pos+pos_future, k = memory+pos_past, v = memory) per pixel (VidHRFormer_modules.py:200-206).
    g: geometry of the target stream, gm: of the memory stream; qadd (T2*H*W, C) = query_pos + pos_future.
    tslma: the same projections around the temporal-spatial window attention instead (TemporalSpatialLocalMultiheadAttention,
    :194-198,219-284): per window, the queries of each future frame attend that window of every memory frame (attention mode 2);
    qadd = query_pos + Tlw_pos[future], mem_k = memory + Tlw_pos[past]; DropPath per clip (the tensor is (N,T2,H,W,C) there)."""
    C = g.C
    _, zq, mean, rstd = ops.layernorm_fwd(x, P.w(ln + ".weight"), P.w(ln + ".bias"), want_y=False, add=qadd, add_div=1, add_mod=qadd.shape[0], round_tf32=RT)
    Wi, bi = P.wr(pre + ".in_proj_weight"), P.w(pre + ".in_proj_bias")
    kv = ops.empty(gm.R, 2 * C, like=x)
    o = ops.empty(g.R, C, like=x)
    use_tc = RT and not tslma and ops.attn_tc_temporal_ok(o, kv[:, :C], kv[:, C:], o, g.T, gm.T, g.nhead, g.d)   # tcgen05 forward (q has o's layout)
    q = ops.gemm(zq, Wi[:C], bias=bi[:C], round_tf32=use_tc)
    ops.gemm(mem_k, Wi[C:2 * C], out=kv[:, :C], bias=bi[C:2 * C], round_tf32=use_tc)
    ops.gemm(mem, Wi[2 * C:], out=kv[:, C:], bias=bi[2 * C:], round_tf32=use_tc)
    amode, aws = (2, g.ws) if tslma else (1, 0)
    if tslma:
        s_attn, dp = D.seed(), D.path()
        rpg = g.T * g.HW
    else:
        s_attn, dp = D.seed(), D.path_per_t(g.T)
        rpg = g.HW                               # DropPath per future-frame index here (see Drop.path_per_t)
    (ops.attn_fwd_tcgen05 if use_tc else ops.attn_fwd)(q, kv[:, :C], kv[:, C:], o, None, amode, g.N, g.H, g.W, aws, g.T, gm.T, g.nhead, g.d, False,
                                                       g.scale, round_tf32=RT, drop_seed=s_attn, drop_p=D.p)
    out = ops.gemm(o, P.wr(pre + ".out_proj.weight"), bias=P.w(pre + ".out_proj.bias"), residual=x, rowscale=dp, rows_per_group=rpg)
    if save is not None:
        if g.lean:
            zq = None
        save.append(("xattn", dict(pre=pre, ln=ln, x=x, mean=mean, rstd=rstd, zq=zq, qadd=qadd, q=q, kv=kv, o=o, g=g, gm=gm, mem=mem,
                                   mem_k=mem_k, s_attn=s_attn, dp=dp, rpg=rpg, p=D.p, amode=amode, aws=aws)))
    return out


def cross_attn_bwd(P, s, dout, dqpos, dmem):
    g, gm, C, pre, ln = s["g"], s["gm"], s["g"].C, s["pre"], s["ln"]
    q, kv, o = s["q"], s["kv"], s["o"]
    dres, dout = dout, _rc(dout, s["dp"], s["rpg"] * C, colsum=P.g(pre + ".out_proj.bias"))
    _wgrad(P, pre + ".out_proj.weight", dout, o)
    do = ops.gemm(dout, P.wr(pre + ".out_proj.weight"), b_mn=True)
    dq = torch.empty_like(q)
    # mode 2: a memory token is a key of every future frame's windows -> dK / dV are accumulated (atomics) and rounded afterwards
    dkv = torch.zeros_like(kv) if s["amode"] == 2 else torch.empty_like(kv)
    gW, gb = P.g(pre + ".in_proj_weight"), P.g(pre + ".in_proj_bias")
    ops.attn_bwd(q, kv[:, :C], kv[:, C:], do, dq, dkv[:, :C], dkv[:, C:], None, None, s["amode"], g.N, g.H, g.W, s["aws"], g.T, gm.T,
                 g.nhead, g.d, False, g.scale, round_tf32=RT, drop_seed=s["s_attn"], drop_p=s["p"])
    if s["amode"] == 2 and RT:
        dkv = ops.round_copy(dkv)
    Wi = P.wr(pre + ".in_proj_weight")
    if gW is not None:
        zq = s["zq"]
        if zq is None:    # lean mode
            zq = ops.layernorm_fwd(s["x"], P.w(ln + ".weight"), P.w(ln + ".bias"), want_y=False, add=s["qadd"], add_div=1,
                                   add_mod=s["qadd"].shape[0], round_tf32=RT, save_stats=False)[1]
        _wgrad_raw(P, dq, zq, gW[:C])
        del zq
        ops.colsum(dq, gb[:C])
        ops.colsum(dkv, gb[C:])
        _wgrad_raw(P, dkv[:, :C], s["mem_k"], gW[C:2 * C])
        _wgrad_raw(P, dkv[:, C:], s["mem"], gW[2 * C:])

    dzq = ops.gemm(dq, Wi[:C], b_mn=True)
    ops.gemm(dkv, Wi[C:], b_mn=True, out=dmem, residual=dmem)     # d(mem) += [dk | dv] [Wk ; Wv]  (memory + pos_past and memory share it)
    if dqpos is not None:
        ops.rowgroup_sum(dzq, dqpos, g.N)
    return ops.layernorm_bwd(dzq, None, s["x"], P.w(ln + ".weight"), P.w(ln + ".bias"), s["mean"], s["rstd"], dres,
                             P.g(ln + ".weight"), P.g(ln + ".bias"))


# =================================================================================================== final norm
def final_norm_fwd(P, ln, x, relu, save, round_out=False):
    y, _, mean, rstd = ops.layernorm_fwd(x, P.w(ln + ".weight"), P.w(ln + ".bias"), relu=relu, round_tf32=round_out and RT)
    if save is not None:
        save.append(("fnorm", dict(ln=ln, x=x, mean=mean, rstd=rstd, relu=relu)))
    return y


def final_norm_bwd(P, s, dout):
    ln = s["ln"]
    return ops.layernorm_bwd(dout, None, s["x"], P.w(ln + ".weight"), P.w(ln + ".bias"), s["mean"], s["rstd"], None,
                             P.g(ln + ".weight"), P.g(ln + ".bias"), relu=s["relu"])


# =================================================================================================== stacks
def lw_table(lw_pos, g):
    """(Hp*Wp, C) table of lw_pos[(h % ws), (w % ws)] over the padded grid (rpe=False path)."""
    ws = g.ws
    hh = torch.arange(g.Hp, device=lw_pos.device) % ws
    ww = torch.arange(g.Wp, device=lw_pos.device) % ws
    return lw_pos[hh][:, ww].reshape(g.Hp * g.Wp, -1).contiguous()


def encoder_fwd(P, bufs, x, g, n_layers, far, rpe, tpos, lw_tab, training, save, D=NO_DROP):
    for i in range(n_layers):
        pre = "transformer.encoder.layers.%d" % i
        x = window_attn_fwd(P, pre + ".SLMHSA", pre + ".norm1", x, g, rpe, None, lw_tab, save, D)
        x = conv_ffn_fwd(P, bufs, pre + ".SpatialFFN", pre + ".norm2", x, g, far, training, save, D)
        x = temporal_attn_fwd(P, pre + ".temporal_MHSA", pre + ".norm3", x, g, tpos, far, save, D)
        x = mlp_fwd(P, pre, pre + ".norm4", x, g, save, D)
    return x


def decoder_fwd(P, bufs, tgt, g, gm, n_layers, rpe, qpos, qadd, tpos_f, mem, mem_k, lw_tab, save, D=NO_DROP, tslma=None):
    """tslma: None, or dict(q_tab (T2*H*W, C) = query_pos + Tlw_pos[future] over the grid, mem_k = memory + Tlw_pos[past]) when the
    blocks carry TemporalSpatialLocalMultiheadAttention instead of the per-pixel encoder-decoder attention (TSLMA_flag)"""
    for i in range(n_layers):
        pre = "transformer.decoder.layers.%d" % i
        tgt = window_attn_fwd(P, pre + ".SLMHSA", pre + ".norm1", tgt, g, rpe, qpos, lw_tab, save, D)
        tgt = conv_ffn_fwd(P, bufs, pre + ".SpatialFFN", pre + ".norm2", tgt, g, True, False, save, D)
        tgt = temporal_attn_fwd(P, pre + ".temporal_MHSA", pre + ".norm3", tgt, g, tpos_f, False, save, D)
        tgt = mlp_fwd(P, pre, pre + ".norm4", tgt, g, save, D)
        if tslma is not None:
            tgt = cross_attn_fwd(P, pre + ".TSLMA.attn", pre + ".norm5", tgt, g, gm, tslma["q_tab"], mem, tslma["mem_k"], save, D, tslma=True)
        else:
            tgt = cross_attn_fwd(P, pre + ".EncDecAttn", pre + ".norm5", tgt, g, gm, qadd, mem, mem_k, save, D)
        tgt = conv_ffn_fwd(P, bufs, pre + ".SpatialFFN1", pre + ".norm6", tgt, g, True, False, save, D)
    return tgt


# Called as GRAD_READY(flat, lo, hi) from the backward thread each time a Transformer layer's parameter gradients are final (they
# are one contiguous slice of the flat gradient buffer); vptr_b200.parallel.GradReducer uses it to start that slice's all-reduce
# while the rest of the backward still runs.
GRAD_READY = None


def _layer_done(P, pre):
    """pre = '<...>.layers.<i>.SLMHSA' of the layer whose backward just finished (its window attention is the first sub-block of
    a layer, so the last one of its backward)"""
    if GRAD_READY is None or P.gflat is None:
        return
    join_side(P)                 # the layer's weight gradients (side stream) must be final before its slice is reduced
    layer = pre.rsplit(".", 1)[0] + "."
    span = [P.goff[k] for k in P.goff if k.startswith(layer)]
    if span:
        GRAD_READY(P.gflat, min(a for a, _ in span), max(b for _, b in span))


def backward_tape(P, save, dout, dqpos=None, dmem=None, stop=0):
    """Walks the saved sub-blocks in reverse down to index `stop`; returns the gradient at that point."""
    d = dout
    while len(save) > stop:
        kind, s = save.pop()
        if kind == "win":
            d = window_attn_bwd(P, s, d, dqpos)
            _layer_done(P, s["pre"])
        elif kind == "ffn":
            d = conv_ffn_bwd(P, s, d)
        elif kind == "tattn":
            d = temporal_attn_bwd(P, s, d)
        elif kind == "mlp":
            d = mlp_bwd(P, s, d)
        elif kind == "xattn":
            d = cross_attn_bwd(P, s, d, dqpos, dmem)
        elif kind == "fnorm":
            d = final_norm_bwd(P, s, d)
        else:
            raise RuntimeError("unknown tape entry " + kind)
        _entry_done(P)
    return d
