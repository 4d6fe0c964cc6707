"""Thin torch-tensor wrappers over the C-ABI (include/vptr_b200.h).  torch is used here only for device memory and
the current CUDA stream; every computation is a kernel of libvptr_b200.so."""
import os

import torch

from . import _lib

_call = _lib.call


def _p(t):
    return 0 if t is None else t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, name="tensor"):
    if t is not None:
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError("vptr_b200: %s must be a CUDA float32 tensor (got %s %s); there is no CPU path" % (name, t.device, t.dtype))
    return t


def empty(*shape, like=None, device=None):
    return torch.empty(*shape, dtype=torch.float32, device=like.device if like is not None else device)


def zeros(*shape, like=None, device=None):
    return torch.zeros(*shape, dtype=torch.float32, device=like.device if like is not None else device)


# --------------------------------------------------------------------------------------------------- GEMM
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
DEBUG_FLAGS = 0     # gemm epilogue debug bits (4: no global store, 8: skip epilogue body)
FORCE_SIMT = False  # tests flip this to cross-check the tensor-core kernel against the fp32 FFMA kernel


def _mat(t):
    """(ptr, pitch) of a 2-D row-major view whose inner stride is 1."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise RuntimeError("vptr_b200.gemm: operand must be 2-D with unit inner stride, got shape %s strides %s" % (tuple(t.shape), t.stride()))
    return t.data_ptr(), (t.stride(0) if t.shape[0] > 1 else t.shape[1])


def gemm(A, B, out=None, a_mn=False, b_mn=False, bias=None, residual=None, alpha=1.0, act=ACT_NONE, accumulate=False,
         round_tf32=False, k_splits=0, rowscale=None, rows_per_group=0, drop_seed=0, drop_p=0.0):
    """out[M,N] (+)= act(alpha * Aop @ Bop^T + bias) + residual.  a_mn: A stored [K,M]; b_mn: B stored [K,N]."""
    _chk(A, "A"); _chk(B, "B"); _chk(bias, "bias"); _chk(residual, "residual")
    if a_mn:
        K, M = A.shape
    else:
        M, K = A.shape
    if b_mn:
        Kb, N = B.shape
    else:
        N, Kb = B.shape
    if K != Kb:
        raise RuntimeError("vptr_b200.gemm: contraction mismatch %d vs %d" % (K, Kb))
    if out is None:
        if accumulate:
            raise RuntimeError("vptr_b200.gemm: accumulate needs an output buffer")
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    if out.shape[0] != M or out.shape[1] != N:
        raise RuntimeError("vptr_b200.gemm: out shape %s != (%d,%d)" % (tuple(out.shape), M, N))
    pa, lda = _mat(A)
    pb, ldb = _mat(B)
    pd, ldd = _mat(out)
    pr, ldr = (0, 0) if residual is None else _mat(residual)
    flags = (1 if accumulate else 0) | (2 if round_tf32 else 0) | DEBUG_FLAGS
    aligned = (lda % 4 == 0 and ldb % 4 == 0 and ldd % 4 == 0 and ldr % 4 == 0 and N % 4 == 0 and pa % 16 == 0 and pb % 16 == 0
               and pd % 16 == 0 and pr % 16 == 0 and _p(bias) % 16 == 0)             # TMA / float4 epilogue: 16-byte rules
    name = "vptr_gemm_tf32" if (aligned and not FORCE_SIMT) else "vptr_gemm_simt"
    if GEMM_3XTF32 and name == "vptr_gemm_tf32" and K % 4 == 0:
        return _gemm_3xtf32(A, B, out, a_mn, b_mn, bias, residual, alpha, act, accumulate, round_tf32, rowscale, rows_per_group, drop_seed, drop_p)
    _call(name, pa, lda, int(a_mn), pb, ldb, int(b_mn), pd, ldd, M, N, K, _p(bias), pr, ldr, float(alpha), int(act), flags,
          int(k_splits), _p(rowscale), int(rows_per_group), int(drop_seed), float(drop_p), _s())
    return out


GEMM_3XTF32 = False   # verification mode (engine.precise_3xtf32): see _gemm_3xtf32


def _gemm_3xtf32(A, B, out, a_mn, b_mn, bias, residual, alpha, act, accumulate, round_tf32, rowscale, rows_per_group, drop_seed, drop_p):
    """The SAME tcgen05 kernel with 3xTF32 operand precision: A = Ah + Al, B = Bh + Bl (tf32 planes), A.B ~ Ah.Bh + Ah.Bl + Al.Bh
    as ONE contraction over 3K on the concatenated operands [Ah | Ah | Al] . [Bh | Bl | Bh]^T.  Tests use it to show that what
    separates the shipped TF32 path from the fp32 reference is operand rounding, not the kernels (tile shapes, pipelines, epilogues,
    split-K reductions are exactly the product's)."""
    def planes(X, mn):
        X = X.contiguous()
        hi = round_copy(X)
        lo = round_copy(X - hi)
        return hi, lo, (0 if mn else 1)
    Ah, Al, da = planes(A, a_mn)
    Bh, Bl, db = planes(B, b_mn)
    A3 = torch.cat([Ah, Ah, Al], dim=da)
    B3 = torch.cat([Bh, Bl, Bh], dim=db)
    global GEMM_3XTF32
    GEMM_3XTF32 = False
    try:
        return gemm(A3, B3, out=out, a_mn=a_mn, b_mn=b_mn, bias=bias, residual=residual, alpha=alpha, act=act, accumulate=accumulate,
                    round_tf32=round_tf32, rowscale=rowscale, rows_per_group=rows_per_group, drop_seed=drop_seed, drop_p=drop_p)
    finally:
        GEMM_3XTF32 = True


# --------------------------------------------------------------------------------------------------- norms
def layernorm_fwd(x, gamma, beta, want_y=True, add=None, add_div=1, add_mod=1, save_stats=True, relu=False, eps=1e-5, round_tf32=False):
    rows, C = x.shape
    y = torch.empty_like(x) if want_y else None
    y2 = torch.empty_like(x) if add is not None else None
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    _call("vptr_layernorm_fwd", _p(_chk(x)), _p(gamma), _p(beta), _p(y), _p(y2), _p(add), int(add_div), int(add_mod), _p(mean),
          _p(rstd), rows, C, float(eps), int(relu), int(round_tf32), _s())
    return y, y2, mean, rstd


def layernorm_bwd(dy1, dy2, x, gamma, beta, mean, rstd, dres, dgamma, dbeta, relu=False, want_dx=True):
    rows, C = x.shape
    dx = torch.empty_like(x) if want_dx else None
    _call("vptr_layernorm_bwd", _p(_chk(dy1)), _p(dy2), _p(x), _p(gamma), _p(beta), _p(mean), _p(rstd), _p(dres), _p(dx), _p(dgamma),
          _p(dbeta), rows, C, int(relu), _s())
    return dx


def bn_stats(x, running_mean, running_var, momentum=0.1, eps=1e-5):
    rows, ch = x.shape
    mean = torch.empty(ch, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    ws = torch.empty(2 * ch, dtype=torch.float64, device=x.device)
    _call("vptr_bn_stats", _p(x), rows, ch, _p(mean), _p(rstd), _p(running_mean), _p(running_var), float(eps), float(momentum), _p(ws), _s())
    return mean, rstd


def bn_eval_stats(running_mean, running_var, eps=1e-5):
    ch = running_mean.numel()
    mean = torch.empty(ch, dtype=torch.float32, device=running_mean.device)
    rstd = torch.empty_like(mean)
    _call("vptr_bn_eval_stats", _p(running_mean), _p(running_var), _p(mean), _p(rstd), ch, float(eps), _s())
    return mean, rstd


def group_stats(x, groups, eps=1e-5):
    gsize = x.numel() // groups
    mean = torch.empty(groups, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    _call("vptr_group_stats", _p(x), groups, gsize, _p(mean), _p(rstd), float(eps), _s())
    return mean, rstd


def norm_act_fwd(x, mean, rstd, gamma, beta, hw, mode, res=None, out=None, round_tf32=False, rowscale=None, rows_per_group=0,
                 drop_seed=0, drop_p=0.0):
    rows, ch = x.shape
    y = torch.empty_like(x) if out is None else out
    _call("vptr_norm_act_fwd", _p(x), _p(y), _p(res), _p(mean), _p(rstd), _p(gamma), _p(beta), rows, ch, hw, mode, int(round_tf32),
          _p(rowscale), int(rows_per_group), int(drop_seed), float(drop_p), _s())
    return y


def norm_act_bwd(dy, x, mean, rstd, gamma, beta, dgamma, dbeta, hw, mode, round_tf32=False, rowscale=None, rows_per_group=0,
                 drop_seed=0, drop_p=0.0, inplace=False, colsum=None):
    """inplace: dx overwrites dy (for temporaries); otherwise a fresh buffer (which also holds the intermediate g0).
    colsum (ch,): += column sums of dx (the bias gradient of the 1x1 conv in front of the norm), out of the final pass"""
    rows, ch = x.shape
    dx = dy if inplace else torch.empty_like(x)
    n_ws = 2 * ch if mode != 1 else 2 * (rows // hw)
    ws = torch.empty(n_ws, dtype=torch.float32, device=x.device)
    _call("vptr_norm_act_bwd_colsum", _p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(dx), _p(dgamma), _p(dbeta), rows, ch, hw, mode,
          _p(ws), int(round_tf32), _p(rowscale), int(rows_per_group), int(drop_seed), float(drop_p), _p(colsum), _s())
    return dx


# --------------------------------------------------------------------------------------------------- attention
def attn_fwd(q, k, v, out, rpe_table, mode, F_or_N, H, W, ws, Tq, Tk, nhead, d, causal, scale, round_tf32=False, drop_seed=0, drop_p=0.0):
    _call("vptr_attn_fwd", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0), _p(rpe_table), mode,
          F_or_N, H, W, ws, Tq, Tk, nhead, d, int(causal), float(scale), int(round_tf32), int(drop_seed), float(drop_p), _s())
    return out


# the engine routes the window-attention forward of the path's shape (8x8 grid, 4x4 windows, 8 heads of 66) to the tcgen05 / TMA /
# TMEM kernel; VPTR_ATTN_TC=0 keeps everything on the warp-level 3xTF32 kernels
ATTN_TC = os.environ.get("VPTR_ATTN_TC", "") != "0"


def attn_tc_window_ok(q, k, v, out, H, W, ws, nhead, d):
    """shape / alignment domain of the tcgen05 window-attention fast path"""
    return (ATTN_TC and not FORCE_SIMT and H == 8 and W == 8 and ws == 4 and d == 66 and 1 <= nhead <= 8
            and all(t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0 for t in (q, k, v, out)))


ATTN_TC_SEQ_SHAPES = ((10, 10), (29, 29), (28, 28), (28, 2), (30, 30), (30, 10))   # (Tq, Tk) with compile-time fast paths: cfg1, cfg2 (FAR), cfg3, cfg4


def attn_tc_temporal_ok(q, k, v, out, Tq, Tk, nhead, d):
    """domain of the tcgen05 temporal / enc-dec fast paths"""
    return (ATTN_TC and not FORCE_SIMT and (Tq, Tk) in ATTN_TC_SEQ_SHAPES and d == 66 and 1 <= nhead <= 8
            and all(t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0 for t in (q, k, v, out)))


def attn_fwd_tcgen05(q, k, v, out, rpe_table, mode, F_or_N, H, W, ws, Tq, Tk, nhead, d, causal, scale, round_tf32=False, drop_seed=0, drop_p=0.0):
    """the tcgen05 / TMA / TMEM forward (opt-in, VPTR_ATTN_TC=1 routes attn_fwd here); raises outside its shape domain"""
    _call("vptr_attn_fwd_tcgen05", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0), _p(rpe_table), mode, F_or_N,
          H, W, ws, Tq, Tk, nhead, d, int(causal), float(scale), int(round_tf32), int(drop_seed), float(drop_p), _s())
    return out


def attn_bwd(q, k, v, do, dq, dk, dv, rpe_table, d_rpe_table, mode, F_or_N, H, W, ws, Tq, Tk, nhead, d, causal, scale, round_tf32=False,
             drop_seed=0, drop_p=0.0, dbq=None, dbk=None, dbv=None):
    """dbq / dbk / dbv (nhead*d,): += column sums of dq / dk / dv (bias gradients of the q / k / v projections)"""
    _call("vptr_attn_bwd_bias", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(do), do.stride(0), _p(dq), dq.stride(0),
          _p(dk), dk.stride(0), _p(dv), dv.stride(0), _p(rpe_table), _p(d_rpe_table), mode, F_or_N, H, W, ws, Tq, Tk, nhead, d,
          int(causal), float(scale), int(round_tf32), int(drop_seed), float(drop_p), _p(dbq), _p(dbk), _p(dbv), _s())


def window_index_maps(F, H, W, ws, device):
    L = ws * ws
    B = F * (H // ws) * (W // ws)
    rpi = torch.empty(L, L, dtype=torch.int64, device=device)
    wmap = torch.empty(L, B, dtype=torch.int64, device=device)
    _call("vptr_window_index_maps", F, H, W, ws, rpi.data_ptr(), wmap.data_ptr(), _s())
    return rpi, wmap


def causal_mask(T, device):
    m = torch.empty(T, T, dtype=torch.uint8, device=device)
    _call("vptr_causal_mask", T, m.data_ptr(), _s())
    return m.bool()


# --------------------------------------------------------------------------------------------------- dw conv / elementwise
def dwconv3x3(x, w9, bias, F, H, W, flip=False):
    y = torch.empty_like(x)
    _call("vptr_dwconv3x3", _p(x), _p(w9), _p(bias), _p(y), F, H, W, x.shape[-1], int(flip), _s())
    return y


def dwconv3x3_stats(x, w9, bias, F, H, W, eps=1e-5):
    """forward depthwise conv that also returns the per-frame (mean, rstd) of its output -- one pass when the streaming kernel
    applies, else dwconv3x3 + group_stats"""
    y = torch.empty_like(x)
    sums = torch.zeros(2 * F, dtype=torch.float64, device=x.device)
    l = _lib.lib()
    rc = l.vptr_dwconv3x3_stats(_p(x), _p(w9), _p(bias), _p(y), F, H, W, x.shape[-1], _p(sums), _s())
    _lib.launch_count += 1
    if rc == -3:
        y = dwconv3x3(x, w9, bias, F, H, W)
        return y, group_stats(y, F, eps)
    if rc != 0:
        raise RuntimeError("vptr_dwconv3x3_stats failed (status %d): %s" % (rc, l.vptr_last_error().decode("utf-8", "replace")))
    mean = torch.empty(F, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    _call("vptr_group_stats_finalize", _p(sums), F, x.numel() // F, _p(mean), _p(rstd), float(eps), _s())
    return y, (mean, rstd)


def dwconv3x3_wgrad(x, dy, dw9, dbias, F, H, W):
    _call("vptr_dwconv3x3_wgrad", _p(x), _p(dy), _p(dw9), _p(dbias), F, H, W, x.shape[-1], _s())


def axpby(a, b, alpha=1.0, beta=1.0, out=None):
    out = torch.empty_like(a) if out is None else out
    _call("vptr_axpby", _p(a), _p(b), _p(out), a.numel(), float(alpha), float(beta), _s())
    return out


def add_rows(x, add, div, mod, round_tf32=False):
    rows, C = x.shape
    out = torch.empty_like(x)
    _call("vptr_add_rows", _p(x), _p(add), _p(out), rows, C, div, mod, int(round_tf32), _s())
    return out


def rowgroup_sum(dy, out, reps):
    _call("vptr_rowgroup_sum", _p(dy), _p(out), out.numel(), reps, _s())


def gelu_fwd(x, round_tf32=False, drop_seed=0, drop_p=0.0):
    y = torch.empty_like(x)
    _call("vptr_gelu_fwd", _p(x), _p(y), x.numel(), int(round_tf32), int(drop_seed), float(drop_p), _s())
    return y


def gelu_bwd(dy, x, out=None, round_tf32=False, drop_seed=0, drop_p=0.0, colsum=None):
    """colsum (C,): += column sums of dx (x is [rows][C]): the bias gradient of the Linear in front of the GELU"""
    dx = torch.empty_like(x) if out is None else out
    if colsum is not None and x.dim() == 2 and x.shape[1] % 4 == 0:
        _call("vptr_gelu_bwd_colsum", _p(dy), _p(x), _p(dx), x.shape[0], x.shape[1], int(round_tf32), int(drop_seed), float(drop_p), _p(colsum), _s())
        return dx
    _call("vptr_gelu_bwd", _p(dy), _p(x), _p(dx), x.numel(), int(round_tf32), int(drop_seed), float(drop_p), _s())
    if colsum is not None:
        globals()["colsum"](dx, colsum)
    return dx


def round_copy(x, do_round=True, rowscale=None, group_elems=0, drop_seed=0, drop_p=0.0, colsum=None):
    """y = [round-to-nearest tf32](x * rowscale[i / group_elems] * dropmask): operands of the tensor-core GEMM are pre-rounded
    so its truncation is exact; the backward of a dropped / DropPath-scaled branch applies the same mask here.
    colsum (C,): += column sums of y ([rows][C]) -- the bias gradient of the Linear whose output gradient y is"""
    y = torch.empty_like(x)
    if colsum is not None and x.dim() == 2 and x.is_contiguous() and x.shape[1] % 4 == 0 and (rowscale is None or group_elems % x.shape[1] == 0):
        _call("vptr_round_copy_colsum", _p(x), _p(y), x.shape[0], x.shape[1], int(do_round), _p(rowscale), int(group_elems) // x.shape[1],
              int(drop_seed), float(drop_p), _p(colsum), _s())
        return y
    _call("vptr_round_copy", _p(x), _p(y), x.numel(), int(do_round), _p(rowscale), int(group_elems), int(drop_seed), float(drop_p), _s())
    if colsum is not None:
        globals()["colsum"](y, colsum)
    return y


_ROUND_TABLES = {}


def round_copy_multi(tensors):
    """tf32-rounded copies of several tensors with ONE launch; returns views of one flat buffer (same shapes).  The device
    pointer table is cached per set of source pointers (parameters keep their storage across steps)."""
    key = tuple((t.data_ptr(), t.numel()) for t in tensors)
    ent = _ROUND_TABLES.get(key)
    if ent is None:
        ends, acc = [], 0
        for t in tensors:
            _chk(t, "weight")
            if t.numel() % 4 or t.data_ptr() % 16 or not t.is_contiguous():
                raise RuntimeError("vptr_b200.round_copy_multi: tensors must be contiguous, 16-byte aligned, numel % 4 == 0")
            acc += t.numel() // 4
            ends.append(acc)
        table = torch.tensor([t.data_ptr() for t in tensors] + ends, dtype=torch.int64).to(tensors[0].device)
        if len(_ROUND_TABLES) > 64:
            _ROUND_TABLES.clear()
        ent = _ROUND_TABLES[key] = (table, ends)
    table, ends = ent
    flat = torch.empty(ends[-1] * 4, dtype=torch.float32, device=tensors[0].device)
    _call("vptr_round_copy_multi", table.data_ptr(), len(tensors), _p(flat), ends[-1], _s())
    out, b = [], 0
    for t, e in zip(tensors, ends):
        out.append(flat[b * 4:e * 4].view(t.shape))
        b = e
    return out


def droppath_scales(n, seed, p, device):
    out = torch.empty(n, dtype=torch.float32, device=device)
    _call("vptr_droppath_scales", _p(out), n, int(seed), float(p), _s())
    return out


def relu_fwd(x, out=None):
    y = torch.empty_like(x) if out is None else out
    _call("vptr_relu_fwd", _p(x), _p(y), x.numel(), _s())
    return y


def relu_bwd(dy, y, out=None):
    dx = torch.empty_like(y) if out is None else out
    _call("vptr_relu_bwd", _p(dy), _p(y), _p(dx), y.numel(), _s())
    return dx


def colsum(x, out):
    """out[c] += sum_r x[r, c]; x may be a column slice of a wider buffer."""
    rows, C = x.shape
    _call("vptr_colsum", _p(x), _p(out), rows, C, x.stride(0), _s())


_TR_TABLES = {}
_PINNED_KEEPALIVE = []      # host tables whose H2D copy was captured into a CUDA graph: replays read them again, never free them


def pinned_table(values, device):
    """int64 device table from a Python list through a pinned host buffer (async copy).  Returns (device tensor, host tensor): the
    caller keeps the host tensor alive at least until the copy ran; if the copy is being captured into a CUDA graph the host
    buffer is kept for the life of the process (a replay re-reads it)."""
    host = torch.tensor(values, dtype=torch.int64).pin_memory()
    dev = host.to(device, non_blocking=True)
    if torch.cuda.is_current_stream_capturing():
        _PINNED_KEEPALIVE.append(host)
    return dev, host


def transpose_multi(site, entries, accumulate=False):
    """entries [(src, dst, R, C)]: dst[c][r] (+)= src[r][c] for every entry in ONE launch.  `site` names the call site: its pointer
    table is rebuilt only when a pointer or shape changes."""
    key = tuple(v for s_, d_, R, C in entries for v in (s_.data_ptr(), d_.data_ptr(), R, C))
    ent = _TR_TABLES.get(site)
    if ent is None or ent[0] != key:
        rows, acc = [], 0
        for s_, d_, R, C in entries:
            acc += ((R + 31) // 32) * ((C + 31) // 32)
            rows += [s_.data_ptr(), d_.data_ptr(), R, C, acc]
        dev, host = pinned_table(rows, entries[0][0].device)
        ent = _TR_TABLES[site] = (key, dev, host, acc)
    _call("vptr_transpose_multi", _p(ent[1]), len(entries), ent[3], int(accumulate), _s())


def transpose(x, batch, R, C, out=None, accumulate=False):
    """[batch][R][C] -> [batch][C][R] (out += when accumulate)."""
    if out is None:
        out = torch.empty(batch * R * C, dtype=torch.float32, device=x.device)
    _call("vptr_transpose", _p(x), _p(out), batch, R, C, int(accumulate), _s())
    return out


def pad_hw(x, F, H, W, Hp, Wp, ph0, pw0):
    C = x.shape[-1]
    out = torch.empty(F * Hp * Wp, C, dtype=torch.float32, device=x.device)
    _call("vptr_pad_crop", _p(x), _p(out), F, H, W, Hp, Wp, ph0, pw0, C, 0, _s())
    return out


def crop_hw(x, F, H, W, Hp, Wp, ph0, pw0):
    C = x.shape[-1]
    out = torch.empty(F * H * W, C, dtype=torch.float32, device=x.device)
    _call("vptr_pad_crop", _p(x), _p(out), F, H, W, Hp, Wp, ph0, pw0, C, 1, _s())
    return out


def sqnorm_accumulate(x, acc):
    _call("vptr_sqnorm_accumulate", _p(x), x.numel(), acc.data_ptr(), _s())


def clip_scale(x, sqnorm, max_norm):
    _call("vptr_clip_scale", _p(x), x.numel(), sqnorm.data_ptr(), float(max_norm), _s())


# --------------------------------------------------------------------------------------------------- ResNet pieces
PAD_MODES = {"zero": 0, "reflect": 1, "replicate": 2}


def im2col(x, F, H, W, Cin, k, stride, pad, pad_mode, mask=None, round_tf32=True):
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    col = torch.empty(F * Ho * Wo, k * k * Cin, dtype=torch.float32, device=x.device)
    _call("vptr_im2col", _p(x), _p(mask), _p(col), F, H, W, Cin, k, stride, pad, pad_mode, int(round_tf32), _s())
    return col, Ho, Wo


def pad_nhwc(x, F, H, W, C, pad, pad_mode, round_tf32=True):
    out = torch.empty(F * (H + 2 * pad) * (W + 2 * pad), C, dtype=torch.float32, device=x.device)
    _call("vptr_pad_nhwc", _p(x), _p(out), F, H, W, C, pad, pad_mode, int(round_tf32), _s())
    return out


def conv3x3_implicit_ok(H, W):
    """does an H x W grid tile into the 128-pixel TMA boxes of vptr_conv3x3_tf32?"""
    px = H * W
    if px <= 128:
        return 128 % px == 0 and H <= 256 and W <= 256
    return 128 % W == 0 and H % (128 // W) == 0


def conv3x3_tf32(xpad, w, F, H, W, C, Cout, bias=None, residual=None, act=ACT_NONE, round_tf32=False, w_planes=1):
    """w: [Cout][9C] (w_planes=1) or the [hi|lo] tf32 split [Cout][2*9C] from split_tf32 (w_planes=2)"""
    out = torch.empty(F * H * W, Cout, dtype=torch.float32, device=xpad.device)
    _call("vptr_conv3x3_tf32", _p(xpad), _p(w), _p(out), F, H, W, C, Cout, _p(bias), _p(residual), int(act), 2 if round_tf32 else 0,
          int(w_planes), _s())
    return out


def conv3x3_quad_ok(H, W):
    """grids the quadrant-tiled raw-tile convolution covers (beyond the native 8x8 one)"""
    return H % 8 == 0 and W % 8 == 0 and (H > 8 or W > 8) and os.environ.get("VPTR_CONV_GENERIC", "") != "1"


def pad_nhwc_quad(x, F, H, W, C, pad_mode, round_tf32=True):
    out = torch.empty(F * (H // 8) * (W // 8) * 100, C, dtype=torch.float32, device=x.device)
    _call("vptr_pad_nhwc_quad", _p(x), _p(out), F, H, W, C, pad_mode, int(round_tf32), _s())
    return out


def conv3x3_tf32_quad(xq, w, F, H, W, C, Cout, bias=None, residual=None, act=ACT_NONE, round_tf32=False, w_planes=1):
    out = torch.empty(F * H * W, Cout, dtype=torch.float32, device=xq.device)
    _call("vptr_conv3x3_tf32_quad", _p(xq), _p(w), _p(out), F, H, W, C, Cout, _p(bias), _p(residual), int(act), 2 if round_tf32 else 0,
          int(w_planes), _s())
    return out


CONV_BF16X3 = os.environ.get("VPTR_CONV_TF32", "") != "1"     # ResnetBlock convs on the bf16x3 raw-tile kernel (else two-plane TF32)


def conv3x3_bf16x3_ok(H, W, C):
    return CONV_BF16X3 and H % 8 == 0 and W % 8 == 0 and C % 8 == 0 and os.environ.get("VPTR_CONV_GENERIC", "") != "1"


def pad_nhwc_quad_bf16x2(x, F, H, W, C, pad_mode):
    out = torch.empty(2, F * (H // 8) * (W // 8) * 100, C, dtype=torch.bfloat16, device=x.device)
    _call("vptr_pad_nhwc_quad_bf16x2", _p(x), _p(out), F, H, W, C, pad_mode, _s())
    return out


def split_bf16x2(w):
    """(rows, K) fp32 -> (rows, 2K) bf16: [hi plane | lo plane] per row"""
    rows, K = w.shape
    out = torch.empty(rows, 2 * K, dtype=torch.bfloat16, device=w.device)
    _call("vptr_split_bf16x2", _p(w), _p(out), rows, K, _s())
    return out


def conv3x3_bf16x3(xq2, w2, F, H, W, C, Cout, bias=None, residual=None, act=ACT_NONE, round_tf32=False):
    out = torch.empty(F * H * W, Cout, dtype=torch.float32, device=xq2.device)
    _call("vptr_conv3x3_bf16x3", _p(xq2), _p(w2), _p(out), F, H, W, C, Cout, _p(bias), _p(residual), int(act), 2 if round_tf32 else 0, _s())
    return out


def split_tf32(w):
    """[rows][K] -> [rows][2K]: tf32 hi plane followed by the tf32 lo plane"""
    rows, K = w.shape
    out = torch.empty(rows, 2 * K, dtype=torch.float32, device=w.device)
    _call("vptr_split_tf32", _p(w), _p(out), rows, K, _s())
    return out


def convT_gather(col, shift, F, H, W, Cout, relu=True):
    out = torch.empty(F * 4 * H * W, Cout, dtype=torch.float32, device=col.device)
    _call("vptr_convT_gather", _p(col), _p(shift), _p(out), F, H, W, Cout, int(relu), _s())
    return out


def bn_fold(bn, eps=None):
    C = bn.weight.numel()
    scale = torch.empty(C, dtype=torch.float32, device=bn.weight.device)
    shift = torch.empty_like(scale)
    _call("vptr_bn_fold", _p(bn.weight.data), _p(bn.bias.data), _p(bn.running_mean), _p(bn.running_var), float(bn.eps if eps is None else eps),
          _p(scale), _p(shift), C, _s())
    return scale, shift


def pack_conv_weight(w, scale, mode):
    """mode 0: Conv2d -> [Co][(kh,kw,ci)]; 1: ConvT -> [(kh,kw,co)][ci]; 2: Conv2d -> [(kh,kw,ci)][co]; 3: Conv2d -> [(kh,kw)][co][ci]."""
    if mode == 1:
        Ci, Co, k, _ = w.shape
    else:
        Co, Ci, k, _ = w.shape
    out = torch.empty(w.numel(), dtype=torch.float32, device=w.device)
    _call("vptr_pack_conv_weight", _p(w), _p(scale), _p(out), Co, Ci, k, mode, _s())
    return out


def stem_conv7x7(x, wpk, shift, F, Ci, H, W, Co):
    out = torch.empty(F * H * W, Co, dtype=torch.float32, device=x.device)
    _call("vptr_stem_conv7x7", _p(x), _p(wpk), _p(shift), _p(out), F, Ci, H, W, Co, _s())
    return out


def head_conv7x7_fwd(x, wpk, bias, F, Ci, Co, H, W, act):
    out = torch.empty(F, Co, H, W, dtype=torch.float32, device=x.device)
    _call("vptr_head_conv7x7_fwd", _p(x), _p(wpk), _p(bias), _p(out), F, Ci, Co, H, W, act, _s())
    return out


def head_conv7x7_bwd(dout, out, w, F, Ci, Co, H, W, act):
    dx = torch.empty(F * H * W, Ci, dtype=torch.float32, device=dout.device)
    ws = torch.empty(F * (H + 6) * (W + 6) * Ci + 49 * Co * Ci, dtype=torch.float32, device=dout.device)
    _call("vptr_head_conv7x7_bwd", _p(dout), _p(out), _p(w), _p(dx), F, Ci, Co, H, W, act, _p(ws), _s())
    return dx


# --------------------------------------------------------------------------------------------------- stage-1 (train-mode) ResNet pieces
def stem_conv7x7_raw(x, wpk, F, Ci, H, W, Co):
    out = torch.empty(F * H * W, Co, dtype=torch.float32, device=x.device)
    _call("vptr_stem_conv7x7_raw", _p(x), _p(wpk), _p(out), F, Ci, H, W, Co, _s())
    return out


def bn_act_fwd(x, mean, rstd, gamma, beta, act, res=None, round_tf32=False):
    """act 0: none, 1: ReLU, 2: ReLU after the residual add"""
    rows, ch = x.shape
    z = torch.empty_like(x)
    _call("vptr_bn_act_fwd", _p(x), _p(z), _p(res), _p(mean), _p(rstd), _p(gamma), _p(beta), rows, ch, int(act), int(round_tf32), _s())
    return z


def bn_act_bwd(dz, x, z, mean, rstd, gamma, dgamma, dbeta, act, round_tf32=False):
    """-> (g0 = activation-masked dz, dx); dgamma / dbeta are accumulated"""
    rows, ch = x.shape
    g0 = torch.empty_like(x)
    dx = torch.empty_like(x)
    ws = torch.empty(2 * ch, dtype=torch.float32, device=x.device)
    _call("vptr_bn_act_bwd", _p(_chk(dz)), _p(x), _p(z), _p(mean), _p(rstd), _p(gamma), _p(g0), _p(dx), _p(dgamma), _p(dbeta), rows, ch, int(act),
          _p(ws), int(round_tf32), _s())
    return g0, dx


def col2im(dcol, F, H, W, C, k, stride, pad, pad_mode):
    dx = torch.zeros(F * H * W, C, dtype=torch.float32, device=dcol.device)
    _call("vptr_col2im", _p(dcol), _p(dx), F, H, W, C, k, stride, pad, pad_mode, _s())
    return dx


def stem_wgrad(x, dy, F, Ci, H, W):
    dw = torch.zeros(49 * Ci, 64, dtype=torch.float32, device=x.device)
    _call("vptr_stem_wgrad", _p(x), _p(dy), _p(dw), F, Ci, H, W, _s())
    return dw


def act_bwd(dout, out, act):
    dpre = torch.empty_like(out)
    _call("vptr_act_bwd", _p(dout), _p(out), _p(dpre), out.numel(), int(act), _s())
    return dpre
