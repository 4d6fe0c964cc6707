"""tcgen05 attention forward (VPTR_ATTN_TC=1) vs an fp32 torch reference on the shapes of the path + timing at cfg1 size."""
import os, sys
os.environ["VPTR_ATTN_TC"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
torch.backends.cuda.matmul.allow_tf32 = False
from vptr_b200 import ops
import vptr_oracle as O


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def core(q, k, v, nhead, scale, bias=None, mask=None):
    return O._mha_core(q * scale, k, v, nhead, bias=bias, mask=mask)


def window_case(Fr, H, W, ws, nhead=8, d=66, p=0.0):
    C, L, rows = nhead * d, ws * ws, Fr * H * W
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv = torch.randn(rows, 3 * C, device="cuda", generator=g)
    table = torch.randn((2 * ws - 1) ** 2, nhead, device="cuda", generator=g) * 0.5
    o = torch.full((rows, C), float("nan"), device="cuda")
    ops.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, table, 0, Fr, H, W, ws, 0, 0, nhead, d, False, d ** -0.5, False, 5, p)
    tmap = O.window_token_map(Fr, H, W, ws).cuda()
    gth = lambda t: t[tmap.t()]
    bias = table[O.relative_position_index(ws).cuda().reshape(-1)].reshape(L, L, nhead).permute(2, 0, 1)
    ob = core(gth(qkv[:, :C]), gth(qkv[:, C:2 * C]), gth(qkv[:, 2 * C:]), nhead, d ** -0.5, bias=bias)
    oref = torch.zeros(rows, C, device="cuda").index_put((tmap.t().reshape(-1),), ob.reshape(-1, C))
    if os.environ.get("TC_DIAG"):
        e = (o - oref).double() ** 2
        bycol = e.view(rows, nhead, d).sum((0,)).sqrt() / oref.double().view(rows, nhead, d).pow(2).sum(0).sqrt()
        print("  err by head (rows) x col: worst cols per head:", [[int(c) for c in bycol[h].topk(3).indices] for h in range(nhead)])
        print("  err per head:", [float("%.3g" % float(bycol[h].mean())) for h in range(nhead)])
        byrow = (e.sum(1).view(-1, 128) if rows % 128 == 0 else e.sum(1)[:128].view(1, 128)).sum(0).sqrt()
        print("  worst rows in tile:", [int(r) for r in byrow.topk(6).indices], [float("%.3g" % float(v)) for v in byrow.topk(6).values], "median", float(byrow.median()))
    return rel(o, oref), bool(torch.isfinite(o).all())


def temporal_case(N, H, W, Tq, Tk, causal, nhead=8, d=66):
    C, HW = nhead * d, H * W
    g = torch.Generator(device="cuda").manual_seed(2)
    q = torch.randn(N * Tq * HW, C, device="cuda", generator=g)
    kv = torch.randn(N * Tk * HW, 2 * C, device="cuda", generator=g)
    o = torch.full_like(q, float("nan"))
    ops.attn_fwd(q, kv[:, :C], kv[:, C:], o, None, 1, N, H, W, 0, Tq, Tk, nhead, d, causal, d ** -0.5)
    seq = lambda t, T: t.view(N, T, HW, -1).permute(0, 2, 1, 3).reshape(N * HW, T, -1)
    mask = O.causal_mask(Tq).cuda() if causal else None
    ob = core(seq(q, Tq), seq(kv[:, :C], Tk), seq(kv[:, C:], Tk), nhead, d ** -0.5, mask=mask)
    oref = ob.view(N, HW, Tq, C).permute(0, 2, 1, 3).reshape(N * Tq * HW, C)
    return rel(o, oref), bool(torch.isfinite(o).all())


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "window"):
    for Fr in (2, 3, 40):
        print("window 8x8 ws4 frames=%d: rel err %.3e finite %s" % ((Fr,) + window_case(Fr, 8, 8, 4)), flush=True)
if which in ("all", "temporal"):
    for (N, H, W, Tq, Tk, c) in ((2, 4, 4, 10, 10, False), (2, 8, 8, 10, 10, False), (1, 8, 8, 29, 29, True), (2, 8, 8, 28, 2, False), (2, 4, 4, 5, 2, False)):
        print("temporal N=%d %dx%d Tq=%d Tk=%d causal=%s: rel err %.3e finite %s" % ((N, H, W, Tq, Tk, c) + temporal_case(N, H, W, Tq, Tk, c)), flush=True)
if which in ("all", "time"):
    N, T, H, W, C, nh = 64, 10, 8, 8, 528, 8
    d, R = C // nh, N * T * H * W
    qkv = torch.randn(R, 3 * C, device="cuda"); o = torch.empty(R, C, device="cuda"); table = torch.randn(49, nh, device="cuda")
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    flush = torch.empty(64 * 1024 * 1024, device="cuda")

    def timeit(fn, n=5):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2] * 1e3
    for p in (0.0, 0.1):
        print("tcgen05 window fwd  p=%.1f  %.0f us" % (p, timeit(lambda: ops.attn_fwd(q, k, v, o, table, 0, N * T, H, W, 4, 0, 0, nh, d, False, d ** -0.5, True, 7, p))))
        print("tcgen05 temporal fwd p=%.1f %.0f us" % (p, timeit(lambda: ops.attn_fwd(q, k, v, o, None, 1, N, H, W, 0, T, T, nh, d, False, d ** -0.5, True, 7, p))))
