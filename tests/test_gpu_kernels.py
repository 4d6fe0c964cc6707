"""GPU: every non-GEMM kernel of libvptr_b200.so against the oracle / torch autograd on the same seeded inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import load_golden, rel_l2

import vptr_oracle as O

pytestmark = pytest.mark.gpu
TOL = 2e-5


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def test_integer_artefacts_bit_exact():
    from vptr_b200 import ops
    z = load_golden("integer_artefacts")
    for key in [k for k in z.files if k.startswith("wmap_")]:
        Fr, H, W, ws = map(int, key.split("_")[1:])
        rpi, wmap = ops.window_index_maps(Fr, H, W, ws, "cuda")
        assert np.array_equal(wmap.cpu().numpy(), z[key]), key
        assert np.array_equal(rpi.cpu().numpy(), z["rpi_%d" % ws]), key
    for ws in (2, 7):
        rpi, _ = ops.window_index_maps(1, ws, ws, ws, "cuda")
        assert np.array_equal(rpi.cpu().numpy(), z["rpi_%d" % ws])
    for T in (1, 5, 29):
        assert np.array_equal(ops.causal_mask(T, "cuda").cpu().numpy(), z["causal_%d" % T])
    for hw in (6, 7, 8, 9):   # PadBlock offsets through the pad kernel
        from vptr_b200.engine import Geom
        g = Geom(1, 1, hw, hw, 4, 1, 4)
        m = ops.pad_hw(torch.ones(hw * hw, 4, device="cuda"), 1, hw, hw, g.Hp, g.Wp, g.ph0, g.pw0).view(g.Hp, g.Wp, 4)[:, :, 0]
        assert np.array_equal(m.cpu().numpy(), z["padmask_%d" % hw]), hw


@pytest.mark.parametrize("rows,C", [(640, 528), (37, 48), (1, 8)])
def test_layernorm_fwd_bwd(rows, C):
    from vptr_b200 import ops
    x, gm, bt = rnd(rows, C, seed=1), rnd(C, seed=2) * 0.2 + 1, rnd(C, seed=3) * 0.1
    add = rnd(5, C, seed=4)
    y, y2, mean, rstd = ops.layernorm_fwd(x, gm, bt, add=add, add_div=2, add_mod=5)
    xr = x.clone().requires_grad_(True); gr = gm.clone().requires_grad_(True); br = bt.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (C,), gr, br)
    idx = (torch.arange(rows, device="cuda") // 2) % 5
    assert rel_l2(y, yr) < TOL and rel_l2(y2, yr + add[idx]) < TOL
    dy1, dy2, dres = rnd(rows, C, seed=5), rnd(rows, C, seed=6), rnd(rows, C, seed=7)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx = ops.layernorm_bwd(dy1, dy2, x, gm, bt, mean, rstd, dres, dg, db)
    (yr * (dy1 + dy2)).sum().backward()
    assert rel_l2(dx, xr.grad + dres) < 5 * TOL and rel_l2(dg, gr.grad) < 5 * TOL and rel_l2(db, br.grad) < 5 * TOL
    # relu flavour (final norm + F.relu_)
    yq, _, mean, rstd = ops.layernorm_fwd(x, gm, bt, relu=True)
    xr.grad = None
    yr = F.relu(F.layer_norm(xr, (C,), gm, bt))
    (yr * dy1).sum().backward()
    dx = ops.layernorm_bwd(dy1, None, x, gm, bt, mean, rstd, None, None, None, relu=True)
    assert rel_l2(yq, yr) < TOL and rel_l2(dx, xr.grad) < 5 * TOL


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_norm_act(mode):
    """MlpDWBN norm + GELU (+ residual): BatchNorm train (0), LayerNorm((ch,H,W)) (1), BatchNorm eval (2)."""
    from vptr_b200 import ops
    Fr, H, W, ch = 6, 4, 4, 48
    hw, rows = H * W, Fr * H * W
    x = rnd(rows, ch, seed=1) * 2 + 0.3
    res, dy = rnd(rows, ch, seed=2), rnd(rows, ch, seed=3)
    xr = x.clone().requires_grad_(True)
    if mode == 1:
        gm, bt = rnd(hw, ch, seed=4) * 0.2 + 1, rnd(hw, ch, seed=5) * 0.1
        gr, br = gm.clone().requires_grad_(True), bt.clone().requires_grad_(True)
        mean, rstd = ops.group_stats(x, Fr)
        z = F.layer_norm(xr.view(Fr, hw, ch), (hw, ch), gr, br).view(rows, ch)
    else:
        gm, bt = rnd(ch, seed=4) * 0.2 + 1, rnd(ch, seed=5) * 0.1
        gr, br = gm.clone().requires_grad_(True), bt.clone().requires_grad_(True)
        rm, rv = rnd(ch, seed=6) * 0.1, rnd(ch, seed=7).abs() + 0.5
        if mode == 0:
            rm2, rv2 = rm.clone(), rv.clone()
            mean, rstd = ops.bn_stats(x, rm2, rv2)
            rm3, rv3 = rm.clone(), rv.clone()
            z = F.batch_norm(xr, rm3, rv3, gr, br, training=True, momentum=0.1, eps=1e-5)
            assert rel_l2(rm2, rm3) < TOL and rel_l2(rv2, rv3) < TOL
        else:
            mean, rstd = ops.bn_eval_stats(rm, rv)
            z = F.batch_norm(xr, rm, rv, gr, br, training=False, eps=1e-5)
    yr = F.gelu(z) + res
    y = ops.norm_act_fwd(x, mean, rstd, gm, bt, hw, 1 if mode == 1 else 0, res=res)
    assert rel_l2(y, yr) < TOL
    (yr * dy).sum().backward()
    dg, db = torch.zeros_like(gm), torch.zeros_like(bt)
    dx = ops.norm_act_bwd(dy, x, mean, rstd, gm, bt, dg, db, hw, mode)
    assert rel_l2(dx, xr.grad) < 1e-4 and rel_l2(dg, gr.grad) < 1e-4 and rel_l2(db, br.grad) < 1e-4


@pytest.mark.parametrize("ch", [528, 2112])
def test_norm_act_bwd_path_frame_sizes(ch):
    """Frame LayerNorm + GELU backward at the path's frame sizes (64 tokens x 528 / 2112 channels): DropPath row scales + dropout
    regenerated from the seed, in place (tf32-rounded, accumulating affine gradients) and out of place."""
    from vptr_b200 import ops
    Fr, hw, P, seed = 21, 64, 0.25, 424242
    rows = Fr * hw
    x = rnd(rows, ch, seed=1) * 2 + 0.3
    dy = rnd(rows, ch, seed=3)
    gm, bt = rnd(hw, ch, seed=4) * 0.2 + 1, rnd(hw, ch, seed=5) * 0.1
    xr, gr, br = x.clone().requires_grad_(True), gm.clone().requires_grad_(True), bt.clone().requires_grad_(True)
    mean, rstd = ops.group_stats(x, Fr)
    z = F.layer_norm(xr.view(Fr, hw, ch), (hw, ch), gr, br).view(rows, ch)
    rs = ops.droppath_scales(7, 99, 0.3, "cuda")                                    # 3 frames per clip
    mask = ops.round_copy(torch.ones_like(x), False, rs, 3 * hw * ch, seed, P)      # keep-scale of every element (dropout x DropPath)
    (F.gelu(z) * mask * dy).sum().backward()
    drop = dict(rowscale=rs, rows_per_group=3 * hw, drop_seed=seed, drop_p=P)
    dg, db = torch.zeros_like(gm), torch.zeros_like(bt)
    dx = ops.norm_act_bwd(dy, x, mean, rstd, gm, bt, dg, db, hw, 1, **drop)
    assert rel_l2(dx, xr.grad) < 1e-4 and rel_l2(dg, gr.grad) < 1e-4 and rel_l2(db, br.grad) < 1e-4
    # in place + tf32-rounded output, accumulating into the same affine gradients
    buf = dy.clone()
    out = ops.norm_act_bwd(buf, x, mean, rstd, gm, bt, dg, db, hw, 1, round_tf32=True, inplace=True, **drop)
    assert out.data_ptr() == buf.data_ptr() and rel_l2(out, dx) < 5e-4
    assert (out.view(torch.int32) & 0x1fff).abs().max().item() == 0
    assert rel_l2(dg, 2 * gr.grad) < 1e-4 and rel_l2(db, 2 * br.grad) < 1e-4
    # no regularisation
    xr.grad = None
    (F.gelu(F.layer_norm(xr.view(Fr, hw, ch), (hw, ch), gm, bt)).view(rows, ch) * dy).sum().backward()
    dx0 = ops.norm_act_bwd(dy, x, mean, rstd, gm, bt, torch.zeros_like(gm), torch.zeros_like(bt), hw, 1)
    assert rel_l2(dx0, xr.grad) < 1e-4


def _attn_ref(q, k, v, nhead, scale, bias=None, mask=None):
    """(B,L,C) oracle core with q scaled first, as the reference does"""
    return O._mha_core(q * scale, k, v, nhead, bias=bias, mask=mask)


@pytest.mark.parametrize("ws,H,W,nhead,d", [(4, 8, 8, 8, 66), (2, 4, 6, 4, 12), (8, 8, 8, 2, 20), (8, 16, 16, 8, 66), (6, 12, 6, 4, 66)])
def test_window_attention_core(ws, H, W, nhead, d):
    from vptr_b200 import ops
    Fr, C, L = 3, nhead * d, ws * ws
    rows = Fr * H * W
    qkv = rnd(rows, 3 * C, seed=1)
    table = rnd((2 * ws - 1) ** 2, nhead, seed=2) * 0.5
    scale = d ** -0.5
    o = torch.empty(rows, C, device="cuda")
    ops.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, table, 0, Fr, H, W, ws, 0, 0, nhead, d, False, scale)
    tmap = O.window_token_map(Fr, H, W, ws).cuda()                  # (L,B)
    qr = qkv.clone().requires_grad_(True); tr = table.clone().requires_grad_(True)
    g = lambda t: t[tmap.t()]                                        # (B,L,C)
    bias = tr[O.relative_position_index(ws).cuda().reshape(-1)].reshape(L, L, nhead).permute(2, 0, 1)
    ob = _attn_ref(g(qr[:, :C]), g(qr[:, C:2 * C]), g(qr[:, 2 * C:]), nhead, scale, bias=bias)
    oref = torch.zeros(rows, C, device="cuda").index_put((tmap.t().reshape(-1),), ob.reshape(-1, C))
    assert rel_l2(o, oref) < TOL
    do = rnd(rows, C, seed=3)
    (oref * do).sum().backward()
    dqkv, dtab = torch.empty_like(qkv), torch.zeros_like(table)
    ops.attn_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], do, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], table, dtab, 0, Fr, H,
                 W, ws, 0, 0, nhead, d, False, scale)
    assert rel_l2(dqkv, qr.grad) < 1e-4 and rel_l2(dtab, tr.grad) < 1e-4


@pytest.mark.parametrize("Tq,Tk,causal", [(10, 10, False), (29, 29, True), (5, 2, False), (1, 1, True), (40, 40, True), (64, 33, False),
                                          (3, 50, False)])
def test_temporal_attention_core(Tq, Tk, causal):
    from vptr_b200 import ops
    N, H, W, nhead, d = 2, 4, 4, 8, 66
    C, HW = nhead * d, H * W
    q, kv = rnd(N * Tq * HW, C, seed=1), rnd(N * Tk * HW, 2 * C, seed=2)
    scale = d ** -0.5
    o = torch.empty_like(q)
    ops.attn_fwd(q, kv[:, :C], kv[:, C:], o, None, 1, N, H, W, 0, Tq, Tk, nhead, d, causal, scale)
    qr, kvr = q.clone().requires_grad_(True), kv.clone().requires_grad_(True)
    seq = lambda t, T: t.view(N, T, HW, -1).permute(0, 2, 1, 3).reshape(N * HW, T, -1)
    mask = O.causal_mask(Tq).cuda() if causal else None
    ob = _attn_ref(seq(qr, Tq), seq(kvr[:, :C], Tk), seq(kvr[:, C:], Tk), nhead, scale, mask=mask)
    oref = ob.view(N, HW, Tq, C).permute(0, 2, 1, 3).reshape(N * Tq * HW, C)
    assert rel_l2(o, oref) < TOL
    do = rnd(N * Tq * HW, C, seed=3)
    (oref * do).sum().backward()
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    ops.attn_bwd(q, kv[:, :C], kv[:, C:], do, dq, dkv[:, :C], dkv[:, C:], None, None, 1, N, H, W, 0, Tq, Tk, nhead, d, causal, scale)
    assert rel_l2(dq, qr.grad) < 1e-4 and rel_l2(dkv, kvr.grad) < 1e-4


@pytest.mark.parametrize("case", ["window", "window_tail", "temporal", "causal29", "cross", "temporal30", "cross30x10"])
def test_attention_tcgen05_forward(case):
    """The tcgen05 / TMA / TMEM forward (vptr_attn_fwd_tcgen05, opt-in VPTR_ATTN_TC=1) against the fp32 oracle core.  Operands enter
    the tensor core as TF32, so q/k/v are pre-rounded (as their producing GEMMs do) and the gate is 1e-3 instead of 2e-5."""
    from vptr_b200 import ops
    nhead, d = 8, 66
    C, scale = nhead * d, d ** -0.5
    if case.startswith("window"):
        Fr, H, W, ws = (5 if case == "window_tail" else 4), 8, 8, 4
        L, rows = ws * ws, Fr * H * W
        qkv = ops.round_copy(rnd(rows, 3 * C, seed=1))
        table = rnd((2 * ws - 1) ** 2, nhead, seed=2) * 0.5
        o = torch.full((rows, C), float("nan"), device="cuda")
        ops.attn_fwd_tcgen05(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, table, 0, Fr, H, W, ws, 0, 0, nhead, d, False, scale)
        tmap = O.window_token_map(Fr, H, W, ws).cuda()
        g = lambda t: t[tmap.t()]
        bias = table[O.relative_position_index(ws).cuda().reshape(-1)].reshape(L, L, nhead).permute(2, 0, 1)
        ob = _attn_ref(g(qkv[:, :C]), g(qkv[:, C:2 * C]), g(qkv[:, 2 * C:]), nhead, scale, bias=bias)
        oref = torch.zeros(rows, C, device="cuda").index_put((tmap.t().reshape(-1),), ob.reshape(-1, C))
    else:
        N, H, W = 2, 8, 8
        Tq, Tk, causal = {"temporal": (10, 10, False), "causal29": (29, 29, True), "cross": (28, 2, False), "temporal30": (30, 30, False),
                          "cross30x10": (30, 10, False)}[case]
        HW = H * W
        q, kv = ops.round_copy(rnd(N * Tq * HW, C, seed=1)), ops.round_copy(rnd(N * Tk * HW, 2 * C, seed=2))
        o = torch.full_like(q, float("nan"))
        ops.attn_fwd_tcgen05(q, kv[:, :C], kv[:, C:], o, None, 1, N, H, W, 0, Tq, Tk, nhead, d, causal, scale)
        seq = lambda t, T: t.view(N, T, HW, -1).permute(0, 2, 1, 3).reshape(N * HW, T, -1)
        mask = O.causal_mask(Tq).cuda() if causal else None
        ob = _attn_ref(seq(q, Tq), seq(kv[:, :C], Tk), seq(kv[:, C:], Tk), nhead, scale, mask=mask)
        oref = ob.view(N, HW, Tq, C).permute(0, 2, 1, 3).reshape(N * Tq * HW, C)
    assert torch.isfinite(o).all() and rel_l2(o, oref) < 1e-3
    # probability dropout: the tcgen05 forward must draw exactly the masks the warp-level kernels (forward and backward) regenerate
    # from (seed, batch, head, i, j) -- compare both forwards under the same seed
    o1, o2 = torch.empty_like(o), torch.empty_like(o)
    if case.startswith("window"):
        args = (qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:])
        tail = (table, 0, Fr, H, W, ws, 0, 0, nhead, d, False, scale)
    else:
        args = (q, kv[:, :C], kv[:, C:])
        tail = (None, 1, N, H, W, 0, Tq, Tk, nhead, d, causal, scale)
    ops.attn_fwd_tcgen05(*args, o1, *tail, drop_seed=4242, drop_p=0.3)
    ops.attn_fwd(*args, o2, *tail, drop_seed=4242, drop_p=0.3)
    assert rel_l2(o1, o2) < 1e-3 and rel_l2(o1, o) > 0.05


def test_attention_tcgen05_rejects_unsupported_shapes():
    from vptr_b200 import ops
    q = torch.zeros(64, 48, device="cuda")
    with pytest.raises(RuntimeError):
        ops.attn_fwd_tcgen05(q, q, q, torch.empty_like(q), None, 1, 1, 2, 2, 0, 16, 16, 4, 12, False, 1.0)


def test_round_copy_multi_matches_single():
    from vptr_b200 import ops
    ws = [rnd(528, 528, seed=1), rnd(2112, 528, seed=2), rnd(8, 4, seed=3)]
    outs = ops.round_copy_multi(ws)
    for w, o in zip(ws, outs):
        assert o.shape == w.shape and torch.equal(o, ops.round_copy(w))


@pytest.mark.parametrize("Fr,H,W,ch", [(5, 8, 6, 48), (3, 8, 8, 528), (150, 8, 8, 352), (2, 4, 4, 704)])
def test_dwconv3x3(Fr, H, W, ch):
    """wide channel counts on small grids take the cp.async double-buffered streaming kernel (with a partial last slab for 528)"""
    from vptr_b200 import ops
    x = rnd(Fr * H * W, ch, seed=1)
    w, b = rnd(ch, 1, 3, 3, seed=2), rnd(ch, seed=3)
    w9 = ops.transpose(w, 1, ch, 9)
    y = ops.dwconv3x3(x, w9, b, Fr, H, W)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv2d(xr.view(Fr, H, W, ch).permute(0, 3, 1, 2), wr, br, padding=1, groups=ch).permute(0, 2, 3, 1).reshape(-1, ch)
    assert rel_l2(y, yr) < TOL
    # fused statistics variant: same output, (mean, rstd) per frame equal to a separate pass
    y2, (m2, r2) = ops.dwconv3x3_stats(x, w9, b, Fr, H, W)
    m1, r1 = ops.group_stats(y, Fr)
    assert torch.equal(y2, y) and rel_l2(m2, m1) < 1e-5 and rel_l2(r2, r1) < 1e-5
    dy = rnd(Fr * H * W, ch, seed=4)
    (yr * dy).sum().backward()
    dx = ops.dwconv3x3(dy, w9, None, Fr, H, W, flip=True)
    dw9, db = torch.zeros(9 * ch, device="cuda"), torch.zeros(ch, device="cuda")
    ops.dwconv3x3_wgrad(x, dy, dw9, db, Fr, H, W)
    dw = torch.zeros_like(w)
    ops.transpose(dw9, 1, 9, ch, out=dw, accumulate=True)
    assert rel_l2(dx, xr.grad) < TOL and rel_l2(dw, wr.grad) < 1e-4 and rel_l2(db, br.grad) < 1e-4


def test_elementwise_helpers():
    from vptr_b200 import ops
    x, y = rnd(300, 48, seed=1), rnd(300, 48, seed=2)
    assert rel_l2(ops.axpby(x, y, 2.0, -0.5), 2 * x - 0.5 * y) < 1e-6
    assert rel_l2(ops.gelu_fwd(x), F.gelu(x)) < 1e-6
    xr = x.clone().requires_grad_(True)
    (F.gelu(xr) * y).sum().backward()
    assert rel_l2(ops.gelu_bwd(y, x), xr.grad) < 1e-5
    assert rel_l2(ops.relu_bwd(y, ops.relu_fwd(x)), y * (x > 0)) < 1e-6
    out = torch.ones(48, device="cuda")
    ops.colsum(x, out)
    assert rel_l2(out, x.sum(0) + 1) < 1e-5
    x50 = rnd(301, 50, seed=4)              # a width that is not a multiple of 4 takes the scalar kernel
    out50 = torch.ones(50, device="cuda")
    ops.colsum(x50, out50)
    assert rel_l2(out50, x50.sum(0) + 1) < 1e-5
    # widths that are multiples of 528 take the tiled kernel, also on a column slice of a wider buffer and a ragged row count
    big = rnd(4099, 1584, seed=9)
    for sl in (slice(0, 528), slice(528, 1584), slice(0, 1584)):
        o = torch.full((sl.stop - sl.start,), 2.0, device="cuda")
        ops.colsum(big[:, sl], o)
        assert rel_l2(o, big[:, sl].double().sum(0) + 2) < 1e-5
    t = ops.transpose(x.view(3, 100, 48), 3, 100, 48).view(3, 48, 100)
    assert torch.equal(t, x.view(3, 100, 48).transpose(1, 2).contiguous())
    add = rnd(5, 48, seed=3)
    idx = (torch.arange(300, device="cuda") // 4) % 5
    assert rel_l2(ops.add_rows(x, add, 4, 5), x + add[idx]) < 1e-6
    acc = torch.zeros(100 * 48, device="cuda")
    ops.rowgroup_sum(x, acc, 3)
    assert rel_l2(acc.view(100, 48), x.view(3, 100, 48).sum(0)) < 1e-5
    p = ops.pad_hw(x[:6 * 7 * 2].contiguous().view(-1, 48)[:84], 2, 6, 7, 8, 8, 1, 0)
    ref = F.pad(x[:84].view(2, 6, 7, 48), (0, 0, 0, 1, 1, 1))
    assert torch.equal(p.view(2, 8, 8, 48), ref)
    assert torch.equal(ops.crop_hw(p, 2, 6, 7, 8, 8, 1, 0), x[:84])
    sq = torch.zeros(1, dtype=torch.float64, device="cuda")
    ops.sqnorm_accumulate(x, sq)
    assert abs(float(sq) - float(x.double().square().sum())) < 1e-6 * float(sq)
    xc = x.clone()
    ops.clip_scale(xc, sq, 1.0)
    assert rel_l2(xc, x * (1.0 / (float(sq) ** 0.5 + 1e-6))) < 1e-5


@pytest.mark.parametrize("k,stride,pad,mode", [(3, 1, 1, "reflect"), (3, 1, 1, "zero"), (3, 2, 1, "zero"), (3, 1, 1, "replicate")])
def test_im2col_gemm_is_conv(k, stride, pad, mode):
    from vptr_b200 import ops
    Fr, H, W, Ci, Co = 2, 8, 8, 16, 24
    x = rnd(Fr * H * W, Ci, seed=1)
    w = rnd(Co, Ci, k, k, seed=2) * 0.1
    scale = rnd(Co, seed=3) * 0.1 + 1
    col, Ho, Wo = ops.im2col(x, Fr, H, W, Ci, k, stride, pad, ops.PAD_MODES[mode], round_tf32=False)
    wpk = ops.pack_conv_weight(w, scale, 0).view(Co, k * k * Ci)
    ops.FORCE_SIMT = True
    try:
        y = ops.gemm(col, wpk)
    finally:
        ops.FORCE_SIMT = False
    xn = x.view(Fr, H, W, Ci).permute(0, 3, 1, 2)
    xp = F.pad(xn, (pad,) * 4, mode={"zero": "constant", "reflect": "reflect", "replicate": "replicate"}[mode])
    yr = (F.conv2d(xp, w, stride=stride) * scale[None, :, None, None]).permute(0, 2, 3, 1).reshape(-1, Co)
    assert rel_l2(y, yr) < TOL


def test_convT_gather_and_backward():
    from vptr_b200 import ops
    Fr, H, W, Ci, Co = 2, 4, 4, 24, 16
    x = rnd(Fr * H * W, Ci, seed=1)
    w = rnd(Ci, Co, 3, 3, seed=2) * 0.2
    scale, shift = rnd(Co, seed=3) * 0.1 + 1, rnd(Co, seed=4) * 0.1
    wpk = ops.pack_conv_weight(w, scale, 1).view(9 * Co, Ci)
    ops.FORCE_SIMT = True
    try:
        col = ops.gemm(x, wpk)
        y = ops.convT_gather(col, shift, Fr, H, W, Co, relu=True)
        xr = x.clone().requires_grad_(True)
        z = F.conv_transpose2d(xr.view(Fr, H, W, Ci).permute(0, 3, 1, 2), w, stride=2, padding=1, output_padding=1)
        yr = F.relu(z * scale[None, :, None, None] + shift[None, :, None, None]).permute(0, 2, 3, 1).reshape(-1, Co)
        assert rel_l2(y, yr) < TOL
        dy = rnd(Fr * 4 * H * W, Co, seed=5)
        (yr * dy).sum().backward()
        colb, _, _ = ops.im2col(dy, Fr, 2 * H, 2 * W, Co, 3, 2, 1, 0, mask=y, round_tf32=False)
        dx = ops.gemm(colb, wpk, b_mn=True)
    finally:
        ops.FORCE_SIMT = False
    assert rel_l2(dx, xr.grad) < TOL


@pytest.mark.parametrize("Ci,H,W", [(1, 32, 24), (3, 32, 24), (1, 64, 64), (3, 16, 32), (1, 40, 16), (1, 128, 128), (3, 24, 192),
                                    (3, 64, 64), (4, 32, 64), (3, 128, 128), (2, 40, 16)])
def test_stem_and_head(Ci, H, W):
    """W in {16, 32, 64} takes the ring-buffered head kernel (head_conv7x7_p_kernel) one frame per block, multiples of 64 beyond
    that (128: cfg4; 192: an interior band with halo columns on both sides) in 64-column bands, other widths the direct kernel"""
    from vptr_b200 import ops
    Fr = 2
    x = rnd(Fr, Ci, H, W, seed=1)
    w = rnd(64, Ci, 7, 7, seed=2) * 0.1
    scale, shift = rnd(64, seed=3) * 0.1 + 1, rnd(64, seed=4) * 0.1
    y = ops.stem_conv7x7(x, ops.pack_conv_weight(w, scale, 2), shift, Fr, Ci, H, W, 64)
    yr = F.relu(F.conv2d(F.pad(x, (3,) * 4, mode="reflect"), w) * scale[None, :, None, None] + shift[None, :, None, None])
    assert rel_l2(y.view(Fr, H, W, 64).permute(0, 3, 1, 2), yr) < TOL
    # head: 64 -> Ci, bias, tanh / sigmoid, and its input gradient
    for act, fn in ((1, torch.tanh), (2, torch.sigmoid)):
        h = rnd(Fr * H * W, 64, seed=5)
        wh, bh = rnd(Ci, 64, 7, 7, seed=6) * 0.05, rnd(Ci, seed=7) * 0.1
        out = ops.head_conv7x7_fwd(h, ops.pack_conv_weight(wh, None, 3), bh, Fr, 64, Ci, H, W, act)
        hr = h.clone().requires_grad_(True)
        outr = fn(F.conv2d(F.pad(hr.view(Fr, H, W, 64).permute(0, 3, 1, 2), (3,) * 4, mode="reflect"), wh, bh))
        assert rel_l2(out, outr) < TOL
        dout = rnd(Fr, Ci, H, W, seed=8)
        (outr * dout).sum().backward()
        dx = ops.head_conv7x7_bwd(dout, out, wh, Fr, 64, Ci, H, W, act)
        assert rel_l2(dx, hr.grad) < 5 * TOL


def test_producers_emit_bias_gradient_column_sums():
    """the kernels that produce a dY also add its column sums to a bias-gradient accumulator (no separate colsum pass)"""
    from vptr_b200 import ops
    rows, C = 640, 528
    x = rnd(rows, C, seed=1)
    rs = ops.droppath_scales(10, 5, 0.3, "cuda")
    acc = torch.full((C,), 2.0, device="cuda")
    y = ops.round_copy(x, True, rs, 64 * C, 77, 0.2, colsum=acc)
    assert torch.equal(y, ops.round_copy(x, True, rs, 64 * C, 77, 0.2))
    assert rel_l2(acc - 2.0, y.double().sum(0)) < 1e-5
    h, du = rnd(rows, 2112, seed=2), rnd(rows, 2112, seed=3)
    acc = torch.zeros(2112, device="cuda")
    dh = ops.gelu_bwd(du, h, round_tf32=True, drop_seed=9, drop_p=0.1, colsum=acc)
    assert torch.equal(dh, ops.gelu_bwd(du, h, round_tf32=True, drop_seed=9, drop_p=0.1))
    assert rel_l2(acc, dh.double().sum(0)) < 1e-5
    for mode, ch, hw in ((1, 2112, 64), (0, 528, 64), (2, 2112, 64), (1, 48, 16)):
        Fr = 5
        r = Fr * hw
        xx, dy = rnd(r, ch, seed=4), rnd(r, ch, seed=5)
        if mode == 1:
            mean, rstd = ops.group_stats(xx, Fr)
            gm, bt = rnd(hw, ch, seed=6) * 0.2 + 1, rnd(hw, ch, seed=7) * 0.1
        else:
            mean, rstd = xx.mean(0).contiguous(), (xx.var(0, unbiased=False) + 1e-5).rsqrt().contiguous()
            gm, bt = rnd(ch, seed=6) * 0.2 + 1, rnd(ch, seed=7) * 0.1
        z = lambda: (torch.zeros_like(gm), torch.zeros_like(bt))
        dg0, db0 = z()
        ref = ops.norm_act_bwd(dy, xx, mean, rstd, gm, bt, dg0, db0, hw, mode, round_tf32=True, drop_seed=3, drop_p=0.1)
        dg1, db1 = z()
        acc = torch.zeros(ch, device="cuda")
        out = ops.norm_act_bwd(dy, xx, mean, rstd, gm, bt, dg1, db1, hw, mode, round_tf32=True, drop_seed=3, drop_p=0.1, colsum=acc)
        assert rel_l2(out, ref) < 1e-6 and rel_l2(dg1, dg0) < 1e-5 and rel_l2(db1, db0) < 1e-5, mode
        # (train-mode BatchNorm: the column sums of dx vanish mathematically, so the gate is relative to the sum of magnitudes)
        assert float((acc - out.double().sum(0)).abs().max()) < 1e-5 * float(out.abs().sum(0).max()), mode


@pytest.mark.parametrize("mode,ws,H,W,Tq,Tk,nhead,d", [(0, 4, 8, 8, 0, 0, 8, 66), (0, 8, 16, 16, 0, 0, 8, 66), (1, 0, 4, 4, 10, 10, 8, 66),
                                                      (1, 0, 4, 4, 28, 2, 8, 66), (0, 2, 4, 6, 0, 0, 4, 12)])
def test_attention_backward_emits_projection_bias_gradients(mode, ws, H, W, Tq, Tk, nhead, d):
    from vptr_b200 import ops
    C = nhead * d
    N = 2
    rq = N * H * W * (Tq if mode else 1)
    rk = N * H * W * (Tk if mode else 1)
    q, k, v, do = rnd(rq, C, seed=1), rnd(rk, C, seed=2), rnd(rk, C, seed=3), rnd(rq, C, seed=4)
    table = rnd((2 * ws - 1) ** 2, nhead, seed=5) * 0.5 if mode == 0 else None
    args = (mode, N, H, W, ws, Tq, Tk, nhead, d, False, d ** -0.5)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    dtab = torch.zeros_like(table) if table is not None else None
    ops.attn_bwd(q, k, v, do, dq, dk, dv, table, dtab, *args)
    dq2, dk2, dv2 = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    bq, bk, bv = (torch.full((C,), 1.0, device="cuda") for _ in range(3))
    dtab2 = torch.zeros_like(table) if table is not None else None
    ops.attn_bwd(q, k, v, do, dq2, dk2, dv2, table, dtab2, *args, dbq=bq, dbk=bk, dbv=bv)
    assert rel_l2(dq2, dq) < 1e-5 and rel_l2(dk2, dk) < 1e-5 and rel_l2(dv2, dv) < 1e-5
    for b, t in ((bq, dq), (bk, dk), (bv, dv)):
        assert float((b - 1.0 - t.double().sum(0)).abs().max()) < 1e-4 * float(t.abs().sum(0).max())
