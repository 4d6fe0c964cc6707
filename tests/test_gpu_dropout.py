"""GPU: train-mode dropout / DropPath (statistical parity only -- the reference's Philox stream cannot be reproduced).
Checks: drop rates and 1/keep scaling, forward/backward mask consistency at every fused site, and a finite-difference
check of the whole Transformer's gradient under a fixed seed (proves the backward regenerates the forward's masks)."""
import pytest
import torch

from helpers import build_former, probe, rel_l2

pytestmark = pytest.mark.gpu
P = 0.25


def test_elementwise_sites_share_masks():
    from vptr_b200 import ops
    x = torch.randn(512, 96, device="cuda")
    seed = 1234567
    ones = torch.ones_like(x)
    mask = ops.round_copy(ones, False, None, 0, seed, P)                       # the mask the backward applies
    keep = 1.0 / (1.0 - P)
    assert set(mask.unique().tolist()) <= {0.0, keep} or torch.allclose(mask[mask > 0], torch.tensor(keep, device="cuda"))
    assert abs(float((mask == 0).float().mean()) - P) < 0.02
    # GELU site
    y0, y1 = ops.gelu_fwd(x), ops.gelu_fwd(x, drop_seed=seed, drop_p=P)
    assert rel_l2(y1, y0 * mask) < 1e-6
    g0, g1 = ops.gelu_bwd(ones, x), ops.gelu_bwd(ones, x, drop_seed=seed, drop_p=P)
    assert rel_l2(g1, g0 * mask) < 1e-6
    # GEMM epilogue site (+ DropPath row scale) : out = rowscale * drop(A W^T + b) + res
    A, W, b, res = torch.randn(512, 64, device="cuda"), torch.randn(96, 64, device="cuda"), torch.randn(96, device="cuda"), torch.randn(512, 96, device="cuda")
    rs = ops.droppath_scales(8, 99, 0.5, "cuda")
    assert set(rs.tolist()) <= {0.0, 2.0}
    base = ops.gemm(A, W, bias=b)
    out = ops.gemm(A, W, bias=b, residual=res, rowscale=rs, rows_per_group=64, drop_seed=seed, drop_p=P)
    ref = base * mask * rs.repeat_interleave(64)[:, None] + res
    assert rel_l2(out, ref) < 1e-6
    ops.FORCE_SIMT = True
    try:
        out2 = ops.gemm(A, W, bias=b, residual=res, rowscale=rs, rows_per_group=64, drop_seed=seed, drop_p=P)
    finally:
        ops.FORCE_SIMT = False
    assert rel_l2(out2, ref) < 2e-3
    # backward operand copy applies the same mask and row scale
    dy = torch.randn(512, 96, device="cuda")
    assert rel_l2(ops.round_copy(dy, False, rs, 64 * 96, seed, P), dy * mask * rs.repeat_interleave(64)[:, None]) < 1e-6
    # norm + GELU site
    mean, rstd = ops.group_stats(x, 8)
    gm, bt = torch.ones(64, 96, device="cuda"), torch.zeros(64, 96, device="cuda")
    n0 = ops.norm_act_fwd(x, mean, rstd, gm, bt, 64, 1)
    n1 = ops.norm_act_fwd(x, mean, rstd, gm, bt, 64, 1, res=res, rowscale=rs, rows_per_group=64, drop_seed=seed, drop_p=P)
    assert rel_l2(n1, n0 * mask * rs.repeat_interleave(64)[:, None] + res) < 1e-6


def test_attention_probability_dropout():
    from vptr_b200 import ops
    Fr, H, W, ws, nhead, d = 2, 8, 8, 4, 4, 12
    C, rows = nhead * d, 2 * 64
    qkv = torch.randn(rows, 3 * C, device="cuda")
    v1 = torch.ones(rows, C, device="cuda")                        # V = 1 -> O = sum_j P_ij * keep-scale_ij
    o = torch.empty(rows, C, device="cuda")
    ops.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], v1, o, None, 0, Fr, H, W, ws, 0, 0, nhead, d, False, d ** -0.5, drop_seed=77, drop_p=P)
    assert abs(float(o.mean()) - 1.0) < 0.05 and float(o.std()) > 0.05      # unbiased, but not identically 1
    ops.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], v1, o, None, 0, Fr, H, W, ws, 0, 0, nhead, d, False, d ** -0.5)
    assert float((o - 1).abs().max()) < 1e-5


@pytest.mark.parametrize("ws,H,W,nhead,d", [(4, 8, 8, 8, 66), (8, 16, 16, 8, 66)])
def test_attention_backward_regenerates_the_forward_dropout_mask(ws, H, W, nhead, d):
    """tensor-core window attention (16-token and 64-token groups) with probability dropout: the backward's dq must be the
    derivative of the forward run with the same seed (central finite difference along a random direction)"""
    from vptr_b200 import ops
    Fr, C, rows = 2, nhead * d, 2 * H * W
    g = torch.Generator().manual_seed(9)
    qkv = torch.randn(rows, 3 * C, generator=g).cuda()
    table = (torch.randn((2 * ws - 1) ** 2, nhead, generator=g) * 0.5).cuda()
    do = torch.randn(rows, C, generator=g).cuda()
    v = torch.randn(rows, 3 * C, generator=g).cuda()
    args = (table, 0, Fr, H, W, ws, 0, 0, nhead, d, False, d ** -0.5)

    def f(x):
        o = torch.empty(rows, C, device="cuda")
        ops.attn_fwd(x[:, :C], x[:, C:2 * C], x[:, 2 * C:], o, *args, drop_seed=123, drop_p=0.25)
        return float((o.double() * do.double()).sum())

    dqkv, dtab = torch.empty_like(qkv), torch.zeros_like(table)
    ops.attn_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], do, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], table, dtab, *args[1:],
                 drop_seed=123, drop_p=0.25)
    eps = 1e-2
    fd = (f(qkv + eps * v) - f(qkv - eps * v)) / (2 * eps)
    an = float((dqkv.double() * v.double()).sum())
    assert abs(fd - an) <= 5e-3 * abs(an) + 1e-3, (fd, an)
    o0 = torch.empty(rows, C, device="cuda")
    ops.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o0, *args)
    o1 = torch.empty(rows, C, device="cuda")
    ops.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o1, *args, drop_seed=123, drop_p=0.25)
    assert rel_l2(o1, o0) > 0.05                                           # dropout really active


@pytest.mark.parametrize("name", ["far_rpe", "nar_rpe"])
def test_transformer_gradient_finite_difference_with_dropout(name):
    from vptr_b200 import engine
    net, x, c = build_former(name, "cuda")
    net.dropout = 0.2
    net.train()
    pr = probe((x.shape[0], c["Tf"] if c["kind"] == "nar" else x.shape[1], *x.shape[2:]), 2).cuda()

    def f(inp):
        torch.manual_seed(5)                   # same seed -> same masks
        with torch.no_grad():
            return float((0.5 * net(inp) ** 2 * pr).double().sum())

    with engine.exact_fp32():
        torch.manual_seed(5)
        xin = x.clone().requires_grad_(True)
        y = net(xin)
        (0.5 * y * y * pr).sum().backward()
        net.eval()
        with torch.no_grad():
            y_eval = net(x)
        net.train()
        assert rel_l2(y, y_eval) > 1e-2                                    # dropout is really active
        torch.manual_seed(6)
        with torch.no_grad():
            assert rel_l2(net(x), y) > 1e-3                                # and seed dependent
        g = xin.grad
        v = g / g.norm() * (g.numel() ** 0.5)          # steepest direction, unit-RMS entries: a well-conditioned derivative
        eps = 2e-5                                      # the NAR stack is strongly curved: error ~ eps^2 (tools/debug_fd.py)
        fd = (f(x + eps * v) - f(x - eps * v)) / (2 * eps)
        an = float((g.double() * v.double()).sum())
    assert abs(fd - an) <= 2e-2 * abs(an), (fd, an)
