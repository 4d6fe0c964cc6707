"""GPU: the tcgen05 TF32 GEMM (through the C-ABI) against fp64 torch.matmul on tf32-exact inputs, and against the
fp32 FFMA kernel on arbitrary inputs.  Covers all operand majors, ragged M/N/K tails, epilogues and split-K."""
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu


def tf32_exact(shape, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(*shape, generator=g)
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32).cuda()   # low 13 mantissa bits cleared: exact in tf32


def rnd(*shape, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g).cuda()


def ref_gemm(A, B, a_mn, b_mn):
    A64 = A.double().t() if a_mn else A.double()
    B64 = B.double() if b_mn else B.double().t()
    return A64 @ B64


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 176, 32), (256, 528, 528), (300, 48, 100), (1000, 2112, 528), (77, 12, 8), (4096, 528, 2112)])
def test_gemm_majors_and_tails(a_mn, b_mn, M, N, K):
    from vptr_b200 import ops
    A = tf32_exact((K, M) if a_mn else (M, K), 1)
    B = tf32_exact((K, N) if b_mn else (N, K), 2)
    D = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    ref = ref_gemm(A, B, a_mn, b_mn)
    assert rel_l2(D, ref) < 1e-5, (M, N, K)   # fp32 accumulation over K


def test_gemm_epilogues():
    from vptr_b200 import ops
    M, N, K = 520, 528, 264
    A, B = tf32_exact((M, K), 3), tf32_exact((N, K), 4)
    bias, res = torch.randn(N, device="cuda"), torch.randn(M, N, device="cuda")
    base = ref_gemm(A, B, False, False)
    D = ops.gemm(A, B, bias=bias, act=ops.ACT_GELU, residual=res, alpha=0.5)
    ref = torch.nn.functional.gelu(0.5 * base + bias.double()) + res.double()
    assert rel_l2(D, ref) < 2e-6
    D = ops.gemm(A, B, bias=bias, act=ops.ACT_RELU)
    assert rel_l2(D, torch.relu(base + bias.double())) < 2e-6
    # strided output / operands (column slices of wider buffers, as the fused qkv buffer uses)
    wide = torch.zeros(M, 3 * N, device="cuda")
    ops.gemm(A, B, out=wide[:, N:2 * N], bias=bias)
    assert rel_l2(wide[:, N:2 * N], base + bias.double()) < 2e-6
    assert float(wide[:, :N].abs().max()) == 0.0 and float(wide[:, 2 * N:].abs().max()) == 0.0
    # in-place accumulate through the residual pointer
    acc = res.clone()
    ops.gemm(A, B, out=acc, residual=acc)
    assert rel_l2(acc, base + res.double()) < 2e-6


def test_gemm_splitk_accumulate():
    from vptr_b200 import ops
    tokens, Nout, Kin = 5000, 528, 2112
    dY, X = tf32_exact((tokens, Nout), 5), tf32_exact((tokens, Kin), 6)
    dW = torch.ones(Nout, Kin, device="cuda")
    ops.gemm(dY, X, out=dW, a_mn=True, b_mn=True, accumulate=True)
    ref = dY.double().t() @ X.double() + 1.0
    assert rel_l2(dW, ref) < 5e-6


def test_gemm_tf32_vs_fp32_kernel():
    """arbitrary fp32 inputs: tensor-core result within tf32 truncation error of the FFMA kernel"""
    from vptr_b200 import ops
    M, N, K = 2048, 528, 528
    A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    D = ops.gemm(A, B)
    ops.FORCE_SIMT = True
    try:
        Ds = ops.gemm(A, B)
    finally:
        ops.FORCE_SIMT = False
    assert rel_l2(Ds, A.double() @ B.double().t()) < 1e-6
    assert rel_l2(D, Ds) < 2e-3


def test_gemm_rejects_cpu_tensor():
    from vptr_b200 import ops
    with pytest.raises(RuntimeError):
        ops.gemm(torch.randn(8, 8), torch.randn(8, 8))


@pytest.mark.parametrize("Fr,H,W,Ci,Co,mode", [(5, 8, 8, 48, 64, "reflect"), (3, 4, 4, 16, 24, "zero"), (2, 16, 16, 32, 176, "replicate"),
                                                (4, 8, 8, 528, 528, "reflect")])
def test_implicit_gemm_conv3x3(Fr, H, W, Ci, Co, mode):
    """4-D TMA implicit-GEMM convolution vs F.conv2d on tf32-exact operands (+ bias, ReLU, residual)"""
    import torch.nn.functional as F
    from vptr_b200 import ops
    x = tf32_exact((Fr * H * W, Ci), 11)
    w = tf32_exact((Co, Ci, 3, 3), 12) * 0.25
    bias, res = torch.randn(Co, device="cuda"), torch.randn(Fr * H * W, Co, device="cuda")
    assert ops.conv3x3_implicit_ok(H, W)
    xpad = ops.pad_nhwc(x, Fr, H, W, Ci, 1, ops.PAD_MODES[mode], round_tf32=False)
    wpk = ops.pack_conv_weight(w, None, 0).view(Co, 9 * Ci)
    y = ops.conv3x3_tf32(xpad, wpk, Fr, H, W, Ci, Co, bias=bias, residual=res, act=ops.ACT_RELU)
    # weight-split variant on NON-tf32-exact weights: the [hi|lo] planes make the weights exact to ~2^-22
    w2 = torch.randn(Co, Ci, 3, 3, device="cuda") * 0.25
    y2 = ops.conv3x3_tf32(xpad, ops.split_tf32(ops.pack_conv_weight(w2, None, 0).view(Co, 9 * Ci)), Fr, H, W, Ci, Co, w_planes=2)
    xn2 = x.view(Fr, H, W, Ci).permute(0, 3, 1, 2).double()
    xp2 = F.pad(xn2, (1,) * 4, mode={"zero": "constant", "reflect": "reflect", "replicate": "replicate"}[mode])
    assert rel_l2(y2, F.conv2d(xp2, w2.double()).permute(0, 2, 3, 1).reshape(-1, Co)) < 3e-5
    xn = x.view(Fr, H, W, Ci).permute(0, 3, 1, 2).double()
    xp = F.pad(xn, (1,) * 4, mode={"zero": "constant", "reflect": "reflect", "replicate": "replicate"}[mode])
    ref = torch.relu(F.conv2d(xp, w.double(), bias.double())).permute(0, 2, 3, 1).reshape(-1, Co) + res.double()
    assert rel_l2(y, ref) < 3e-5     # fp32 accumulation over K = 9*Ci


@pytest.mark.parametrize("Fr,H,W,Ci,Co,mode", [(3, 16, 16, 48, 64, "reflect"), (1, 16, 24, 32, 180, "zero"), (5, 24, 8, 16, 24, "replicate"),
                                                (2, 16, 16, 528, 528, "reflect")])
def test_quadrant_conv3x3(Fr, H, W, Ci, Co, mode):
    """grids beyond 8x8 on the raw-tile kernel: every 8x8 quadrant with its own halo (vptr_pad_nhwc_quad) is one "frame" of
    conv3x3_w8_kernel, the epilogue maps rows back into the full frame -- vs F.conv2d (+ bias, ReLU, residual, [hi|lo] weights),
    odd quadrant counts exercising the masked last pair tile"""
    import torch.nn.functional as F
    from vptr_b200 import ops
    import os
    assert ops.conv3x3_quad_ok(H, W) == (os.environ.get("VPTR_CONV_GENERIC", "") != "1")     # the host-side dispatch predicate
    x = tf32_exact((Fr * H * W, Ci), 21)
    w = tf32_exact((Co, Ci, 3, 3), 22) * 0.25
    bias, res = torch.randn(Co, device="cuda"), torch.randn(Fr * H * W, Co, device="cuda")
    xq = ops.pad_nhwc_quad(x, Fr, H, W, Ci, ops.PAD_MODES[mode], round_tf32=False)
    xp = F.pad(x.view(Fr, H, W, Ci).permute(0, 3, 1, 2).double(), (1,) * 4,
               mode={"zero": "constant", "reflect": "reflect", "replicate": "replicate"}[mode])
    # the tiled copy itself, bit for bit: quadrant (qy, qx) = the 10x10 patch of the padded frame at (8 qy, 8 qx)
    patches = xp.float().unfold(2, 10, 8).unfold(3, 10, 8).permute(0, 2, 3, 4, 5, 1).reshape(-1, Ci)
    assert torch.equal(xq, patches)
    wpk = ops.pack_conv_weight(w, None, 0).view(Co, 9 * Ci)
    y = ops.conv3x3_tf32_quad(xq, wpk, Fr, H, W, Ci, Co, bias=bias, residual=res, act=ops.ACT_RELU)
    ref = torch.relu(F.conv2d(xp, w.double(), bias.double())).permute(0, 2, 3, 1).reshape(-1, Co) + res.double()
    assert rel_l2(y, ref) < 3e-5
    w2 = torch.randn(Co, Ci, 3, 3, device="cuda") * 0.25
    y2 = ops.conv3x3_tf32_quad(xq, ops.split_tf32(ops.pack_conv_weight(w2, None, 0).view(Co, 9 * Ci)), Fr, H, W, Ci, Co, w_planes=2)
    assert rel_l2(y2, F.conv2d(xp, w2.double()).permute(0, 2, 3, 1).reshape(-1, Co)) < 3e-5
    # and agrees with the generic per-tap implicit GEMM where that one applies
    if ops.conv3x3_implicit_ok(H, W):
        xpad = ops.pad_nhwc(x, Fr, H, W, Ci, 1, ops.PAD_MODES[mode], round_tf32=False)
        yg = ops.conv3x3_tf32(xpad, wpk, Fr, H, W, Ci, Co, bias=bias, residual=res, act=ops.ACT_RELU)
        assert rel_l2(y, yg) < 2e-5     # two fp32 summation orders over K = 9*Ci
    with pytest.raises(RuntimeError):      # the raw-tile epilogue has no GELU: refused, not silently skipped
        ops.conv3x3_tf32_quad(xq, wpk, Fr, H, W, Ci, Co, act=ops.ACT_GELU)


@pytest.mark.parametrize("Fr,H,W,Ci,Co,mode", [(5, 8, 8, 48, 64, "reflect"), (3, 16, 16, 64, 180, "zero"), (7, 24, 8, 16, 24, "replicate"),
                                                (4, 8, 8, 528, 528, "reflect"), (2, 16, 16, 528, 528, "reflect")])
def test_conv3x3_bf16x3(Fr, H, W, Ci, Co, mode):
    """the ResnetBlock convolution with both operands as two bf16 planes and three bf16 tensor-core passes (hi*hi + lo*hi + hi*lo)
    on the raw-tile kernel: arbitrary fp32 inputs (nothing pre-rounded), within ~2^-16 of fp64 -- 8x8 grid and quadrant grids,
    channel counts that end inside a 64-channel slice, odd quadrant counts (masked last pair tile), bias / ReLU / residual"""
    import torch.nn.functional as F
    from vptr_b200 import ops
    x = rnd(Fr * H * W, Ci, seed=31)
    w = rnd(Co, Ci, 3, 3, seed=32) * 0.25
    bias, res = torch.randn(Co, device="cuda"), torch.randn(Fr * H * W, Co, device="cuda")
    xq2 = ops.pad_nhwc_quad_bf16x2(x, Fr, H, W, Ci, ops.PAD_MODES[mode])
    xp = F.pad(x.view(Fr, H, W, Ci).permute(0, 3, 1, 2).double(), (1,) * 4,
               mode={"zero": "constant", "reflect": "reflect", "replicate": "replicate"}[mode])
    # the two planes: hi = bf16(x), lo = bf16(x - hi), tiled like vptr_pad_nhwc_quad
    patches = xp.float().unfold(2, 10, 8).unfold(3, 10, 8).permute(0, 2, 3, 4, 5, 1).reshape(-1, Ci)
    hi = patches.bfloat16()
    assert torch.equal(xq2[0], hi) and torch.equal(xq2[1], (patches - hi.float()).bfloat16())
    w2 = ops.split_bf16x2(ops.pack_conv_weight(w, None, 0).view(Co, 9 * Ci))
    y = ops.conv3x3_bf16x3(xq2, w2, Fr, H, W, Ci, Co, bias=bias, residual=res, act=ops.ACT_RELU)
    ref = torch.relu(F.conv2d(xp, w.double(), bias.double())).permute(0, 2, 3, 1).reshape(-1, Co) + res.double()
    err = rel_l2(y, ref)
    print("bf16x3 conv rel_l2 vs fp64: %.2e" % err)
    assert err < 3e-5
    y0 = ops.conv3x3_bf16x3(xq2, w2, Fr, H, W, Ci, Co)
    assert rel_l2(y0, F.conv2d(xp, w.double()).permute(0, 2, 3, 1).reshape(-1, Co)) < 3e-5
    with pytest.raises(RuntimeError):
        ops.conv3x3_bf16x3(xq2, w2, Fr, H, W, Ci, Co, act=ops.ACT_GELU)
