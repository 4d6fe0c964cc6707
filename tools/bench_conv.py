"""Micro-benchmark (GPU box): implicit-GEMM 3x3 conv (1 / 2 weight planes) vs the plain GEMM of the same contraction, the
fused-epilogue GEMM variants, and a per-tile clock64 timeline of the conv kernel's CTA 0."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vptr_b200 import ops, _lib

flush = torch.empty(64 * 1024 * 1024, device="cuda")


def timeit(fn, n=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e-3


F_, H, W, C = 640, 8, 8, 528
x = torch.randn(F_ * H * W, C, device="cuda")
w = torch.randn(C, 9 * C, device="cuda") * 0.02
bias = torch.randn(C, device="cuda")
xpad = ops.pad_nhwc(x, F_, H, W, C, 1, 1, round_tf32=True)
w1 = ops.round_copy(w)
w2 = ops.split_tf32(w)
fl = 2.0 * F_ * H * W * C * 9 * C
# correctness of the implicit path vs im2col + plain GEMM (same tf32 operands)
colr, _, _ = ops.im2col(x, F_, H, W, C, 3, 1, 1, 1, round_tf32=True)
ref = ops.gemm(colr, w1, bias=bias)
got = ops.conv3x3_tf32(xpad, w1, F_, H, W, C, C, bias=bias, w_planes=1)
res_in = torch.randn_like(ref)
got2 = ops.conv3x3_tf32(xpad, w2, F_, H, W, C, C, bias=bias, residual=res_in, act=0, w_planes=2)
ref2 = ops.gemm(torch.cat([colr, colr], 1), w2, bias=bias, residual=res_in)
torch.cuda.synchronize()
print("implicit vs im2col: rel err planes=1 %.3e, planes=2(+res) %.3e" % (float((got - ref).norm() / ref.norm()), float((got2 - ref2).norm() / ref2.norm())))
for planes, ww in ((1, w1), (2, w2)):
    for act in (0, 2):
        t = timeit(lambda: ops.conv3x3_tf32(xpad, ww, F_, H, W, C, C, bias=bias, act=act, w_planes=planes))
        print("conv3x3 implicit planes=%d act=%d   %7.1f us  %6.1f TFLOP/s (x planes)" % (planes, act, t * 1e6, planes * fl / t / 1e12))
if os.environ.get("CONV_ONLY"):
    sys.exit(0)
col = torch.randn(F_ * H * W, 9 * C, device="cuda")
t = timeit(lambda: ops.gemm(col, w1, bias=bias))
print("plain GEMM K=4752               %7.1f us  %6.1f TFLOP/s" % (t * 1e6, fl / t / 1e12))
col2 = torch.randn(F_ * H * W, 18 * C, device="cuda")
t = timeit(lambda: ops.gemm(col2, w2, bias=bias))
print("plain GEMM K=9504               %7.1f us  %6.1f TFLOP/s" % (t * 1e6, 2 * fl / t / 1e12))

# timeline of CTA 0
buf = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
_lib.lib().vptr_gemm_debug_buffer(buf.data_ptr())
ops.conv3x3_tf32(xpad, w2, F_, H, W, C, C, bias=bias, act=2, w_planes=2)
torch.cuda.synchronize()
_lib.lib().vptr_gemm_debug_buffer(None)
tl = buf.view(-1, 8).cpu()
t0 = int(tl[0, 0])
print("conv planes=2 tile timeline (cycles): mma_wait_start mma_wait_end mma_issued | epi_wait_start epi_full epi_done")
for i in range(8):
    if int(tl[i, 0]) == 0 and i > 0:
        break
    print(i, [int(v) - t0 for v in tl[i, :6]])

# fused-epilogue GEMM variants (M=40960, N=528, K=528)
M, N, K = 40960, 528, 528
A, B, res = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.randn(M, N, device="cuda")
rs = ops.droppath_scales(64, 3, 0.1, "cuda")
for name, kw in (("bias", {}), ("bias+res", dict(residual=res)), ("bias+res+dropout", dict(residual=res, drop_seed=5, drop_p=0.1)),
                 ("bias+res+droppath", dict(residual=res, rowscale=rs, rows_per_group=640)), ("bias+round", dict(round_tf32=True))):
    t = timeit(lambda: ops.gemm(A, B, bias=bias, **kw))
    print("gemm 40960x528x528 %-18s %7.1f us  %6.1f TFLOP/s" % (name, t * 1e6, 2.0 * M * N * K / t / 1e12))
