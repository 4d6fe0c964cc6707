"""ncu target: the bf16x3 raw-tile convolution at cfg1's encoder shape (1280 frames of 8x8, 528 -> 528)"""
import torch
from vptr_b200 import ops
F_, H, W, C = 1280, 8, 8, 528
x = torch.randn(F_ * H * W, C, device="cuda"); w2 = ops.split_bf16x2(torch.randn(C, 9 * C, device="cuda") * 0.02); b = torch.randn(C, device="cuda")
for _ in range(4):
    xq2 = ops.pad_nhwc_quad_bf16x2(x, F_, H, W, C, 1)
    y = ops.conv3x3_bf16x3(xq2, w2, F_, H, W, C, C, bias=b, act=2)
torch.cuda.synchronize()
