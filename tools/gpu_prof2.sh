#!/bin/bash
# round-2 final ncu --set full captures: GEMMs with the TMA-store epilogue (qk projection, fc1, 528+residual) and the quadrant conv
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
GEMM_CASES=1 timeout 600 $NCU -k regex:gemm_tf32_2cta -c 1 -s 3 -f -o gpurun_out/prof2_gemm_qk python tools/bench_gemm.py > gpurun_out/prof2_gemm_qk.log 2>&1; echo "gemm qk rc=$?"
GEMM_CASES=2 timeout 600 $NCU -k regex:gemm_tf32_2cta -c 1 -s 3 -f -o gpurun_out/prof2_gemm_fc1 python tools/bench_gemm.py > gpurun_out/prof2_gemm_fc1.log 2>&1; echo "gemm fc1 rc=$?"
GEMM_CASES=0 timeout 600 $NCU -k regex:gemm_tf32_2cta -c 1 -s 3 -f -o gpurun_out/prof2_gemm_528res python tools/bench_gemm.py > gpurun_out/prof2_gemm_528res.log 2>&1; echo "gemm 528 res rc=$?"
timeout 600 $NCU -k regex:conv3x3_w8 -c 1 -s 2 -f -o gpurun_out/prof2_conv_quad python - > gpurun_out/prof2_conv_quad.log 2>&1 <<'PY'
import torch
from vptr_b200 import ops
F_, H, W, C = 640, 16, 16, 528
x = torch.randn(F_*H*W, C, device="cuda"); w2 = ops.split_tf32(torch.randn(C, 9*C, device="cuda")*0.02); b = torch.randn(C, device="cuda")
for _ in range(4):
    xq = ops.pad_nhwc_quad(x, F_, H, W, C, 1, round_tf32=True)
    y = ops.conv3x3_tf32_quad(xq, w2, F_, H, W, C, C, bias=b, act=2, w_planes=2)
torch.cuda.synchronize()
PY
echo "conv quad rc=$?"
ls -la gpurun_out/prof2_*.ncu-rep
