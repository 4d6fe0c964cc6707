#!/bin/bash
# round-2 ncu --set full captures of the kernels the bench line talks about (one GPU; summaries go to profiles/ via tools/ncu_summary.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:attn_mma64 -c 2 -s 2 -f -o gpurun_out/prof_attn_mma64 python tools/prof_attn64.py > gpurun_out/prof_attn64.log 2>&1; echo "attn64 rc=$?"
GEMM_CASES=2 timeout 600 $NCU -k regex:gemm_tf32_2cta -c 1 -s 3 -f -o gpurun_out/prof_gemm_fc1 python tools/bench_gemm.py > gpurun_out/prof_gemm_fc1.log 2>&1; echo "gemm fc1 rc=$?"
GEMM_CASES=0 timeout 600 $NCU -k regex:gemm_tf32_2cta -c 1 -s 3 -f -o gpurun_out/prof_gemm_528res python tools/bench_gemm.py > gpurun_out/prof_gemm_528res.log 2>&1; echo "gemm 528 res rc=$?"
GEMM_CASES=8 timeout 600 $NCU -k regex:gemm_tf32_2cta -c 1 -s 3 -f -o gpurun_out/prof_gemm_wgrad528 python tools/bench_gemm.py > gpurun_out/prof_gemm_wgrad528.log 2>&1; echo "gemm wgrad rc=$?"
timeout 600 $NCU -k regex:"attn_mma_kernel|attn_tc_fwd" -c 2 -s 2 -f -o gpurun_out/prof_attn_cfg1 python tools/prof_attn.py > gpurun_out/prof_attn_cfg1.log 2>&1; echo "attn cfg1 rc=$?"
timeout 600 $NCU -k regex:"ln3_act_bwd|norm_act_bwd_dx" -c 2 -s 2 -f -o gpurun_out/prof_norm python tools/bench_norm.py > gpurun_out/prof_norm.log 2>&1; echo "norm rc=$?"
ls -la gpurun_out/*.ncu-rep
