#!/bin/bash
# compute-sanitizer passes over the GPU tests (run on the GPU box): memcheck on the kernel / GEMM / tail / model / stage-1 / rollout
# tests, racecheck on the kernels that share tiles between warps.  Round 1: memcheck 0 errors (89 tests); racecheck 0 hazards on the
# attention, depthwise conv, LayerNorm, norm+GELU and head kernels (the only reports were inside torch's own layer_norm backward,
# used by the reference side of the tests).  Round 2 (TMA-store GEMM epilogue, quadrant conv, tail, train-mode BN): see
# profiles/r02_sanitizer.log.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-700}
timeout $T compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py tests/test_gpu_tail.py \
    tests/test_gpu_models.py tests/test_gpu_stage1.py tests/test_gpu_rollout.py -q -m gpu 2>&1 | grep -v "Host Frame\|Device Frame" | tail -8 | tee gpurun_out/sanitizer.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu \
    -k "attention_core or dwconv or layernorm or norm_act or head or elementwise" 2>&1 | grep -v "Host Frame\|Device Frame" | tail -8 | tee -a gpurun_out/sanitizer.log
