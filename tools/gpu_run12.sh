#!/bin/bash
# round 2: TMA-store GEMM epilogue: parity tests, per-shape timing (TMA store vs per-lane stores), cfg1 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== gemm tests"
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_models.py tests/test_gpu_fullsize.py tests/test_gpu_dropout.py -q -m gpu -x 2>&1 | grep -v Warning | grep -E "^E  |^>|passed|failed|Error|error|^FAILED" | head -30
echo "=== bench_gemm TMA-store epilogue"
timeout 300 python tools/bench_gemm.py 2>&1 | tail -12
echo "=== bench_gemm per-lane store epilogue"
VPTR_GEMM_EPI_STG=1 timeout 300 python tools/bench_gemm.py 2>&1 | tail -12
if [ "$1" = bench ]; then
echo "=== bench cfg1"
timeout 600 python bench.py --config cfg1 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg1_tmaepi.json 2> gpurun_out/bench_cfg1_tmaepi.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg1_tmaepi.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['peak'])"; tail -2 gpurun_out/bench_cfg1_tmaepi.err
fi
