#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_bwd_parity.py tests/test_gpu_dropout.py -q -m gpu 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head
timeout 900 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_tc30.json 2> gpurun_out/bench_cfg4_tc30.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg4_tc30.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'), d['roofline']['achieved'], d['roofline']['peak'])"; tail -2 gpurun_out/bench_cfg4_tc30.err
