#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== model / parity / graph tests"
timeout 1800 python -m pytest tests/test_gpu_models.py tests/test_gpu_bwd_parity.py tests/test_gpu_graph.py tests/test_gpu_dropout.py tests/test_gpu_tail.py tests/test_gpu_dropin.py -q -m gpu 2>&1 | grep -v Warning | grep -E "^E  |^>|passed|failed|Error|error|^FAILED" | head -40
for cfg in cfg1 cfg2; do
echo "=== bench $cfg graph + overlap"
timeout 600 python bench.py --config $cfg --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${cfg}_ov.json 2> gpurun_out/bench_${cfg}_ov.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_${cfg}_ov.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['gpu_launches'], d['config']['cuda_graph'][:30], d['config']['peak_mem_gib'])"; tail -2 gpurun_out/bench_${cfg}_ov.err
done
echo "=== bench cfg1 eager + overlap"
timeout 600 python bench.py --config cfg1 --steps 10 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/bench_cfg1_ov_eager.json 2> gpurun_out/bench_cfg1_ov_eager.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg1_ov_eager.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
