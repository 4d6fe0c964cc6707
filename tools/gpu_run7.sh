#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py -q -m gpu -s 2>&1 | grep -v Warning | grep -E "^E  |^>|passed|failed|Error|error|^FAILED" | head -40
for cfg in cfg1 cfg2; do
echo "=== bench $cfg graph"
timeout 600 python bench.py --config $cfg --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${cfg}_graph.json 2> gpurun_out/bench_${cfg}_graph.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_${cfg}_graph.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['gpu_launches'], d['config']['cuda_graph'], d['config']['peak_mem_gib'])"; tail -3 gpurun_out/bench_${cfg}_graph.err
echo "=== bench $cfg eager"
timeout 600 python bench.py --config $cfg --steps 10 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/bench_${cfg}_eager.json 2> gpurun_out/bench_${cfg}_eager.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_${cfg}_eager.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['gpu_launches'], d['config']['cuda_graph'])"
done
