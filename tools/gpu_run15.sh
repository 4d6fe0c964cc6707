#!/bin/bash
# round 2: batched layout transposes of the conv-FFN parameters: tests + cfg1/cfg2 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests"
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v Warning | grep -E "^E  |^>|passed|failed|Error|error|^FAILED" | head -30
for cfg in cfg1 cfg2; do
echo "=== bench $cfg"
timeout 600 python bench.py --config $cfg --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${cfg}_trm.json 2> gpurun_out/bench_${cfg}_trm.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_${cfg}_trm.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['peak'], d['gpu_launches'])"; tail -2 gpurun_out/bench_${cfg}_trm.err
done
