#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v Warning | grep -E "^E  |^>|passed|failed|Error|error|^FAILED" | head -40
echo "=== op table cfg1"
timeout 600 python tools/op_table.py --config cfg1 --out gpurun_out/op_table_cfg1.txt > /dev/null 2> gpurun_out/op_table_cfg1.err; echo rc=$?; head -40 gpurun_out/op_table_cfg1.txt; tail -3 gpurun_out/op_table_cfg1.err
echo "=== bench cfg1"
timeout 600 python bench.py --config cfg1 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg1_d.json 2> gpurun_out/bench_cfg1_d.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg1_d.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['peak'], d['gpu_launches'])"; tail -3 gpurun_out/bench_cfg1_d.err
