#!/bin/bash
# round 2 final pass (1 GPU): full GPU test suite, smoke, the four BASELINE.json configs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | grep -E "^E  |^>|passed|failed|Error|error|^FAILED" | head -40
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
for cfg in cfg1 cfg2 cfg3 cfg4; do
  echo "=== bench $cfg"
  extra="--no-cpu-baseline"; [ $cfg = cfg1 ] && extra=""
  timeout 1200 python bench.py --config $cfg --steps 10 --warmup 5 $extra > gpurun_out/final_bench_${cfg}_1gpu.json 2> gpurun_out/final_bench_${cfg}_1gpu.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/final_bench_${cfg}_1gpu.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'), 'roof', d['roofline']['achieved'], d['roofline']['peak'], d['roofline']['frac'], 'mem', d['config'].get('peak_mem_gib'), 'graph:', d['config'].get('cuda_graph'), 'cpu', d.get('cpu_baseline'))
except Exception as e:
    print('FAILED', e)
PY
  tail -2 gpurun_out/final_bench_${cfg}_1gpu.err
done
