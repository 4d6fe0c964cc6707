#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -20
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for c in cfg1 cfg2 cfg3 cfg4; do extra="--no-cpu-baseline"; [ $c = cfg1 ] && extra=""; timeout 900 python bench.py --config $c --steps 10 --warmup 5 $extra > gpurun_out/final_bench_${c}_1gpu.json 2> gpurun_out/final_bench_${c}_1gpu.err; python -c "
import json; d=json.loads(open('gpurun_out/final_bench_${c}_1gpu.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$c', d['value'], d['ms_per_step'], d['e2e']['value'], r['achieved'], r['peak'], r['frac'], d['gpu_launches'], d['config']['peak_mem_gib'], d.get('cpu_baseline',{}).get('value'))
k=d['kernel_rooflines'].get('encoder_conv3x3'); print('    conv', k and (k['avg_launch_us'], k['achieved'], k['executed'], k['frac'], k['frac_executed']))"; tail -1 gpurun_out/final_bench_${c}_1gpu.err | cut -c1-160; done
