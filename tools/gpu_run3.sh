#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in tests/test_gpu_kernels.py tests/test_gpu_dropout.py tests/test_gpu_tail.py tests/test_gpu_rollout.py tests/test_gpu_gemm.py tests/test_gpu_fullsize.py; do
  echo "=== $f"
  timeout 900 python -m pytest "$f" -q -m gpu 2>&1 | grep -v Warning | grep -E "^E  |^>|passed|failed|Error|error|^FAILED" | head -30
done
echo "=== bwd parity"
timeout 900 python -m pytest tests/test_gpu_bwd_parity.py -q -m gpu -s 2>&1 | grep -E "tf32\]|passed|failed|^E " | head
for cfg in cfg4 cfg1; do
  echo "=== op table $cfg"
  timeout 600 python tools/op_table.py --config $cfg --out gpurun_out/op_table_$cfg.txt > /dev/null 2> gpurun_out/op_table_$cfg.err; echo rc=$?; head -16 gpurun_out/op_table_$cfg.txt; tail -3 gpurun_out/op_table_$cfg.err
done
grep "a_mn=1 b_mn=1" gpurun_out/op_table_cfg1.txt | head
echo "=== bench cfg1 / cfg4"
timeout 600 python bench.py --config cfg1 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg1_c.json 2> gpurun_out/bench_cfg1_c.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg1_c.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['peak'])"
timeout 600 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_c.json 2> gpurun_out/bench_cfg4_c.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg4_c.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['peak_mem_gib'])"
