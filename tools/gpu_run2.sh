#!/bin/bash
# round-2 GPU batch 2: fixed tests, rollout, op tables per config, ncu launch list + GEMM traffic of one cfg1 step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in tests/test_gpu_tail.py tests/test_gpu_bwd_parity.py tests/test_gpu_rollout.py; do
  echo "=== $f"
  timeout 900 python -m pytest "$f" -q -m gpu -s 2>&1 | grep -v Warning | grep -E "^E  |^>|passed|failed|tf32\]|Error|error" | head -40
done
for cfg in cfg1 cfg2 cfg4; do
  echo "=== op table $cfg"
  timeout 600 python tools/op_table.py --config $cfg --out gpurun_out/op_table_$cfg.txt > /dev/null 2> gpurun_out/op_table_$cfg.err; echo rc=$?; head -28 gpurun_out/op_table_$cfg.txt; tail -3 gpurun_out/op_table_$cfg.err
done
echo "=== bench cfg1 10 steps"
timeout 600 python bench.py --config cfg1 --steps 10 --warmup 5 > gpurun_out/bench_cfg1_b.json 2> gpurun_out/bench_cfg1_b.err; tail -c 2500 gpurun_out/bench_cfg1_b.json
echo "=== bench cfg3 10 steps"
timeout 600 python bench.py --config cfg3 --steps 6 --warmup 4 --no-cpu-baseline > gpurun_out/bench_cfg3_b.json 2> gpurun_out/bench_cfg3_b.err; tail -c 900 gpurun_out/bench_cfg3_b.json
echo "=== ncu launch list (one cfg1 step)"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_cfg1.csv python bench.py --config cfg1 --ncu-step --warmup 2 > gpurun_out/ncu_step.log 2>&1; echo rc=$?; wc -l gpurun_out/launches_cfg1.csv
python tools/summarize_launches.py gpurun_out/launches_cfg1.csv 45 > gpurun_out/launches_cfg1_summary.txt 2>&1; head -50 gpurun_out/launches_cfg1_summary.txt
