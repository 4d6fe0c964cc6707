#!/bin/bash
# round-2 GPU batch 1: new tests first (each file under its own timeout), then the old suite, then short bench lines per config
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in tests/test_gpu_tail.py tests/test_gpu_dropin.py tests/test_gpu_bwd_parity.py; do
  echo "=== $f"
  timeout 600 python -m pytest "$f" -q -m gpu -s 2>&1 | grep -v Warning | tail -25
done
echo "=== old suite"
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_tail.py --deselect tests/test_gpu_dropin.py --deselect tests/test_gpu_bwd_parity.py 2>&1 | tail -5
for cfg in cfg1 cfg2 cfg3 cfg4; do
  echo "=== bench $cfg"
  timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  echo "rc=$?"; tail -c 1500 gpurun_out/bench_$cfg.json; tail -3 gpurun_out/bench_$cfg.err
done
echo "=== bench cfg1 torch tail"
timeout 600 python bench.py --config cfg1 --steps 3 --warmup 3 --no-cpu-baseline --torch-tail > gpurun_out/bench_cfg1_torchtail.json 2>&1; tail -c 600 gpurun_out/bench_cfg1_torchtail.json
