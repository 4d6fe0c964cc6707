#!/bin/bash
# what the driver runs at round end, in one go: GPU test suite, smoke(), default bench, reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -20
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench (default flags)"; timeout 1200 python bench.py > gpurun_out/final_default_bench.json 2> gpurun_out/final_default_bench.err; tail -c 600 gpurun_out/final_default_bench.json; echo; tail -2 gpurun_out/final_default_bench.err
echo "=== bench --impl reference"; timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 700
