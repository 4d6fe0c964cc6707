#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== full gpu suite"
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | grep -E "^E  |^>|passed|failed|Error|error|^FAILED" | head -40
echo "=== dp test"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/test_dp.py > gpurun_out/test_dp.log 2>&1; grep -n "DP OK\|Assert\|Error" gpurun_out/test_dp.log | head -5
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
