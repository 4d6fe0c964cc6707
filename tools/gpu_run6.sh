#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage1.py -q -m gpu -s 2>&1 | grep -v Warning | grep -E "^E  |^>|passed|failed|Error|error|^FAILED|worst" | head -40
