#!/bin/bash
# Runs each GPU test file in its own process under a timeout so a hung kernel cannot take the whole call down.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for f in "$@"; do
  echo "=== $f"
  timeout 300 python -m pytest "$f" -x -q -m gpu 2>&1 | tail -40
done
