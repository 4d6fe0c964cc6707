#!/bin/bash
# 8-GPU weak-scaling lines for every BASELINE.json config (one node, one rank per GPU, NCCL over NVLink / NVSwitch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NMAX=${1:-8}
for spec in cfg1:2 cfg1:4 cfg1:$NMAX cfg2:$NMAX cfg3:$NMAX cfg4:$NMAX; do
  cfg=${spec%%:*}; N=${spec##*:}
  [ $N -gt $NMAX ] && continue
  echo "=== $cfg x $N GPUs"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --config $cfg --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${cfg}_${N}gpu.json 2> gpurun_out/bench_${cfg}_${N}gpu.err
  echo "rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${cfg}_${N}gpu.json").read().strip().splitlines()[-1])
    print(d["config"]["config"], "n_gpus", d["n_gpus"], "frames/s", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "mem", d["config"]["peak_mem_gib"], "clocks", d["clocks"]["sm_mhz"])
except Exception as e:
    print("parse failed", e)
PY
  grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/bench_${cfg}_${N}gpu.err | tail -3
done
