#!/bin/bash
# round 2 (final code): ncu launch list of one cfg1 step (per-launch time + DRAM bytes) -> summary + GEMM traffic json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_cfg1.csv python bench.py --config cfg1 --ncu-step --warmup 2 > gpurun_out/ncu_step.log 2>&1; echo rc=$?; wc -l gpurun_out/launches_cfg1.csv
python tools/summarize_launches.py gpurun_out/launches_cfg1.csv 60 > gpurun_out/launches_cfg1_summary.txt 2>&1; head -64 gpurun_out/launches_cfg1_summary.txt
gzip -kf gpurun_out/launches_cfg1.csv
