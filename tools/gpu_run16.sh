#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head
PYTHONPATH=. timeout 300 python - <<'PY'
import torch
from vptr_b200 import ops
for Fr, H, W, Co in ((640, 128, 128, 3), (960, 64, 64, 3), (1280, 64, 64, 1)):
    h = torch.randn(Fr*H*W, 64, device="cuda"); wh = torch.randn(Co, 64, 7, 7, device="cuda")*0.05; bh = torch.zeros(Co, device="cuda")
    wp = ops.pack_conv_weight(wh, None, 3)
    for _ in range(2): ops.head_conv7x7_fwd(h, wp, bh, Fr, 64, Co, H, W, 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.head_conv7x7_fwd(h, wp, bh, Fr, 64, Co, H, W, 2)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)/5
    print(f"head fwd F={Fr} {H}x{W} Co={Co}: {t:.3f} ms, {h.numel()*4/t/1e6:.0f} GB/s")
PY
timeout 900 python bench.py --config cfg3 --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/bench_cfg4_head.json 2> gpurun_out/bench_cfg4_head.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg4_head.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'))"; tail -2 gpurun_out/bench_cfg4_head.err
