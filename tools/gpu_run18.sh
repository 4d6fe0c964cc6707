#!/bin/bash
# round 2: bf16x3 raw-tile conv: parity, timing vs the two-plane TF32 kernel, model tests, cfg1 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -s -k "bf16x3" 2>&1 | grep -E "bf16x3 conv|^E  |passed|failed|^FAILED" | head -20
PYTHONPATH=. timeout 300 python - <<'PY'
import torch
from vptr_b200 import ops
for F_, H in ((1280, 8), (640, 16)):
    W, C = H, 528
    x = torch.randn(F_*H*W, C, device="cuda"); w = torch.randn(C, 9*C, device="cuda")*0.02; b = torch.randn(C, device="cuda")
    w2t, w2b = ops.split_tf32(w), ops.split_bf16x2(w)
    def t(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
    if H == 8:
        xp = ops.pad_nhwc(x, F_, H, W, C, 1, 1, round_tf32=True)
        tf = lambda: ops.conv3x3_tf32(xp, w2t, F_, H, W, C, C, bias=b, act=2, w_planes=2)
    else:
        xp = ops.pad_nhwc_quad(x, F_, H, W, C, 1, round_tf32=True)
        tf = lambda: ops.conv3x3_tf32_quad(xp, w2t, F_, H, W, C, C, bias=b, act=2, w_planes=2)
    xq2 = ops.pad_nhwc_quad_bf16x2(x, F_, H, W, C, 1)
    bf = lambda: ops.conv3x3_bf16x3(xq2, w2b, F_, H, W, C, C, bias=b, act=2)
    a, c = tf(), bf()
    print(f"F={F_} {H}x{W}: tf32x2 {t(tf):.0f} us, bf16x3 {t(bf):.0f} us, pad tf32 {t(lambda: ops.pad_nhwc_quad(x, F_, H, W, C, 1, round_tf32=True)):.0f} us, pad bf16x2 {t(lambda: ops.pad_nhwc_quad_bf16x2(x, F_, H, W, C, 1)):.0f} us, rel diff {float((a-c).norm()/a.norm()):.2e}")
PY
if [ "$1" = full ]; then
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_models.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py tests/test_gpu_switches.py -q -m gpu 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head
timeout 600 python bench.py --config cfg1 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg1_bf16.json 2> gpurun_out/bench_cfg1_bf16.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg1_bf16.json').read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['achieved'], r['peak'], r['frac']); print(d['kernel_rooflines'].get('encoder_conv3x3'))"; tail -1 gpurun_out/bench_cfg1_bf16.err | cut -c1-200
fi
