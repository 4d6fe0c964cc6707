#!/bin/bash
# round 2: quadrant conv for the 16x16 grid (cfg4): parity, op timing generic vs quad, cfg4 bench both ways
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== conv tests"
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | grep -v Warning | grep -E "^E  |^>|passed|failed|Error|error|^FAILED" | head -30
echo "=== conv timing 1280 frames 16x16 528->528"
timeout 300 python - <<'PY'
import torch
from vptr_b200 import ops
F_, H, W, C = 1280, 16, 16, 528
x = torch.randn(F_*H*W, C, device="cuda"); w = torch.randn(C, 9*C, device="cuda")*0.02
w2 = ops.split_tf32(w); b = torch.randn(C, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
fl = 2.0*F_*H*W*C*9*C*2
def gen():
    xp = ops.pad_nhwc(x, F_, H, W, C, 1, 1, round_tf32=True); return ops.conv3x3_tf32(xp, w2, F_, H, W, C, C, bias=b, act=2, w_planes=2)
def quad():
    xq = ops.pad_nhwc_quad(x, F_, H, W, C, 1, round_tf32=True); return ops.conv3x3_tf32_quad(xq, w2, F_, H, W, C, C, bias=b, act=2, w_planes=2)
a, q = gen(), quad()
print("max diff", (a-q).abs().max().item())
tg, tq = t(gen), t(quad)
print(f"generic {tg:.3f} ms {fl/tg/1e9:.0f} TF/s executed | quad {tq:.3f} ms {fl/tq/1e9:.0f} TF/s executed")
xp = ops.pad_nhwc(x, F_, H, W, C, 1, 1, round_tf32=True); xq = ops.pad_nhwc_quad(x, F_, H, W, C, 1, round_tf32=True)
print("pad only", t(lambda: ops.pad_nhwc(x, F_, H, W, C, 1, 1, round_tf32=True)), t(lambda: ops.pad_nhwc_quad(x, F_, H, W, C, 1, round_tf32=True)))
print("conv only", t(lambda: ops.conv3x3_tf32(xp, w2, F_, H, W, C, C, bias=b, act=2, w_planes=2)), t(lambda: ops.conv3x3_tf32_quad(xq, w2, F_, H, W, C, C, bias=b, act=2, w_planes=2)))
PY
for mode in quad generic; do
  echo "=== bench cfg4 $mode"
  if [ $mode = generic ]; then export VPTR_CONV_GENERIC=1; else unset VPTR_CONV_GENERIC; fi
  timeout 900 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_$mode.json 2> gpurun_out/bench_cfg4_$mode.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg4_$mode.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step'))"; tail -2 gpurun_out/bench_cfg4_$mode.err
done
