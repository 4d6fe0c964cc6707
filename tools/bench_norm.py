"""Micro-benchmark (GPU box): frame LayerNorm + GELU backward at cfg1's sizes (640 frames x 64 tokens x 2112 / 528 channels),
reported against the 3-pass HBM minimum (read dy, read x, write dx); the shipped two-kernel path moves 6 passes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vptr_b200 import ops

flush = torch.empty(64 * 1024 * 1024, device="cuda")
for ch in (2112, 528):
    Fr, hw = 640, 64
    rows = Fr * hw
    x, dy = torch.randn(rows, ch, device="cuda"), torch.randn(rows, ch, device="cuda")
    gm, bt = torch.rand(hw, ch, device="cuda") + 0.5, torch.randn(hw, ch, device="cuda") * 0.1
    mean, rstd = ops.group_stats(x, Fr)
    dg, db = torch.zeros_like(gm), torch.zeros_like(bt)
    rs = ops.droppath_scales(64, 3, 0.1, "cuda")
    for name, kw in (("plain", {}), ("dropout+droppath", dict(rowscale=rs, rows_per_group=10 * hw, drop_seed=5, drop_p=0.1))):
        ts = []
        for it in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.norm_act_bwd(dy, x, mean, rstd, gm, bt, dg, db, hw, 1, round_tf32=True, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts[2:])[len(ts[2:]) // 2] * 1e-3
        print("norm_act_bwd ln3 ch=%4d %-17s %7.1f us   %.2f TB/s of the 3-pass minimum" % (ch, name, t * 1e6, 3 * rows * ch * 4 / t / 1e12))

    # forward (read x, write y = 2 passes)
    y = torch.empty_like(x)
    for name, kw in (("plain", {}), ("dropout+droppath+res", dict(rowscale=rs, rows_per_group=10 * hw, drop_seed=5, drop_p=0.1, res=dy))):
        ts = []
        for it in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.norm_act_fwd(x, mean, rstd, gm, bt, hw, 1, out=y, round_tf32=True, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts[2:])[len(ts[2:]) // 2] * 1e-3
        print("norm_act_fwd ln3 ch=%4d %-22s %7.1f us   %.2f TB/s of the 2-pass minimum" % (ch, name, t * 1e6, 2 * rows * ch * 4 / t / 1e12))
