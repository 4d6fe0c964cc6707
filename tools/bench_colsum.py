"""column-sum (bias-gradient) kernel timing at the path's shapes (GPU box)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vptr_b200 import ops
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for rows, width, sl in ((40960, 528, None), (40960, 1584, (0, 528)), (40960, 1584, None), (122880, 528, None), (122880, 2112, None), (300, 48, None)):
    big = torch.randn(rows, width, device="cuda")
    x = big if sl is None else big[:, sl[0]:sl[1]]
    out = torch.zeros(x.shape[1], device="cuda")
    ops.colsum(x, out)
    ref = x.double().sum(0)
    err = float((out.double() - ref).norm() / ref.norm())
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.colsum(x, out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[5] * 1e-3
    print("colsum %6d x %4d (of %4d): %6.1f us  %5.0f GB/s  rel err %.1e" % (rows, x.shape[1], width, t * 1e6, x.numel() * 4 / t / 1e9, err))
