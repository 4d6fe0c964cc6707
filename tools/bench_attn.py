"""Attention-core micro-benchmark on the cfg1 shapes (GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vptr_b200 import ops
N, T, H, W, C, nh = 64, 10, 8, 8, 528, 8
d = C // nh
R = N * T * H * W
qkv = torch.randn(R, 3 * C, device="cuda"); o = torch.empty(R, C, device="cuda"); do = torch.randn(R, C, device="cuda")
dqkv = torch.empty_like(qkv); table = torch.randn(49, nh, device="cuda"); dtab = torch.zeros(49, nh, device="cuda")
flush = torch.empty(64 * 1024 * 1024, device="cuda")
def timeit(fn, n=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3
q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
sc = d ** -0.5
for p in (0.0, 0.1):
    print("dropout", p)
    print("  window fwd  %.0f us" % timeit(lambda: ops.attn_fwd(q, k, v, o, table, 0, N * T, H, W, 4, 0, 0, nh, d, False, sc, True, 7, p)))
    print("  window bwd  %.0f us" % timeit(lambda: ops.attn_bwd(q, k, v, do, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], table, dtab, 0, N * T, H, W, 4, 0, 0, nh, d, False, sc, True, 7, p)))
    print("  temporal fwd %.0f us" % timeit(lambda: ops.attn_fwd(q, k, v, o, None, 1, N, H, W, 0, T, T, nh, d, False, sc, True, 7, p)))
    print("  temporal bwd %.0f us" % timeit(lambda: ops.attn_bwd(q, k, v, do, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], None, None, 1, N, H, W, 0, T, T, nh, d, False, sc, True, 7, p)))
print("ideal HBM: fwd %.0f us, bwd %.0f us" % (4 * R * C * 4 / 6.55e12 * 1e6, 7 * R * C * 4 / 6.55e12 * 1e6))
