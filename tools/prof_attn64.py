"""One launch each of the wide (64-token) window attention forward / backward at the cfg4 decoder shape (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vptr_b200 import ops
N, T, H, W, C, nh, ws = 4, 30, 16, 16, 528, 8, 8
d = C // nh
R = N * T * H * W
qkv = torch.randn(R, 3 * C, device="cuda"); o = torch.empty(R, C, device="cuda"); do = torch.randn(R, C, device="cuda")
dqkv = torch.empty_like(qkv); table = torch.randn((2 * ws - 1) ** 2, nh, device="cuda"); dtab = torch.zeros_like(table)
q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
sc = d ** -0.5
for _ in range(2):
    ops.attn_fwd(q, k, v, o, table, 0, N * T, H, W, ws, 0, 0, nh, d, False, sc, True, 7, 0.1)
    ops.attn_bwd(q, k, v, do, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], table, dtab, 0, N * T, H, W, ws, 0, 0, nh, d, False, sc, True, 7, 0.1)
torch.cuda.synchronize()
