"""Debug aid (GPU box): per-tile clock64 timeline of CTA 0 of the tcgen05 GEMM."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vptr_b200 import ops, _lib
M, N, K = 40960, int(sys.argv[1]) if len(sys.argv) > 1 else 2112, int(sys.argv[2]) if len(sys.argv) > 2 else 528
A, B, D = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.empty(M, N, device="cuda")
RES = torch.randn(M, N, device="cuda") if os.environ.get("RES") == "1" else None
for _ in range(2):
    ops.gemm(A, B, out=D, residual=RES)
buf = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
_lib.lib().vptr_gemm_debug_buffer(buf.data_ptr())
ops.gemm(A, B, out=D, residual=RES)
torch.cuda.synchronize()
_lib.lib().vptr_gemm_debug_buffer(None)
t = buf.view(-1, 8).cpu()
t0 = int(t[0, 0])
print("tile: mma_wait_start mma_wait_end mma_issued | epi_wait_start epi_full epi_done   (cycles since start)")
for i in range(12):
    if int(t[i, 0]) == 0 and i > 0:
        break
    print(i, [int(v) - t0 for v in t[i, :6]])
