"""GEMM micro-benchmark on the cfg1 shapes (GPU box): CUDA-event time per launch, TFLOP/s, effective GB/s."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vptr_b200 import ops

R = 40960
cases = [
    ("fwd 528->528 +bias+res", dict(M=R, N=528, K=528, res=True)),
    ("fwd 528->1056 qk", dict(M=R, N=1056, K=528)),
    ("fwd 528->2112", dict(M=R, N=2112, K=528)),
    ("fwd 2112->528 +res", dict(M=R, N=528, K=2112, res=True)),
    ("conv3x3 K=4752", dict(M=R, N=528, K=4752)),
    ("dgrad 528<-528", dict(M=R, N=528, K=528, b_mn=True)),
    ("dgrad 528<-1056 +res", dict(M=R, N=528, K=1056, b_mn=True, res=True)),
    ("dgrad 528<-2112", dict(M=R, N=528, K=2112, b_mn=True)),
    ("dgrad 2112<-528", dict(M=R, N=2112, K=528, b_mn=True)),
    ("wgrad 528x528", dict(M=528, N=528, K=R, a_mn=True, b_mn=True, acc=True)),
    ("wgrad 2112x528", dict(M=2112, N=528, K=R, a_mn=True, b_mn=True, acc=True)),
    ("wgrad 528x2112", dict(M=528, N=2112, K=R, a_mn=True, b_mn=True, acc=True)),
]
flush = torch.empty(64 * 1024 * 1024, device="cuda")
ops.DEBUG_FLAGS = int(os.environ.get("GEMM_DEBUG", "0"))
if os.environ.get("GEMM_CASES"):
    cases = [c for i, c in enumerate(cases) if str(i) in os.environ["GEMM_CASES"].split(",")]
for name, c in cases:
    M, N, K = c["M"], c["N"], c["K"]
    A = torch.randn((K, M) if c.get("a_mn") else (M, K), device="cuda")
    B = torch.randn((K, N) if c.get("b_mn") else (N, K), device="cuda")
    D = torch.zeros(M, N, device="cuda")
    res = torch.randn(M, N, device="cuda") if c.get("res") else None
    bias = torch.randn(N, device="cuda") if not c.get("acc") else None
    kw = dict(a_mn=bool(c.get("a_mn")), b_mn=bool(c.get("b_mn")), out=D, bias=bias, residual=res, accumulate=bool(c.get("acc")))
    for _ in range(3):
        ops.gemm(A, B, **kw)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(A, B, **kw); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2] * 1e-3
    byts = 4 * (A.numel() + B.numel() + D.numel() * (2 if c.get("acc") else 1) + (res.numel() if res is not None else 0))
    print("%-26s %7.1f us  %6.1f TFLOP/s  %6.0f GB/s (algorithmic)" % (name, t * 1e6, 2.0 * M * N * K / t / 1e12, byts / t / 1e9))
