"""Debug aid (GPU box): per-parameter gradient error of the CUDA path vs the CPU oracle for one golden case."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_former, oracle_former, probe, rel_l2  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else "far_rpe"
simt = len(sys.argv) > 2 and sys.argv[2] == "simt"
from vptr_b200 import engine, ops  # noqa: E402
if simt:
    ops.FORCE_SIMT = True
    engine.ROUND_TF32 = engine.RT = False
net, x, c = build_former(name, "cuda")
net.train()
xin = x.clone().requires_grad_(True)
y = net(xin)
(y * probe(y.shape, 2).cuda()).sum().backward()
sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
params = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.named_parameters()}
sd.update(params)
xo = x.cpu().clone().requires_grad_(True)
yo, _ = oracle_former(name, sd, xo, training=True)
(yo * probe(yo.shape, 2)).sum().backward()
print("case", name, "simt" if simt else "tf32", "y", rel_l2(y, yo), "dx", rel_l2(xin.grad, xo.grad))
gmax = max(float(p.grad.abs().sum()) for p in params.values() if p.grad is not None)
for k, p in net.named_parameters():
    go = params[k].grad
    if go is None:
        continue
    if p.grad is None:
        print("%-70s MISSING" % k)
        continue
    e = rel_l2(p.grad, go)
    flag = "" if (e < 3e-3 or float(go.abs().sum()) < 1e-5 * gmax) else "   <<<<<<"
    print("%-70s %.2e  |g|1=%.2e%s" % (k, e, float(go.abs().sum()), flag))
