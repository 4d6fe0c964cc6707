"""Debug aid (GPU box): finite-difference check of the Transformer gradient with dropout variants."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_former, probe, rel_l2
from vptr_b200 import engine
name = sys.argv[1]
for p_elem, p_path in ((0.0, 0.0), (0.2, 0.0), (0.0, 0.2), (0.2, 0.2)):
    net, x, c = build_former(name, "cuda")
    net.dropout = 0.2
    net.train()
    orig_init = engine.Drop.__init__
    def patched(self, p, n, dev, _pe=p_elem, _pp=p_path):
        orig_init(self, p, n, dev)
        self.p, self.p_path = _pe, _pp
    engine.Drop.__init__ = patched
    pr = probe((x.shape[0], c["Tf"] if c["kind"] == "nar" else x.shape[1], *x.shape[2:]), 2).cuda()
    def f(inp):
        torch.manual_seed(5)
        with torch.no_grad():
            return float((0.5 * net(inp) ** 2 * pr).double().sum())
    with engine.exact_fp32():
        torch.manual_seed(5)
        xin = x.clone().requires_grad_(True)
        y = net(xin)
        (0.5 * y * y * pr).sum().backward()
        g = xin.grad
        v = g / g.norm() * (g.numel() ** 0.5)
        for eps in (4e-4, 2e-4, 1e-4, 5e-5, 2e-5):
            fd = (f(x + eps * v) - f(x - eps * v)) / (2 * eps)
            an = float((xin.grad.double() * v.double()).sum())
            print(name, "p_elem", p_elem, "p_path", p_path, "eps", eps, "fd %.4f an %.4f rel %.3e" % (fd, an, abs(fd - an) / max(abs(an), 1e-9)))
    engine.Drop.__init__ = orig_init
