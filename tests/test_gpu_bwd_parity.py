"""GPU: forward + BACKWARD parity of the shipped TF32 tensor-core path at the REAL model width (d_model 528, 8 heads of 66,
FFN 2112), train mode (dropout 0, BatchNorm batch statistics), against autograd through the CPU oracle (oracle/vptr_oracle.py,
pinned to the unmodified reference by tests/golden).  Reference path: train_NAR.py:63-85 / train_FAR.py:62-82.

Cotangent: L = sum(0.5 * y^2 * probe), i.e. dL/dy = y * probe, which vanishes where the final ReLU has its kink -- the gradient
is then a continuous function of the forward values, so a ~5e-4 forward perturbation cannot flip a mask into an O(1) gradient
difference (with a linear cotangent ~4e-4 of the ReLU outputs flip and move every gradient by ~3 % regardless of kernel accuracy).

For every case the per-tensor table (rel-L2 of dx and of each parameter gradient, product path vs oracle) is written to
gpurun_out/bwd_parity_<case>.txt and summarised in profiles/; the gates below are the ones BASELINE.md 5 states (1e-3 on
outputs) and, for gradients, the measured TF32 envelope stated next to each assert."""
import os

import pytest
import torch

from helpers import ROOT, probe, rel_l2

import vptr_oracle as O

pytestmark = pytest.mark.gpu

OUT_DIR = os.path.join(ROOT, "gpurun_out")


def _run_case(case, kind, Tp, Tf, enc_layers, dec_layers, encH, ws, n_clips, T_in=None, seed=5):
    from vptr_b200.model import VPTRFormerFAR, VPTRFormerNAR
    torch.manual_seed(2021)
    if kind == "nar":
        net = VPTRFormerNAR(Tp, Tf, encH=encH, encW=encH, d_model=528, nhead=8, num_encoder_layers=enc_layers,
                            num_decoder_layers=dec_layers, dropout=0.0, window_size=ws, rpe=True)
        T_in = Tp
    else:
        net = VPTRFormerFAR(Tp, Tf, encH=encH, encW=encH, d_model=528, nhead=8, num_encoder_layers=enc_layers, dropout=0.0,
                            window_size=ws, rpe=True)
    x = torch.rand(n_clips, T_in, 528, encH, encH, generator=torch.Generator().manual_seed(seed))
    # --- oracle (CPU fp32 autograd)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    params = {k: v.detach().clone().requires_grad_(True) for k, v in net.named_parameters() if not k.startswith("NCE_projector")}
    sd.update(params)
    xo = x.clone().requires_grad_(True)
    if kind == "nar":
        yo = O.vptr_former_nar(sd, xo, nhead=8, ws=ws, rpe=True, training=True, bn_updates={})
    else:
        yo = O.vptr_former_far(sd, xo, nhead=8, ws=ws, rpe=True, training=True)
    pr = probe(yo.shape, 2)
    (0.5 * yo * yo * pr).sum().backward()
    # --- product path (TF32 tcgen05 GEMMs, tensor-core attention): the default engine mode, nothing switched
    net = net.cuda().train()
    xin = x.cuda().requires_grad_(True)
    y = net(xin)
    (0.5 * y * y * pr.cuda()).sum().backward()
    rows = [("<output y>", rel_l2(y, yo), float(yo.norm())), ("<dx>", rel_l2(xin.grad, xo.grad), float(xo.grad.norm()))]
    gmax = max(float(p.grad.norm()) for p in params.values() if p.grad is not None)
    for k, p in net.named_parameters():
        if k.startswith("NCE_projector"):
            continue
        go = params[k].grad
        if go is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        rows.append((k, rel_l2(p.grad, go), float(go.norm())))
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, "bwd_parity_%s.txt" % case), "w") as f:
        f.write("# %s: rel-L2(product TF32 path, CPU oracle autograd); train mode, dropout 0; %d clip(s); cotangent y*probe\n" % (case, n_clips))
        f.write("# largest parameter-gradient norm %.4e; tensors whose gradient norm is < 1e-4 of it are rounding noise of a mathematical zero\n" % gmax)
        for k, e, n in sorted(rows, key=lambda r: -r[1]):
            f.write("%-72s rel_l2 %.3e   |ref| %.3e%s\n" % (k, e, n, "   (noise-level gradient)" if n < 1e-4 * gmax else ""))
    sig = [(k, e) for k, e, n in rows[2:] if n >= 1e-4 * gmax]
    return rows[0][1], rows[1][1], sig


def _summ(sig):
    es = sorted(e for _, e in sig)
    return es[len(es) // 2], es[-1], max(sig, key=lambda t: t[1])[0]


# Gates.  Outputs: 1e-3.  Gradients: every GEMM operand is rounded to tf32 (2^-11 relative) and the backward chains
# 3 GEMMs per sub-block through up to 12 + 8 blocks; the measured envelope of the product path is recorded in
# profiles/r02_bwd_parity.md and the gates sit ~1.5x above it.
def test_nar_cfg1_full_depth_backward():
    ey, edx, sig = _run_case("cfg1_nar_4enc_8dec", "nar", 10, 10, 4, 8, 8, 4, n_clips=2)
    med, worst, who = _summ(sig)
    print("cfg1: y %.2e dx %.2e grads median %.2e worst %.2e (%s)" % (ey, edx, med, worst, who))
    assert ey < 1e-3
    assert edx < 5e-3 and med < 3e-3 and worst < 2e-2, (edx, med, worst, who)


def test_far_cfg2_full_depth_backward():
    ey, edx, sig = _run_case("cfg2_far_12enc_T29", "far", 10, 20, 12, 0, 8, 4, n_clips=1, T_in=29)
    med, worst, who = _summ(sig)
    print("cfg2: y %.2e dx %.2e grads median %.2e worst %.2e (%s)" % (ey, edx, med, worst, who))
    assert ey < 1e-3
    assert edx < 5e-3 and med < 3e-3 and worst < 2e-2, (edx, med, worst, who)


def test_nar_cfg3_slice_backward():
    ey, edx, sig = _run_case("cfg3_nar_2to28_1enc_2dec", "nar", 2, 28, 1, 2, 8, 4, n_clips=2)
    med, worst, who = _summ(sig)
    print("cfg3: y %.2e dx %.2e grads median %.2e worst %.2e (%s)" % (ey, edx, med, worst, who))
    assert ey < 1e-3
    assert edx < 5e-3 and med < 3e-3 and worst < 2e-2, (edx, med, worst, who)


def test_nar_cfg4_slice_backward():
    ey, edx, sig = _run_case("cfg4_nar_10to30_grid16_ws8_1enc_1dec", "nar", 10, 30, 1, 1, 16, 8, n_clips=1)
    med, worst, who = _summ(sig)
    print("cfg4: y %.2e dx %.2e grads median %.2e worst %.2e (%s)" % (ey, edx, med, worst, who))
    assert ey < 1e-3
    assert edx < 5e-3 and med < 3e-3 and worst < 2e-2, (edx, med, worst, who)


def test_lean_memory_mode_matches_fast():
    """lean mode re-derives GEMM operands in the backward with the same kernels and dropout seeds: identical outputs, and gradients
    equal up to the summation order of the split-K weight-gradient atomics"""
    from vptr_b200 import engine
    from vptr_b200.model import VPTRFormerNAR
    torch.manual_seed(3)
    net = VPTRFormerNAR(2, 3, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dropout=0.1,
                        window_size=4, rpe=True).cuda().train()
    x = torch.rand(2, 2, 528, 8, 8, generator=torch.Generator().manual_seed(1)).cuda()
    res = []
    for mode in ("fast", "lean"):
        engine.set_memory_mode(mode)
        try:
            torch.manual_seed(77)                     # same dropout seeds in both runs (engine.Drop draws its base from torch's RNG)
            net.zero_grad(set_to_none=True)
            xin = x.clone().requires_grad_(True)
            y = net(xin)
            (y * y).sum().backward()
            res.append((y.detach().clone(), xin.grad.clone(), [p.grad.clone() for p in net.parameters() if p.grad is not None]))
        finally:
            engine.set_memory_mode("auto")
    assert torch.equal(res[0][0], res[1][0]) and rel_l2(res[1][1], res[0][1]) < 1e-5
    assert len(res[0][2]) == len(res[1][2])
    for a, b in zip(res[0][2], res[1][2]):
        assert rel_l2(b, a) < 1e-5
