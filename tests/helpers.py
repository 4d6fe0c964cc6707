"""Shared test helpers: rebuild the closed-form golden cases (tests/golden/fixtures.py) without the reference."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle"), GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)

from fixtures import CASES, closed_form, fill_state_dict_, grad_signature, probe  # noqa: E402


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def build_former(name, device="cpu"):
    """Our module for a golden transformer case, filled with the closed-form weights make_golden.py used."""
    from vptr_b200.model import VPTRFormerFAR, VPTRFormerNAR
    c = CASES[name]
    if c["kind"] == "nar":
        net = VPTRFormerNAR(c["Tp"], c["Tf"], encH=c["encH"], encW=c["encW"], d_model=c["d_model"], nhead=c["nhead"],
                            num_encoder_layers=c["enc_layers"], num_decoder_layers=c["dec_layers"], dropout=0.0, window_size=c["ws"],
                            TSLMA_flag=c.get("tslma", False), rpe=c["rpe"])
        T_in = c["Tp"]
    else:
        net = VPTRFormerFAR(c["Tp"], c["Tf"], encH=c["encH"], encW=c["encW"], d_model=c["d_model"], nhead=c["nhead"],
                            num_encoder_layers=c["enc_layers"], dropout=0.0, window_size=c["ws"], rpe=c["rpe"])
        T_in = c["T_in"]
    with torch.no_grad():
        fill_state_dict_(net.state_dict(), 0)
    x = closed_form((c["N"], T_in, c["d_model"], c["encH"], c["encW"]), 3, 1.0).abs()
    return net.to(device), x.to(device), c


def build_ae(name, device="cpu"):
    from vptr_b200.model import VPTRDec, VPTREnc
    c = CASES[name]
    enc = VPTREnc(c["img_channels"], feat_dim=c["feat_dim"], n_downsampling=c["n_down"], padding_type=c["padding_type"]).eval()
    dec = VPTRDec(c["img_channels"], feat_dim=c["feat_dim"], n_downsampling=c["n_down"], out_layer=c["out_layer"],
                  padding_type=c["padding_type"]).eval()
    with torch.no_grad():
        fill_state_dict_(enc.state_dict(), 0)
        fill_state_dict_(dec.state_dict(), 500)
    x = closed_form((c["N"], c["T"], c["img_channels"], c["HW"], c["HW"]), 7, 0.5, 0.5)
    return enc.to(device), dec.to(device), x.to(device), c


def oracle_former(name, sd, x, training):
    """Runs oracle/vptr_oracle.py on a reference-format state_dict (CPU, fp32)."""
    import vptr_oracle as O
    c = CASES[name]
    sd = {k: (v if v.device.type == 'cpu' else v.detach().cpu()) for k, v in sd.items()}
    if c["kind"] == "nar":
        bu = {}
        y = O.vptr_former_nar(sd, x, nhead=c["nhead"], ws=c["ws"], rpe=c["rpe"], training=training, bn_updates=bu)
        return y, bu
    return O.vptr_former_far(sd, x, nhead=c["nhead"], ws=c["ws"], rpe=c["rpe"], training=training), {}
