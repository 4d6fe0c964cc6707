"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on the CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The reference has no tests or golden vectors of its own, so these fixtures --
outputs of the reference itself on closed-form weights/inputs (fixtures.py) -- are what pins
oracle/vptr_oracle.py, and through it the CUDA path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

from fixtures import CASES, closed_form, fill_state_dict_, grad_signature, probe  # noqa: E402
from ref_loader import load_reference  # noqa: E402


def ae_case(model, name, c):
    enc = model.VPTREnc(c["img_channels"], feat_dim=c["feat_dim"], n_downsampling=c["n_down"], padding_type=c["padding_type"]).eval()
    dec = model.VPTRDec(c["img_channels"], feat_dim=c["feat_dim"], n_downsampling=c["n_down"], out_layer=c["out_layer"],
                        padding_type=c["padding_type"]).eval()
    with torch.no_grad():
        fill_state_dict_(enc.state_dict(), 0)
        fill_state_dict_(dec.state_dict(), 500)
    x = closed_form((c["N"], c["T"], c["img_channels"], c["HW"], c["HW"]), 7, 0.5, 0.5)
    with torch.no_grad():
        feat = enc(x)
    feat_in = feat.clone().requires_grad_(True)
    rec = dec(feat_in)
    (rec * probe(rec.shape, 1)).sum().backward()
    out = {"feat": feat.numpy(), "rec": rec.detach().numpy(), "dfeat": feat_in.grad.numpy(),
           "enc_keys": np.array(list(enc.state_dict().keys())), "dec_keys": np.array(list(dec.state_dict().keys()))}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, feat.shape, rec.shape, float(feat.abs().mean()), float(rec.mean()))


def former_case(model, name, c):
    if c["kind"] == "nar":
        net = model.VPTRFormerNAR(c["Tp"], c["Tf"], encH=c["encH"], encW=c["encW"], d_model=c["d_model"], nhead=c["nhead"],
                                  num_encoder_layers=c["enc_layers"], num_decoder_layers=c["dec_layers"], dropout=0.0,
                                  window_size=c["ws"], TSLMA_flag=c.get("tslma", False), rpe=c["rpe"])
        T_in = c["Tp"]
    else:
        net = model.VPTRFormerFAR(c["Tp"], c["Tf"], encH=c["encH"], encW=c["encW"], d_model=c["d_model"], nhead=c["nhead"],
                                  num_encoder_layers=c["enc_layers"], dropout=0.0, window_size=c["ws"], rpe=c["rpe"])
        T_in = c["T_in"]
    with torch.no_grad():
        fill_state_dict_(net.state_dict(), 0)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    x = closed_form((c["N"], T_in, c["d_model"], c["encH"], c["encW"]), 3, 1.0).abs()   # encoder features are post-ReLU
    out = {"keys": np.array(list(sd0.keys()))}
    for b in ("temporal_pos", "lw_pos", "Tlw_pos"):
        if b in sd0:
            out["buf_" + b] = sd0[b].numpy()
    net.eval()
    with torch.no_grad():
        out["y_eval"] = net(x).contiguous().numpy()
    net.train()
    xin = x.clone().requires_grad_(True)
    y = net(xin)
    (y * probe(y.shape, 2)).sum().backward()
    out["y_train"] = y.detach().contiguous().numpy()
    out["dx"] = xin.grad.numpy()
    names, sigs = [], []
    for i, (k, p) in enumerate(net.named_parameters()):
        if p.grad is None:
            continue
        names.append(k)
        sigs.append(grad_signature(p.grad, i))
    out["grad_names"] = np.array(names)
    out["grad_sigs"] = np.stack(sigs)
    # BatchNorm side effects of the train-mode forward
    sd1 = net.state_dict()
    bn = [k for k in sd1 if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked")]
    out["bn_keys"] = np.array(bn)
    for i, k in enumerate(bn):
        out["bn_%d" % i] = sd1[k].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, y.shape, float(y.abs().mean()), len(names), "param grads")


def integer_artefacts(model):
    import importlib
    vm = importlib.import_module("model.VidHRFormer_modules")
    rpe_mod = importlib.import_module("model.MultiHeadAttentionRPE")
    out = {}
    for ws in (2, 4, 7, 8):
        m = rpe_mod.MultiheadAttentionRPE(embed_dim=8, num_heads=2, rpe=True, window_size=ws)
        out["rpi_%d" % ws] = m.relative_position_index.numpy()
    for (Fr, H, W, ws) in ((2, 8, 8, 4), (1, 16, 16, 8), (3, 8, 12, 4)):
        ids = torch.arange(Fr * H * W, dtype=torch.float32).reshape(Fr, H, W, 1)
        perm = vm.LocalPermuteModule(ws).permute(ids, ids.size())          # (L, B, 1)
        out["wmap_%d_%d_%d_%d" % (Fr, H, W, ws)] = perm[..., 0].long().numpy()
    for T in (1, 5, 29):
        out["causal_%d" % T] = (torch.triu(torch.ones(T, T), diagonal=1) == 1).numpy()
    pb = vm.PadBlock(4)
    for hw in (6, 7, 8, 9):
        x = torch.ones(1, hw, hw, 1)
        xp = pb.pad_if_needed(x, x.size())
        out["padmask_%d" % hw] = xp[0, :, :, 0].numpy()
    np.savez_compressed(os.path.join(HERE, "integer_artefacts.npz"), **out)
    print("integer artefacts", len(out))


def pos_528(model):
    net = model.VPTRFormerFAR(10, 10, d_model=528, nhead=8, num_encoder_layers=1, dropout=0.0, window_size=4, rpe=True)
    sd = net.state_dict()
    np.savez_compressed(os.path.join(HERE, "pos_528.npz"), temporal_pos=sd["temporal_pos"].numpy(), lw_pos=sd["lw_pos"].numpy())


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    model = load_reference("cpu")
    only = sys.argv[1:]                 # optional: regenerate just the named fixtures
    for name, c in CASES.items():
        if only and name not in only:
            continue
        if name.startswith("ae_"):
            ae_case(model, name, c)
        else:
            former_case(model, name, c)
    if not only:
        integer_artefacts(model)
        pos_528(model)
