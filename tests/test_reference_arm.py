"""CPU: the reference arm of bench.py / the drop-in harness.  Checks that (1) the tree staged by oracle/make_ref.sh is a
byte-for-byte copy of the reference, (2) the reference's unmodified train_NAR.single_iter / train_FAR.single_iter run on the host
through oracle/ref_loader.py, and (3) the oracle port of one training iteration (oracle/train_step.py) reproduces the loss of the
reference's own single_iter on the same weights and clips at dropout 0 -- which pins the port to the reference end to end."""
import hashlib
import math
import os

import pytest
import torch

import ref_loader as RL

pytestmark = pytest.mark.skipif(RL.ref_root() is None, reason="reference tree not available (neither /root/reference nor oracle/_ref)")


def test_staged_tree_is_an_unmodified_copy():
    if not os.path.isdir(RL.STAGED):
        pytest.skip("oracle/_ref not staged")
    sums = {}
    with open(os.path.join(RL.STAGED, "SHA256SUMS")) as f:
        for line in f:
            h, name = line.split()
            sums[name] = h
    assert len(sums) >= 20
    for name, h in sums.items():
        with open(os.path.join(RL.STAGED, name), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == h, name
        if os.path.isdir("/root/reference"):
            with open(os.path.join("/root/reference", name), "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == h, name


def _tiny(model, kind):
    enc = model.VPTREnc(1, feat_dim=48, n_downsampling=3).eval()
    dec = model.VPTRDec(1, feat_dim=48, n_downsampling=3, out_layer="Sigmoid").eval()
    if kind == "nar":
        T = model.VPTRFormerNAR(2, 2, encH=8, encW=8, d_model=48, nhead=4, num_encoder_layers=1, num_decoder_layers=1, dropout=0.0,
                                window_size=4, Spatial_FFN_hidden_ratio=4, TSLMA_flag=False, rpe=True)
    else:
        T = model.VPTRFormerFAR(2, 2, encH=8, encW=8, d_model=48, nhead=4, num_encoder_layers=2, dropout=0.0, window_size=4,
                                Spatial_FFN_hidden_ratio=4, rpe=True)
    return enc, dec, T


def test_reference_single_iter_on_cpu_and_port_agree():
    import train_step as TS
    dev = torch.device("cpu")
    try:
        tn, model = RL.load_train_script("train_NAR", dropin=False, device="cpu")
        assert os.path.abspath(model.__file__).startswith(os.path.abspath(RL.ref_root()))
        torch.manual_seed(5)
        enc, dec, T = _tiny(model, "nar")
        for k, v in dict(mse_loss=model.MSELoss(), gdl_loss=model.GDL(alpha=1), bpnce=model.BiPatchNCE(2, 2, 8, 8, 1.0), lam_pc=0.1,
                         lam_gan=None, max_grad_norm=1.0).items():
            setattr(tn, k, v)
        g = torch.Generator().manual_seed(1)
        past, fut = torch.rand(2, 2, 1, 64, 64, generator=g), torch.rand(2, 2, 1, 64, 64, generator=g)
        sd_e = {k: v.detach().clone() for k, v in enc.state_dict().items()}
        sd_d = {k: v.detach().clone() for k, v in dec.state_dict().items()}
        sd_T = {k: v.detach().clone() for k, v in T.state_dict().items()}
        params = {k: v.detach().clone().requires_grad_(True) for k, v in T.named_parameters()}
        opt_ref = torch.optim.AdamW(T.parameters(), lr=1e-4)
        opt_port = torch.optim.AdamW(list(params.values()), lr=1e-4)
        for it in range(2):
            d = tn.single_iter(enc, dec, None, T, opt_ref, None, (past, fut), dev, train_flag=True)
            lp = TS.nar_step(sd_e, sd_d, sd_T, params, opt_port, past, fut, nhead=4, ws=4)
            assert math.isfinite(d["T_total"])
            assert abs(lp - d["T_total"]) <= 2e-5 * abs(d["T_total"]), (it, lp, d["T_total"])
        worst = max(float((params[k].detach() - v.detach()).abs().max()) for k, v in T.named_parameters())
        assert worst < 1e-5, worst                      # after two optimizer steps the parameters still coincide
        tf, model = RL.load_train_script("train_FAR", dropin=False, device="cpu")
        enc, dec, T = _tiny(model, "far")
        for k, v in dict(mse_loss=model.MSELoss(), gdl_loss=model.GDL(alpha=1), max_grad_norm=1.0).items():
            setattr(tf, k, v)
        d = tf.single_iter(enc, dec, None, T, torch.optim.AdamW(T.parameters(), lr=1e-4), None, (past, fut), dev, None, train_flag=True)
        assert math.isfinite(d["T_total"])
    finally:
        RL.unload()
