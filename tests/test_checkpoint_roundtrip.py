"""CPU: checkpoint adjacency.  The reference's own, unmodified `save_ckpt` / `load_ckpt` / `resume_training`
(utils/train_summary.py:10-38,130-160) write and restore the drop-in modules and the fused optimizer: the `.tar` container
layout, strict `load_state_dict`, the retry that strips DistributedDataParallel's 7-character `module.` prefix (:16-21), the
optimizer state (`optimizer_T.state_dict()`), and interchange with a checkpoint produced by the reference's own modules."""
import os

import pytest
import torch

import ref_loader as RL

pytestmark = pytest.mark.skipif(RL.ref_root() is None, reason="reference tree not available")


def _former(model):
    return model.VPTRFormerNAR(2, 2, encH=8, encW=8, d_model=48, nhead=4, num_encoder_layers=1, num_decoder_layers=1, dropout=0.0,
                               window_size=4, Spatial_FFN_hidden_ratio=4, TSLMA_flag=False, rpe=True)


def test_reference_checkpoint_functions_round_trip_dropin_modules(tmp_path):
    try:
        ref_model = RL.load_reference("cpu")
        import utils as ref_utils                                            # the reference's utils package
        assert os.path.abspath(ref_utils.__file__).startswith(os.path.abspath(RL.ref_root()))
        import vptr_b200.model as ours
        torch.manual_seed(3)
        enc, dec, T = ours.VPTREnc(1, feat_dim=48), ours.VPTRDec(1, feat_dim=48, out_layer="Sigmoid"), _former(ours)
        opt = torch.optim.AdamW(T.parameters(), lr=1e-4)
        for p in T.parameters():                                             # give the optimizer a state to save
            p.grad = torch.randn_like(p) * 1e-3
        opt.step()
        # torch >= 2.6 defaults torch.load to weights_only=True; the reference's container holds plain python objects
        loss_dict = {"T_total": [[1.0], [2.0]], "epochs": 1}
        # --- (1) plain keys
        ref_utils.save_ckpt({"VPTR_Enc": enc, "VPTR_Dec": dec, "VPTR_Transformer": T}, {"optimizer_T": opt}, 7, loss_dict, str(tmp_path))
        f = tmp_path / "epoch_7.tar"
        assert f.exists()
        ck = torch.load(str(f), map_location="cpu", weights_only=False)
        assert set(ck) == {"epoch", "loss_dict", "Module_state_dict", "optimizer_state_dict", "code"}
        assert set(ck["Module_state_dict"]) == {"VPTR_Enc", "VPTR_Dec", "VPTR_Transformer"}
        enc2, dec2, T2 = ours.VPTREnc(1, feat_dim=48), ours.VPTRDec(1, feat_dim=48, out_layer="Sigmoid"), _former(ours)
        opt2 = torch.optim.AdamW(T2.parameters(), lr=1e-4)
        orig_load = torch.load
        torch.load = lambda *a, **k: orig_load(*a, **{**k, "weights_only": False})
        try:
            epoch, hist = ref_utils.resume_training({"VPTR_Enc": enc2, "VPTR_Dec": dec2, "VPTR_Transformer": T2}, {"optimizer_T": opt2},
                                                    str(f), map_location="cpu")
            assert epoch == 7 and hist["epochs"] == 1
            for a, b in ((enc, enc2), (dec, dec2), (T, T2)):
                for (k, v), (k2, v2) in zip(a.state_dict().items(), b.state_dict().items()):
                    assert k == k2 and torch.equal(v, v2), k
            s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
            assert set(s1) == set(s2) and all(torch.equal(s1[k]["exp_avg"], s2[k]["exp_avg"]) for k in s1)
            # --- (2) a checkpoint saved from DistributedDataParallel-wrapped modules: every key carries `module.`
            class Wrapped(torch.nn.Module):                                    # what DDP's state_dict looks like
                def __init__(self, m):
                    super().__init__()
                    self.module = m
            ref_utils.save_ckpt({"VPTR_Transformer": Wrapped(T)}, {}, 8, loss_dict, str(tmp_path))
            T3 = _former(ours)
            ref_utils.resume_training({"VPTR_Transformer": T3}, {}, str(tmp_path / "epoch_8.tar"), map_location="cpu")
            assert all(torch.equal(v, T3.state_dict()[k]) for k, v in T.state_dict().items())
            # --- (3) a checkpoint written from the REFERENCE's modules loads strictly into ours (and back)
            Tr = _former(ref_model)
            ref_utils.save_ckpt({"VPTR_Transformer": Tr}, {}, 9, loss_dict, str(tmp_path))
            T4 = _former(ours)
            ref_utils.resume_training({"VPTR_Transformer": T4}, {}, str(tmp_path / "epoch_9.tar"), map_location="cpu")
            assert all(torch.equal(v, T4.state_dict()[k]) for k, v in Tr.state_dict().items())
            Tr.load_state_dict(T.state_dict(), strict=True)
        finally:
            torch.load = orig_load
    finally:
        RL.unload()
