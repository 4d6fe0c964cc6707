"""GPU: the drop-in claim as a test (north_star: "train_{FAR,NAR}.py import and run unchanged").

The reference's UNMODIFIED training scripts (byte-for-byte copies staged by oracle/make_ref.sh into oracle/_ref/, or
/root/reference in the build container) are imported with `vptr_b200/` first on sys.path, so their own
`from model import VPTREnc, VPTRDec, VPTRDisc, init_weights, VPTRFormerNAR` (train_NAR.py:13-14) binds to vptr_b200/model.
Their `single_iter` (train_NAR.py:49-107, train_FAR.py:48-101) then drives our modules on the B200 for two iterations with the
`__main__` globals injected: finite losses, parameters updated, AdamW/clip untouched.  A second test
wraps the Transformer in DistributedDataParallel and reaches `.module.NCE_projector` as train_NAR_mp.py:163-164 does."""
import math
import os

import pytest
import torch

import ref_loader as RL

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(RL.ref_root() is None, reason="reference tree not staged (oracle/make_ref.sh)")]


def _models(model, kind, dev, layers=1):
    enc = model.VPTREnc(1, feat_dim=528, n_downsampling=3).to(dev).eval()
    dec = model.VPTRDec(1, feat_dim=528, n_downsampling=3, out_layer="Sigmoid").to(dev).eval()
    model.init_weights(enc)
    model.init_weights(dec)
    if kind == "nar":
        T = model.VPTRFormerNAR(4, 4, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=layers, num_decoder_layers=layers,
                                dropout=0.1, window_size=4, Spatial_FFN_hidden_ratio=4, TSLMA_flag=False, rpe=True).to(dev)
    else:
        T = model.VPTRFormerFAR(4, 4, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=layers, dropout=0.1,
                                window_size=4, Spatial_FFN_hidden_ratio=4, rpe=False).to(dev)     # rpe=False is train_FAR.py:165's default
    return enc, dec, T


def _sample(n=2, T=4):
    g = torch.Generator().manual_seed(2021)
    return torch.rand(n, T, 1, 64, 64, generator=g), torch.rand(n, T, 1, 64, 64, generator=g)


def test_reference_train_nar_single_iter_runs_on_dropin_modules():
    dev = torch.device("cuda:0")
    try:
        tn, model = RL.load_train_script("train_NAR", dropin=True)
        assert model.__name__ == "vptr_b200.model" and "vptr_b200" in model.VPTRFormerNAR.__module__
        assert os.path.abspath(tn.__file__).startswith(os.path.abspath(RL.ref_root()))
        assert tn.VPTRFormerNAR is model.VPTRFormerNAR                 # the script's own import bound to the drop-in class
        torch.manual_seed(2021)
        enc, dec, T = _models(model, "nar", dev)
        opt = torch.optim.AdamW(params=T.parameters(), lr=1e-4)       # train_NAR.py:205
        for k, v in dict(mse_loss=model.MSELoss(), gdl_loss=model.GDL(alpha=1), bpnce=model.BiPatchNCE(2, 4, 8, 8, 1.0).to(dev),
                         lam_pc=0.1, lam_gan=None, max_grad_norm=1.0).items():
            setattr(tn, k, v)                                          # train_NAR.py:160-216 globals
        before = {k: v.detach().clone() for k, v in T.named_parameters()}
        losses = []
        for _ in range(2):
            d = tn.single_iter(enc, dec, None, T, opt, None, _sample(), dev, train_flag=True)
            losses.append(d["T_total"])
            assert all(math.isfinite(v) for v in d.values()), d
        moved = sum(int(not torch.equal(before[k], v.detach())) for k, v in T.named_parameters())
        assert moved == len(before), "every Transformer parameter must have been stepped (%d of %d)" % (moved, len(before))
        d = tn.single_iter(enc, dec, None, T, opt, None, _sample(), dev, train_flag=False)      # the validation branch
        assert math.isfinite(d["T_total"])
    finally:
        RL.unload()


def test_reference_train_far_single_iter_runs_on_dropin_modules():
    dev = torch.device("cuda:0")
    try:
        tf, model = RL.load_train_script("train_FAR", dropin=True)
        assert tf.VPTRFormerFAR is model.VPTRFormerFAR
        torch.manual_seed(2021)
        enc, dec, T = _models(model, "far", dev, layers=2)
        opt = torch.optim.AdamW(params=T.parameters(), lr=1e-4)
        for k, v in dict(mse_loss=model.MSELoss(), gdl_loss=model.GDL(alpha=1), max_grad_norm=1.0).items():
            setattr(tf, k, v)
        before = {k: v.detach().clone() for k, v in T.named_parameters()}
        for _ in range(2):
            d = tf.single_iter(enc, dec, None, T, opt, None, _sample(), dev, None, train_flag=True)
            assert all(math.isfinite(v) for v in d.values()), d
        moved = sum(int(not torch.equal(before[k], v.detach())) for k, v in T.named_parameters())
        assert moved == len(before)
    finally:
        RL.unload()


def test_distributed_data_parallel_wrap_and_module_attribute():
    """train_NAR_mp.py:118,163-164: DDP(VPTR_Transformer) forward/backward + `.module.NCE_projector`"""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from vptr_b200.model import VPTRFormerNAR
    dev = torch.device("cuda:0")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29700 + os.getpid() % 200))
    own = not dist.is_initialized()
    if own:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        torch.manual_seed(1)
        T = VPTRFormerNAR(2, 2, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dropout=0.1,
                          window_size=4, rpe=True).to(dev)
        ddp = DDP(T, device_ids=[0])
        x = torch.rand(2, 2, 528, 8, 8, device=dev)
        for _ in range(2):
            ddp.zero_grad(set_to_none=True)
            y = ddp(x)
            pf = ddp.module.NCE_projector(y.permute(0, 1, 3, 4, 2)).permute(0, 1, 4, 2, 3)
            (y.square().mean() + pf.square().mean()).backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in T.parameters())
    finally:
        if own:
            dist.destroy_process_group()
