"""GPU: parity at the REAL model dimensions (d_model 528, 8 heads of 66, 4x4 windows on an 8x8 grid, LayerNorm((2112,8,8)),
T = 10 / 29) against the CPU oracle on seeded inputs -- one clip, so the oracle finishes in seconds.  Random (xavier) weights,
eval mode; forward outputs of the TF32 path within 1e-3."""
import pytest
import torch

from helpers import rel_l2

import vptr_oracle as O

pytestmark = pytest.mark.gpu


def test_nar_cfg1_shape_one_clip():
    from vptr_b200.model import VPTRFormerNAR
    torch.manual_seed(2021)
    net = VPTRFormerNAR(10, 10, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=4, num_decoder_layers=8, dropout=0.1,
                        window_size=4, rpe=True).eval()
    x = torch.rand(1, 10, 528, 8, 8, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        yo = O.vptr_former_nar({k: v for k, v in net.state_dict().items()}, x, nhead=8, ws=4, rpe=True, training=False)
        y = net.cuda()(x.cuda())
    assert tuple(y.shape) == (1, 10, 528, 8, 8)
    assert rel_l2(y, yo) < 1e-3


def test_nar_cfg3_shape_one_clip():
    """BAIR-shape 2 -> 28 (cfg3): temporal groups of 28 and enc-dec attention with 28 queries x 2 keys take the 32-row
    instantiation of the tensor-core attention kernels"""
    from vptr_b200.model import VPTRFormerNAR
    torch.manual_seed(2021)
    net = VPTRFormerNAR(2, 28, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=4, num_decoder_layers=8, dropout=0.1,
                        window_size=4, rpe=True).eval()
    x = torch.rand(1, 2, 528, 8, 8, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        yo = O.vptr_former_nar({k: v for k, v in net.state_dict().items()}, x, nhead=8, ws=4, rpe=True, training=False)
        y = net.cuda()(x.cuda())
    assert tuple(y.shape) == (1, 28, 528, 8, 8)
    assert rel_l2(y, yo) < 1e-3


def test_nar_cfg4_geometry_one_clip():
    """the stress geometry of cfg4 (16x16 feature grid, 8x8 windows = 64-token groups, 10 -> 30 frames, LayerNorm((2112,16,16))) at
    d_model 528 with one encoder and one decoder layer: the wide tensor-core window attention (attn_mma64_kernel), the 30 x 30 / 30 x 10
    temporal shapes and the register-window depthwise conv at the real dimensions"""
    from vptr_b200.model import VPTRFormerNAR
    torch.manual_seed(2021)
    net = VPTRFormerNAR(10, 30, encH=16, encW=16, d_model=528, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dropout=0.1,
                        window_size=8, rpe=True).eval()
    x = torch.rand(1, 10, 528, 16, 16, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        yo = O.vptr_former_nar({k: v for k, v in net.state_dict().items()}, x, nhead=8, ws=8, rpe=True, training=False)
        y = net.cuda()(x.cuda())
    assert tuple(y.shape) == (1, 30, 528, 16, 16)
    assert rel_l2(y, yo) < 1e-3


def test_nar_tslma_real_width_one_clip():
    """TSLMA_flag=True (temporal-spatial window cross-attention, reference VidHRFormer_modules.py:219-284) at d_model 528, 10 -> 10
    frames: 16 queries x 160 keys per (window, future frame) through attention mode 2"""
    from vptr_b200.model import VPTRFormerNAR
    torch.manual_seed(2021)
    net = VPTRFormerNAR(10, 10, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=1, num_decoder_layers=2, dropout=0.1,
                        window_size=4, TSLMA_flag=True, rpe=True).eval()
    assert any("TSLMA.attn.in_proj_weight" in k for k in net.state_dict()) and not any("EncDecAttn" in k for k in net.state_dict())
    x = torch.rand(1, 10, 528, 8, 8, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        yo = O.vptr_former_nar({k: v for k, v in net.state_dict().items()}, x, nhead=8, ws=4, rpe=True, training=False)
        y = net.cuda()(x.cuda())
    assert rel_l2(y, yo) < 1e-3


def test_far_cfg2_shape_one_clip():
    from vptr_b200.model import VPTRFormerFAR
    torch.manual_seed(2021)
    net = VPTRFormerFAR(10, 20, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=12, dropout=0.1, window_size=4, rpe=True).eval()
    x = torch.rand(1, 29, 528, 8, 8, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        yo = O.vptr_former_far({k: v for k, v in net.state_dict().items()}, x, nhead=8, ws=4, rpe=True, training=False)
        y = net.cuda()(x.cuda())
    assert rel_l2(y, yo) < 1e-3
    # causality of the CUDA path: perturbing frame t must not change outputs before t
    x2 = x.clone()
    x2[:, 20:] += 1.0
    with torch.no_grad():
        y2 = net(x2.cuda())
    assert torch.equal(y2[:, :20], y[:, :20])
    assert not torch.equal(y2[:, 20:], y[:, 20:])


def test_autoencoder_528_channels():
    from vptr_b200.model import VPTRDec, VPTREnc, init_weights
    import contextlib, io
    torch.manual_seed(2021)
    enc, dec = VPTREnc(1, feat_dim=528, n_downsampling=3).eval(), VPTRDec(1, feat_dim=528, n_downsampling=3, out_layer="Sigmoid").eval()
    with contextlib.redirect_stdout(io.StringIO()):
        init_weights(enc)
        init_weights(dec)
    x = torch.rand(1, 3, 1, 64, 64, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        fo = O.resnet_encoder(enc.state_dict(), x, 3, "reflect")
        ro = O.resnet_decoder(dec.state_dict(), fo, 3, "Sigmoid")
        f = enc.cuda()(x.cuda())
        r = dec.cuda()(f)
    assert rel_l2(f, fo) < 1e-3 and rel_l2(r, ro) < 1e-3
