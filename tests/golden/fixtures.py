"""Closed-form, platform-independent weights / inputs shared by make_golden.py (which runs the
unmodified reference in the build container) and by the tests (which rebuild the very same
tensors for the oracle and for the CUDA path).  Nothing random: fixtures stay tiny because
only *outputs* have to be stored."""
import numpy as np
import torch


def closed_form(shape, salt, scale=1.0, offset=0.0):
    n = int(np.prod(shape)) if len(shape) else 1
    i = np.arange(n, dtype=np.float64)
    v = np.sin(0.37 * i + 1.3 * salt + 0.11 * np.sqrt(i + 1.0)) * scale + offset
    return torch.from_numpy(v.astype(np.float32)).reshape(shape)


def fill_state_dict_(sd, salt0=0):
    """Deterministically overwrite every floating-point parameter / running stat of a
    reference-format state_dict (positional buffers and integer buffers are left alone)."""
    keep = ("temporal_pos", "lw_pos", "Tlw_pos")
    for salt, (k, v) in enumerate(sd.items()):
        if not v.dtype.is_floating_point or k in keep:
            continue
        s = salt + salt0
        leaf = k.rsplit(".", 1)[-1]
        if leaf == "running_var":
            new = closed_form(v.shape, s, 0.4, 1.0)
        elif leaf == "running_mean":
            new = closed_form(v.shape, s, 0.1)
        elif leaf == "bias" or leaf == "in_proj_bias":
            new = closed_form(v.shape, s, 0.1)
        elif leaf == "weight" and (v.dim() == 1 or ("norm" in k and v.dim() == 3)):
            new = closed_form(v.shape, s, 0.2, 1.0)              # norm scales around 1
        elif leaf == "relative_position_bias_table":
            new = closed_form(v.shape, s, 0.5)
        elif leaf == "frame_queries":
            new = closed_form(v.shape, s, 0.5)
        else:                                                    # conv / linear weights
            fan_in = int(np.prod(v.shape[1:])) if v.dim() > 1 else v.shape[0]
            new = closed_form(v.shape, s, 1.0 / np.sqrt(max(fan_in, 1)))
        v.copy_(new)
    return sd


def probe(shape, salt):
    """Closed-form cotangent / probe vector."""
    return closed_form(shape, 1000 + salt, 1.0)


def grad_signature(g, salt):
    """Three numbers that pin a gradient tensor without storing it."""
    g64 = g.detach().double().reshape(-1)
    p = probe(g.shape, salt).double().reshape(-1)
    return np.array([g64.sum().item(), g64.abs().sum().item(), (g64 * p).sum().item()], dtype=np.float64)


CASES = {
    "ae_reflect": dict(img_channels=1, feat_dim=24, n_down=3, padding_type="reflect", out_layer="Sigmoid", N=1, T=2, HW=32),
    "ae_zero": dict(img_channels=3, feat_dim=16, n_down=3, padding_type="zero", out_layer="Tanh", N=1, T=2, HW=32),
    "nar_rpe": dict(kind="nar", Tp=2, Tf=3, encH=8, encW=8, d_model=48, nhead=4, enc_layers=2, dec_layers=2, ws=4, rpe=True, N=2),
    "far_rpe": dict(kind="far", Tp=2, Tf=3, encH=8, encW=8, d_model=48, nhead=4, enc_layers=2, ws=4, rpe=True, N=2, T_in=4),
    "far_norpe_pad": dict(kind="far", Tp=2, Tf=2, encH=6, encW=6, d_model=48, nhead=4, enc_layers=1, ws=4, rpe=False, N=1, T_in=3),
    # TSLMA_flag=True: the decoder's encoder-decoder attention is TemporalSpatialLocalMultiheadAttention (off in every reference script)
    "nar_tslma": dict(kind="nar", Tp=2, Tf=3, encH=8, encW=8, d_model=48, nhead=4, enc_layers=1, dec_layers=2, ws=4, rpe=True, N=2, tslma=True),
}
