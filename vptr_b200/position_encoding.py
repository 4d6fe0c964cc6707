"""Sinusoidal positional buffers, evaluated once at construction (restates utils/position_encoding.py:13-160 of the
reference: positions start at 1, no normalisation, interleaved sin/cos; 2-D = [y | x] halves, 3-D = [t | y | x] thirds).
Returned channel-last."""
import torch


def _dim_t(E, temperature=10000.0):
    i = torch.arange(E, dtype=torch.float32)
    return temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / E)


def _interleave(p):
    return torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=-1).flatten(-2)


def pos_1d(L, E):
    """(L, E); utils/position_encoding.py:29-49"""
    pos = torch.arange(1, L + 1, dtype=torch.float32)
    return _interleave(pos[:, None] / _dim_t(E))


def pos_2d(E, H, W):
    """(H, W, E); utils/position_encoding.py:67-93"""
    y = torch.arange(1, H + 1, dtype=torch.float32)[:, None].expand(H, W)
    x = torch.arange(1, W + 1, dtype=torch.float32)[None, :].expand(H, W)
    d = _dim_t(E // 2)
    return torch.cat((_interleave(y[..., None] / d), _interleave(x[..., None] / d)), dim=-1)


def pos_3d(E, T, H, W):
    """(T, H, W, E); utils/position_encoding.py:117-158 (needs E % 3 == 0, :129)"""
    assert E % 3 == 0, "d_model must be divisible by 3 for the 3-D positional embedding"
    t = torch.arange(1, T + 1, dtype=torch.float32)[:, None, None].expand(T, H, W)
    y = torch.arange(1, H + 1, dtype=torch.float32)[None, :, None].expand(T, H, W)
    x = torch.arange(1, W + 1, dtype=torch.float32)[None, None, :].expand(T, H, W)
    d = _dim_t(E // 3)
    return torch.cat((_interleave(t[..., None] / d), _interleave(y[..., None] / d), _interleave(x[..., None] / d)), dim=-1)
