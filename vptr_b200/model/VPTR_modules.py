"""Drop-in replacements for the reference's public model API (model/VPTR_modules.py:10-197 of XiYe20/VPTR):
VPTREnc, VPTRDec, VPTRFormerNAR, VPTRFormerFAR (+ VPTRDisc, which stays a plain PyTorch module: disabled in every
stage-2 script).  Constructor signatures, forward shapes, attributes and state_dict keys/shapes are
the reference's; parameters live in ordinary nn.* holder modules so init_weights / optimizers / checkpoints / DDP see
what they expect, while `forward` runs the hand-written sm_100a kernels of libvptr_b200.so through vptr_b200.engine.
There is no PyTorch or CPU fallback: inputs must be CUDA float32 tensors."""
import functools

import torch
import torch.nn as nn

from .. import engine as E
from .. import ops
from ..position_encoding import pos_1d, pos_2d, pos_3d
from .ResNetAutoEncoder import (ResnetDecoder, ResnetEncoder, decoder_backward, decoder_backward_train, decoder_forward,
                                decoder_forward_train, encoder_backward_train, encoder_forward, encoder_forward_train)


# ===================================================================================================== ResNet wrappers
class VPTREnc(nn.Module):
    def __init__(self, img_channels, feat_dim=528, n_downsampling=3, padding_type='reflect'):
        super().__init__()
        self.feat_dim = feat_dim
        self.encoder = ResnetEncoder(input_nc=img_channels, out_dim=feat_dim, n_downsampling=n_downsampling, padding_type=padding_type)

    def forward(self, x):
        """x (N, T, img_channels, H, W) -> (N, T, feat_dim, H/2^n, W/2^n).  Forward only (stage 2 calls it under
        no_grad with BatchNorm in eval mode, train_NAR.py:54-56,190)."""
        N, T = x.shape[:2]
        if self.training:       # stage-1 autoencoder training (train_AutoEncoder.py:53-57): batch statistics + every weight gradient
            _check_input(x, "VPTREnc")
            params = list(self.encoder.parameters())
            record = torch.is_grad_enabled() and any(p.requires_grad for p in params)
            feat = _EncTrainFunction.apply(self.encoder, x.flatten(0, 1), record, *params)
            H, W = x.shape[-2] >> self.encoder.n_downsampling, x.shape[-1] >> self.encoder.n_downsampling
            return feat.view(N, T, H, W, self.feat_dim).permute(0, 1, 4, 2, 3)
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("vptr_b200.VPTREnc in eval mode is forward-only (the reference runs it under torch.no_grad())")
        feat, H, W = encoder_forward(self.encoder, x.flatten(0, 1))          # (F*H*W, C) channel-last
        return feat.view(N, T, H, W, self.feat_dim).permute(0, 1, 4, 2, 3)   # same values/shape as the reference's NCHW tensor


class _DecFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, feat):
        N, T, C, H, W = feat.shape
        feat_cl = _channel_last(feat.flatten(0, 1)).view(N * T * H * W, C)
        out, saved = decoder_forward(dec, feat_cl, N * T, H, W, save=True)
        ctx.dec, ctx.saved, ctx.shape = dec, saved, (N, T, C, H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        N, T, C, H, W = ctx.shape
        dfeat = decoder_backward(ctx.dec, ctx.saved, dout.contiguous())          # (F*H*W, C) channel-last
        ctx.saved = None
        return None, dfeat.view(N, T, H, W, C).permute(0, 1, 4, 2, 3)


class _EncTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, frames, record, *params):
        tape = [] if record else None
        feat, _, _ = encoder_forward_train(enc, frames, tape)
        ctx.tape, ctx.params = tape, params
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        grads = encoder_backward_train(ctx.tape, dfeat.contiguous())
        ctx.tape = None
        return (None, None, None) + tuple(grads.get(p) for p in ctx.params)


class _DecTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, feat, record, *params):
        N, T, C, H, W = feat.shape
        feat_cl = _channel_last(feat.flatten(0, 1)).view(N * T * H * W, C)
        tape = [] if record else None
        out = decoder_forward_train(dec, feat_cl, N * T, H, W, tape)
        ctx.tape, ctx.params, ctx.shape = tape, params, (N, T, C, H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        N, T, C, H, W = ctx.shape
        dfeat, grads = decoder_backward_train(ctx.tape, dout)
        ctx.tape = None
        return (None, dfeat.view(N, T, H, W, C).permute(0, 1, 4, 2, 3), None) + tuple(grads.get(p) for p in ctx.params)


class VPTRDec(nn.Module):
    def __init__(self, img_channels, feat_dim=528, n_downsampling=3, out_layer='Tanh', padding_type='reflect'):
        super().__init__()
        self.decoder = ResnetDecoder(output_nc=img_channels, feat_dim=feat_dim, n_downsampling=n_downsampling, out_layer=out_layer,
                                     padding_type=padding_type)

    def forward(self, feat):
        """feat (N, T, feat_dim, h, w) -> (N, T, img_channels, H, W).  Differentiable w.r.t. feat; the decoder's own
        weight gradients (computed but never consumed by the reference) are not produced."""
        N, T, C, H, W = feat.shape
        if not feat.is_cuda or feat.dtype != torch.float32:
            raise RuntimeError("vptr_b200.VPTRDec: input must be a CUDA float32 tensor (got %s, %s); there is no CPU fallback" % (feat.device, feat.dtype))
        if self.training:       # stage-1 autoencoder training: batch statistics, input AND weight gradients
            params = list(self.decoder.parameters())
            record = torch.is_grad_enabled() and (feat.requires_grad or any(p.requires_grad for p in params))
            out = _DecTrainFunction.apply(self.decoder, feat, record, *params)
            return out.view(N, T, *out.shape[1:])
        if torch.is_grad_enabled() and feat.requires_grad:
            out = _DecFunction.apply(self.decoder, feat)
        else:
            feat_cl = _channel_last(feat.flatten(0, 1)).view(N * T * H * W, C)   # (F*H*W, C) contiguous
            out, _ = decoder_forward(self.decoder, feat_cl, N * T, H, W, save=False)
        return out.view(N, T, *out.shape[1:])


def _channel_last(x):
    """(F, C, H, W) -> contiguous (F, H, W, C); free when x is already a channel-last view (as our own modules emit)."""
    xp = x.permute(0, 2, 3, 1)
    if xp.is_contiguous():
        return xp
    F_, C, H, W = x.shape
    xc = x.contiguous()
    return ops.transpose(xc, F_, C, H * W).view(F_, H, W, C)


class VPTRDisc(nn.Module):
    """PatchGAN discriminator (reference model/VPTR_modules.py:49-95).  Plain PyTorch: it is disabled
    (`VPTR_Disc = None`) in every stage-2 script and is not on the hot path."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d):
        super().__init__()
        inst = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
        use_bias = inst == nn.InstanceNorm2d
        layers = [nn.Conv2d(input_nc, ndf, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(0.2, True)]
        mult = 1
        for n in range(1, n_layers + 1):
            prev, mult = mult, min(2 ** n, 8)
            stride = 2 if n < n_layers else 1
            layers += [nn.Conv2d(ndf * prev, ndf * mult, kernel_size=4, stride=stride, padding=1, bias=use_bias),
                       norm_layer(ndf * mult), nn.LeakyReLU(0.2, True)]
        layers += [nn.Conv2d(ndf * mult, 1, kernel_size=4, stride=1, padding=1)]
        self.model = nn.Sequential(*layers)

    def forward(self, input):
        return self.model(input)


# ===================================================================================================== parameter holders
class _RPEAttnHolder(nn.Module):
    """Parameters of MultiheadAttentionRPE (reference model/MultiHeadAttentionRPE.py:50-53,359-388)."""

    def __init__(self, embed_dim, num_heads, window_size):
        super().__init__()
        ws = window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), num_heads))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)
        i = torch.arange(ws * ws)
        ih, iw = i // ws, i % ws
        idx = (ih[:, None] - ih[None, :] + ws - 1) * (2 * ws - 1) + (iw[:, None] - iw[None, :] + ws - 1)
        self.register_buffer("relative_position_index", idx.to(torch.int64))
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, embed_dim)


class _SLMHSAHolder(nn.Module):
    def __init__(self, embed_dim, num_heads, window_size, dropout, rpe):
        super().__init__()
        self.attn = _RPEAttnHolder(embed_dim, num_heads, window_size) if rpe else nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout)


class _MlpDWBNHolder(nn.Module):
    """Parameters of MlpDWBN (reference model/VidHRFormer_modules.py:380-422)."""

    def __init__(self, encH, encW, in_features, hidden_features, layer_norm):
        super().__init__()
        norm = (lambda ch: nn.LayerNorm((ch, encH, encW))) if layer_norm else (lambda ch: nn.BatchNorm2d(ch))
        self.fc1 = nn.Conv2d(in_features, hidden_features, kernel_size=1)
        self.norm1 = norm(hidden_features)
        self.dw3x3 = nn.Conv2d(hidden_features, hidden_features, kernel_size=3, stride=1, groups=hidden_features, padding=1)
        self.norm2 = norm(hidden_features)
        self.fc2 = nn.Conv2d(hidden_features, in_features, kernel_size=1)
        self.norm3 = norm(in_features)


class _EncBlockHolder(nn.Module):
    def __init__(self, encH, encW, d_model, nhead, dim_feedforward, dropout, window_size, ffn_ratio, far, rpe):
        super().__init__()
        self.SLMHSA = _SLMHSAHolder(d_model, nhead, window_size, dropout, rpe)
        self.SpatialFFN = _MlpDWBNHolder(encH, encW, d_model, d_model * ffn_ratio, layer_norm=far)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.temporal_MHSA = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm4 = nn.LayerNorm(d_model)


class _TSLMAHolder(nn.Module):
    """Parameters of TemporalSpatialLocalMultiheadAttention (reference model/VidHRFormer_modules.py:219-246)."""

    def __init__(self, embed_dim, num_heads, dropout):
        super().__init__()
        self.attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout)


class _DecBlockHolder(nn.Module):
    def __init__(self, encH, encW, d_model, nhead, dim_feedforward, dropout, window_size, ffn_ratio, rpe, tslma=False):
        super().__init__()
        self.SLMHSA = _SLMHSAHolder(d_model, nhead, window_size, dropout, rpe)
        self.SpatialFFN = _MlpDWBNHolder(encH, encW, d_model, d_model * ffn_ratio, layer_norm=True)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.temporal_MHSA = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm4 = nn.LayerNorm(d_model)
        if tslma:        # reference :154-158: TSLMA replaces EncDecAttn (and its state_dict keys)
            self.TSLMA = _TSLMAHolder(d_model, nhead, dropout)
        else:
            self.EncDecAttn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.SpatialFFN1 = _MlpDWBNHolder(encH, encW, d_model, d_model * ffn_ratio, layer_norm=True)
        self.norm5 = nn.LayerNorm(d_model)
        self.norm6 = nn.LayerNorm(d_model)


class _Stack(nn.Module):
    def __init__(self, layers, d_model):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        self.norm = nn.LayerNorm(d_model)


class _TransformerHolder(nn.Module):
    def __init__(self, encoder, decoder=None):
        super().__init__()
        self.encoder = encoder
        if decoder is not None:
            self.decoder = decoder


class _GemmLinear(nn.Linear):
    """nn.Linear whose forward/backward run on the tcgen05 GEMM (used for NCE_projector, reference VPTR_modules.py:133-135)."""

    def forward(self, x):
        return _LinearFunction.apply(x, self.weight, self.bias)


class _LinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        x2, w = E._rc(x2), E._rc(w)
        y = ops.gemm(x2, w, bias=b)
        ctx.save_for_backward(x2, w)
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dy2 = E._rc(dy2)
        dx = ops.gemm(dy2, w, b_mn=True).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1]:
            dw = torch.zeros_like(w)
            ops.gemm(dy2, x2, out=dw, a_mn=True, b_mn=True, accumulate=True)
        if ctx.needs_input_grad[2]:
            db = torch.zeros_like(w[:, 0])
            ops.colsum(dy2, db)
        return dx, dw, db


class _GemmReLU(nn.ReLU):
    def forward(self, x):
        return _ReLUFunction.apply(x)


class _ReLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = ops.relu_fwd(x.contiguous())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return ops.relu_bwd(dy.contiguous(), y)


# ===================================================================================================== transformer functions
def _check_input(x, what):
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError("vptr_b200.%s: input must be a CUDA float32 tensor (got %s, %s); there is no CPU fallback" % (what, x.device, x.dtype))


def _drop_ctx(mod, n_clips, device):
    """train mode with p > 0: dropout + DropPath (rate = dropout, reference VPTR_modules.py:114) fused into the kernels"""
    if mod.training and mod.dropout > 0:
        return E.Drop(mod.dropout, n_clips, device)
    return E.NO_DROP


def _tokens(x):
    """(N,T,C,H,W) -> contiguous token-major (N*T*H*W, C)."""
    N, T, C, H, W = x.shape
    return _channel_last(x.flatten(0, 1)).view(N * T * H * W, C)


class _FARFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, names, record, *params):
        N, T, C, H, W = x.shape
        want = record and any(ctx.needs_input_grad)      # grad mode is always off in here: the module's forward passes it in
        P = E.Params(zip(names, params), want_grads=False)
        bufs = dict(mod.named_buffers())
        lean = want and E.lean_for(N * T * H * W, mod.num_encoder_layers, 0, 0, x.device)
        g = E.Geom(N, T, H, W, C, mod.nhead, mod.window_size, lean=lean)
        save = [] if want else None
        lw_tab = None if mod.rpe else E.lw_table(mod.lw_pos, g)
        tpos = mod.temporal_pos[:T].contiguous()
        D = _drop_ctx(mod, N, x.device)
        h = E.encoder_fwd(P, bufs, _tokens(x), g, mod.num_encoder_layers, True, mod.rpe, tpos, lw_tab, mod.training, save, D)
        y = E.final_norm_fwd(P, "transformer.encoder.norm", h, True, save)
        ctx.save, ctx.names, ctx.params, ctx.shape, ctx.mod = save, names, params, (N, T, C, H, W), mod
        ctx.rounded = P.rounded if (want and E.ROUND_TF32) else None
        return y.view(N, T, H, W, C).permute(0, 1, 4, 2, 3)

    @staticmethod
    def backward(ctx, dout):
        N, T, C, H, W = ctx.shape
        P = E.Params(zip(ctx.names, ctx.params), want_grads=True, rounded=ctx.rounded if E.ROUND_TF32 else None, holder=ctx.mod)
        ctx.rounded = None
        d = _tokens(dout)
        dx = E.backward_tape(P, ctx.save, d)
        E.join_side(P)
        ctx.save = None
        grads = tuple(P.g(n) for n in ctx.names)
        return (None, dx.view(N, T, H, W, C).permute(0, 1, 4, 2, 3), None, None) + grads


class _NARFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, names, record, *params):
        N, Tp, C, H, W = x.shape
        Tf = mod.num_future_frames
        want = record and any(ctx.needs_input_grad)
        P = E.Params(zip(names, params), want_grads=False)
        bufs = dict(mod.named_buffers())
        lean = want and E.lean_for(N * Tp * H * W, mod.num_encoder_layers, N * Tf * H * W, mod.num_decoder_layers, x.device)
        ge = E.Geom(N, Tp, H, W, C, mod.nhead, mod.window_size, lean=lean)
        gd = E.Geom(N, Tf, H, W, C, mod.nhead, mod.window_size, lean=lean)
        save = [] if want else None
        lw_tab = None if mod.rpe else E.lw_table(mod.lw_pos, ge)
        tpos_p = mod.temporal_pos[:Tp].contiguous()
        tpos_f = mod.temporal_pos[Tp:Tp + Tf].contiguous()
        D = _drop_ctx(mod, N, x.device)
        h = E.encoder_fwd(P, bufs, _tokens(x), ge, mod.num_encoder_layers, False, mod.rpe, tpos_p, lw_tab, mod.training, save, D)
        mem = E.final_norm_fwd(P, "transformer.encoder.norm", h, False, save, round_out=True)
        n_enc = len(save) if want else 0
        qpos = P.w("frame_queries").reshape(Tf * H * W, C)                       # query_pos (VidHRFormer.py:46)
        qadd = ops.add_rows(qpos, tpos_f, H * W, Tf)                             # query_pos + pos_future (VidHRFormer_modules.py:200)
        mem_k = ops.add_rows(mem, tpos_p, H * W, Tp, round_tf32=E.ROUND_TF32)                             # memory + pos_past
        tgt = ops.zeros(gd.R, C, like=x)                                         # init_tgt = zeros (VidHRFormer.py:48)
        tslma = None
        if mod.TSLMA_flag:   # Tlw_pos (T, ws, ws, C) laid over the grid: row (t, h, w) -> Tlw_pos[t, h % ws, w % ws]
            ws = mod.window_size
            hh, ww = torch.arange(H, device=x.device) % ws, torch.arange(W, device=x.device) % ws
            grid = mod.Tlw_pos[:, hh][:, :, ww]                                  # (Tp+Tf, H, W, C)
            k_tab = grid[:Tp].reshape(Tp * H * W, C).contiguous()
            q_tab = (grid[Tp:Tp + Tf].reshape(Tf * H * W, C) + qpos).contiguous() # query = LN5(tgt) + query_pos (+ Tlw_pos after the permute)
            tslma = dict(q_tab=q_tab, mem_k=ops.add_rows(mem, k_tab, 1, k_tab.shape[0], round_tf32=E.ROUND_TF32))
        tgt = E.decoder_fwd(P, bufs, tgt, gd, ge, mod.num_decoder_layers, mod.rpe, qpos, qadd, tpos_f, mem, mem_k, lw_tab, save, D, tslma=tslma)
        y = E.final_norm_fwd(P, "transformer.decoder.norm", tgt, True, save)
        ctx.save, ctx.names, ctx.params, ctx.n_enc, ctx.mod = save, names, params, n_enc, mod
        ctx.rounded = P.rounded if (want and E.ROUND_TF32) else None
        ctx.shape = (N, Tp, Tf, C, H, W)
        return y.view(N, Tf, H, W, C).permute(0, 1, 4, 2, 3)

    @staticmethod
    def backward(ctx, dout):
        N, Tp, Tf, C, H, W = ctx.shape
        P = E.Params(zip(ctx.names, ctx.params), want_grads=True, rounded=ctx.rounded if E.ROUND_TF32 else None, holder=ctx.mod)
        ctx.rounded = None
        dq = P.g("frame_queries")
        dqpos = dq.view(Tf * H * W, C) if dq is not None else None
        dmem = ops.zeros(N * Tp * H * W, C, like=dout)
        E.backward_tape(P, ctx.save, _tokens(dout), dqpos=dqpos, dmem=dmem, stop=ctx.n_enc)   # decoder (+ its final norm)
        dx = E.backward_tape(P, ctx.save, dmem)                                                # encoder norm + encoder
        E.join_side(P)
        ctx.save = None
        grads = tuple(P.g(n) for n in ctx.names)
        return (None, dx.view(N, Tp, H, W, C).permute(0, 1, 4, 2, 3), None, None) + grads


def _fwd_params(mod):
    names, params = [], []
    for k, v in mod.named_parameters():
        if k.startswith("NCE_projector"):
            continue
        names.append(k)
        params.append(v)
    return tuple(names), params


# ===================================================================================================== public transformers
class VPTRFormerNAR(nn.Module):
    def __init__(self, num_past_frames, num_future_frames, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=6,
                 num_decoder_layers=6, dropout=0.1, window_size=4, Spatial_FFN_hidden_ratio=4, TSLMA_flag=False, rpe=True):
        super().__init__()
        if TSLMA_flag and (encH % window_size or encW % window_size):
            raise NotImplementedError("vptr_b200: TSLMA_flag=True needs a feature grid that is a multiple of the window "
                                      "(%dx%d grid, window %d)" % (encH, encW, window_size))
        self.TSLMA_flag = TSLMA_flag
        self.num_past_frames, self.num_future_frames = num_past_frames, num_future_frames
        self.nhead, self.d_model = nhead, d_model
        self.num_encoder_layers, self.num_decoder_layers = num_encoder_layers, num_decoder_layers
        self.dropout, self.window_size, self.Spatial_FFN_hidden_ratio = dropout, window_size, Spatial_FFN_hidden_ratio
        self.rpe = rpe
        ff = d_model * Spatial_FFN_hidden_ratio
        enc = _Stack([_EncBlockHolder(encH, encW, d_model, nhead, ff, dropout, window_size, Spatial_FFN_hidden_ratio, False, rpe)
                      for _ in range(num_encoder_layers)], d_model)
        dec = _Stack([_DecBlockHolder(encH, encW, d_model, nhead, ff, dropout, window_size, Spatial_FFN_hidden_ratio, rpe, TSLMA_flag)
                      for _ in range(num_decoder_layers)], d_model)
        self.transformer = _TransformerHolder(enc, dec)
        T = num_past_frames + num_future_frames
        self.register_buffer('temporal_pos', pos_1d(T, d_model))
        self.register_buffer('lw_pos', pos_2d(d_model, window_size, window_size))
        self.register_buffer('Tlw_pos', pos_3d(d_model, T, window_size, window_size))
        self.frame_queries = nn.Parameter(torch.randn(num_future_frames, encH, encW, d_model), requires_grad=True)
        self.NCE_projector = nn.Sequential(_GemmLinear(d_model, d_model), _GemmReLU(inplace=True), _GemmLinear(d_model, d_model))
        self._reset_parameters()

    def forward(self, past_gt_feat):
        """past_gt_feat (N, Tp, C, H, W) -> predicted future features (N, Tf, C, H, W)."""
        _check_input(past_gt_feat, "VPTRFormerNAR")
        names, params = _fwd_params(self)
        return _NARFunction.apply(self, past_gt_feat, names, torch.is_grad_enabled(), *params)

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class VPTRFormerFAR(nn.Module):
    def __init__(self, num_past_frames, num_future_frames, encH=8, encW=8, d_model=528, nhead=8, num_encoder_layers=6, dropout=0.1,
                 window_size=4, Spatial_FFN_hidden_ratio=4, rpe=True):
        super().__init__()
        self.num_past_frames, self.num_future_frames = num_past_frames, num_future_frames
        self.nhead, self.d_model, self.num_encoder_layers = nhead, d_model, num_encoder_layers
        self.dropout, self.window_size, self.Spatial_FFN_hidden_ratio = dropout, window_size, Spatial_FFN_hidden_ratio
        self.rpe = rpe
        ff = d_model * Spatial_FFN_hidden_ratio
        enc = _Stack([_EncBlockHolder(encH, encW, d_model, nhead, ff, dropout, window_size, Spatial_FFN_hidden_ratio, True, rpe)
                      for _ in range(num_encoder_layers)], d_model)
        self.transformer = _TransformerHolder(enc)
        T = num_past_frames + num_future_frames
        self.register_buffer('temporal_pos', pos_1d(T, d_model))
        self.register_buffer('lw_pos', pos_2d(d_model, window_size, window_size))
        self._reset_parameters()

    def forward(self, input_feats):
        """input_feats (N, T, C, H, W), any T <= Tp+Tf -> same shape; output t predicts frame t+1 (causal in time)."""
        _check_input(input_feats, "VPTRFormerFAR")
        names, params = _fwd_params(self)
        return _FARFunction.apply(self, input_feats, names, torch.is_grad_enabled(), *params)

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
