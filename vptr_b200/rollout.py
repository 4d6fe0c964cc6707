"""Inference rollouts.

`far_rollout` is the reference's `FAR_show_sample` test-phase loop (train_FAR.py:103-134) with an incremental Transformer: the
reference re-runs the WHOLE VidHRFormerFAR stack on the growing sequence for every new frame (O(T^2) layer passes).  The model is
causal in time -- the window attention, the LayerNorm-flavoured conv FFN and the MLP act per frame, and the temporal attention is
masked (VidHRFormer_modules.py:76-82) -- so the outputs of earlier positions never change.  Here the temporal
K / V projections of every layer are cached and each new frame costs ONE pass of ONE frame through the stack, attending to the
cache (Tq = 1, Tk = t + 1).  The results equal the full recompute up to fp32 rounding (tests/test_gpu_rollout.py).

`nar_rollout` chains NAR predictions for horizons beyond num_future_frames (Test_VPTR.ipynb cell 5, NAR_test_single_iter)."""
import torch

from . import engine as E
from . import ops


class FARCache:
    """per-layer temporal K / V of the frames seen so far: (N, T_seen, HW, C) each, plus the running output features"""

    def __init__(self):
        self.k, self.v, self.t = [], [], 0


def _frame_step(mod, P, bufs, x, cache, lw_tab):
    """one new frame x (N*HW, C) at temporal position cache.t through every layer; appends to the caches; returns (N*HW, C)"""
    N = x.shape[0] // (mod._rollout_hw[0] * mod._rollout_hw[1])
    H, W = mod._rollout_hw
    C = x.shape[1]
    g1 = E.Geom(N, 1, H, W, C, mod.nhead, mod.window_size)
    t = cache.t
    pos_t = mod.temporal_pos[t:t + 1].contiguous()
    RT = E.ROUND_TF32
    for i in range(mod.num_encoder_layers):
        pre = "transformer.encoder.layers.%d" % i
        x = E.window_attn_fwd(P, pre + ".SLMHSA", pre + ".norm1", x, g1, mod.rpe, None, lw_tab, None)
        x = E.conv_ffn_fwd(P, bufs, pre + ".SpatialFFN", pre + ".norm2", x, g1, True, False, None)
        # temporal attention of the new frame against the cache
        tp = pre + ".temporal_MHSA"
        z, zp, _, _ = ops.layernorm_fwd(x, P.w(pre + ".norm3.weight"), P.w(pre + ".norm3.bias"), add=pos_t, add_div=g1.HW, add_mod=1,
                                        round_tf32=RT, save_stats=False)
        Wi, bi = P.wr(tp + ".in_proj_weight"), P.w(tp + ".in_proj_bias")
        qkv = ops.empty(g1.R, 3 * C, like=x)
        ops.gemm(zp, Wi[:2 * C], out=qkv[:, :2 * C], bias=bi[:2 * C])      # (the 3xTF32 attention kernel takes unrounded q / k / v)
        ops.gemm(z, Wi[2 * C:], out=qkv[:, 2 * C:], bias=bi[2 * C:])
        k_new = qkv[:, C:2 * C].reshape(N, 1, g1.HW, C)
        v_new = qkv[:, 2 * C:].reshape(N, 1, g1.HW, C)
        if len(cache.k) <= i:
            cache.k.append(k_new.contiguous())
            cache.v.append(v_new.contiguous())
        else:
            cache.k[i] = torch.cat([cache.k[i], k_new], dim=1)
            cache.v[i] = torch.cat([cache.v[i], v_new], dim=1)
        Tk = cache.k[i].shape[1]
        q = qkv[:, :C].contiguous()
        o = ops.empty(g1.R, C, like=x)
        ops.attn_fwd(q, cache.k[i].view(N * Tk * g1.HW, C), cache.v[i].view(N * Tk * g1.HW, C), o, None, 1, N, H, W, 0, 1, Tk,
                     mod.nhead, g1.d, False, g1.scale, round_tf32=RT)
        x = ops.gemm(o, P.wr(tp + ".out_proj.weight"), bias=P.w(tp + ".out_proj.bias"), residual=x)
        x = E.mlp_fwd(P, pre, pre + ".norm4", x, g1, None)
    cache.t = t + 1
    return E.final_norm_fwd(P, "transformer.encoder.norm", x, True, None)


def _prefill(mod, P, bufs, feats, cache, lw_tab):
    """the given frames in one pass (the ordinary causal forward), keeping every layer's temporal K / V"""
    N, T, C, H, W = feats.shape
    from .model.VPTR_modules import _tokens
    g = E.Geom(N, T, H, W, C, mod.nhead, mod.window_size)
    save = []
    tpos = mod.temporal_pos[:T].contiguous()
    h = E.encoder_fwd(P, bufs, _tokens(feats), g, mod.num_encoder_layers, True, mod.rpe, tpos, lw_tab, False, save)
    y = E.final_norm_fwd(P, "transformer.encoder.norm", h, True, None)
    for kind, s in save:
        if kind == "tattn":
            qkv = s["qkv"]
            cache.k.append(qkv[:, C:2 * C].reshape(N, T, H * W, C).contiguous())
            cache.v.append(qkv[:, 2 * C:].reshape(N, T, H * W, C).contiguous())
    save.clear()
    cache.t = T
    return y.view(N, T, H, W, C)


@torch.no_grad()
def far_rollout(enc, dec, transformer, past_frames, num_pred):
    """past_frames (N, Tp, Cimg, H, W) -> (pred_frames (N, Tp - 1 + num_pred, Cimg, H, W), pred_feats), exactly the tensors
    FAR_show_sample's test phase ends with (train_FAR.py:111-126): frame t of the output predicts input frame t + 1; the last
    num_pred entries are the future.  The first generated feature is fed back directly, later ones through Dec -> Enc (:115-121)."""
    mod = transformer
    if mod.training or enc.training or dec.training:
        raise RuntimeError("vptr_b200.rollout.far_rollout: call .eval() on the modules first (the reference does, train_FAR.py:104)")
    names, params = [], []
    for k, v in mod.named_parameters():
        if not k.startswith("NCE_projector"):
            names.append(k)
            params.append(v)
    P = E.Params(zip(names, params), want_grads=False)
    bufs = dict(mod.named_buffers())
    feats = enc(past_frames)                                       # (N, Tp, C, h, w) channel-last view
    N, Tp, C, H, W = feats.shape
    if Tp + num_pred - 1 > mod.temporal_pos.shape[0]:
        raise RuntimeError("vptr_b200.rollout.far_rollout: %d frames exceed the model's %d temporal positions" % (Tp + num_pred - 1, mod.temporal_pos.shape[0]))
    mod._rollout_hw = (H, W)
    lw_tab = None if mod.rpe else E.lw_table(mod.lw_pos, E.Geom(N, 1, H, W, C, mod.nhead, mod.window_size))
    cache = FARCache()
    outs = [_prefill(mod, P, bufs, feats, cache, lw_tab)]          # (N, Tp, H, W, C)
    last = outs[0][:, -1]                                           # prediction of frame Tp
    for i in range(num_pred - 1):
        if i == 0:
            nxt = last.reshape(N * H * W, C)                        # fed back as a feature (train_FAR.py:116)
        else:
            frame = dec(last.permute(0, 3, 1, 2).unsqueeze(1))      # Dec -> Enc round trip (:118-120)
            nxt = enc(frame)[:, 0].permute(0, 2, 3, 1).reshape(N * H * W, C)
        y = _frame_step(mod, P, bufs, nxt.contiguous(), cache, lw_tab).view(N, 1, H, W, C)
        outs.append(y)
        last = y[:, 0]
    pred_feats = torch.cat(outs, dim=1).permute(0, 1, 4, 2, 3)      # (N, Tp - 1 + num_pred, C, H, W)
    return dec(pred_feats), pred_feats


@torch.no_grad()
def far_rollout_recompute(enc, dec, transformer, past_frames, num_pred):
    """the reference's literal loop (whole-sequence recompute per frame) on the same modules: the checker for far_rollout"""
    past_gt_feats = enc(past_frames)
    pred_feats = transformer(past_gt_feats)
    input_feats = None
    for i in range(num_pred - 1):
        if i == 0:
            input_feats = torch.cat([past_gt_feats, pred_feats[:, -1:, ...]], dim=1)
        else:
            pred_future_feat = enc(dec(pred_feats[:, -1:, ...]))
            input_feats = torch.cat([input_feats, pred_future_feat], dim=1)
        pred_feats = transformer(input_feats)
    return dec(pred_feats), pred_feats


@torch.no_grad()
def nar_rollout(enc, dec, transformer, past_frames, num_blocks):
    """chained NAR prediction (Test_VPTR.ipynb cell 5): predict Tf frames, feed the last Tp predicted frames back, repeat"""
    Tp = transformer.num_past_frames
    frames, cur = [], past_frames
    for _ in range(num_blocks):
        pred = dec(transformer(enc(cur)))
        frames.append(pred)
        cur = torch.cat([cur, pred], dim=1)[:, -Tp:]
    return torch.cat(frames, dim=1)
