"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the VPTR stage-2 hot path.

A from-scratch, functional, channel-last restatement (plain torch fp32/fp64 tensor
arithmetic on the CPU) of the reference algorithm:  ResNet encoder -> VidHRFormer
(NAR / FAR) -> ResNet decoder.  It consumes a *reference-format* state_dict (Appendix B of
SURVEY.md) and is differentiable through torch autograd, so it is the checker for outputs,
input gradients and parameter gradients of the CUDA path.

Who may import this:  tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline /
--impl reference legs.  The product package vptr_b200/ never does (and fails loudly when
its CUDA library is missing instead of falling back to anything in here).

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the unmodified reference itself, generated in the build container
by tests/golden/make_golden.py and committed under tests/golden/*.npz
(tests/test_oracle_golden.py).

Every function cites the reference file:line it restates (paths relative to /root/reference).
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------
# integer artefacts (bit-exact contract)
# ----------------------------------------------------------------------------------------

def relative_position_index(ws):
    """model/MultiHeadAttentionRPE.py:373-387.  idx[i,j] for in-window tokens i=(ih,iw), j=(jh,jw):
    ((ih-jh)+ws-1)*(2ws-1) + (iw-jw)+ws-1, int64 (L,L)."""
    L = ws * ws
    i = torch.arange(L)
    ih, iw = i // ws, i % ws
    dh = ih[:, None] - ih[None, :] + ws - 1
    dw = iw[:, None] - iw[None, :] + ws - 1
    return (dh * (2 * ws - 1) + dw).to(torch.int64)


def pad_offsets(size, ws):
    """model/VidHRFormer_modules.py:538-550: centre zero-pad to a multiple of ws.
    returns (before, after)."""
    pad = math.ceil(size / ws) * ws - size
    return pad // 2, pad - pad // 2


def window_token_map(F_, H, W, ws):
    """model/VidHRFormer_modules.py:503-513 ("n (qh ph) (qw pw) c -> (ph pw) (n qh qw) c").
    Returns int64 (L, B) of flat token indices into a (F_, H, W) grid, H,W multiples of ws:
    map[l, b] with l = ph*ws+pw and b = (n*(H/ws) + qh)*(W/ws) + qw."""
    qh_n, qw_n = H // ws, W // ws
    l = torch.arange(ws * ws)
    ph, pw = l // ws, l % ws
    b = torch.arange(F_ * qh_n * qw_n)
    n = b // (qh_n * qw_n)
    qh = (b // qw_n) % qh_n
    qw = b % qw_n
    hh = qh[None, :] * ws + ph[:, None]
    ww = qw[None, :] * ws + pw[:, None]
    return (n[None, :] * H + hh) * W + ww


def causal_mask(T):
    """model/VidHRFormer_modules.py:78: triu(ones(T,T),1)==1 -> True where key j > query i."""
    i = torch.arange(T)
    return i[None, :] > i[:, None]


# ----------------------------------------------------------------------------------------
# positional encodings (utils/position_encoding.py)
# ----------------------------------------------------------------------------------------

def pos_embed_1d(L, E, temperature=10000.0):
    """utils/position_encoding.py:29-49; positions start at 1; returns (L,E)."""
    pos = torch.arange(1, L + 1, dtype=torch.float32)
    dim_t = torch.arange(E, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / E)
    p = pos[:, None] / dim_t
    out = torch.stack((p[:, 0::2].sin(), p[:, 1::2].cos()), dim=2).flatten(1)
    return out


def _sincos(emb, dim_t):
    p = emb[..., None] / dim_t
    return torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=-1).flatten(-2)


def pos_embed_2d(E, H, W, temperature=10000.0):
    """utils/position_encoding.py:67-93 -> returned channel-last (H,W,E): first E/2 from y, last from x."""
    y = torch.arange(1, H + 1, dtype=torch.float32)[:, None].expand(H, W)
    x = torch.arange(1, W + 1, dtype=torch.float32)[None, :].expand(H, W)
    dim_t = torch.arange(E // 2, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / (E // 2))
    return torch.cat((_sincos(y, dim_t), _sincos(x, dim_t)), dim=-1)


def pos_embed_3d(E, T, H, W, temperature=10000.0):
    """utils/position_encoding.py:117-158 -> channel-last (T,H,W,E): thirds t,y,x."""
    assert E % 3 == 0
    t = torch.arange(1, T + 1, dtype=torch.float32)[:, None, None].expand(T, H, W)
    y = torch.arange(1, H + 1, dtype=torch.float32)[None, :, None].expand(T, H, W)
    x = torch.arange(1, W + 1, dtype=torch.float32)[None, None, :].expand(T, H, W)
    dim_t = torch.arange(E // 3, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / (E // 3))
    return torch.cat((_sincos(t, dim_t), _sincos(y, dim_t), _sincos(x, dim_t)), dim=-1)


# ----------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------

def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def _ln(x, sd, name, eps=1e-5):
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def _pad2d(x, p, mode):
    """x NCHW."""
    if p == 0:
        return x
    if mode == "zero":
        return F.pad(x, (p, p, p, p))
    if mode == "reflect":
        return F.pad(x, (p, p, p, p), mode="reflect")
    if mode == "replicate":
        return F.pad(x, (p, p, p, p), mode="replicate")
    raise NotImplementedError("padding [%s] is not implemented" % mode)


def _bn_eval(x, sd, name, eps=1e-5):
    """BatchNorm2d with running statistics (stage 2 keeps Enc/Dec in .eval(), train_NAR.py:190-191)."""
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    m, v = sd[name + ".running_mean"], sd[name + ".running_var"]
    s = w / torch.sqrt(v + eps)
    return x * s[None, :, None, None] + (b - m * s)[None, :, None, None]


# ----------------------------------------------------------------------------------------
# ResNet encoder / decoder  (model/ResNetAutoEncoder.py, model/VPTR_modules.py:10-47)
# ----------------------------------------------------------------------------------------

def resnet_encoder(sd, x, n_downsampling=3, padding_type="reflect"):
    """VPTREnc.forward (VPTR_modules.py:16-29) -> ResnetEncoder (ResNetAutoEncoder.py:26-51), eval-mode BN.
    x: (N,T,Cimg,H,W) -> (N,T,feat,H/2^n,W/2^n).  sd keys 'encoder.model.<i>...'."""
    N, T = x.shape[:2]
    h = x.flatten(0, 1)
    p = "encoder.model."
    h = F.conv2d(_pad2d(h, 3, "reflect"), sd[p + "1.weight"])
    h = F.relu(_bn_eval(h, sd, p + "2"))
    idx = 4
    for _ in range(n_downsampling):
        h = F.conv2d(h, sd[p + "%d.weight" % idx], stride=2, padding=1)
        h = F.relu(_bn_eval(h, sd, p + "%d" % (idx + 1)))
        idx += 3
    # nine residual blocks (ResNetAutoEncoder.py:104-158); key indices depend on padding_type
    c1, b1, c2, b2 = (0, 1, 3, 4) if padding_type == "zero" else (1, 2, 5, 6)
    for blk in range(9):
        q = p + "%d.conv_block." % (idx + blk)
        r = F.conv2d(_pad2d(h, 1, padding_type), sd[q + "%d.weight" % c1])
        r = F.relu(_bn_eval(r, sd, q + "%d" % b1))
        r = F.conv2d(_pad2d(r, 1, padding_type), sd[q + "%d.weight" % c2])
        r = _bn_eval(r, sd, q + "%d" % b2)
        h = h + r
    h = F.relu(h)
    return h.reshape(N, T, *h.shape[1:])


def resnet_decoder(sd, feat, n_downsampling=3, out_layer="Tanh"):
    """VPTRDec.forward (VPTR_modules.py:36-47) -> ResnetDecoder (ResNetAutoEncoder.py:70-101), eval-mode BN."""
    N, T = feat.shape[:2]
    h = feat.flatten(0, 1)
    p = "decoder.model."
    idx = 0
    for _ in range(n_downsampling):
        h = F.conv_transpose2d(h, sd[p + "%d.weight" % idx], stride=2, padding=1, output_padding=1)
        h = F.relu(_bn_eval(h, sd, p + "%d" % (idx + 1)))
        idx += 3
    h = F.conv2d(_pad2d(h, 3, "reflect"), sd[p + "%d.weight" % (idx + 1)], sd[p + "%d.bias" % (idx + 1)])
    if out_layer == "Tanh":
        h = torch.tanh(h)
    elif out_layer == "Sigmoid":
        h = torch.sigmoid(h)
    else:
        raise ValueError("Unsupported output layer")
    return h.reshape(N, T, *h.shape[1:])


# ----------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------

def _mha_core(q, k, v, nhead, bias=None, mask=None):
    """q (B,Lq,C) already projected+scaled, k,v (B,Lk,C).  Per head: softmax(q k^T + bias, mask->-inf) v.
    MultiHeadAttentionRPE.py:586-590,623,635-686 / torch F.multi_head_attention_forward."""
    B, Lq, C = q.shape
    Lk = k.shape[1]
    d = C // nhead
    qh = q.reshape(B, Lq, nhead, d).permute(0, 2, 1, 3)
    kh = k.reshape(B, Lk, nhead, d).permute(0, 2, 1, 3)
    vh = v.reshape(B, Lk, nhead, d).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2)                      # (B,h,Lq,Lk)
    if bias is not None:
        s = s + bias[None]
    if mask is not None:
        s = s.masked_fill(mask[None, None], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = p @ vh                                         # (B,h,Lq,d)
    return o.permute(0, 2, 1, 3).reshape(B, Lq, C)


def window_attention(sd, xq, xv, ws, nhead, rpe=True, lw_pos=None):
    """SpatialLocalMultiheadAttention.forward (VidHRFormer_modules.py:321-357) with
    MultiheadAttentionRPE (MultiHeadAttentionRPE.py:527-697).  xq (query=key source) and xv
    (value source) are (F,H,W,C) channel-last.  sd keys 'attn....'.  Returns (F,H,W,C)."""
    Fr, H, W, C = xq.shape
    ph0, ph1 = pad_offsets(H, ws)
    pw0, pw1 = pad_offsets(W, ws)
    if ph0 + ph1 + pw0 + pw1 > 0:                      # centre zero-pad; padded tokens attend unmasked
        xq = F.pad(xq, (0, 0, pw0, pw1, ph0, ph1))
        xv = F.pad(xv, (0, 0, pw0, pw1, ph0, ph1))
    Hp, Wp = xq.shape[1:3]
    tmap = window_token_map(Fr, Hp, Wp, ws)            # (L,B)
    L, B = tmap.shape
    q_in = xq.reshape(-1, C)[tmap.t()]                 # (B,L,C)
    v_in = xv.reshape(-1, C)[tmap.t()]
    d = C // nhead
    if rpe:
        q = (q_in @ sd["attn.q_proj.weight"].t() + sd["attn.q_proj.bias"]) * (float(d) ** -0.5)
        k = q_in @ sd["attn.k_proj.weight"].t() + sd["attn.k_proj.bias"]
        v = v_in @ sd["attn.v_proj.weight"].t() + sd["attn.v_proj.bias"]
        idx = relative_position_index(ws)
        bias = sd["attn.relative_position_bias_table"][idx.reshape(-1)].reshape(L, L, nhead).permute(2, 0, 1)
        o = _mha_core(q, k, v, nhead, bias=bias)
    else:                                              # VidHRFormer_modules.py:341, plain nn.MultiheadAttention
        q_in = q_in + lw_pos.reshape(L, C)[None]
        Wi, bi = sd["attn.in_proj_weight"], sd["attn.in_proj_bias"]
        q = (q_in @ Wi[:C].t() + bi[:C]) * (float(d) ** -0.5)
        k = q_in @ Wi[C:2 * C].t() + bi[C:2 * C]
        v = v_in @ Wi[2 * C:].t() + bi[2 * C:]
        o = _mha_core(q, k, v, nhead)
    o = o @ sd["attn.out_proj.weight"].t() + sd["attn.out_proj.bias"]
    out = torch.zeros(Fr * Hp * Wp, C, dtype=o.dtype)
    out = out.index_put((tmap.t().reshape(-1),), o.reshape(-1, C))
    out = out.reshape(Fr, Hp, Wp, C)
    return out[:, ph0:ph0 + H, pw0:pw0 + W, :]


def temporal_attention(sd, q_in, k_in, v_in, nhead, causal=False):
    """nn.MultiheadAttention call sites VidHRFormer_modules.py:79-84,185-187,204-205 (packed in_proj).
    q_in (N,Tq,H,W,C); k_in, v_in (N,Tk,H,W,C): one sequence per (n,h,w) pixel.  Returns (N,Tq,H,W,C)."""
    N, Tq, H, W, C = q_in.shape
    Tk = k_in.shape[1]
    d = C // nhead
    seq = lambda t: t.permute(0, 2, 3, 1, 4).reshape(N * H * W, t.shape[1], C)
    Wi, bi = sd["in_proj_weight"], sd["in_proj_bias"]
    q = (seq(q_in) @ Wi[:C].t() + bi[:C]) * (float(d) ** -0.5)
    k = seq(k_in) @ Wi[C:2 * C].t() + bi[C:2 * C]
    v = seq(v_in) @ Wi[2 * C:].t() + bi[2 * C:]
    mask = causal_mask(Tq) if causal else None
    o = _mha_core(q, k, v, nhead, mask=mask)
    o = o @ sd["out_proj.weight"].t() + sd["out_proj.bias"]
    return o.reshape(N, H, W, Tq, C).permute(0, 3, 1, 2, 4)


# ----------------------------------------------------------------------------------------
# conv feed-forward (MlpDWBN, VidHRFormer_modules.py:376-442)
# ----------------------------------------------------------------------------------------

def _ffn_norm(u, sd, name, layer_norm, training, bn_updates, eps=1e-5):
    """u (F,H,W,ch) channel-last.  layer_norm: LayerNorm((ch,H,W)) per frame with affine (ch,H,W)
    (VidHRFormer_modules.py:397-398); else BatchNorm2d(ch) (batch stats if training, :399-400)."""
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    if layer_norm:
        mu = u.mean(dim=(1, 2, 3), keepdim=True)
        var = ((u - mu) ** 2).mean(dim=(1, 2, 3), keepdim=True)
        return (u - mu) / torch.sqrt(var + eps) * w.permute(1, 2, 0)[None] + b.permute(1, 2, 0)[None]
    if training:
        mu = u.mean(dim=(0, 1, 2))
        var = ((u - mu) ** 2).mean(dim=(0, 1, 2))
        if bn_updates is not None:                    # momentum 0.1, unbiased variance (torch BatchNorm2d)
            n = u.numel() // u.shape[-1]
            bn_updates[name] = (mu.detach(), (var * n / max(n - 1, 1)).detach())
    else:
        mu, var = sd[name + ".running_mean"], sd[name + ".running_var"]
    return (u - mu) / torch.sqrt(var + eps) * w + b


def mlp_dwbn(sd, x, layer_norm, training=False, bn_updates=None, prefix=""):
    """MlpDWBN.forward (VidHRFormer_modules.py:424-442), x (F,H,W,C) channel-last; dropout = identity."""
    u = x @ sd["fc1.weight"].flatten(1).t() + sd["fc1.bias"]
    u = _gelu(_ffn_norm(u, sd, "norm1", layer_norm, training, bn_updates))
    ch = u.shape[-1]
    u = F.conv2d(u.permute(0, 3, 1, 2), sd["dw3x3.weight"], sd["dw3x3.bias"], padding=1, groups=ch).permute(0, 2, 3, 1)
    u = _gelu(_ffn_norm(u, sd, "norm2", layer_norm, training, bn_updates))
    u = u @ sd["fc2.weight"].flatten(1).t() + sd["fc2.bias"]
    u = _gelu(_ffn_norm(u, sd, "norm3", layer_norm, training, bn_updates))
    return u


# ----------------------------------------------------------------------------------------
# transformer blocks
# ----------------------------------------------------------------------------------------

def enc_block(sd, x, temporal_pos, ws, nhead, far, rpe=True, lw_pos=None, training=False, bn_updates=None):
    """VidHRFormerBlockEnc.forward (VidHRFormer_modules.py:60-93).  x (N,T,H,W,C); dropout/DropPath identity."""
    N, T, H, W, C = x.shape
    a = _ln(x, sd, "norm1").flatten(0, 1)
    x = x + window_attention(_sub(sd, "SLMHSA."), a, a, ws, nhead, rpe, lw_pos).reshape(N, T, H, W, C)
    bu = {} if bn_updates is not None else None
    x = x + mlp_dwbn(_sub(sd, "SpatialFFN."), _ln(x, sd, "norm2").flatten(0, 1), far, training, bu).reshape(N, T, H, W, C)
    if bn_updates is not None:
        bn_updates.update({"SpatialFFN." + k: v for k, v in bu.items()})
    z = _ln(x, sd, "norm3")
    zp = z + temporal_pos[None, :T, None, None, :]
    x = x + temporal_attention(_sub(sd, "temporal_MHSA."), zp, zp, z, nhead, causal=far)
    y = _ln(x, sd, "norm4")
    y = _gelu(y @ sd["linear1.weight"].t() + sd["linear1.bias"]) @ sd["linear2.weight"].t() + sd["linear2.bias"]
    return x + y


def tslma_attention(sd, memory, query, Tlw_pos, ws, nhead):
    """TemporalSpatialLocalMultiheadAttention.forward (VidHRFormer_modules.py:246-284) with TemporalLocalPermuteModule (:444-484):
    per spatial window, the T2*ws^2 query tokens attend the T1*ws^2 memory tokens; Tlw_pos (T1+T2, ws, ws, C) is added to the
    key (first T1 entries) and query (next T2) after the window permutation.  memory (N,T1,H,W,C), query (N,T2,H,W,C);
    grids that are not multiples of the window are centre zero-padded first (PadBlock, :527-561)."""
    N, T1, H, W, C = memory.shape
    T2 = query.shape[1]
    (t0, t1), (l0, l1) = pad_offsets(H, ws), pad_offsets(W, ws)
    mp = F.pad(memory, (0, 0, l0, l1, t0, t1))
    qp = F.pad(query, (0, 0, l0, l1, t0, t1))
    Hp, Wp = mp.shape[2], mp.shape[3]

    def perm(x):      # "n t (qh ph) (qw pw) c -> (n qh qw) (t ph pw) c"  (batch-first form of the reference's seq-first permute)
        n, t = x.shape[:2]
        x = x.reshape(n, t, Hp // ws, ws, Wp // ws, ws, C).permute(0, 2, 4, 1, 3, 5, 6)
        return x.reshape(n * (Hp // ws) * (Wp // ws), t * ws * ws, C)
    km, qm = perm(mp), perm(qp)
    q_in = qm + Tlw_pos[T1:T1 + T2].reshape(1, T2 * ws * ws, C)
    k_in = km + Tlw_pos[:T1].reshape(1, T1 * ws * ws, C)
    Wi, bi = sd["attn.in_proj_weight"], sd["attn.in_proj_bias"]
    d = C // nhead
    q = (q_in @ Wi[:C].t() + bi[:C]) * (d ** -0.5)
    k = k_in @ Wi[C:2 * C].t() + bi[C:2 * C]
    v = km @ Wi[2 * C:].t() + bi[2 * C:]
    o = _mha_core(q, k, v, nhead) @ sd["attn.out_proj.weight"].t() + sd["attn.out_proj.bias"]
    o = o.reshape(N, Hp // ws, Wp // ws, T2, ws, ws, C).permute(0, 3, 1, 4, 2, 5, 6).reshape(N, T2, Hp, Wp, C)
    return o[:, :, t0:t0 + H, l0:l0 + W]


def dec_block(sd, tgt, query_pos, memory, pos_future, pos_past, ws, nhead, rpe=True, lw_pos=None, Tlw_pos=None):
    """VidHRFormerBlockDecNAR.forward (VidHRFormer_modules.py:164-211); the encoder-decoder attention is the per-pixel temporal
    cross-attention (:200-206) or, when the block carries TSLMA.* parameters (TSLMA_flag=True, :194-198), tslma_attention.
    tgt, query_pos (N,T2,H,W,C); memory (N,T1,H,W,C).  MlpDWBN here is the LayerNorm flavour (:136,159,390)."""
    N, T2, H, W, C = tgt.shape
    a = _ln(tgt, sd, "norm1")
    x = tgt + window_attention(_sub(sd, "SLMHSA."), (a + query_pos).flatten(0, 1), a.flatten(0, 1), ws, nhead, rpe, lw_pos).reshape(N, T2, H, W, C)
    x = x + mlp_dwbn(_sub(sd, "SpatialFFN."), _ln(x, sd, "norm2").flatten(0, 1), True).reshape(N, T2, H, W, C)
    z = _ln(x, sd, "norm3")
    zp = z + pos_future[None, :, None, None, :]
    x = x + temporal_attention(_sub(sd, "temporal_MHSA."), zp, zp, z, nhead)
    y = _ln(x, sd, "norm4")
    y = _gelu(y @ sd["linear1.weight"].t() + sd["linear1.bias"]) @ sd["linear2.weight"].t() + sd["linear2.bias"]
    x = x + y
    z = _ln(x, sd, "norm5")
    if any(k.startswith("TSLMA.") for k in sd):
        x = x + tslma_attention(_sub(sd, "TSLMA."), memory, z + query_pos, Tlw_pos, ws, nhead)
    else:
        q = z + query_pos + pos_future[None, :, None, None, :]
        k = memory + pos_past[None, :, None, None, :]
        x = x + temporal_attention(_sub(sd, "EncDecAttn."), q, k, memory, nhead)
    x = x + mlp_dwbn(_sub(sd, "SpatialFFN1."), _ln(x, sd, "norm6").flatten(0, 1), True).reshape(N, T2, H, W, C)
    return x


def _num_layers(sd, prefix):
    n = 0
    while any(k.startswith("%s%d." % (prefix, n)) for k in sd):
        n += 1
    return n


def vptr_former_far(sd, feats, nhead=8, ws=4, rpe=True, training=False):
    """VPTRFormerFAR.forward (VPTR_modules.py:185-197) -> VidHRFormerFAR.forward (VidHRFormer.py:71-88).
    feats (N,T,C,H,W) -> (N,T,C,H,W)."""
    x = feats.permute(0, 1, 3, 4, 2)
    T = x.shape[1]
    tp = sd["temporal_pos"][:T]
    p = "transformer.encoder.layers."
    for i in range(_num_layers(sd, p)):
        x = enc_block(_sub(sd, "%s%d." % (p, i)), x, tp, ws, nhead, far=True, rpe=rpe, lw_pos=sd["lw_pos"], training=training)
    x = _ln(x, sd, "transformer.encoder.norm")
    return F.relu(x.permute(0, 1, 4, 2, 3))


def vptr_former_nar(sd, past_feats, nhead=8, ws=4, rpe=True, training=False, bn_updates=None):
    """VPTRFormerNAR.forward (VPTR_modules.py:140-147) -> VidHRFormerNAR.forward (VidHRFormer.py:28-53).
    past_feats (N,Tp,C,H,W) -> (N,Tf,C,H,W).  training=True uses batch statistics in the encoder's
    BatchNorm2d and records the running-stat updates in bn_updates."""
    x = past_feats.permute(0, 1, 3, 4, 2)
    N, Tp = x.shape[:2]
    tpos = sd["temporal_pos"]
    p = "transformer.encoder.layers."
    for i in range(_num_layers(sd, p)):
        bu = {} if bn_updates is not None else None
        x = enc_block(_sub(sd, "%s%d." % (p, i)), x, tpos[:Tp], ws, nhead, far=False, rpe=rpe, lw_pos=sd["lw_pos"],
                      training=training, bn_updates=bu)
        if bn_updates is not None:
            bn_updates.update({"%s%d.%s" % (p, i, k): v for k, v in bu.items()})
    memory = _ln(x, sd, "transformer.encoder.norm")
    qp = sd["frame_queries"][None].expand(N, -1, -1, -1, -1)
    tgt = torch.zeros_like(qp)
    p = "transformer.decoder.layers."
    for i in range(_num_layers(sd, p)):
        tgt = dec_block(_sub(sd, "%s%d." % (p, i)), tgt, qp, memory, tpos[Tp:], tpos[:Tp], ws, nhead, rpe=rpe, lw_pos=sd["lw_pos"],
                        Tlw_pos=sd.get("Tlw_pos"))
    out = _ln(tgt, sd, "transformer.decoder.norm")
    return F.relu(out.permute(0, 1, 4, 2, 3))


# ----------------------------------------------------------------------------------------
# losses used by the timed step (model/criterion.py) -- restated for the step-level checks
# ----------------------------------------------------------------------------------------

def mse_loss(gt, pred):
    """criterion.py:105-132 (no temporal weight)."""
    return ((pred - gt) ** 2).mean()


def gdl_loss(gt, pred):
    """criterion.py:134-204, alpha = 1."""
    g, p = gt.flatten(0, -4), pred.flatten(0, -4)
    t1 = (g[:, :, 1:, :] - g[:, :, :-1, :]).abs()
    t2 = (p[:, :, 1:, :] - p[:, :, :-1, :]).abs()
    t3 = (g[:, :, :, :-1] - g[:, :, :, 1:]).abs()
    t4 = (p[:, :, :, :-1] - p[:, :, :, 1:]).abs()
    return (t1 - t2).abs().mean() + (t3 - t4).abs().mean()
