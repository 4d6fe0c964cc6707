"""ResNet conv encoder / transposed-conv decoder (reference model/ResNetAutoEncoder.py:8-158) on channel-last
activations.  The nn.Sequential layout (and therefore every state_dict key, including the padding_type-dependent
indices of the ResnetBlocks) is the reference's; compute is libvptr_b200.so:
  7x7 stem / head      -> direct kernels (csrc/conv.cu)
  3x3 (stride 1 | 2)   -> im2col gather + tcgen05 TF32 GEMM with the eval-BatchNorm folded into weights (scale) and
                          epilogue (shift, ReLU, residual)
  ConvTranspose2d      -> GEMM to per-tap columns + output gather (shift + ReLU); its input gradient is the stride-2
                          im2col of the ReLU-masked output gradient times the same packed weight.
BatchNorm runs with running statistics (stage 2 keeps both modules in .eval(), train_NAR.py:190-191)."""
import functools
import os

import torch
import torch.nn as nn
from torch.nn import init

from .. import engine as E
from .. import ops


class ResnetBlock(nn.Module):
    """conv_block = [pad?, Conv3x3, BN, ReLU, pad?, Conv3x3, BN]; out = x + conv_block(x) (reference :104-158)."""

    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        super().__init__()
        self.padding_type = padding_type
        block = []
        p = 0
        for half in range(2):
            if padding_type == 'reflect':
                block += [nn.ReflectionPad2d(1)]
            elif padding_type == 'replicate':
                block += [nn.ReplicationPad2d(1)]
            elif padding_type == 'zero':
                p = 1
            else:
                raise NotImplementedError('padding [%s] is not implemented' % padding_type)
            block += [nn.Conv2d(dim, dim, kernel_size=3, padding=p, bias=use_bias), norm_layer(dim)]
            if half == 0:
                block += [nn.ReLU(True)]
                if use_dropout:
                    block += [nn.Dropout(0.5)]
        self.conv_block = nn.Sequential(*block)

    def convs(self):
        mods = [m for m in self.conv_block if isinstance(m, (nn.Conv2d, nn.BatchNorm2d))]
        return mods[0], mods[1], mods[2], mods[3]


def _use_bias(norm_layer):
    inst = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
    return inst == nn.InstanceNorm2d


class ResnetEncoder(nn.Module):
    def __init__(self, input_nc, ngf=64, out_dim=528, n_downsampling=2, norm_layer=nn.BatchNorm2d, use_dropout=False,
                 padding_type='reflect'):
        super().__init__()
        use_bias = _use_bias(norm_layer)
        self.padding_type, self.n_downsampling = padding_type, n_downsampling
        model = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, bias=use_bias), norm_layer(ngf), nn.ReLU(True)]
        ch = ngf
        for i in range(n_downsampling):
            nxt = out_dim if i == n_downsampling - 1 else ch * 2
            model += [nn.Conv2d(ch, nxt, kernel_size=3, stride=2, padding=1, bias=use_bias), norm_layer(nxt), nn.ReLU(True)]
            ch = nxt
        for _ in range(9):
            model += [ResnetBlock(out_dim, padding_type=padding_type, norm_layer=norm_layer, use_dropout=use_dropout, use_bias=use_bias)]
        model += [nn.ReLU()]
        self.model = nn.Sequential(*model)


class ResnetDecoder(nn.Module):
    def __init__(self, output_nc, ngf=64, feat_dim=528, n_downsampling=2, norm_layer=nn.BatchNorm2d, use_dropout=False,
                 padding_type='reflect', out_layer='Tanh'):
        super().__init__()
        use_bias = _use_bias(norm_layer)
        self.n_downsampling, self.out_layer = n_downsampling, out_layer
        model = []
        ch = feat_dim
        for i in range(n_downsampling):
            nxt = ngf * 2 ** (n_downsampling - 1 - i)
            model += [nn.ConvTranspose2d(ch, nxt, kernel_size=3, stride=2, padding=1, output_padding=1, bias=use_bias), norm_layer(nxt),
                      nn.ReLU(True)]
            ch = nxt
        model += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0)]
        if out_layer == 'Tanh':
            model += [nn.Tanh()]
        elif out_layer == 'Sigmoid':
            model += [nn.Sigmoid()]
        else:
            raise ValueError("Unsupported output layer")
        self.model = nn.Sequential(*model)


def init_weights(net, init_type='normal', init_gain=0.02):
    """Same policy as the reference's init_weights (:160-189): Conv/Linear weights ~ init_type, biases 0,
    BatchNorm2d weight ~ N(1, gain), bias 0; matched on class names."""
    def init_func(m):
        cls = m.__class__.__name__
        if hasattr(m, 'weight') and ('Conv' in cls or 'Linear' in cls):
            if init_type == 'normal':
                init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == 'xavier':
                init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == 'kaiming':
                init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
            elif init_type == 'orthogonal':
                init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
            if getattr(m, 'bias', None) is not None:
                init.constant_(m.bias.data, 0.0)
        elif 'BatchNorm2d' in cls:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)

    print('initialize network with %s' % init_type)
    net.apply(init_func)
    clear_packed_weights(net)      # the writes above go through .data and bypass the version counters _packed keys on


# ----------------------------------------------------------------------------------------------- functional compute
_NO_WEIGHT_CACHE = os.environ.get("VPTR_NO_WEIGHT_CACHE", "") not in ("", "0")


def _packed(mod, tag, norm, tensors, build):
    """Kernel-layout copy of a frozen layer's weights (BatchNorm folded in, tap-major packing, tf32 hi/lo planes), kept on the
    module across calls.  Stage 2 runs the autoencoder in eval mode with constant weights (train_NAR.py:190-191), so packing them
    again every step only costs launches.  The copy is rebuilt whenever a source tensor is replaced or modified in place
    (data_ptr / autograd version counter of the weights and BatchNorm statistics: optimizer steps, load_state_dict, .to()),
    when the precision mode changes, and always while the norm layer is in training mode.  Writes through `.data` bypass the
    version counter: call `vptr_b200.model.clear_packed_weights(module)` after those (or set VPTR_NO_WEIGHT_CACHE=1)."""
    if _NO_WEIGHT_CACHE or (norm is not None and norm.training):
        return build()
    key = (tag, E.ROUND_TF32, ops.FORCE_SIMT) + tuple((t.data_ptr(), t._version) for t in tensors)
    ent = mod.__dict__.get("_vptr_packed")
    if ent is not None and ent[0] == key:
        return ent[1]
    val = build()
    mod.__dict__["_vptr_packed"] = (key, val)
    return val


def clear_packed_weights(module):
    """Drop the cached kernel-layout weights of every layer under `module` (see _packed)."""
    for m in module.modules():
        m.__dict__.pop("_vptr_packed", None)


def _bn_tensors(bn):
    return (bn.weight, bn.bias, bn.running_mean, bn.running_var)


def _conv3x3(x, F_, H, W, conv, bn, stride, pad_mode, relu, residual=None, act_after_residual=False):
    """x (F*H*W, Cin) channel-last -> (F*Ho*Wo, Cout).  y = act(conv(x)*scale + shift) (+ residual)."""
    Cout, Cin = conv.weight.shape[:2]
    implicit = stride == 1 and not ops.FORCE_SIMT and ops.conv3x3_implicit_ok(H, W) and Cin % 4 == 0 and Cout % 4 == 0

    bf16x3 = implicit and ops.conv3x3_bf16x3_ok(H, W, Cin)

    def build():
        scale, shift = ops.bn_fold(bn)
        wraw = ops.pack_conv_weight(conv.weight.data, scale, 0).view(Cout, 9 * Cin)
        if bf16x3:
            return ops.split_bf16x2(wraw), shift
        if implicit:
            return (ops.split_tf32(wraw) if E.ROUND_TF32 else wraw), shift
        return E._rc(wraw), shift

    wk, shift = _packed(conv, "conv3x3-bf16x3" if bf16x3 else ("conv3x3-implicit" if implicit else "conv3x3-gemm"), bn,
                        (conv.weight,) + _bn_tensors(bn), build)
    if bf16x3:
        # raw-tile implicit GEMM with both operands as two bf16 planes and three bf16 tensor-core passes: ~2^-16 per product (the
        # two-plane TF32 form below rounds the activations to tf32) at 3/4 of its tensor-pipe time
        xq2 = ops.pad_nhwc_quad_bf16x2(x, F_, H, W, Cin, pad_mode)
        y = ops.conv3x3_bf16x3(xq2, wk, F_, H, W, Cin, Cout, bias=shift, residual=residual, act=ops.ACT_RELU if relu else ops.ACT_NONE)
        if residual is not None and act_after_residual:
            y = ops.relu_fwd(y, out=y)
        return y, H, W
    if implicit:
        # implicit GEMM: 4-D TMA boxes of the padded activation feed the tcgen05 kernel directly (no im2col matrix)
        # The frozen encoder chains 21 convolutions; with plain tf32 weights its features land at 1.16e-3 relative (just outside
        # the 1e-3 gate), so the weights enter as two tf32 planes [hi|lo] (2x MMA work, weight rounding error removed).
        w2 = wk
        if ops.conv3x3_quad_ok(H, W):      # 16x16 (and larger) grids: quadrant-tiled padded copy -> the 8x8 raw-tile kernel
            xq = ops.pad_nhwc_quad(x, F_, H, W, Cin, pad_mode, round_tf32=E.ROUND_TF32)
            y = ops.conv3x3_tf32_quad(xq, w2, F_, H, W, Cin, Cout, bias=shift, residual=residual, act=ops.ACT_RELU if relu else ops.ACT_NONE,
                                      w_planes=2 if E.ROUND_TF32 else 1)
        else:
            xpad = ops.pad_nhwc(x, F_, H, W, Cin, 1, pad_mode, round_tf32=E.ROUND_TF32)
            y = ops.conv3x3_tf32(xpad, w2, F_, H, W, Cin, Cout, bias=shift, residual=residual, act=ops.ACT_RELU if relu else ops.ACT_NONE,
                                 w_planes=2 if E.ROUND_TF32 else 1)
        if residual is not None and act_after_residual:
            y = ops.relu_fwd(y, out=y)
        return y, H, W
    wpk = wk
    col, Ho, Wo = ops.im2col(x, F_, H, W, Cin, 3, stride, 1, pad_mode, round_tf32=E.ROUND_TF32)
    if residual is not None and act_after_residual:
        y = ops.gemm(col, wpk, bias=shift, residual=residual)
        y = ops.relu_fwd(y, out=y)
    else:
        y = ops.gemm(col, wpk, bias=shift, act=ops.ACT_RELU if relu else ops.ACT_NONE, residual=residual)
    return y, Ho, Wo


def encoder_forward(enc, x):
    """x (F, Cimg, H, W) NCHW frames -> ((F*h*w, C) channel-last features, h, w).  Reference :26-51."""
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError("vptr_b200.VPTREnc: input must be a CUDA float32 tensor (got %s, %s); there is no CPU fallback" % (x.device, x.dtype))
    x = x.contiguous()
    F_, Ci, H, W = x.shape
    m = enc.model
    def build_stem():
        scale, shift = ops.bn_fold(m[2])
        return ops.pack_conv_weight(m[1].weight.data, scale, 2), shift

    wpk, shift = _packed(m[1], "stem7x7", m[2], (m[1].weight,) + _bn_tensors(m[2]), build_stem)
    h = ops.stem_conv7x7(x, wpk, shift, F_, Ci, H, W, m[1].weight.shape[0])
    idx = 4
    for _ in range(enc.n_downsampling):
        h, H, W = _conv3x3(h, F_, H, W, m[idx], m[idx + 1], 2, 0, True)
        idx += 3
    pad_mode = ops.PAD_MODES[enc.padding_type]
    for b in range(9):
        c1, n1, c2, n2 = m[idx + b].convs()
        r, _, _ = _conv3x3(h, F_, H, W, c1, n1, 1, pad_mode, True)
        h, _, _ = _conv3x3(r, F_, H, W, c2, n2, 1, pad_mode, False, residual=h, act_after_residual=(b == 8))   # final nn.ReLU of :48
    return h, H, W


_ACT = {"Tanh": 1, "Sigmoid": 2}


def decoder_forward(dec, feat, F_, H, W, save):
    """feat (F*H*W, C) channel-last -> frames (F, Cimg, 2^n H, 2^n W) NCHW.  Reference :70-101."""
    m = dec.model
    h = feat
    saved = []
    idx = 0
    for _ in range(dec.n_downsampling):
        convT, bn = m[idx], m[idx + 1]
        Cin, Cout = convT.weight.shape[:2]
        def build_up(convT=convT, bn=bn, Cin=Cin, Cout=Cout):
            scale, shift = ops.bn_fold(bn)
            return E._rc(ops.pack_conv_weight(convT.weight.data, scale, 1)).view(9 * Cout, Cin), shift

        wpk, shift = _packed(convT, "convT3x3", bn, (convT.weight,) + _bn_tensors(bn), build_up)
        col = ops.gemm(E._rc(h), wpk)      # operands of the tf32 GEMM are pre-rounded (engine.ROUND_TF32)
        y = ops.convT_gather(col, shift, F_, H, W, Cout, relu=True)
        if save:
            saved.append((wpk, y, H, W, Cin, Cout))
        h, H, W = y, 2 * H, 2 * W
        idx += 3
    head = m[idx + 1]
    Co, Ci = head.weight.shape[:2]
    wpk = _packed(head, "head7x7", None, (head.weight,), lambda: ops.pack_conv_weight(head.weight.data, None, 3))
    act = _ACT[dec.out_layer]
    out = ops.head_conv7x7_fwd(h, wpk, head.bias.data, F_, Ci, Co, H, W, act)
    return out, ((saved, out, head, act, F_, H, W) if save else None)


def decoder_backward(dec, saved_all, dout):
    saved, out, head, act, F_, H, W = saved_all
    Co, Ci = head.weight.shape[:2]
    d = ops.head_conv7x7_bwd(dout, out, head.weight.data, F_, Ci, Co, H, W, act)     # (F*H*W, 64) gradient at the last ReLU output
    for wpk, y, h_in, w_in, Cin, Cout in reversed(saved):
        # y = relu(convT(x)*scale + shift): dx = conv_s2(d * (y > 0)) with the same packed weight
        col, _, _ = ops.im2col(d, F_, 2 * h_in, 2 * w_in, Cout, 3, 2, 1, 0, mask=y, round_tf32=E.ROUND_TF32)
        d = ops.gemm(col, wpk, b_mn=True)                                             # (F*h*w, Cin)
    return d


# =============================================================================================== stage-1 training (train-mode BatchNorm)
# train_AutoEncoder.py:44-86 trains VPTREnc / VPTRDec with BatchNorm2d in TRAIN mode (batch statistics, running-stat updates) and needs
# every weight gradient.  Same kernels as above for the contractions -- im2col + tcgen05 GEMM (forward and weight
# gradient), GEMM + col2im scatter (input gradient), ConvTranspose2d as GEMM + gather -- with vptr_bn_stats / vptr_bn_act_{fwd,bwd}
# between them instead of the folded eval-mode scale / shift.  Nothing is cached: the weights change every step.
def _bn_train_fwd(y, bn, act, res=None):
    mean, rstd = ops.bn_stats(y, bn.running_mean, bn.running_var, momentum=bn.momentum if bn.momentum is not None else 0.1, eps=bn.eps)
    bn.num_batches_tracked.add_(1)
    z = ops.bn_act_fwd(y, mean, rstd, bn.weight.data, bn.bias.data, act, res)
    return z, mean, rstd


def _acc(grads, p, g):
    grads[p] = g if p not in grads else grads[p] + g


def _conv_train_fwd(x, F_, H, W, conv, bn, stride, pad_mode, act, res, tape):
    Cout, Cin = conv.weight.shape[:2]
    col, Ho, Wo = ops.im2col(x, F_, H, W, Cin, 3, stride, 1, pad_mode, round_tf32=E.ROUND_TF32)
    wpk = E._rc(ops.pack_conv_weight(conv.weight.data, None, 0).view(Cout, 9 * Cin))
    y = ops.gemm(col, wpk)
    del col
    z, mean, rstd = _bn_train_fwd(y, bn, act, res)
    if tape is not None:
        tape.append(dict(kind="conv", x=x, y=y, z=z, mean=mean, rstd=rstd, wpk=wpk, conv=conv, bn=bn, act=act, has_res=res is not None,
                         geom=(F_, H, W, Cin, Cout, stride, pad_mode)))
    return z, Ho, Wo


def _conv_train_bwd(s, dz, grads):
    """-> (dx, gradient of the residual operand or None)"""
    F_, H, W, Cin, Cout, stride, pad_mode = s["geom"]
    conv, bn = s["conv"], s["bn"]
    dg, db = torch.zeros_like(bn.weight.data), torch.zeros_like(bn.bias.data)
    g0, dy = ops.bn_act_bwd(dz, s["y"], s["z"], s["mean"], s["rstd"], bn.weight.data, dg, db, s["act"], round_tf32=E.ROUND_TF32)
    _acc(grads, bn.weight, dg)
    _acc(grads, bn.bias, db)
    col, _, _ = ops.im2col(s["x"], F_, H, W, Cin, 3, stride, 1, pad_mode, round_tf32=E.ROUND_TF32)     # recomputed: 9x the activation is not kept
    dW = torch.zeros(Cout, 9 * Cin, dtype=torch.float32, device=dy.device)
    ops.gemm(dy, col, out=dW, a_mn=True, b_mn=True, accumulate=True)
    del col
    _acc(grads, conv.weight, dW.view(Cout, 3, 3, Cin).permute(0, 3, 1, 2).contiguous())
    dcol = ops.gemm(dy, s["wpk"], b_mn=True)
    dx = ops.col2im(dcol, F_, H, W, Cin, 3, stride, 1, pad_mode)
    dres = None
    if s["has_res"]:
        dres = g0 if s["act"] == 2 else dz
    return dx, dres


def encoder_forward_train(enc, x, tape):
    """train-mode forward of ResnetEncoder (reference :26-51); tape: list that receives what the backward needs, or None"""
    x = x.contiguous()
    F_, Ci, H, W = x.shape
    m = enc.model
    conv, bn = m[1], m[2]
    wpk = ops.pack_conv_weight(conv.weight.data, None, 2)                       # [(kh,kw,ci)][64]
    y = ops.stem_conv7x7_raw(x, wpk, F_, Ci, H, W, conv.weight.shape[0])
    h, mean, rstd = _bn_train_fwd(y, bn, 1)
    if tape is not None:
        tape.append(dict(kind="stem", x=x, y=y, z=h, mean=mean, rstd=rstd, conv=conv, bn=bn, geom=(F_, Ci, H, W)))
    idx = 4
    for _ in range(enc.n_downsampling):
        h, H, W = _conv_train_fwd(h, F_, H, W, m[idx], m[idx + 1], 2, 0, 1, None, tape)
        idx += 3
    pad_mode = ops.PAD_MODES[enc.padding_type]
    for b in range(9):
        c1, n1, c2, n2 = m[idx + b].convs()
        if tape is not None:
            tape.append(dict(kind="block_in"))
        r, _, _ = _conv_train_fwd(h, F_, H, W, c1, n1, 1, pad_mode, 1, None, tape)
        h, _, _ = _conv_train_fwd(r, F_, H, W, c2, n2, 1, pad_mode, 2 if b == 8 else 0, h, tape)
    return h, H, W


def encoder_backward_train(tape, dfeat):
    """-> {parameter: gradient}"""
    grads = {}
    d = dfeat
    pending = None                                        # gradient of the current block's residual operand
    while tape:
        s = tape.pop()
        if s["kind"] == "conv":
            d, dres = _conv_train_bwd(s, d, grads)
            if dres is not None:
                pending = dres
        elif s["kind"] == "block_in":
            d = ops.axpby(d, pending)                     # block input feeds conv1 and the residual add
            pending = None
        else:                                             # stem: BatchNorm backward + weight gradient (no input gradient: x is the image)
            F_, Ci, H, W = s["geom"]
            bn, conv = s["bn"], s["conv"]
            dg, db = torch.zeros_like(bn.weight.data), torch.zeros_like(bn.bias.data)
            _, dy = ops.bn_act_bwd(d, s["y"], s["z"], s["mean"], s["rstd"], bn.weight.data, dg, db, 1)
            _acc(grads, bn.weight, dg)
            _acc(grads, bn.bias, db)
            dw = ops.stem_wgrad(s["x"], dy, F_, Ci, H, W)                        # [(kh,kw,ci)][co]
            _acc(grads, conv.weight, dw.view(7, 7, Ci, -1).permute(3, 2, 0, 1).contiguous())
    return grads


def decoder_forward_train(dec, feat, F_, H, W, tape):
    m = dec.model
    h = feat
    idx = 0
    for _ in range(dec.n_downsampling):
        convT, bn = m[idx], m[idx + 1]
        Cin, Cout = convT.weight.shape[:2]
        wpk = E._rc(ops.pack_conv_weight(convT.weight.data, None, 1)).view(9 * Cout, Cin)
        xr = E._rc(h)
        col = ops.gemm(xr, wpk)
        y = ops.convT_gather(col, torch.zeros(Cout, dtype=torch.float32, device=h.device), F_, H, W, Cout, relu=False)
        del col
        z, mean, rstd = _bn_train_fwd(y, bn, 1)
        if tape is not None:
            tape.append(dict(kind="up", x=xr, y=y, z=z, mean=mean, rstd=rstd, wpk=wpk, convT=convT, bn=bn, geom=(F_, H, W, Cin, Cout)))
        h, H, W = z, 2 * H, 2 * W
        idx += 3
    head = m[idx + 1]
    Co, Ci = head.weight.shape[:2]
    act = _ACT[dec.out_layer]
    out = ops.head_conv7x7_fwd(h, ops.pack_conv_weight(head.weight.data, None, 3), head.bias.data, F_, Ci, Co, H, W, act)
    if tape is not None:
        tape.append(dict(kind="head", x=h, out=out, head=head, act=act, geom=(F_, H, W, Ci, Co)))
    return out


def decoder_backward_train(tape, dout):
    """-> (gradient of the input features (F*h*w, C), {parameter: gradient})"""
    grads = {}
    s = tape.pop()
    F_, H, W, Ci, Co = s["geom"]
    head, out, act, x = s["head"], s["out"], s["act"], s["x"]
    dout = dout.contiguous()
    d = ops.head_conv7x7_bwd(dout, out, head.weight.data, F_, Ci, Co, H, W, act)
    dpre = ops.act_bwd(dout, out, act)                                            # (F, Co, H, W)
    dpre_cl = dpre.view(F_ * H * W, 1) if Co == 1 else ops.transpose(dpre, F_, Co, H * W).view(F_ * H * W, Co)
    gb = torch.zeros(Co, dtype=torch.float32, device=d.device)
    ops.colsum(dpre_cl, gb)
    _acc(grads, head.bias, gb)
    dW = torch.zeros(Co, 49 * Ci, dtype=torch.float32, device=d.device)
    fpc = max(1, (64 << 20) // (H * W * 49 * Ci))                                 # frames per chunk: the 49x im2col stays <= 256 MB
    for f0 in range(0, F_, fpc):
        f1 = min(F_, f0 + fpc)
        col, _, _ = ops.im2col(x[f0 * H * W:f1 * H * W], f1 - f0, H, W, Ci, 7, 1, 3, 1, round_tf32=False)
        ops.gemm(dpre_cl[f0 * H * W:f1 * H * W], col, out=dW, a_mn=True, b_mn=True, accumulate=True)
        del col
    _acc(grads, head.weight, dW.view(Co, 7, 7, Ci).permute(0, 3, 1, 2).contiguous())
    while tape:
        s = tape.pop()
        F_, H, W, Cin, Cout = s["geom"]
        bn, convT = s["bn"], s["convT"]
        dg, db = torch.zeros_like(bn.weight.data), torch.zeros_like(bn.bias.data)
        _, dy = ops.bn_act_bwd(d, s["y"], s["z"], s["mean"], s["rstd"], bn.weight.data, dg, db, 1, round_tf32=E.ROUND_TF32)
        _acc(grads, bn.weight, dg)
        _acc(grads, bn.bias, db)
        dcol, _, _ = ops.im2col(dy, F_, 2 * H, 2 * W, Cout, 3, 2, 1, 0, round_tf32=E.ROUND_TF32)       # adjoint of the output gather
        dW = torch.zeros(9 * Cout, Cin, dtype=torch.float32, device=d.device)
        ops.gemm(dcol, s["x"], out=dW, a_mn=True, b_mn=True, accumulate=True)
        _acc(grads, convT.weight, dW.view(3, 3, Cout, Cin).permute(3, 2, 0, 1).contiguous())
        d = ops.gemm(dcol, s["wpk"], b_mn=True)
    return d, grads
