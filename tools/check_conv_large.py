"""Large-frame-count check of the three 3x3 convolution paths (every CTA of the persistent kernels walks several tiles) against
cuDNN fp32 (TF32 off): generic per-tap implicit GEMM, 8x8 raw-tile kernel, quadrant-tiled raw-tile kernel."""
import torch
import torch.nn.functional as F
from vptr_b200 import ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def ref(x, w, b, Fr, H, W, C, Co):
    xn = x.view(Fr, H, W, C).permute(0, 3, 1, 2)
    out = []
    for i in range(0, Fr, 64):
        xp = F.pad(xn[i:i + 64], (1,) * 4, mode="reflect")
        out.append(torch.relu(F.conv2d(xp, w, b)).permute(0, 2, 3, 1).reshape(-1, Co))
    return torch.cat(out)


def report(name, y, r):
    d = (y - r).abs()
    bad = (d > 1e-2 * r.abs().max()).nonzero()
    print(f"{name}: rel_l2 {((y - r).norm() / r.norm()).item():.3e} max {d.max().item():.3e} bad rows {bad[:, 0].unique().numel()}"
          + (f" first {bad[0].tolist()} last {bad[-1].tolist()}" if bad.numel() else ""))


for Fr, H, W in ((1280, 16, 16), (2560, 8, 8), (330, 16, 16), (75, 16, 16)):
    C = Co = 528
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(Fr * H * W, C, device="cuda", generator=g)
    w = torch.randn(Co, C, 3, 3, device="cuda", generator=g) * 0.02
    b = torch.randn(Co, device="cuda", generator=g)
    w2 = ops.split_tf32(ops.pack_conv_weight(w, None, 0).view(Co, 9 * C))
    xr = ops.round_copy(x) if hasattr(ops, "round_copy") else x
    r = ref(xr, w, b, Fr, H, W, C, Co)
    xp = ops.pad_nhwc(x, Fr, H, W, C, 1, 1, round_tf32=True)
    report(f"F={Fr} {H}x{W} native ", ops.conv3x3_tf32(xp, w2, Fr, H, W, C, Co, bias=b, act=ops.ACT_RELU, w_planes=2), r)
    if ops.conv3x3_quad_ok(H, W):
        xq = ops.pad_nhwc_quad(x, Fr, H, W, C, 1, round_tf32=True)
        report(f"F={Fr} {H}x{W} quad   ", ops.conv3x3_tf32_quad(xq, w2, Fr, H, W, C, Co, bias=b, act=ops.ACT_RELU, w_planes=2), r)
