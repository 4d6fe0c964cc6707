// Train-mode pieces of the ResNet autoencoder (stage-1 training, reference train_AutoEncoder.py:44-86): the
// convolutions are GEMMs on the tcgen05 kernel over im2col operands (conv.cu, gemm_tcgen05.cu); this file holds what stage 2 never
// needs -- BatchNorm2d with BATCH statistics followed by ReLU / nothing (forward and backward), the adjoint of the im2col gather
// (input gradient of a k x k convolution under zero / reflect / replicate padding), and the weight gradients of the 7x7 stem and
// the derivative of the head's output activation.   model/ResNetAutoEncoder.py:26-48,70-98,104-158.
#include "common.cuh"

extern "C" int vptr_axpby(const float* a, const float* b, float* out, long long n, float alpha, float beta, cudaStream_t stream);

namespace {

int ew_grid(long long n, int block) {
    long long g = (n + block - 1) / block;
    long long cap = 148LL * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}
__device__ __forceinline__ int pad_index(int i, int n, int mode) {  // returns -1 for a zero tap (same rule as conv.cu)
    if (i >= 0 && i < n) return i;
    if (mode == 1) return i < 0 ? -i : 2 * (n - 1) - i;
    if (mode == 2) return i < 0 ? 0 : n - 1;
    return -1;
}

// z = act(xhat * gamma + beta) [+ res];  act 0: none, 1: ReLU before the residual add, 2: ReLU after it (last ResnetBlock + nn.ReLU)
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const float* __restrict__ x, float* __restrict__ z, const float* __restrict__ res,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta, long long n4, int C4,
                                                         int act, int round_tf32) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + c4), r = __ldg(reinterpret_cast<const float4*>(rstd) + c4);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
        float4 o = make_float4((v.x - m.x) * r.x * g.x + b.x, (v.y - m.y) * r.y * g.y + b.y, (v.z - m.z) * r.z * g.z + b.z, (v.w - m.w) * r.w * g.w + b.w);
        if (act == 1) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (res) { const float4 q = reinterpret_cast<const float4*>(res)[i]; o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
        if (act == 2) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (round_tf32) { o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w); }
        reinterpret_cast<float4*>(z)[i] = o;
    }
}
// pass A: g0 = dz * act'(.) (kept in g0, may alias dz), per channel S1 = sum g0, S2 = sum g0 * xhat (-> dbeta, dgamma).
// z: the forward OUTPUT (ReLU mask: z > 0 for act 1 and 2); for act 2 the masked dz is also the residual branch's gradient.
__global__ void __launch_bounds__(128) bn_act_bwd_a_kernel(const float* __restrict__ dz, const float* __restrict__ x, const float* __restrict__ z,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ g0,
                                                           float* __restrict__ s1, float* __restrict__ s2, long long rows, int ch, int act,
                                                           int rows_per_block) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (c >= ch) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, rows);
    const float4 m = *reinterpret_cast<const float4*>(mean + c), r = *reinterpret_cast<const float4*>(rstd + c);
    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
    for (long long row = r0; row < r1; ++row) {
        const long long e = row * ch + c;
        const float4 xv = *reinterpret_cast<const float4*>(x + e);
        float4 d = *reinterpret_cast<const float4*>(dz + e);
        if (act) {
            const float4 zv = *reinterpret_cast<const float4*>(z + e);
            d.x = zv.x > 0.f ? d.x : 0.f; d.y = zv.y > 0.f ? d.y : 0.f; d.z = zv.z > 0.f ? d.z : 0.f; d.w = zv.w > 0.f ? d.w : 0.f;
        }
        *reinterpret_cast<float4*>(g0 + e) = d;
        a1.x += d.x; a1.y += d.y; a1.z += d.z; a1.w += d.w;
        a2.x = fmaf(d.x, (xv.x - m.x) * r.x, a2.x); a2.y = fmaf(d.y, (xv.y - m.y) * r.y, a2.y);
        a2.z = fmaf(d.z, (xv.z - m.z) * r.z, a2.z); a2.w = fmaf(d.w, (xv.w - m.w) * r.w, a2.w);
    }
    atomicAdd(s1 + c, a1.x); atomicAdd(s1 + c + 1, a1.y); atomicAdd(s1 + c + 2, a1.z); atomicAdd(s1 + c + 3, a1.w);
    atomicAdd(s2 + c, a2.x); atomicAdd(s2 + c + 1, a2.y); atomicAdd(s2 + c + 2, a2.z); atomicAdd(s2 + c + 3, a2.w);
}
// pass B: dx = gamma * rstd * (g0 - S1/n - xhat * S2/n)
__global__ void __launch_bounds__(256) bn_act_bwd_b_kernel(const float* __restrict__ g0, const float* __restrict__ x, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ s1, const float* __restrict__ s2, float* __restrict__ dx,
                                                           long long n4, int C4, float inv_n, int round_tf32) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const float4 d = reinterpret_cast<const float4*>(g0)[i], xv = reinterpret_cast<const float4*>(x)[i];
        const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + c4), r = __ldg(reinterpret_cast<const float4*>(rstd) + c4);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        const float4 u1 = __ldg(reinterpret_cast<const float4*>(s1) + c4), u2 = __ldg(reinterpret_cast<const float4*>(s2) + c4);
        float4 o;
        o.x = g.x * r.x * (d.x - u1.x * inv_n - (xv.x - m.x) * r.x * u2.x * inv_n);
        o.y = g.y * r.y * (d.y - u1.y * inv_n - (xv.y - m.y) * r.y * u2.y * inv_n);
        o.z = g.z * r.z * (d.z - u1.z * inv_n - (xv.z - m.z) * r.z * u2.z * inv_n);
        o.w = g.w * r.w * (d.w - u1.w * inv_n - (xv.w - m.w) * r.w * u2.w * inv_n);
        if (round_tf32) { o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w); }
        reinterpret_cast<float4*>(dx)[i] = o;
    }
}

// dx[f][ih][iw][c] += dcol[(f,oh,ow)][(kh,kw,c)] with (ih, iw) = pad_index(oh*s + kh - p, ow*s + kw - p): the adjoint of im2col.
// Under reflect / replicate padding several taps of one output pixel can land on the same input pixel, so it is a scatter.
__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ dcol, float* __restrict__ dx, long long total4, int H, int W, int C4,
                                                     int Ho, int Wo, int k, int stride, int pad, int pad_mode) {
    const int row4 = k * k * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int slot = (int)(i % row4);
        long long t = i / row4;
        const int ow = (int)(t % Wo); t /= Wo;
        const int oh = (int)(t % Ho);
        const long long f = t / Ho;
        const int tap = slot / C4, c = slot - tap * C4, kh = tap / k, kw = tap - kh * k;
        const int ih = pad_index(oh * stride + kh - pad, H, pad_mode), iw = pad_index(ow * stride + kw - pad, W, pad_mode);
        if (ih < 0 || iw < 0) continue;
        const float4 v = reinterpret_cast<const float4*>(dcol)[i];
        float* d = dx + (((f * H + ih) * W + iw) * C4 + c) * 4;
        atomicAdd(d, v.x); atomicAdd(d + 1, v.y); atomicAdd(d + 2, v.z); atomicAdd(d + 3, v.w);
    }
}

// dW[(kh,kw,ci)][co] += sum_{f,h,w} x[f][ci][refl(h+kh-3)][refl(w+kw-3)] * dy[f][h][w][co]   (7x7 stem, x NCHW, dy NHWC, 64 outputs)
__global__ void __launch_bounds__(64) stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int F,
                                                        int Ci, int H, int W, int pix_per_block) {
    const int tapci = blockIdx.x;                      // (kh*7 + kw)*Ci + ci
    const int ci = tapci % Ci, tap = tapci / Ci, kh = tap / 7, kw = tap - kh * 7;
    const int co = threadIdx.x;
    const long long total = (long long)F * H * W;
    const long long p0 = (long long)blockIdx.y * pix_per_block, p1 = min(p0 + pix_per_block, total);
    float acc = 0.f;
    for (long long p = p0; p < p1; ++p) {
        const int w = (int)(p % W), h = (int)((p / W) % H);
        const long long f = p / ((long long)H * W);
        const int ih = pad_index(h + kh - 3, H, 1), iw = pad_index(w + kw - 3, W, 1);
        acc = fmaf(__ldg(x + ((f * Ci + ci) * H + ih) * W + iw), dy[p * 64 + co], acc);
    }
    atomicAdd(dw + (long long)tapci * 64 + co, acc);
}
// dpre = dout * act'(out) for the head's output activation (1 tanh: 1 - out^2, 2 sigmoid: out (1 - out))
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* __restrict__ dpre, long long n,
                                                      int act) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float o = out[i];
        dpre[i] = dout[i] * (act == 1 ? 1.f - o * o : (act == 2 ? o * (1.f - o) : 1.f));
    }
}

}  // namespace

extern "C" int vptr_bn_act_fwd(const float* x, float* z, const float* res, const float* mean, const float* rstd, const float* gamma,
                               const float* beta, long long rows, int ch, int act, int round_tf32, cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && ch > 0 && ch % 4 == 0 && act >= 0 && act <= 2, VPTR_ERR_SHAPE, "vptr_bn_act_fwd: rows=%lld ch=%d act=%d", rows, ch, act);
    const long long n4 = rows * (ch / 4);
    bn_act_fwd_kernel<<<ew_grid(n4, 256), 256, 0, stream>>>(x, z, res, mean, rstd, gamma, beta, n4, ch / 4, act, round_tf32);
    return vptr_check_launch("bn_act_fwd_kernel");
}
// Backward of z = act(BN_batchstats(x)) [+ res].  g0 (rows x ch) receives the activation-masked dz (for act 2 it IS the residual
// branch's gradient; it may alias dz); dx may alias g0.  dgamma / dbeta are accumulated (+=).  ws: 2*ch floats of scratch.
extern "C" int vptr_bn_act_bwd(const float* dz, const float* x, const float* z, const float* mean, const float* rstd, const float* gamma,
                               float* g0, float* dx, float* dgamma, float* dbeta, long long rows, int ch, int act, float* ws, int round_tf32,
                               cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && ch > 0 && ch % 4 == 0 && act >= 0 && act <= 2 && (act == 0 || z != nullptr), VPTR_ERR_SHAPE,
                 "vptr_bn_act_bwd: rows=%lld ch=%d act=%d", rows, ch, act);
    cudaMemsetAsync(ws, 0, sizeof(float) * 2 * ch, stream);
    const int rpb = 128;
    dim3 ga(vptr_cdiv(ch / 4, 128), vptr_cdiv(rows, rpb));
    bn_act_bwd_a_kernel<<<ga, 128, 0, stream>>>(dz, x, z, mean, rstd, g0, ws, ws + ch, rows, ch, act, rpb);
    const long long n4 = rows * (ch / 4);
    bn_act_bwd_b_kernel<<<ew_grid(n4, 256), 256, 0, stream>>>(g0, x, mean, rstd, gamma, ws, ws + ch, dx, n4, ch / 4, 1.0f / (float)rows, round_tf32);
    int rc = vptr_check_launch("bn_act_bwd");
    if (rc) return rc;
    // dbeta += S1, dgamma += S2
    if (dbeta) { rc = vptr_axpby(dbeta, ws, dbeta, ch, 1.f, 1.f, stream); if (rc) return rc; }
    if (dgamma) { rc = vptr_axpby(dgamma, ws + ch, dgamma, ch, 1.f, 1.f, stream); if (rc) return rc; }
    return VPTR_OK;
}
// dx (F*H*W x C, zeroed by the caller) += adjoint of vptr_im2col applied to dcol (F*Ho*Wo x k*k*C)
extern "C" int vptr_col2im(const float* dcol, float* dx, int F, int H, int W, int C, int k, int stride, int pad, int pad_mode, cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && k > 0 && stride > 0, VPTR_ERR_SHAPE, "vptr_col2im: F=%d H=%d W=%d C=%d k=%d", F, H, W, C, k);
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    const long long total4 = (long long)F * Ho * Wo * k * k * (C / 4);
    col2im_kernel<<<ew_grid(total4, 256), 256, 0, stream>>>(dcol, dx, total4, H, W, C / 4, Ho, Wo, k, stride, pad, pad_mode);
    return vptr_check_launch("col2im_kernel");
}
// dw ((49*Ci) x 64, zeroed by the caller, layout of vptr_pack_conv_weight mode 2) += weight gradient of the 7x7 stem
extern "C" int vptr_stem_wgrad(const float* x, const float* dy, float* dw, int F, int Ci, int H, int W, cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && Ci > 0 && Ci <= 4 && H > 3 && W > 3, VPTR_ERR_SHAPE, "vptr_stem_wgrad: F=%d Ci=%d H=%d W=%d", F, Ci, H, W);
    const long long total = (long long)F * H * W;
    int chunks = (int)((total + 4095) / 4096);
    if (chunks > 512) chunks = 512;
    const int ppb = (int)((total + chunks - 1) / chunks);
    dim3 grid(49 * Ci, vptr_cdiv(total, ppb));
    stem_wgrad_kernel<<<grid, 64, 0, stream>>>(x, dy, dw, F, Ci, H, W, ppb);
    return vptr_check_launch("stem_wgrad_kernel");
}
extern "C" int vptr_act_bwd(const float* dout, const float* out, float* dpre, long long n, int act, cudaStream_t stream) {
    VPTR_REQUIRE(n > 0 && act >= 0 && act <= 2, VPTR_ERR_SHAPE, "vptr_act_bwd: n=%lld act=%d", n, act);
    act_bwd_kernel<<<ew_grid(n, 256), 256, 0, stream>>>(dout, out, dpre, n, act);
    return vptr_check_launch("act_bwd_kernel");
}
