// ResNet encoder / decoder support kernels (model/ResNetAutoEncoder.py:26-48 encoder, :70-98 decoder), channel-last.
// The k x k convolutions run as GEMMs on the tcgen05 kernel (gemm_tcgen05.cu); this file holds what surrounds them:
//   * im2col gather (zero / reflect / replicate padding, stride 1|2, optional ReLU-mask for the decoder's backward),
//   * the ConvTranspose2d(3, s2, p1, op1) output gather ("col2im") with folded eval-BatchNorm shift + ReLU,
//   * weight re-layout with the eval-BatchNorm scale folded in,
//   * the 7x7 stem (Cimg -> 64, reads the NCHW frames directly) and the 7x7 head (64 -> Cimg, + bias, Tanh|Sigmoid,
//     writes NCHW frames) as direct HBM-bound kernels, and the head's input gradient.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

int ew_grid(long long n, int block) {
    long long g = (n + block - 1) / block;
    long long cap = 148LL * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

__device__ __forceinline__ int pad_index(int i, int n, int mode) {  // returns -1 for a zero tap
    if (i >= 0 && i < n) return i;
    if (mode == 1) return i < 0 ? -i : 2 * (n - 1) - i;  // reflect (no edge repeat), as nn.ReflectionPad2d
    if (mode == 2) return i < 0 ? 0 : n - 1;             // replicate
    return -1;
}

// col[(f,oh,ow)][(kh,kw,ci)] = x[f][oh*s+kh-p][ow*s+kw-p][ci] (* (mask > 0))
// Grid = (output rows (f, oh), chunks of ow); a thread owns one (tap, float4 channel group) slot of the k*k*C4-wide column row
// and walks its block's output pixels, so the per-element index arithmetic is 32-bit adds (the flat-index version did six
// 64-bit divisions per float4 and ran at a third of its HBM bound).
__global__ void __launch_bounds__(288) im2col_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ col,
                                                     int H, int W, int C4, int Ho, int Wo, int k, int stride, int pad, int pad_mode,
                                                     int round_tf32, int ow_per_block) {
    const int row4 = k * k * C4;                            // float4 per column row
    const int f = blockIdx.x / Ho, oh = blockIdx.x - f * Ho;
    const int ow0 = blockIdx.y * ow_per_block, ow1 = min(ow0 + ow_per_block, Wo);
    for (int slot = threadIdx.x; slot < row4; slot += blockDim.x) {
        const int tap = slot / C4, c = slot - tap * C4;
        const int kh = tap / k, kw = tap - kh * k;
        const int ih = pad_index(oh * stride + kh - pad, H, pad_mode);
        const long long src_row = ((long long)f * H + (ih >= 0 ? ih : 0)) * W;
        float4* dst = reinterpret_cast<float4*>(col) + ((long long)blockIdx.x * Wo + ow0) * row4 + slot;
        for (int ow = ow0; ow < ow1; ++ow, dst += row4) {
            const int iw = pad_index(ow * stride + kw - pad, W, pad_mode);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ih >= 0 && iw >= 0) {
                const long long src = (src_row + iw) * C4 + c;
                v = reinterpret_cast<const float4*>(x)[src];
                if (mask) {
                    const float4 m = reinterpret_cast<const float4*>(mask)[src];
                    v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
                }
            }
            if (round_tf32) { v.x = vptr_round_tf32(v.x); v.y = vptr_round_tf32(v.y); v.z = vptr_round_tf32(v.z); v.w = vptr_round_tf32(v.w); }
            *dst = v;
        }
    }
}

// xpad[f][h][w][c] = x[f][pad_index(h - pad)][pad_index(w - pad)][c]  (zero / reflect / replicate), optionally rounded to tf32:
// the explicit padded copy that the implicit-GEMM convolution's TMA boxes read (1.56x the activation instead of a 9x im2col)
// IDX = unsigned when the element count fits 32 bits (always, at the path's sizes): 64-bit div/mod cost ~5x more instructions
template <typename IDX>
__global__ void __launch_bounds__(256) pad_nhwc_kernel(const float* __restrict__ x, float* __restrict__ out, long long total4, int H, int W,
                                                       int C4, int pad, int pad_mode, int round_tf32) {
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < (IDX)total4; i += (IDX)gridDim.x * blockDim.x) {
        const int c = (int)(i % (IDX)C4);
        IDX t = i / (IDX)C4;
        const int w = (int)(t % (IDX)Wp); t /= (IDX)Wp;
        const int h = (int)(t % (IDX)Hp);
        const long long f = (long long)(t / (IDX)Hp);
        const int ih = pad_index(h - pad, H, pad_mode), iw = pad_index(w - pad, W, pad_mode);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ih >= 0 && iw >= 0) v = reinterpret_cast<const float4*>(x)[((f * H + ih) * W + iw) * C4 + c];
        if (round_tf32) { v.x = vptr_round_tf32(v.x); v.y = vptr_round_tf32(v.y); v.z = vptr_round_tf32(v.z); v.w = vptr_round_tf32(v.w); }
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

// Quadrant-tiled padded copy for vptr_conv3x3_tf32_quad: out[(f, qy, qx)][ph][pw][c] = x[f][pad_index(qy*8 + ph - 1)][pad_index(qx*8 + pw - 1)][c],
// ph, pw in 0..9 -- every 8x8 quadrant with its own 1-pixel halo (neighbouring quadrant, or the frame border's padding rule)
__global__ void __launch_bounds__(256) pad_nhwc_quad_kernel(const float* __restrict__ x, float* __restrict__ out, unsigned total4, int H, int W,
                                                            int C4, int pad_mode, int round_tf32) {
    const int qw = W / 8, qh = H / 8;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
        const int c = (int)(i % (unsigned)C4);
        unsigned t = i / (unsigned)C4;
        const int pw = (int)(t % 10u); t /= 10u;
        const int ph = (int)(t % 10u); t /= 10u;
        const int qx = (int)(t % (unsigned)qw); t /= (unsigned)qw;
        const int qy = (int)(t % (unsigned)qh);
        const long long f = (long long)(t / (unsigned)qh);
        const int ih = pad_index(qy * 8 + ph - 1, H, pad_mode), iw = pad_index(qx * 8 + pw - 1, W, pad_mode);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ih >= 0 && iw >= 0) v = reinterpret_cast<const float4*>(x)[((f * H + ih) * W + iw) * C4 + c];
        if (round_tf32) { v.x = vptr_round_tf32(v.x); v.y = vptr_round_tf32(v.y); v.z = vptr_round_tf32(v.z); v.w = vptr_round_tf32(v.w); }
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

// ConvTranspose2d(k3,s2,p1,op1) output gather: out[f][oh][ow][co] = relu( sum_{kh,kw} col[(f,ih,iw)][(kh,kw,co)] + shift[co] )
// with oh = 2*ih - 1 + kh.
template <typename IDX>
__global__ void __launch_bounds__(256) convT_gather_kernel(const float* __restrict__ col, const float* __restrict__ shift,
                                                           float* __restrict__ out, long long total4, int H, int W, int C4, int relu) {
    const int Ho = 2 * H, Wo = 2 * W;
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < (IDX)total4; i += (IDX)gridDim.x * blockDim.x) {
        const int c = (int)(i % (IDX)C4);
        IDX t = i / (IDX)C4;
        const int ow = (int)(t % (IDX)Wo); t /= (IDX)Wo;
        const int oh = (int)(t % (IDX)Ho);
        const long long f = (long long)(t / (IDX)Ho);
        float4 acc = shift ? __ldg(reinterpret_cast<const float4*>(shift) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int th = oh + 1 - kh;
            if (th < 0 || (th & 1)) continue;
            const int ih = th >> 1;
            if (ih >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int tw = ow + 1 - kw;
                if (tw < 0 || (tw & 1)) continue;
                const int iw = tw >> 1;
                if (iw >= W) continue;
                float4 v = reinterpret_cast<const float4*>(col)[(((f * H + ih) * W + iw) * 9 + kh * 3 + kw) * C4 + c];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
        reinterpret_cast<float4*>(out)[i] = acc;
    }
}

// eval BatchNorm fold: scale = gamma / sqrt(var + eps); shift = beta - mean * scale
__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, float* scale, float* shift, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float s = gamma[c] / sqrtf(rv[c] + eps);
    scale[c] = s;
    shift[c] = beta[c] - rm[c] * s;
}

// mode 0: Conv2d weight [Co][Ci][k][k]           -> out[co][(kh,kw,ci)] * scale[co]        (GEMM B operand, K-major)
// mode 1: ConvTranspose2d weight [Ci][Co][k][k]  -> out[(kh,kw,co)][ci] * scale[co]        (GEMM B operand, K-major, N = k*k*Co)
// mode 2: Conv2d weight [Co][Ci][k][k]           -> out[(kh,kw,ci)][co] * scale[co]        (stem: tap-major, Co contiguous)
// mode 3: Conv2d weight [Co][Ci][k][k]           -> out[(kh,kw)][co][ci]                   (head)
__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale, float* __restrict__ out, int Co, int Ci,
                                   int k, int mode) {
    const long long total = (long long)Co * Ci * k * k;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kw = (int)(i % k);
        long long t = i / k;
        const int kh = (int)(t % k); t /= k;
        int co, ci;
        if (mode == 1) { co = (int)(t % Co); ci = (int)(t / Co); } else { ci = (int)(t % Ci); co = (int)(t / Ci); }
        const float v = w[i] * (scale ? scale[co] : 1.f);
        long long dst;
        if (mode == 0) dst = ((long long)co * k * k + kh * k + kw) * Ci + ci;
        else if (mode == 1) dst = ((long long)(kh * k + kw) * Co + co) * Ci + ci;
        else if (mode == 2) dst = ((long long)(kh * k + kw) * Ci + ci) * Co + co;
        else dst = ((long long)(kh * k + kw) * Co + co) * Ci + ci;
        out[dst] = v;
    }
}

// 7x7 stem: x NCHW [F][Ci][H][W] (reflect pad 3) -> out NHWC [F][H][W][64] = relu(conv + shift); wpk [(kh,kw,ci)][64] (scale folded)
constexpr int STEM_CO = 64;
__global__ void __launch_bounds__(256) stem_conv7x7_kernel(const float* __restrict__ x, const float* __restrict__ wpk,
                                                           const float* __restrict__ shift, float* __restrict__ out, int Ci, int H, int W,
                                                           int relu) {
    extern __shared__ float sm[];
    float* sw = sm;                          // [49*Ci][64]
    float* sp = sm + 49 * Ci * STEM_CO;      // [Ci][22][22]
    const int f = blockIdx.z;
    const int oh0 = blockIdx.y * 16, ow0 = blockIdx.x * 16;
    for (int e = threadIdx.x; e < 49 * Ci * STEM_CO; e += 256) sw[e] = wpk[e];
    for (int e = threadIdx.x; e < Ci * 22 * 22; e += 256) {
        const int ci = e / 484, r = e % 484;
        const int ph = r / 22, pw = r % 22;
        int ih = pad_index(oh0 + ph - 3, H, 1), iw = pad_index(ow0 + pw - 3, W, 1);
        ih = min(max(ih, 0), H - 1); iw = min(max(iw, 0), W - 1);  // tile overhang past the image (never used by valid outputs)
        sp[e] = x[(((long long)f * Ci + ci) * H + ih) * W + iw];
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int oh = oh0 + ty, ow = ow0 + tx;
    float acc[STEM_CO];
#pragma unroll
    for (int c = 0; c < STEM_CO; ++c) acc[c] = shift ? shift[c] : 0.f;
    for (int kh = 0; kh < 7; ++kh)
        for (int kw = 0; kw < 7; ++kw)
            for (int ci = 0; ci < Ci; ++ci) {
                const float v = sp[ci * 484 + (ty + kh) * 22 + tx + kw];
                const float4* wrow = reinterpret_cast<const float4*>(sw + ((kh * 7 + kw) * Ci + ci) * STEM_CO);
#pragma unroll
                for (int c4 = 0; c4 < STEM_CO / 4; ++c4) {
                    const float4 k = wrow[c4];
                    acc[c4 * 4 + 0] = fmaf(v, k.x, acc[c4 * 4 + 0]);
                    acc[c4 * 4 + 1] = fmaf(v, k.y, acc[c4 * 4 + 1]);
                    acc[c4 * 4 + 2] = fmaf(v, k.z, acc[c4 * 4 + 2]);
                    acc[c4 * 4 + 3] = fmaf(v, k.w, acc[c4 * 4 + 3]);
                }
            }
    if (oh < H && ow < W) {
        float4* o = reinterpret_cast<float4*>(out + (((long long)f * H + oh) * W + ow) * STEM_CO);
#pragma unroll
        for (int c4 = 0; c4 < STEM_CO / 4; ++c4)
            o[c4] = relu ? make_float4(fmaxf(acc[c4 * 4], 0.f), fmaxf(acc[c4 * 4 + 1], 0.f), fmaxf(acc[c4 * 4 + 2], 0.f), fmaxf(acc[c4 * 4 + 3], 0.f))
                         : make_float4(acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]);
    }
}

// 7x7 head: x NHWC [F][H][W][Ci] (reflect pad 3) -> out NCHW [F][Co][H][W] = act(conv + bias); wpk [(kh,kw)][co][ci]
constexpr int HEAD_MAXCO = 4;
__device__ __forceinline__ float head_act(float v, int act) {
    if (act == 1) return tanhf(v);
    if (act == 2) return 1.f / (1.f + __expf(-v));
    return v;
}
__global__ void __launch_bounds__(256) head_conv7x7_kernel(const float* __restrict__ x, const float* __restrict__ wpk,
                                                           const float* __restrict__ bias, float* __restrict__ out, int Ci, int Co, int H,
                                                           int W, int act) {
    extern __shared__ float sm[];
    constexpr int CC = 16;                 // channel chunk
    float* sp = sm;                        // [22][22][CC]
    float* sw = sm + 22 * 22 * CC;         // [49][Co][CC]
    const int f = blockIdx.z;
    const int oh0 = blockIdx.y * 16, ow0 = blockIdx.x * 16;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float acc[HEAD_MAXCO] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < Ci; c0 += CC) {
        __syncthreads();
        for (int e = threadIdx.x; e < 484 * (CC / 4); e += 256) {
            const int c4 = e % (CC / 4), r = e / (CC / 4);
            const int ph = r / 22, pw = r % 22;
            int ih = pad_index(oh0 + ph - 3, H, 1), iw = pad_index(ow0 + pw - 3, W, 1);
            ih = min(max(ih, 0), H - 1); iw = min(max(iw, 0), W - 1);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + c4 * 4 < Ci) v = *reinterpret_cast<const float4*>(x + (((long long)f * H + ih) * W + iw) * Ci + c0 + c4 * 4);
            reinterpret_cast<float4*>(sp)[e] = v;
        }
        for (int e = threadIdx.x; e < 49 * Co * CC; e += 256) {
            const int c = e % CC, r = e / CC;  // r = tap*Co + co
            sw[e] = (c0 + c < Ci) ? wpk[(long long)r * Ci + c0 + c] : 0.f;
        }
        __syncthreads();
        for (int kh = 0; kh < 7; ++kh)
            for (int kw = 0; kw < 7; ++kw) {
                const float4* p = reinterpret_cast<const float4*>(sp + ((ty + kh) * 22 + tx + kw) * CC);
                float4 v[CC / 4];
#pragma unroll
                for (int q = 0; q < CC / 4; ++q) v[q] = p[q];
                for (int co = 0; co < Co; ++co) {
                    const float4* wv = reinterpret_cast<const float4*>(sw + ((kh * 7 + kw) * Co + co) * CC);
                    float s = 0.f;
#pragma unroll
                    for (int q = 0; q < CC / 4; ++q) {
                        const float4 k = wv[q];
                        s = fmaf(v[q].x, k.x, s); s = fmaf(v[q].y, k.y, s); s = fmaf(v[q].z, k.z, s); s = fmaf(v[q].w, k.w, s);
                    }
                    acc[co] += s;
                }
            }
    }
    const int oh = oh0 + ty, ow = ow0 + tx;
    if (oh < H && ow < W)
        for (int co = 0; co < Co; ++co)
            out[(((long long)f * Co + co) * H + oh) * W + ow] = head_act(acc[co] + bias[co], act);
}

// 7x7 head, fast path for Ci == 64 and W in {16, 32, 64} or a multiple of 64: the channel contraction is separated from the
// spatial gather.
//   phase 1 (per input pixel, one thread): P[pixel][tap] = sum_c x[pixel][c] * w[tap][c] for all 49 taps -- the pixel's 64
//           channels sit in registers, the weights are warp-broadcast float4 shared loads (4 FMAs per load);
//   phase 2 (per output pixel): out = act(bias + sum_tap P[reflect(y+ky-3)][reflect(x+kx-3)][tap]) -- 49 conflict-free loads.
// A block walks one frame (W <= 64) or one 64-column band of it (wider frames: cfg4's 128 x 128) top to bottom, ROWS image rows
// per step, keeping P of the last ROWS + 6 rows in a shared ring [row][tap][column]; reflect padding is index arithmetic.  A
// band needs the 3 columns on either side of it that exist in the image (the reflected ones map back into the band), so a wide
// frame recomputes 3-6 of 64 columns of P; x is otherwise read exactly once.
// (The direct kernel above spends 4.2 ms on the cfg1 decoder output, 40x its HBM floor: 16-way bank conflicts on the patch
// loads and two shared loads per FMA pair -- and 16.8 ms on cfg4's, which it served until the banded form existed.)
__global__ void __launch_bounds__(512) head_conv7x7_p_kernel(const float* __restrict__ x, const float* __restrict__ wpk,
                                                             const float* __restrict__ bias, float* __restrict__ out, int Co, int H, int W,
                                                             int act, int TW, int ROWS, int NCP) {
    extern __shared__ __align__(16) float sm[];
    const int RING = ROWS + 6;
    float* sw = sm;                   // [49][64]
    float* ring = sm + 49 * 64;       // [RING][49][NCP]
    const int f = blockIdx.x, co = blockIdx.y;
    const int x0 = blockIdx.z * TW;                                   // first output column of this band
    const int c_lo = max(x0 - 3, 0), ncol = min(x0 + TW + 3, W) - c_lo;   // input columns whose P this band needs
    for (int e = threadIdx.x; e < 49 * 64; e += 512) sw[e] = wpk[((long long)(e >> 6) * Co + co) * 64 + (e & 63)];
    __syncthreads();
    const int lr = threadIdx.x / ncol, xx = threadIdx.x - lr * ncol;      // phase 1: (row in step, input column)
    const int orows = 512 / TW, lr2 = threadIdx.x / TW, ox = threadIdx.x - lr2 * TW;   // phase 2: (row, output column)
    const float b = bias[co];
    int emitted = 0;
    for (int r0 = 0; r0 < H; r0 += ROWS) {
        const int row = r0 + lr;
        if (lr < ROWS && row < H) {
            float4 v[16];
            const float4* px = reinterpret_cast<const float4*>(x + (((long long)f * H + row) * W + c_lo + xx) * 64);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = __ldg(px + q);
            float* dst = ring + (size_t)(row % RING) * 49 * NCP + xx;
#pragma unroll 7
            for (int tap = 0; tap < 49; ++tap) {
                const float4* wv = reinterpret_cast<const float4*>(sw + tap * 64);
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float4 k = wv[q];
                    s0 = fmaf(v[q].x, k.x, s0); s1 = fmaf(v[q].y, k.y, s1); s2 = fmaf(v[q].z, k.z, s2); s3 = fmaf(v[q].w, k.w, s3);
                }
                dst[tap * NCP] = (s0 + s1) + (s2 + s3);
            }
        }
        __syncthreads();
        const int r_hi = min(r0 + ROWS, H) - 1;
        const int ready_to = (r_hi == H - 1) ? H - 1 : r_hi - 3;
        for (int orow = emitted + lr2; orow <= ready_to; orow += orows) {
            float acc = b;
#pragma unroll
            for (int ky = 0; ky < 7; ++ky) {
                int ir = orow + ky - 3;
                ir = ir < 0 ? -ir : (ir >= H ? 2 * (H - 1) - ir : ir);
                const float* rp = ring + (size_t)(ir % RING) * 49 * NCP + ky * 7 * NCP - c_lo;
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) {
                    int ic = x0 + ox + kx - 3;
                    ic = ic < 0 ? -ic : (ic >= W ? 2 * (W - 1) - ic : ic);
                    acc += rp[kx * NCP + ic];
                }
            }
            out[(((long long)f * Co + co) * H + orow) * W + x0 + ox] = head_act(acc, act);
        }
        emitted = ready_to + 1;
        __syncthreads();
    }
}

// Head input gradient, step 1: full correlation into the reflect-PADDED frame
//   dxpad[f][p][q][ci] = sum_{kh,kw,co} dz[f][co][p-kh][q-kw] * W[co][ci][kh][kw],   p in [0,H+6), q in [0,W+6),
// with dz = dout * act'(out) (zero outside the image).  Same structure as the stem: a 1..4-channel image correlated with
// 64 filters, 16x16 output tile per block, 64 accumulators per thread.  wf = flipped weights [(kh',kw',co)][64] with
// kh' = 6-kh, so that dxpad[p][q] = sum dzz[p+kh'][q+kw'] * wf, dzz = dz zero-padded by 6.
__global__ void __launch_bounds__(256) head_bwd_corr_kernel(const float* __restrict__ dout, const float* __restrict__ outv,
                                                            const float* __restrict__ wf, float* __restrict__ dxpad, int Co, int H, int W,
                                                            int act) {
    extern __shared__ float sm[];
    float* sw = sm;                          // [49*Co][64]
    float* sp = sm + 49 * Co * STEM_CO;      // [Co][22][22]
    const int Hp = H + 6, Wp = W + 6;
    const int f = blockIdx.z;
    const int p0 = blockIdx.y * 16, q0 = blockIdx.x * 16;
    for (int e = threadIdx.x; e < 49 * Co * STEM_CO; e += 256) sw[e] = wf[e];
    for (int e = threadIdx.x; e < Co * 484; e += 256) {
        const int co = e / 484, r = e % 484;
        const int ih = p0 + r / 22 - 6, iw = q0 + r % 22 - 6;
        float g = 0.f;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
            const long long o = (((long long)f * Co + co) * H + ih) * W + iw;
            g = dout[o];
            if (act == 1) { const float y = outv[o]; g *= (1.f - y * y); }
            else if (act == 2) { const float y = outv[o]; g *= y * (1.f - y); }
        }
        sp[e] = g;
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float acc[STEM_CO];
#pragma unroll
    for (int c = 0; c < STEM_CO; ++c) acc[c] = 0.f;
    for (int kh = 0; kh < 7; ++kh)
        for (int kw = 0; kw < 7; ++kw)
            for (int co = 0; co < Co; ++co) {
                const float v = sp[co * 484 + (ty + kh) * 22 + tx + kw];
                const float4* wrow = reinterpret_cast<const float4*>(sw + ((kh * 7 + kw) * Co + co) * STEM_CO);
#pragma unroll
                for (int c4 = 0; c4 < STEM_CO / 4; ++c4) {
                    const float4 k = wrow[c4];
                    acc[c4 * 4 + 0] = fmaf(v, k.x, acc[c4 * 4 + 0]);
                    acc[c4 * 4 + 1] = fmaf(v, k.y, acc[c4 * 4 + 1]);
                    acc[c4 * 4 + 2] = fmaf(v, k.z, acc[c4 * 4 + 2]);
                    acc[c4 * 4 + 3] = fmaf(v, k.w, acc[c4 * 4 + 3]);
                }
            }
    const int p = p0 + ty, q = q0 + tx;
    if (p < Hp && q < Wp) {
        float4* o = reinterpret_cast<float4*>(dxpad + (((long long)f * Hp + p) * Wp + q) * STEM_CO);
#pragma unroll
        for (int c4 = 0; c4 < STEM_CO / 4; ++c4) o[c4] = make_float4(acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]);
    }
}

// step 2: fold the reflect padding (pad 3): source pixel i receives padded rows i+3, 3-i (1<=i<=3) and 2H+1-i (H-4<=i<=H-2)
__global__ void __launch_bounds__(256) head_bwd_fold_kernel(const float* __restrict__ dxpad, float* __restrict__ dx, long long total4, int H,
                                                            int W) {
    const int Hp = H + 6, Wp = W + 6, C4 = STEM_CO / 4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(t % C4);
        long long r = t / C4;
        const int j = (int)(r % W); r /= W;
        const int i = (int)(r % H);
        const long long f = r / H;
        int ps[3], qs[3], np = 1, nq = 1;
        ps[0] = i + 3; qs[0] = j + 3;
        if (i >= 1 && i <= 3) ps[np++] = 3 - i;
        if (i >= H - 4 && i <= H - 2) ps[np++] = 2 * H + 1 - i;
        if (j >= 1 && j <= 3) qs[nq++] = 3 - j;
        if (j >= W - 4 && j <= W - 2) qs[nq++] = 2 * W + 1 - j;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int a = 0; a < np; ++a)
            for (int b = 0; b < nq; ++b) {
                const float4 v = reinterpret_cast<const float4*>(dxpad)[((f * Hp + ps[a]) * Wp + qs[b]) * C4 + c];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        reinterpret_cast<float4*>(dx)[t] = acc;
    }
}

// head weight [Co][64][7][7] -> flipped, tap-major [(6-kh, 6-kw, co)][64]
__global__ void head_bwd_pack_kernel(const float* __restrict__ w, float* __restrict__ wf, int Co) {
    const int total = Co * STEM_CO * 49;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kw = e % 7, kh = (e / 7) % 7, ci = (e / 49) % STEM_CO, co = e / (49 * STEM_CO);
        wf[(((6 - kh) * 7 + (6 - kw)) * Co + co) * STEM_CO + ci] = w[e];
    }
}

}  // namespace

extern "C" int vptr_im2col(const float* x, const float* mask, float* col, int F, int H, int W, int Cin, int k, int stride, int pad,
                           int pad_mode, int round_tf32, cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && H > 0 && W > 0 && Cin > 0 && Cin % 4 == 0 && k > 0 && stride > 0, VPTR_ERR_SHAPE,
                 "vptr_im2col: F=%d H=%d W=%d Cin=%d k=%d stride=%d", F, H, W, Cin, k, stride);
    VPTR_REQUIRE(pad_mode == 0 || (pad < H && pad < W), VPTR_ERR_SHAPE, "vptr_im2col: reflect/replicate pad %d too large for %dx%d", pad, H, W);
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    VPTR_REQUIRE((long long)F * Ho < 2147483647LL, VPTR_ERR_SHAPE, "vptr_im2col: F*Ho too large");
    const int row4 = k * k * (Cin / 4);
    const int threads = row4 < 288 ? ((row4 + 31) / 32) * 32 : 288;
    // enough blocks to fill the GPU: split each output row into ow chunks when there are few rows
    int chunks = 1;
    while ((long long)F * Ho * chunks < 148 * 8 && chunks < Wo) chunks *= 2;
    const int owpb = (Wo + chunks - 1) / chunks;
    dim3 grid(F * Ho, (Wo + owpb - 1) / owpb);
    im2col_kernel<<<grid, threads, 0, stream>>>(x, mask, col, H, W, Cin / 4, Ho, Wo, k, stride, pad, pad_mode, round_tf32, owpb);
    return vptr_check_launch("im2col_kernel");
}

extern "C" int vptr_pad_nhwc(const float* x, float* out, int F, int H, int W, int C, int pad, int pad_mode, int round_tf32,
                             cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && pad >= 0, VPTR_ERR_SHAPE, "vptr_pad_nhwc: F=%d H=%d W=%d C=%d pad=%d", F, H, W, C, pad);
    VPTR_REQUIRE(pad_mode == 0 || (pad < H && pad < W), VPTR_ERR_SHAPE, "vptr_pad_nhwc: reflect/replicate pad %d too large for %dx%d", pad, H, W);
    const long long total4 = (long long)F * (H + 2 * pad) * (W + 2 * pad) * (C / 4);
    if (total4 < 0x7fffffffLL) pad_nhwc_kernel<unsigned><<<ew_grid(total4, 256), 256, 0, stream>>>(x, out, total4, H, W, C / 4, pad, pad_mode, round_tf32);
    else pad_nhwc_kernel<long long><<<ew_grid(total4, 256), 256, 0, stream>>>(x, out, total4, H, W, C / 4, pad, pad_mode, round_tf32);
    return vptr_check_launch("pad_nhwc_kernel");
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): ~16 significant bits in two bf16 planes (operands of the bf16x3 convolution)
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// quadrant-tiled padded copy as TWO bf16 planes: out[plane][(f, qy, qx)][ph][pw][c], plane 0 = hi, 1 = lo (vptr_conv3x3_bf16x3)
__global__ void __launch_bounds__(256) pad_nhwc_quad_bf16x2_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, unsigned total4,
                                                                   int H, int W, int C4, int pad_mode) {
    const int qw = W / 8, qh = H / 8;
    const size_t plane = (size_t)total4 * 4;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
        const int c = (int)(i % (unsigned)C4);
        unsigned t = i / (unsigned)C4;
        const int pw = (int)(t % 10u); t /= 10u;
        const int ph = (int)(t % 10u); t /= 10u;
        const int qx = (int)(t % (unsigned)qw); t /= (unsigned)qw;
        const int qy = (int)(t % (unsigned)qh);
        const long long f = (long long)(t / (unsigned)qh);
        const int ih = pad_index(qy * 8 + ph - 1, H, pad_mode), iw = pad_index(qx * 8 + pw - 1, W, pad_mode);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ih >= 0 && iw >= 0) v = reinterpret_cast<const float4*>(x)[((f * H + ih) * W + iw) * C4 + c];
        __nv_bfloat16 h[4], l[4];
        split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
        *reinterpret_cast<uint2*>(out + (size_t)i * 4) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(out + plane + (size_t)i * 4) = *reinterpret_cast<const uint2*>(l);
    }
}
// out[r] = [ bf16 hi plane of w[r][0..K) | bf16 lo plane ]  (weights of vptr_conv3x3_bf16x3)
__global__ void __launch_bounds__(256) split_bf16x2_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, long long rows, long long K) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * K; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / K, k = i - r * K;
        __nv_bfloat16 h, l;
        split_bf16(w[i], h, l);
        out[r * 2 * K + k] = h;
        out[r * 2 * K + K + k] = l;
    }
}

extern "C" int vptr_pad_nhwc_quad_bf16x2(const float* x, void* out, int F, int H, int W, int C, int pad_mode, cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0 && C > 0 && C % 8 == 0, VPTR_ERR_SHAPE,
                 "vptr_pad_nhwc_quad_bf16x2: F=%d H=%d W=%d C=%d (H, W, C multiples of 8)", F, H, W, C);
    const long long total4 = (long long)F * (H / 8) * (W / 8) * 100 * (C / 4);
    VPTR_REQUIRE(total4 < 0xffffffffLL, VPTR_ERR_SHAPE, "vptr_pad_nhwc_quad_bf16x2: tensor too large for 32-bit indexing");
    pad_nhwc_quad_bf16x2_kernel<<<ew_grid(total4, 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), (unsigned)total4, H, W, C / 4, pad_mode);
    return vptr_check_launch("pad_nhwc_quad_bf16x2_kernel");
}
extern "C" int vptr_split_bf16x2(const float* w, void* out, long long rows, long long K, cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && K > 0, VPTR_ERR_SHAPE, "vptr_split_bf16x2: rows=%lld K=%lld", rows, K);
    split_bf16x2_kernel<<<ew_grid(rows * K, 256), 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(out), rows, K);
    return vptr_check_launch("split_bf16x2_kernel");
}

extern "C" int vptr_pad_nhwc_quad(const float* x, float* out, int F, int H, int W, int C, int pad_mode, int round_tf32, cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0 && C > 0 && C % 4 == 0, VPTR_ERR_SHAPE,
                 "vptr_pad_nhwc_quad: F=%d H=%d W=%d C=%d (H, W multiples of 8)", F, H, W, C);
    const long long total4 = (long long)F * (H / 8) * (W / 8) * 100 * (C / 4);
    VPTR_REQUIRE(total4 < 0xffffffffLL, VPTR_ERR_SHAPE, "vptr_pad_nhwc_quad: tensor too large for 32-bit indexing");
    pad_nhwc_quad_kernel<<<ew_grid(total4, 256), 256, 0, stream>>>(x, out, (unsigned)total4, H, W, C / 4, pad_mode, round_tf32);
    return vptr_check_launch("pad_nhwc_quad_kernel");
}

extern "C" int vptr_convT_gather(const float* col, const float* shift, float* out, int F, int H, int W, int Cout, int relu,
                                 cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && H > 0 && W > 0 && Cout > 0 && Cout % 4 == 0, VPTR_ERR_SHAPE, "vptr_convT_gather: F=%d H=%d W=%d Cout=%d", F, H, W, Cout);
    const long long total4 = (long long)F * 4 * H * W * (Cout / 4);
    if (total4 < 0x7fffffffLL) convT_gather_kernel<unsigned><<<ew_grid(total4, 256), 256, 0, stream>>>(col, shift, out, total4, H, W, Cout / 4, relu);
    else convT_gather_kernel<long long><<<ew_grid(total4, 256), 256, 0, stream>>>(col, shift, out, total4, H, W, Cout / 4, relu);
    return vptr_check_launch("convT_gather_kernel");
}

extern "C" int vptr_bn_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                            float* scale, float* shift, int C, cudaStream_t stream) {
    bn_fold_kernel<<<vptr_cdiv(C, 128), 128, 0, stream>>>(gamma, beta, running_mean, running_var, eps, scale, shift, C);
    return vptr_check_launch("bn_fold_kernel");
}

extern "C" int vptr_pack_conv_weight(const float* w, const float* scale, float* out, int Co, int Ci, int k, int mode, cudaStream_t stream) {
    VPTR_REQUIRE(Co > 0 && Ci > 0 && k > 0 && mode >= 0 && mode <= 3, VPTR_ERR_SHAPE, "vptr_pack_conv_weight: Co=%d Ci=%d k=%d mode=%d", Co, Ci, k, mode);
    pack_weight_kernel<<<ew_grid((long long)Co * Ci * k * k, 256), 256, 0, stream>>>(w, scale, out, Co, Ci, k, mode);
    return vptr_check_launch("pack_weight_kernel");
}

extern "C" int vptr_stem_conv7x7(const float* x, const float* wpk, const float* shift, float* out, int F, int Ci, int H, int W, int Co,
                                 cudaStream_t stream) {
    VPTR_REQUIRE(Co == STEM_CO, VPTR_ERR_UNSUPPORTED, "vptr_stem_conv7x7: Co=%d (only %d, the reference's fixed ngf)", Co, STEM_CO);
    VPTR_REQUIRE(F > 0 && F < 65536 && Ci > 0 && Ci <= 4 && H > 3 && W > 3, VPTR_ERR_SHAPE, "vptr_stem_conv7x7: F=%d Ci=%d H=%d W=%d", F, Ci, H, W);
    size_t smem = sizeof(float) * ((size_t)49 * Ci * STEM_CO + (size_t)Ci * 484);
    if (smem > 48 * 1024) cudaFuncSetAttribute(stem_conv7x7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(vptr_cdiv(W, 16), vptr_cdiv(H, 16), F);
    stem_conv7x7_kernel<<<grid, 256, smem, stream>>>(x, wpk, shift, out, Ci, H, W, 1);
    return vptr_check_launch("stem_conv7x7_kernel");
}
// the same convolution without the folded BatchNorm shift and ReLU (train-mode BatchNorm follows as its own kernels)
extern "C" int vptr_stem_conv7x7_raw(const float* x, const float* wpk, float* out, int F, int Ci, int H, int W, int Co, cudaStream_t stream) {
    VPTR_REQUIRE(Co == STEM_CO, VPTR_ERR_UNSUPPORTED, "vptr_stem_conv7x7_raw: Co=%d (only %d, the reference's fixed ngf)", Co, STEM_CO);
    VPTR_REQUIRE(F > 0 && F < 65536 && Ci > 0 && Ci <= 4 && H > 3 && W > 3, VPTR_ERR_SHAPE, "vptr_stem_conv7x7_raw: F=%d Ci=%d H=%d W=%d", F, Ci, H, W);
    size_t smem = sizeof(float) * ((size_t)49 * Ci * STEM_CO + (size_t)Ci * 484);
    if (smem > 48 * 1024) cudaFuncSetAttribute(stem_conv7x7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(vptr_cdiv(W, 16), vptr_cdiv(H, 16), F);
    stem_conv7x7_kernel<<<grid, 256, smem, stream>>>(x, wpk, nullptr, out, Ci, H, W, 0);
    return vptr_check_launch("stem_conv7x7_kernel");
}

extern "C" int vptr_head_conv7x7_fwd(const float* x, const float* wpk, const float* bias, float* out, int F, int Ci, int Co, int H, int W,
                                     int act, cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && F < 65536 && Ci > 0 && Ci % 4 == 0 && Co > 0 && Co <= HEAD_MAXCO && H > 3 && W > 3, VPTR_ERR_SHAPE,
                 "vptr_head_conv7x7_fwd: F=%d Ci=%d Co=%d H=%d W=%d", F, Ci, Co, H, W);
    const bool one_band = W == 16 || W == 32 || W == 64;
    if (Ci == 64 && (one_band || (W > 64 && W % 64 == 0)) && H > 6) {
        const int TW = one_band ? W : 64;                       // output columns per block
        const int ncol_max = one_band ? W : (W > 128 ? 70 : 67);    // band + the in-image halo columns
        const int ROWS = 512 / ncol_max, NCP = (ncol_max + 3) & ~3;
        if (H >= ROWS) {
            const size_t psm = sizeof(float) * ((size_t)49 * 64 + (size_t)(ROWS + 6) * 49 * NCP);
            VPTR_REQUIRE(psm <= 200 * 1024, VPTR_ERR_SHAPE, "vptr_head_conv7x7_fwd: ring of %zu bytes", psm);
            static bool attr_set = false;
            if (!attr_set) {
                cudaError_t e = cudaFuncSetAttribute(head_conv7x7_p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                VPTR_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(head_conv7x7_p): %s", cudaGetErrorString(e));
                attr_set = true;
            }
            head_conv7x7_p_kernel<<<dim3(F, Co, W / TW), 512, psm, stream>>>(x, wpk, bias, out, Co, H, W, act, TW, ROWS, NCP);
            return vptr_check_launch("head_conv7x7_p_kernel");
        }
    }
    size_t smem = sizeof(float) * ((size_t)484 * 16 + (size_t)49 * Co * 16);
    dim3 grid(vptr_cdiv(W, 16), vptr_cdiv(H, 16), F);
    head_conv7x7_kernel<<<grid, 256, smem, stream>>>(x, wpk, bias, out, Ci, Co, H, W, act);
    return vptr_check_launch("head_conv7x7_kernel");
}

// workspace `ws`: F*(H+6)*(W+6)*Ci + 49*Co*Ci floats
extern "C" int vptr_head_conv7x7_bwd(const float* dout, const float* out, const float* w, float* dx, int F, int Ci, int Co, int H, int W,
                                     int act, float* ws, cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && F < 65536 && Ci == STEM_CO && Co > 0 && Co <= HEAD_MAXCO && H > 6 && W > 6, VPTR_ERR_SHAPE,
                 "vptr_head_conv7x7_bwd: F=%d Ci=%d (must be %d) Co=%d H=%d W=%d", F, Ci, STEM_CO, Co, H, W);
    VPTR_REQUIRE(ws != nullptr, VPTR_ERR_SHAPE, "vptr_head_conv7x7_bwd: workspace missing");
    float* dxpad = ws;
    float* wf = ws + (long long)F * (H + 6) * (W + 6) * Ci;
    head_bwd_pack_kernel<<<vptr_cdiv(Co * Ci * 49, 256), 256, 0, stream>>>(w, wf, Co);
    size_t smem = sizeof(float) * ((size_t)49 * Co * STEM_CO + (size_t)Co * 484);
    if (smem > 48 * 1024) cudaFuncSetAttribute(head_bwd_corr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(vptr_cdiv(W + 6, 16), vptr_cdiv(H + 6, 16), F);
    head_bwd_corr_kernel<<<grid, 256, smem, stream>>>(dout, out, wf, dxpad, Co, H, W, act);
    const long long total4 = (long long)F * H * W * (Ci / 4);
    head_bwd_fold_kernel<<<ew_grid(total4, 256), 256, 0, stream>>>(dxpad, dx, total4, H, W);
    return vptr_check_launch("vptr_head_conv7x7_bwd");
}
