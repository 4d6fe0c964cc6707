// Depthwise 3x3 convolution of MlpDWBN (VidHRFormer_modules.py:405-410: groups = hidden, padding 1, bias) on
// channel-last activations [F][H][W][ch].  Weights are taken tap-major [9][ch] (engine layout; the module's
// (ch,1,3,3) parameter is re-laid once per step).  HBM-bound: every element is read once from DRAM (the 9-tap
// reuse is served by L1/L2) and written once.
#include "common.cuh"
#include <stdlib.h>

namespace {

int ew_grid(long long n, int block) {
    long long g = (n + block - 1) / block;
    long long cap = 148LL * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// One block row per (frame, image row h); thread = one float4 channel group, walking along w with a 3x3 register window so
// every input element is loaded once per output row it touches (3 loads per output instead of 9).
// flip = 0: y[h][w] = b + sum_t w[t] x[h+dh-1][w+dw-1]  (forward)
// flip = 1: the adjoint w.r.t. x (input gradient): uses tap 8-t and no bias.
__device__ __forceinline__ float4 ld_col(const float* __restrict__ x, long long frame_base, int hh, int j, int H, int W, int C4, int c) {
    if (hh < 0 || hh >= H || j < 0 || j >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
    return reinterpret_cast<const float4*>(x)[(frame_base + (long long)hh * W + j) * C4 + c];
}
__device__ __forceinline__ void fma4(float4& a, const float4& k, const float4& v) {
    a.x = fmaf(k.x, v.x, a.x); a.y = fmaf(k.y, v.y, a.y); a.z = fmaf(k.z, v.z, a.z); a.w = fmaf(k.w, v.w, a.w);
}
// Two output rows per thread: a 4-row x 3-column register window (every input row is loaded for 2 outputs instead of 1, and
// the two accumulator chains are independent); the next column's loads are issued before the current column's FMAs.
constexpr int DW_THREADS = 176;   // 528 float4 channel groups = 3 blocks exactly (128-thread blocks left the 5th 87 % idle)
__global__ void __launch_bounds__(DW_THREADS) dwconv3x3_kernel(const float* __restrict__ x, const float* __restrict__ w9,
                                                               const float* __restrict__ bias, float* __restrict__ y, int H, int W, int C4,
                                                               int flip) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= C4) return;
    const int hp = (H + 1) >> 1;                      // row pairs per frame
    const int f = blockIdx.x / hp, h = (blockIdx.x - f * hp) * 2;
    const long long fb = (long long)f * H * W;
    float4 k[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = __ldg(reinterpret_cast<const float4*>(w9) + (flip ? 8 - t : t) * C4 + c);
    const float4 b = (bias && !flip) ? __ldg(reinterpret_cast<const float4*>(bias) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 win[4][3], nxt[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        win[r][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        win[r][1] = ld_col(x, fb, h + r - 1, 0, H, W, C4, c);
        win[r][2] = ld_col(x, fb, h + r - 1, 1, H, W, C4, c);
    }
    const bool two = h + 1 < H;
    for (int w = 0; w < W; ++w) {
#pragma unroll
        for (int r = 0; r < 4; ++r) nxt[r] = ld_col(x, fb, h + r - 1, w + 2, H, W, C4, c);
        float4 a0 = b, a1 = b;
#pragma unroll
        for (int dh = 0; dh < 3; ++dh)
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
                fma4(a0, k[dh * 3 + dw], win[dh][dw]);
                fma4(a1, k[dh * 3 + dw], win[dh + 1][dw]);
            }
        reinterpret_cast<float4*>(y)[(fb + (long long)h * W + w) * C4 + c] = a0;
        if (two) reinterpret_cast<float4*>(y)[(fb + (long long)(h + 1) * W + w) * C4 + c] = a1;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            win[r][0] = win[r][1];
            win[r][1] = win[r][2];
            win[r][2] = nxt[r];
        }
    }
}

// Streaming variant for small frames (H*W*SLAB float4 <= ~100 KB, i.e. the 8x8 grid of every 64x64 config): persistent blocks,
// work item = (frame, slab of SLAB float4 channel groups), the item's whole input tile [H*W][SLAB] is brought in with cp.async
// into one of two shared buffers while the previous item is computed -- ~90 KB of loads in flight per SM instead of the
// 4 float4 per thread of the register-window kernel above (which was memory-latency bound at 52 % of HBM peak).  A thread owns
// one channel group and two output rows and slides the same 4x3 register window, now fed from shared memory.
__global__ void __launch_bounds__(352, 1) dwconv3x3_stream_kernel(const float* __restrict__ x, const float* __restrict__ w9,
                                                                  const float* __restrict__ bias, float* __restrict__ y, int F, int H, int W,
                                                                  int C4, int slab, int nslab, int flip, double* __restrict__ sums) {
    extern __shared__ __align__(16) float4 dsm[];
    const int HW = H * W, tile4 = HW * slab;
    const int c = threadIdx.x % slab, rg = threadIdx.x / slab;       // channel group in slab, output row pair
    const int items = F * nslab;
    auto issue = [&](int item, float4* buf) {
        const int f = item / nslab, s0 = (item - f * nslab) * slab;
        const float4* src = reinterpret_cast<const float4*>(x) + (long long)f * HW * C4 + s0;
        for (int e = threadIdx.x; e < tile4; e += blockDim.x) {
            const int px = e / slab, cc = e - px * slab;
            if (s0 + cc < C4) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf + e);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (long long)px * C4 + cc) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int cur = 0;
    if ((int)blockIdx.x < items) issue(blockIdx.x, dsm);
    for (int item = blockIdx.x; item < items; item += gridDim.x, cur ^= 1) {
        const int next = item + gridDim.x;
        if (next < items) issue(next, dsm + (cur ^ 1) * tile4);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const int f = item / nslab, s0 = (item - f * nslab) * slab;
        const int h = rg * 2;
        float st_s = 0.f, st_q = 0.f;        // per-frame sum / sum of squares of the outputs (LayerNorm((ch,H,W)) statistics of the next op)
        if (s0 + c < C4 && h < H) {
            const float4* tile = dsm + cur * tile4;
            float4 k[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) k[t] = __ldg(reinterpret_cast<const float4*>(w9) + (flip ? 8 - t : t) * C4 + s0 + c);
            const float4 b = (bias && !flip) ? __ldg(reinterpret_cast<const float4*>(bias) + s0 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            auto ld = [&](int hh, int j) -> float4 {
                if (hh < 0 || hh >= H || j < 0 || j >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
                return tile[(hh * W + j) * slab + c];
            };
            float4 win[4][3];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                win[r][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                win[r][1] = ld(h + r - 1, 0);
                win[r][2] = ld(h + r - 1, 1);
            }
            const bool two = h + 1 < H;
            float4* out = reinterpret_cast<float4*>(y) + ((long long)f * HW + (long long)h * W) * C4 + s0 + c;
            for (int w = 0; w < W; ++w) {
                float4 a0 = b, a1 = b;
#pragma unroll
                for (int dh = 0; dh < 3; ++dh)
#pragma unroll
                    for (int dw = 0; dw < 3; ++dw) {
                        fma4(a0, k[dh * 3 + dw], win[dh][dw]);
                        fma4(a1, k[dh * 3 + dw], win[dh + 1][dw]);
                    }
                out[(long long)w * C4] = a0;
                st_s += (a0.x + a0.y) + (a0.z + a0.w);
                st_q += (a0.x * a0.x + a0.y * a0.y) + (a0.z * a0.z + a0.w * a0.w);
                if (two) {
                    out[(long long)(W + w) * C4] = a1;
                    st_s += (a1.x + a1.y) + (a1.z + a1.w);
                    st_q += (a1.x * a1.x + a1.y * a1.y) + (a1.z * a1.z + a1.w * a1.w);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    win[r][0] = win[r][1];
                    win[r][1] = win[r][2];
                    win[r][2] = ld(h + r - 1, w + 2);
                }
            }
        }
        if (sums) {   // one pair of fp64 atomics per warp and item
            st_s = warp_sum(st_s);
            st_q = warp_sum(st_q);
            if ((threadIdx.x & 31) == 0) { atomicAdd(sums + f, (double)st_s); atomicAdd(sums + F + f, (double)st_q); }
        }
        __syncthreads();      // the buffer just read is the target of the prefetch issued in the next iteration
    }
}

// dW9[t][c] += sum_{f,h,w} dy[f,h,w,c] * x[f,h+dh-1,w+dw-1,c] ; dbias[c] += sum dy.  Thread = float4 channel group, a chunk of
// frames per blockIdx.y, same sliding window over x.
__global__ void __launch_bounds__(DW_THREADS) dwconv3x3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                     float* __restrict__ dw9, float* __restrict__ dbias, int F, int H, int W,
                                                                     int C4, int frames_per_block) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C4) return;
    const int f0 = blockIdx.y * frames_per_block;
    const int f1 = min(f0 + frames_per_block, F);
    float4 acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ab = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int f = f0; f < f1; ++f) {
        const long long fb = (long long)f * H * W;
        for (int h = 0; h < H; h += 2) {                    // two dy rows per pass over a 4-row x window
            const bool two = h + 1 < H;
            float4 win[4][3], nxt[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                win[r][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                win[r][1] = ld_col(x, fb, h + r - 1, 0, H, W, C4, c);
                win[r][2] = ld_col(x, fb, h + r - 1, 1, H, W, C4, c);
            }
            for (int w = 0; w < W; ++w) {
#pragma unroll
                for (int r = 0; r < 4; ++r) nxt[r] = ld_col(x, fb, h + r - 1, w + 2, H, W, C4, c);
                const float4 g0 = reinterpret_cast<const float4*>(dy)[(fb + (long long)h * W + w) * C4 + c];
                const float4 g1 = two ? reinterpret_cast<const float4*>(dy)[(fb + (long long)(h + 1) * W + w) * C4 + c]
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                ab.x += g0.x + g1.x; ab.y += g0.y + g1.y; ab.z += g0.z + g1.z; ab.w += g0.w + g1.w;
#pragma unroll
                for (int dh = 0; dh < 3; ++dh)
#pragma unroll
                    for (int dw = 0; dw < 3; ++dw) {
                        fma4(acc[dh * 3 + dw], g0, win[dh][dw]);
                        fma4(acc[dh * 3 + dw], g1, win[dh + 1][dw]);
                    }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    win[r][0] = win[r][1];
                    win[r][1] = win[r][2];
                    win[r][2] = nxt[r];
                }
            }
        }
    }
    const int ch = C4 * 4;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        float* d = dw9 + t * ch + c * 4;
        atomicAdd(d, acc[t].x); atomicAdd(d + 1, acc[t].y); atomicAdd(d + 2, acc[t].z); atomicAdd(d + 3, acc[t].w);
    }
    float* d = dbias + c * 4;
    atomicAdd(d, ab.x); atomicAdd(d + 1, ab.y); atomicAdd(d + 2, ab.z); atomicAdd(d + 3, ab.w);
}

}  // namespace

// launches the streaming kernel when the shape allows it; returns 1 if it did not (caller falls back), <0 / >1 on error
static int dwconv_stream_launch(const float* x, const float* w9, const float* bias, float* y, int F, int H, int W, int ch, int flip,
                                double* sums, cudaStream_t stream) {
    static const bool off = [] { const char* e = getenv("VPTR_DWCONV_NOSTREAM"); return e && e[0] == '1'; }();
    const int C4 = ch / 4, HW = H * W, pairs = (H + 1) / 2;
    int slab = 352 / pairs;                                   // one thread per (channel group, row pair)
    if (slab > C4) slab = C4;
    const size_t smem = (size_t)2 * HW * slab * sizeof(float4);
    if (off || slab < 32 || smem > 200 * 1024 || ((uintptr_t)x % 16) || ((uintptr_t)y % 16)) return 1;
    const int nslab = vptr_cdiv(C4, slab);
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(dwconv3x3_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        VPTR_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(dwconv3x3_stream, smem=%zu): %s", smem, cudaGetErrorString(e));
        attr = smem;
    }
    const long long items = (long long)F * nslab;
    const int blocks = (int)(items < 148 ? items : 148);
    dwconv3x3_stream_kernel<<<blocks, slab * pairs, smem, stream>>>(x, w9, bias, y, F, H, W, C4, slab, nslab, flip, sums);
    return vptr_check_launch("dwconv3x3_stream_kernel");
}

extern "C" int vptr_dwconv3x3(const float* x, const float* w9, const float* bias, float* y, int F, int H, int W, int ch, int flip,
                              cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && H > 0 && W > 0 && ch > 0 && ch % 4 == 0, VPTR_ERR_SHAPE, "vptr_dwconv3x3: F=%d H=%d W=%d ch=%d", F, H, W, ch);
    VPTR_REQUIRE((long long)F * H < 2147483647LL, VPTR_ERR_SHAPE, "vptr_dwconv3x3: F*H too large");
    const int rc = dwconv_stream_launch(x, w9, bias, y, F, H, W, ch, flip, nullptr, stream);
    if (rc != 1) return rc;
    dim3 grid(F * ((H + 1) / 2), vptr_cdiv(ch / 4, DW_THREADS));
    dwconv3x3_kernel<<<grid, DW_THREADS, 0, stream>>>(x, w9, bias, y, H, W, ch / 4, flip);
    return vptr_check_launch("dwconv3x3_kernel");
}

// Forward depthwise conv that also accumulates, per frame, the sum and the sum of squares of its outputs into
// sums[0:F] / sums[F:2F] (fp64, zeroed by the caller): the statistics of the LayerNorm((ch,H,W)) that follows (MlpDWBN norm2,
// VidHRFormer_modules.py:436) without a separate pass over the tensor.  Returns -3 when the shape is outside the streaming
// kernel's domain (callers then use vptr_dwconv3x3 + vptr_group_stats).  vptr_group_stats_finalize turns sums into mean / rstd.
extern "C" int vptr_dwconv3x3_stats(const float* x, const float* w9, const float* bias, float* y, int F, int H, int W, int ch,
                                    double* sums, cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && H > 0 && W > 0 && ch > 0 && ch % 4 == 0 && sums != nullptr, VPTR_ERR_SHAPE, "vptr_dwconv3x3_stats: F=%d H=%d W=%d ch=%d", F, H, W, ch);
    const int rc = dwconv_stream_launch(x, w9, bias, y, F, H, W, ch, 0, sums, stream);
    if (rc == 1) { vptr_set_error("vptr_dwconv3x3_stats: shape outside the streaming kernel's domain"); return VPTR_ERR_UNSUPPORTED; }
    return rc;
}

extern "C" int vptr_dwconv3x3_wgrad(const float* x, const float* dy, float* dw9, float* dbias, int F, int H, int W, int ch,
                                    cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && H > 0 && W > 0 && ch > 0 && ch % 4 == 0, VPTR_ERR_SHAPE, "vptr_dwconv3x3_wgrad: F=%d H=%d W=%d ch=%d", F, H, W, ch);
    int fpb = (256 + H * W - 1) / (H * W);  // ~256 pixels per block
    if (fpb < 1) fpb = 1;
    dim3 grid(vptr_cdiv(ch / 4, DW_THREADS), vptr_cdiv(F, fpb));
    dwconv3x3_wgrad_kernel<<<grid, DW_THREADS, 0, stream>>>(x, dy, dw9, dbias, F, H, W, ch / 4, fpb);
    return vptr_check_launch("dwconv3x3_wgrad_kernel");
}
