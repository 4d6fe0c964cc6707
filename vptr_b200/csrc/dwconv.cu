// Depthwise 3x3 convolution of MlpDWBN (VidHRFormer_modules.py:405-410: groups = hidden, padding 1, bias) on
// channel-last activations [F][H][W][ch].  Weights are taken tap-major [9][ch] (engine layout; the module's
// (ch,1,3,3) parameter is re-laid once per step).  HBM-bound: every element is read once from DRAM (the 9-tap
// reuse is served by L1/L2) and written once.
#include "common.cuh"

namespace {

int ew_grid(long long n, int block) {
    long long g = (n + block - 1) / block;
    long long cap = 148LL * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// flip = 0: y[h][w] = b + sum_t w[t] x[h+dh-1][w+dw-1]  (forward)
// flip = 1: the adjoint w.r.t. x (input gradient): uses tap 8-t and no bias.
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const float* __restrict__ x, const float* __restrict__ w9,
                                                        const float* __restrict__ bias, float* __restrict__ y, long long total4,
                                                        int H, int W, int C4, int flip) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        long long t = i / C4;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const long long f = t / H;
        float4 acc = (bias && !flip) ? __ldg(reinterpret_cast<const float4*>(bias) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
            const int hh = h + dh - 1;
            if (hh < 0 || hh >= H) continue;
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
                const int ww = w + dw - 1;
                if (ww < 0 || ww >= W) continue;
                const int tap = flip ? 8 - (dh * 3 + dw) : dh * 3 + dw;
                float4 k = __ldg(reinterpret_cast<const float4*>(w9) + tap * C4 + c);
                float4 v = reinterpret_cast<const float4*>(x)[((f * H + hh) * W + ww) * C4 + c];
                acc.x = fmaf(k.x, v.x, acc.x); acc.y = fmaf(k.y, v.y, acc.y);
                acc.z = fmaf(k.z, v.z, acc.z); acc.w = fmaf(k.w, v.w, acc.w);
            }
        }
        reinterpret_cast<float4*>(y)[i] = acc;
    }
}

// dW9[t][c] += sum_{f,h,w} dy[f,h,w,c] * x[f,h+dh-1,w+dw-1,c] ; dbias[c] += sum dy.  Thread per channel, a chunk of
// frames per blockIdx.y.
__global__ void __launch_bounds__(128) dwconv3x3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              float* __restrict__ dw9, float* __restrict__ dbias, int F, int H, int W,
                                                              int ch, int frames_per_block) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ch) return;
    const int f0 = blockIdx.y * frames_per_block;
    const int f1 = min(f0 + frames_per_block, F);
    float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float ab = 0.f;
    for (int f = f0; f < f1; ++f)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const float g = dy[(((long long)f * H + h) * W + w) * ch + c];
                ab += g;
#pragma unroll
                for (int dh = 0; dh < 3; ++dh) {
                    const int hh = h + dh - 1;
                    if (hh < 0 || hh >= H) continue;
#pragma unroll
                    for (int dw = 0; dw < 3; ++dw) {
                        const int ww = w + dw - 1;
                        if (ww < 0 || ww >= W) continue;
                        acc[dh * 3 + dw] = fmaf(g, x[(((long long)f * H + hh) * W + ww) * ch + c], acc[dh * 3 + dw]);
                    }
                }
            }
#pragma unroll
    for (int t = 0; t < 9; ++t) atomicAdd(dw9 + t * ch + c, acc[t]);
    atomicAdd(dbias + c, ab);
}

}  // namespace

extern "C" int vptr_dwconv3x3(const float* x, const float* w9, const float* bias, float* y, int F, int H, int W, int ch, int flip,
                              cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && H > 0 && W > 0 && ch > 0 && ch % 4 == 0, VPTR_ERR_SHAPE, "vptr_dwconv3x3: F=%d H=%d W=%d ch=%d", F, H, W, ch);
    long long total4 = (long long)F * H * W * (ch / 4);
    dwconv3x3_kernel<<<ew_grid(total4, 256), 256, 0, stream>>>(x, w9, bias, y, total4, H, W, ch / 4, flip);
    return vptr_check_launch("dwconv3x3_kernel");
}

extern "C" int vptr_dwconv3x3_wgrad(const float* x, const float* dy, float* dw9, float* dbias, int F, int H, int W, int ch,
                                    cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && H > 0 && W > 0 && ch > 0, VPTR_ERR_SHAPE, "vptr_dwconv3x3_wgrad: F=%d H=%d W=%d ch=%d", F, H, W, ch);
    int fpb = (256 + H * W - 1) / (H * W);  // ~256 pixels per block
    if (fpb < 1) fpb = 1;
    dim3 grid(vptr_cdiv(ch, 128), vptr_cdiv(F, fpb));
    dwconv3x3_wgrad_kernel<<<grid, 128, 0, stream>>>(x, dy, dw9, dbias, F, H, W, ch, fpb);
    return vptr_check_launch("dwconv3x3_wgrad_kernel");
}
