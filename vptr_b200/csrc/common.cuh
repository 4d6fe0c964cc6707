// Shared device/host helpers for libvptr_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define VPTR_OK 0
#define VPTR_ERR_SHAPE (-1)
#define VPTR_ERR_ALIGN (-2)
#define VPTR_ERR_UNSUPPORTED (-3)
#define VPTR_ERR_DRIVER (-4)

void vptr_set_error(const char* fmt, ...);
int vptr_check_launch(const char* what);   // cudaGetLastError -> status (+ message)

#define VPTR_REQUIRE(cond, code, ...)                     \
    do {                                                  \
        if (!(cond)) {                                    \
            vptr_set_error(__VA_ARGS__);                  \
            return (code);                                \
        }                                                 \
    } while (0)

static inline int vptr_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
__device__ __forceinline__ float vptr_gelu(float x) {
    // exact (erf) GELU, as torch.nn.GELU() default
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float vptr_gelu_grad(float x) {
    const float kInvSqrt2Pi = 0.39894228040143267794f;
    float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    return cdf + x * kInvSqrt2Pi * __expf(-0.5f * x * x);
}
__device__ __forceinline__ float vptr_round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// Counter-based RNG for dropout / DropPath: splitmix64 of (seed, element index) -> uniform [0,1).  Stateless, so the
// backward pass regenerates the very mask the forward used instead of storing it.
__device__ __forceinline__ float vptr_uniform(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = idx * 0x9E3779B97F4A7C15ULL + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}
// dropout keep-scale of element idx: 0 with probability p, else 1/(1-p)  (p == 0 -> 1)
__device__ __forceinline__ float vptr_drop_scale(unsigned long long seed, unsigned long long idx, float p) {
    if (p <= 0.f) return 1.f;
    return vptr_uniform(seed, idx) >= p ? 1.f / (1.f - p) : 0.f;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum of a float; `red` is >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
#endif
