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
// erf(|z|) tail: 1 - erf(|z|) = poly(t) * exp(-z^2), t = 1/(1 + p|z|) (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7 + fp32
// rounding ~ 6e-7 absolute): one MUFU.RCP + one MUFU.EX2 + 6 FMA, branch-free.  libdevice's erff is two divergent branches of
// ~40 instructions; with four of them per float4 the norm+GELU kernels were ALU-bound, not HBM-bound.  The resulting GELU
// differs from the erff one by < 3e-7 absolute (2e-8 relative L2), far inside the parity gates.  `e` returns exp(-z^2).
__device__ __forceinline__ float vptr_erfc_abs(float az, float& e) {
    const float t = __frcp_rn(fmaf(0.3275911f, az, 1.f));
    const float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
    e = __expf(-az * az);
    return poly * e;
}
__device__ __forceinline__ float vptr_gelu(float x) {
    // exact-erf GELU, as torch.nn.GELU() default: 0.5 x (1 + erf(x / sqrt 2))
    float e;
    const float c = vptr_erfc_abs(fabsf(x) * 0.70710678118654752440f, e);     // 1 - erf(|z|)
    const float cdf = x >= 0.f ? 1.f - 0.5f * c : 0.5f * c;
    return x * cdf;
}
__device__ __forceinline__ float vptr_gelu_grad(float x) {
    const float kInvSqrt2Pi = 0.39894228040143267794f;
    float e;                                                                  // = exp(-x^2 / 2): shared with the pdf term
    const float c = vptr_erfc_abs(fabsf(x) * 0.70710678118654752440f, e);
    const float cdf = x >= 0.f ? 1.f - 0.5f * c : 0.5f * c;
    return cdf + x * kInvSqrt2Pi * e;
}
// round to nearest tf32 (ties away from zero, like cvt.rna.tf32.f32): add half an ulp of the 10-bit mantissa and truncate.
// cvt.rna.tf32.f32 itself is EMULATED on sm_100 with four instructions (VIADD, FSETP |x| < inf, SEL, LOP3); the two-instruction
// form below differs only for NaN payloads and for finite values within half a tf32 ulp of FLT_MAX (which round to inf).
__device__ __forceinline__ float vptr_round_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// Counter-based RNG for dropout / DropPath: ONE splitmix64 hash per group of four consecutive elements, 16 random bits per
// element (the fused sites are float4-vectorised, and a hash per element made the 4-warp GEMM epilogue compute-bound).
// Stateless, so the backward pass regenerates the very mask the forward used instead of storing it.
// g_vptr_rng_epoch: a per-translation-unit device counter mixed into every seed.  It stays 0 in eager use (the host draws fresh seeds
// per step); when a whole training step is replayed as a CUDA graph the kernel arguments -- seeds included -- are frozen, so the
// graph starts with vptr_rng_advance(), which bumps the epoch of every translation unit on the device: each replay draws new masks,
// and the forward and backward of one replay still agree because the epoch only changes between steps.
static __device__ unsigned long long g_vptr_rng_epoch = 0;
#define VPTR_RNG_EPOCH_ACCESSOR(tu)                                                   \
    extern "C" unsigned long long* vptr_rng_epoch_addr_##tu(void) {                   \
        unsigned long long* p = nullptr;                                              \
        cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_vptr_rng_epoch);         \
        return p;                                                                     \
    }
__device__ __forceinline__ unsigned long long vptr_hash4(unsigned long long seed, unsigned long long idx4) {
    unsigned long long z = idx4 * 0x9E3779B97F4A7C15ULL + seed + g_vptr_rng_epoch * 0xD1B54A32D192ED03ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned vptr_drop_threshold(float p) { return (unsigned)(p * 65536.f); }
// dropout keep-scale of element idx: 0 with probability p, else 1/(1-p)  (p == 0 -> 1)
__device__ __forceinline__ float vptr_drop_scale(unsigned long long seed, unsigned long long idx, float p) {
    if (p <= 0.f) return 1.f;
    const unsigned r = (unsigned)(vptr_hash4(seed, idx >> 2) >> (16 * (unsigned)(idx & 3))) & 0xFFFFu;
    return r >= vptr_drop_threshold(p) ? 1.f / (1.f - p) : 0.f;
}
// keep-scales of elements 4*idx4 .. 4*idx4+3 (identical to four vptr_drop_scale calls, one hash)
__device__ __forceinline__ float4 vptr_drop_scale4(unsigned long long seed, unsigned long long idx4, float p) {
    if (p <= 0.f) return make_float4(1.f, 1.f, 1.f, 1.f);
    const unsigned long long z = vptr_hash4(seed, idx4);
    const unsigned thr = vptr_drop_threshold(p), lo = (unsigned)z, hi = (unsigned)(z >> 32);
    const float inv = 1.f / (1.f - p);
    return make_float4((lo & 0xFFFFu) >= thr ? inv : 0.f, (lo >> 16) >= thr ? inv : 0.f, (hi & 0xFFFFu) >= thr ? inv : 0.f,
                       (hi >> 16) >= thr ? inv : 0.f);
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
