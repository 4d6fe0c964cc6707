// Plain fp32 FFMA GEMM with the same contract as vptr_gemm_tf32 (gemm_tcgen05.cu).  It exists for
// (1) shapes the TMA path cannot address (pitches that are not multiples of 16 bytes, N % 4 != 0),
// and (2) as the on-device fp32 cross-check of the tensor-core kernel in tests/.  It is not a
// fallback for a missing tensor-core path: both live in the same library.
#include "common.cuh"

namespace {
constexpr int TM = 64, TN = 64, TK = 16;

struct SimtParams {
    const float* A; long long sa_m, sa_k;
    const float* B; long long sb_n, sb_k;
    float* D; long long ldd;
    int M, N, K;
    const float* bias; const float* residual; long long ldr;
    float alpha; int act; int flags;
    const float* rowscale; int rows_per_group; unsigned long long drop_seed; float drop_p;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtParams p) {
    __shared__ float sA[TK][TM + 1];
    __shared__ float sB[TK][TN + 1];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < p.K; k0 += TK) {
        for (int e = tid; e < TM * TK; e += 256) {
            int mm, kk;
            if (p.sa_k == 1) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
            int m = m0 + mm, k = k0 + kk;
            sA[kk][mm] = (m < p.M && k < p.K) ? p.A[m * p.sa_m + k * p.sa_k] : 0.f;
        }
        for (int e = tid; e < TN * TK; e += 256) {
            int nn, kk;
            if (p.sb_k == 1) { kk = e % TK; nn = e / TK; } else { nn = e % TN; kk = e / TN; }
            int n = n0 + nn, k = k0 + kk;
            sB[kk][nn] = (n < p.N && k < p.K) ? p.B[n * p.sb_n + k * p.sb_k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float o = acc[i][j] * p.alpha;
            if (p.bias) o += p.bias[n];
            if (p.act == 1) o = vptr_gelu(o);
            else if (p.act == 2) o = fmaxf(o, 0.f);
            if (p.drop_p > 0.f) o *= vptr_drop_scale(p.drop_seed, (unsigned long long)m * p.N + n, p.drop_p);
            if (p.rowscale) o *= p.rowscale[m / p.rows_per_group];
            if (p.residual) o += p.residual[m * p.ldr + n];
            if (p.flags & 2) o = vptr_round_tf32(o);
            float* d = p.D + m * p.ldd + n;
            if (p.flags & 1) *d += o; else *d = o;
        }
    }
}
}  // namespace

extern "C" int vptr_gemm_simt(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, float* D,
                              long long ldd, int M, int N, int K, const float* bias, const float* residual, long long ldr,
                              float alpha, int act, int flags, int k_splits, const float* rowscale, int rows_per_group,
                              unsigned long long drop_seed, float drop_p, cudaStream_t stream) {
    (void)k_splits;
    VPTR_REQUIRE(M > 0 && N > 0 && K > 0, VPTR_ERR_SHAPE, "vptr_gemm_simt: empty problem M=%d N=%d K=%d", M, N, K);
    SimtParams p;
    p.A = A; p.sa_m = a_mn ? 1 : lda; p.sa_k = a_mn ? lda : 1;
    p.B = B; p.sb_n = b_mn ? 1 : ldb; p.sb_k = b_mn ? ldb : 1;
    p.D = D; p.ldd = ldd; p.M = M; p.N = N; p.K = K;
    p.bias = bias; p.residual = residual; p.ldr = ldr; p.alpha = alpha; p.act = act; p.flags = flags;
    p.rowscale = rowscale; p.rows_per_group = rows_per_group; p.drop_seed = drop_seed; p.drop_p = drop_p;
    dim3 grid(vptr_cdiv(N, TN), vptr_cdiv(M, TM));
    gemm_simt_kernel<<<grid, 256, 0, stream>>>(p);
    return vptr_check_launch("gemm_simt_kernel");
}

VPTR_RNG_EPOCH_ACCESSOR(gemm_simt)
