// Attention cores of VidHRFormer on token-major activations (rows ordered (n, t, h, w), C = nhead*d columns):
//   mode 0 -- spatial local-window attention with learned relative-position bias
//             (SpatialLocalMultiheadAttention VidHRFormer_modules.py:321-357 + MultiheadAttentionRPE
//             MultiHeadAttentionRPE.py:586-590,623,635-650,677-686): batch b = window (f, qh, qw),
//             positions l = ph*ws+pw; the window gather/scatter ("n (qh ph) (qw pw) c -> (ph pw) (n qh qw) c",
//             VidHRFormer_modules.py:503-525) is pure index arithmetic here -- no permuted copy exists.
//   mode 1 -- temporal attention per pixel (nn.MultiheadAttention call sites VidHRFormer_modules.py:79-84,
//             185-187, 204-205): batch b = (n, h, w), query positions t over Tq, key positions over Tk,
//             optional causal mask (key j > query i -> -inf, :78).
// Q is taken unscaled; scores are scale*(Q.K) + bias, which equals the reference's (q*scale).k + bias.
#include "common.cuh"

namespace {

struct AttnGeom {
    int mode;
    int H, W, ws, nwh, nww;  // window mode
    int Tq, Tk, HW;          // temporal mode
    int Lq, Lk;
    int nhead, d;
    int causal;
    float scale;
    int round_tf32;  // round stored outputs (they only feed tf32 contractions)
    unsigned long long drop_seed;  // attention-probability dropout (MultiHeadAttentionRPE.py:678 / nn.MultiheadAttention)
    float drop_p;
};

// dropout keep-scale of probability (b, h, i, j)
__device__ __forceinline__ float prob_drop(const AttnGeom& g, int b, int h, int i, int j) {
    return vptr_drop_scale(g.drop_seed, (((unsigned long long)b * g.nhead + h) * g.Lq + i) * g.Lk + j, g.drop_p);
}

__device__ __forceinline__ long long window_row(const AttnGeom& g, int b, int l) {
    const int per = g.nwh * g.nww;
    const int f = b / per, r = b - f * per;
    const int qh = r / g.nww, qw = r - qh * g.nww;
    const int ph = l / g.ws, pw = l - ph * g.ws;
    return ((long long)f * g.H + qh * g.ws + ph) * g.W + qw * g.ws + pw;
}
__device__ __forceinline__ long long q_row(const AttnGeom& g, int b, int i) {
    if (g.mode == 0) return window_row(g, b, i);
    const int n = b / g.HW, p = b - n * g.HW;
    return ((long long)n * g.Tq + i) * g.HW + p;
}
__device__ __forceinline__ long long k_row(const AttnGeom& g, int b, int j) {
    if (g.mode == 0) return window_row(g, b, j);
    const int n = b / g.HW, p = b - n * g.HW;
    return ((long long)n * g.Tk + j) * g.HW + p;
}
// relative_position_index[i][j] (MultiHeadAttentionRPE.py:373-387)
__device__ __forceinline__ int rel_pos_index(int ws, int i, int j) {
    const int ih = i / ws, iw = i - ih * ws, jh = j / ws, jw = j - jh * ws;
    return (ih - jh + ws - 1) * (2 * ws - 1) + (iw - jw + ws - 1);
}

// scores + softmax into sS[Lq][Lk+1]; all threads participate
__device__ __forceinline__ void scores_softmax(const AttnGeom& g, int h, const float* sQ, const float* sK, float* sS,
                                               const float* __restrict__ rpe_table) {
    const int dp = g.d + 1, lp = g.Lk + 1;
    for (int e = threadIdx.x; e < g.Lq * g.Lk; e += blockDim.x) {
        const int i = e / g.Lk, j = e - i * g.Lk;
        float s = 0.f;
        const float* q = sQ + i * dp;
        const float* k = sK + j * dp;
        for (int c = 0; c < g.d; ++c) s = fmaf(q[c], k[c], s);
        s *= g.scale;
        if (rpe_table) s += __ldg(rpe_table + rel_pos_index(g.ws, i, j) * g.nhead + h);
        if (g.causal && j > i) s = -INFINITY;
        sS[i * lp + j] = s;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = w; i < g.Lq; i += nw) {
        float m = -INFINITY;
        for (int j = lane; j < g.Lk; j += 32) m = fmaxf(m, sS[i * lp + j]);
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < g.Lk; j += 32) {
            float p = __expf(sS[i * lp + j] - m);
            sS[i * lp + j] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int j = lane; j < g.Lk; j += 32) sS[i * lp + j] *= inv;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(128) attn_fwd_kernel(const float* __restrict__ Q, long long ldq, const float* __restrict__ K,
                                                       long long ldk, const float* __restrict__ V, long long ldv,
                                                       float* __restrict__ O, long long ldo, const float* __restrict__ rpe_table,
                                                       const AttnGeom g, int batches) {
    extern __shared__ float sm[];
    const int dp = g.d + 1, lp = g.Lk + 1;
    float* sQ = sm;
    float* sK = sQ + g.Lq * dp;
    float* sV = sK + g.Lk * dp;
    float* sS = sV + g.Lk * dp;
    const int h = blockIdx.y;
    const int col0 = h * g.d;
    for (int b = blockIdx.x; b < batches; b += gridDim.x) {
        for (int e = threadIdx.x; e < g.Lq * g.d; e += blockDim.x) {
            const int i = e / g.d, c = e - i * g.d;
            sQ[i * dp + c] = Q[q_row(g, b, i) * ldq + col0 + c];
        }
        for (int e = threadIdx.x; e < g.Lk * g.d; e += blockDim.x) {
            const int j = e / g.d, c = e - j * g.d;
            const long long r = k_row(g, b, j);
            sK[j * dp + c] = K[r * ldk + col0 + c];
            sV[j * dp + c] = V[r * ldv + col0 + c];
        }
        __syncthreads();
        scores_softmax(g, h, sQ, sK, sS, rpe_table);
        if (g.drop_p > 0.f) {
            for (int e = threadIdx.x; e < g.Lq * g.Lk; e += blockDim.x) {
                const int i = e / g.Lk, j = e - i * g.Lk;
                sS[i * lp + j] *= prob_drop(g, b, h, i, j);
            }
            __syncthreads();
        }
        for (int e = threadIdx.x; e < g.Lq * g.d; e += blockDim.x) {
            const int i = e / g.d, c = e - i * g.d;
            float o = 0.f;
            for (int j = 0; j < g.Lk; ++j) o = fmaf(sS[i * lp + j], sV[j * dp + c], o);
            O[q_row(g, b, i) * ldo + col0 + c] = g.round_tf32 ? vptr_round_tf32(o) : o;
        }
        __syncthreads();
    }
}

// Backward: recomputes P from Q,K; writes dQ,dK,dV (disjoint per (b,h)); accumulates the bias gradient over the
// batches this CTA visits in shared memory and flushes it with one atomic per (i,j) at the end.
__global__ void __launch_bounds__(128) attn_bwd_kernel(const float* __restrict__ Q, long long ldq, const float* __restrict__ K,
                                                       long long ldk, const float* __restrict__ V, long long ldv,
                                                       const float* __restrict__ dO, long long ldo, float* __restrict__ dQ,
                                                       long long lddq, float* __restrict__ dK, long long lddk,
                                                       float* __restrict__ dV, long long lddv,
                                                       const float* __restrict__ rpe_table, float* __restrict__ d_rpe_table,
                                                       const AttnGeom g, int batches) {
    extern __shared__ float sm[];
    const int dp = g.d + 1, lp = g.Lk + 1;
    float* sQ = sm;
    float* sK = sQ + g.Lq * dp;
    float* sV = sK + g.Lk * dp;
    float* sdO = sV + g.Lk * dp;
    float* sP = sdO + g.Lq * dp;
    float* sdS = sP + g.Lq * lp;
    float* sdB = sdS + g.Lq * lp;  // [Lq][Lk] bias-gradient accumulator (only when d_rpe_table)
    float* sM = sdB + g.Lq * g.Lk;  // [Lq][Lk] dropout keep-scales of this (b, h) (only when drop_p > 0)
    const int h = blockIdx.y;
    const int col0 = h * g.d;
    if (d_rpe_table)
        for (int e = threadIdx.x; e < g.Lq * g.Lk; e += blockDim.x) sdB[e] = 0.f;
    for (int b = blockIdx.x; b < batches; b += gridDim.x) {
        for (int e = threadIdx.x; e < g.Lq * g.d; e += blockDim.x) {
            const int i = e / g.d, c = e - i * g.d;
            const long long r = q_row(g, b, i);
            sQ[i * dp + c] = Q[r * ldq + col0 + c];
            sdO[i * dp + c] = dO[r * ldo + col0 + c];
        }
        for (int e = threadIdx.x; e < g.Lk * g.d; e += blockDim.x) {
            const int j = e / g.d, c = e - j * g.d;
            const long long r = k_row(g, b, j);
            sK[j * dp + c] = K[r * ldk + col0 + c];
            sV[j * dp + c] = V[r * ldv + col0 + c];
        }
        __syncthreads();
        scores_softmax(g, h, sQ, sK, sP, rpe_table);
        // dP = (dO V^T) * keep-scale   (gradient w.r.t. the undropped probabilities)
        const bool drop = g.drop_p > 0.f;
        for (int e = threadIdx.x; e < g.Lq * g.Lk; e += blockDim.x) {
            const int i = e / g.Lk, j = e - i * g.Lk;
            float s = 0.f;
            for (int c = 0; c < g.d; ++c) s = fmaf(sdO[i * dp + c], sV[j * dp + c], s);
            if (drop) {
                const float m = prob_drop(g, b, h, i, j);
                sM[e] = m;
                s *= m;
            }
            sdS[i * lp + j] = s;
        }
        __syncthreads();
        // dS = P * (dP - rowsum(P*dP))
        {
            const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
            for (int i = w; i < g.Lq; i += nw) {
                float t = 0.f;
                for (int j = lane; j < g.Lk; j += 32) t = fmaf(sP[i * lp + j], sdS[i * lp + j], t);
                t = warp_sum(t);
                for (int j = lane; j < g.Lk; j += 32) {
                    float ds = sP[i * lp + j] * (sdS[i * lp + j] - t);
                    sdS[i * lp + j] = ds;
                    if (d_rpe_table) sdB[i * g.Lk + j] += ds;
                }
            }
        }
        __syncthreads();
        // dV = P^T dO ; dK = scale * dS^T Q
        for (int e = threadIdx.x; e < g.Lk * g.d; e += blockDim.x) {
            const int j = e / g.d, c = e - j * g.d;
            float dv = 0.f, dk = 0.f;
            for (int i = 0; i < g.Lq; ++i) {
                dv = fmaf(drop ? sP[i * lp + j] * sM[i * g.Lk + j] : sP[i * lp + j], sdO[i * dp + c], dv);
                dk = fmaf(sdS[i * lp + j], sQ[i * dp + c], dk);
            }
            const long long r = k_row(g, b, j);
            dk *= g.scale;
            dV[r * lddv + col0 + c] = g.round_tf32 ? vptr_round_tf32(dv) : dv;
            dK[r * lddk + col0 + c] = g.round_tf32 ? vptr_round_tf32(dk) : dk;
        }
        // dQ = scale * dS K
        for (int e = threadIdx.x; e < g.Lq * g.d; e += blockDim.x) {
            const int i = e / g.d, c = e - i * g.d;
            float dq = 0.f;
            for (int j = 0; j < g.Lk; ++j) dq = fmaf(sdS[i * lp + j], sK[j * dp + c], dq);
            dq *= g.scale;
            dQ[q_row(g, b, i) * lddq + col0 + c] = g.round_tf32 ? vptr_round_tf32(dq) : dq;
        }
        __syncthreads();
    }
    if (d_rpe_table) {
        for (int e = threadIdx.x; e < g.Lq * g.Lk; e += blockDim.x) {
            const int i = e / g.Lk, j = e - i * g.Lk;
            atomicAdd(d_rpe_table + rel_pos_index(g.ws, i, j) * g.nhead + h, sdB[e]);
        }
    }
}

__global__ void index_maps_kernel(int F, int H, int W, int ws, long long* rpi, long long* wmap) {
    AttnGeom g{};
    g.mode = 0; g.H = H; g.W = W; g.ws = ws; g.nwh = H / ws; g.nww = W / ws;
    const int L = ws * ws;
    const int B = F * g.nwh * g.nww;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < L * L; e += gridDim.x * blockDim.x)
        if (rpi) rpi[e] = rel_pos_index(ws, e / L, e % L);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < L * B; e += gridDim.x * blockDim.x)
        if (wmap) wmap[e] = window_row(g, e % B, e / B);  // layout (L, B) like the reference's permuted tensor
}

__global__ void causal_mask_kernel(int T, unsigned char* mask) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < T * T; e += gridDim.x * blockDim.x) mask[e] = (e % T) > (e / T);
}

int fill_geom(AttnGeom& g, int mode, int F_or_N, int H, int W, int ws, int Tq, int Tk, int nhead, int d, int causal, float scale, int* batches) {
    g = AttnGeom{};
    g.mode = mode; g.nhead = nhead; g.d = d; g.causal = causal; g.scale = scale;
    if (mode == 0) {
        VPTR_REQUIRE(ws > 0 && H % ws == 0 && W % ws == 0, VPTR_ERR_SHAPE, "window attention: H=%d W=%d not multiples of ws=%d (pad first)", H, W, ws);
        g.H = H; g.W = W; g.ws = ws; g.nwh = H / ws; g.nww = W / ws; g.Lq = g.Lk = ws * ws;
        *batches = F_or_N * g.nwh * g.nww;
    } else {
        VPTR_REQUIRE(Tq > 0 && Tk > 0, VPTR_ERR_SHAPE, "temporal attention: Tq=%d Tk=%d", Tq, Tk);
        VPTR_REQUIRE(!causal || Tq == Tk, VPTR_ERR_SHAPE, "causal temporal attention needs Tq == Tk");
        g.Tq = Tq; g.Tk = Tk; g.HW = H * W; g.Lq = Tq; g.Lk = Tk;
        *batches = F_or_N * H * W;
    }
    return VPTR_OK;
}

}  // namespace

// mode 0: F_or_N = number of frames (N*T); mode 1: F_or_N = number of clips N.
extern "C" int vptr_attn_fwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv, float* O,
                             long long ldo, const float* rpe_table, int mode, int F_or_N, int H, int W, int ws, int Tq, int Tk,
                             int nhead, int d, int causal, float scale, int round_tf32, unsigned long long drop_seed, float drop_p,
                             cudaStream_t stream) {
    AttnGeom g;
    int batches = 0;
    int rc = fill_geom(g, mode, F_or_N, H, W, ws, Tq, Tk, nhead, d, causal, scale, &batches);
    if (rc) return rc;
    g.round_tf32 = round_tf32;
    g.drop_seed = drop_seed;
    g.drop_p = drop_p;
    VPTR_REQUIRE(batches > 0 && nhead > 0 && d > 0, VPTR_ERR_SHAPE, "vptr_attn_fwd: empty problem");
    size_t smem = sizeof(float) * ((size_t)(g.Lq + 2 * g.Lk) * (d + 1) + (size_t)g.Lq * (g.Lk + 1));
    VPTR_REQUIRE(smem <= 200 * 1024, VPTR_ERR_UNSUPPORTED, "vptr_attn_fwd: tile too large (%zu B of shared memory)", smem);
    if (smem > 48 * 1024) cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(batches < 65535 * 8 ? batches : 65535 * 8, nhead);
    attn_fwd_kernel<<<grid, 128, smem, stream>>>(Q, ldq, K, ldk, V, ldv, O, ldo, rpe_table, g, batches);
    return vptr_check_launch("attn_fwd_kernel");
}

extern "C" int vptr_attn_bwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv,
                             const float* dO, long long ldo, float* dQ, long long lddq, float* dK, long long lddk, float* dV,
                             long long lddv, const float* rpe_table, float* d_rpe_table, int mode, int F_or_N, int H, int W, int ws,
                             int Tq, int Tk, int nhead, int d, int causal, float scale, int round_tf32, unsigned long long drop_seed,
                             float drop_p, cudaStream_t stream) {
    AttnGeom g;
    int batches = 0;
    int rc = fill_geom(g, mode, F_or_N, H, W, ws, Tq, Tk, nhead, d, causal, scale, &batches);
    if (rc) return rc;
    g.round_tf32 = round_tf32;
    g.drop_seed = drop_seed;
    g.drop_p = drop_p;
    VPTR_REQUIRE(batches > 0 && nhead > 0 && d > 0, VPTR_ERR_SHAPE, "vptr_attn_bwd: empty problem");
    size_t smem = sizeof(float) * ((size_t)(2 * g.Lq + 2 * g.Lk) * (d + 1) + (size_t)2 * g.Lq * (g.Lk + 1) + (size_t)2 * g.Lq * g.Lk);
    VPTR_REQUIRE(smem <= 200 * 1024, VPTR_ERR_UNSUPPORTED, "vptr_attn_bwd: tile too large (%zu B of shared memory)", smem);
    if (smem > 48 * 1024) cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int gx = batches;
    if (d_rpe_table && gx > 148 * 4) gx = 148 * 4;  // bound the number of bias-gradient flushes
    dim3 grid(gx, nhead);
    attn_bwd_kernel<<<grid, 128, smem, stream>>>(Q, ldq, K, ldk, V, ldv, dO, ldo, dQ, lddq, dK, lddk, dV, lddv, rpe_table,
                                                 d_rpe_table, g, batches);
    return vptr_check_launch("attn_bwd_kernel");
}

// Integer artefacts produced by the very index functions the kernels use (bit-exact contract):
// rpi int64 [L][L]; wmap int64 [L][B] flat token index of (position l, window b); either may be NULL.
extern "C" int vptr_window_index_maps(int F, int H, int W, int ws, long long* rpi, long long* wmap, cudaStream_t stream) {
    VPTR_REQUIRE(ws > 0 && H % ws == 0 && W % ws == 0, VPTR_ERR_SHAPE, "vptr_window_index_maps: H=%d W=%d ws=%d", H, W, ws);
    index_maps_kernel<<<64, 256, 0, stream>>>(F, H, W, ws, rpi, wmap);
    return vptr_check_launch("index_maps_kernel");
}

extern "C" int vptr_causal_mask(int T, unsigned char* mask, cudaStream_t stream) {
    causal_mask_kernel<<<8, 256, 0, stream>>>(T, mask);
    return vptr_check_launch("causal_mask_kernel");
}
