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
#include <stdlib.h>

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
    float *dbq, *dbk, *dbv;  // backward, tensor-core paths: column sums of dQ / dK / dV are added here ([nhead*d] each; NULL = off)
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
// mode 2 (TemporalSpatialLocalMultiheadAttention, model/VidHRFormer_modules.py:219-284,444-484): batch entry b = (clip n, future
// frame t2, window): its ws^2 queries are that window of frame n*Tq + t2 (exactly mode 0's rows over N*Tq frames) and its Tk*ws^2
// keys are the same window of ALL Tk memory frames of clip n, key j = (t1, in-window position l) = (j / ws^2, j % ws^2).  The
// reference attends all Tq*ws^2 queries of a window in one sequence; softmax is per query, so splitting them by frame is exact.
__device__ __forceinline__ long long q_row(const AttnGeom& g, int b, int i) {
    if (g.mode != 1) return window_row(g, b, i);
    const int n = b / g.HW, p = b - n * g.HW;
    return ((long long)n * g.Tq + i) * g.HW + p;
}
__device__ __forceinline__ long long k_row(const AttnGeom& g, int b, int j) {
    if (g.mode == 0) return window_row(g, b, j);
    if (g.mode == 2) {
        const int per = g.nwh * g.nww, L = g.ws * g.ws;
        const int fq = b / per, r = b - fq * per, n = fq / g.Tq;
        const int t1 = j / L, l = j - t1 * L;
        const int qh = r / g.nww, qw = r - qh * g.nww, ph = l / g.ws, pw = l - ph * g.ws;
        return (((long long)n * g.Tk + t1) * g.H + qh * g.ws + ph) * g.W + qw * g.ws + pw;
    }
    const int n = b / g.HW, p = b - n * g.HW;
    return ((long long)n * g.Tk + j) * g.HW + p;
}
// relative_position_index[i][j] (MultiHeadAttentionRPE.py:373-387)
__device__ __forceinline__ int rel_pos_index(int ws, int i, int j) {
    const int ih = i / ws, iw = i - ih * ws, jh = j / ws, jw = j - jh * ws;
    return (ih - jh + ws - 1) * (2 * ws - 1) + (iw - jw + ws - 1);
}

// Shared tiles use a row pitch dp = (d rounded up to 4) + 4 floats: rows stay 16-byte aligned for float4 access and the
// quarter-warp bank pattern is conflict-free; pad columns are zero-filled so whole float4 dot products are exact.
__device__ __forceinline__ int tile_pitch(int d) { return ((d + 3) & ~3) + 4; }

// rows [0, L) x d of head h -> tile; one warp per row (row index computed once per row, float2 global loads: h*d*4 bytes is
// only 8-byte aligned for odd heads)
__device__ __forceinline__ void load_rows(float* tile, const float* __restrict__ src, long long ld, const AttnGeom& g, int b, int L,
                                          bool is_q, int col0, int dp) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int l = w; l < L; l += nw) {
        const float* r = src + (is_q ? q_row(g, b, l) : k_row(g, b, l)) * ld + col0;
        for (int c2 = lane; 2 * c2 < dp; c2 += 32) {
            float2 v = make_float2(0.f, 0.f);
            if (2 * c2 < g.d) v = __ldg(reinterpret_cast<const float2*>(r + 2 * c2));
            *reinterpret_cast<float2*>(tile + l * dp + 2 * c2) = v;
        }
    }
}
__device__ __forceinline__ void store_rows(const float* tile, float* __restrict__ dst, long long ld, const AttnGeom& g, int b, int L,
                                           bool is_q, int col0, int dp, float mul) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int l = w; l < L; l += nw) {
        float* r = dst + (is_q ? q_row(g, b, l) : k_row(g, b, l)) * ld + col0;
        for (int c2 = lane; 2 * c2 < g.d; c2 += 32) {
            float2 v = *reinterpret_cast<const float2*>(tile + l * dp + 2 * c2);
            v.x *= mul; v.y *= mul;
            if (g.mode == 2 && !is_q) {   // a memory token is a key of every future frame's batch entry: accumulate (caller zeroes dK / dV)
                atomicAdd(r + 2 * c2, v.x);
                atomicAdd(r + 2 * c2 + 1, v.y);
                continue;
            }
            if (g.round_tf32) { v.x = vptr_round_tf32(v.x); v.y = vptr_round_tf32(v.y); }
            *reinterpret_cast<float2*>(r + 2 * c2) = v;
        }
    }
}
__device__ __forceinline__ float dot_rows(const float* a, const float* b, int dp4) {
    float s = 0.f;
    for (int c = 0; c < dp4; ++c) {
        const float4 x = reinterpret_cast<const float4*>(a)[c], y = reinterpret_cast<const float4*>(b)[c];
        s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
    }
    return s;
}

// scores + softmax into sS[Lq][Lk+1]; all threads participate
__device__ __forceinline__ void scores_softmax(const AttnGeom& g, int h, const float* sQ, const float* sK, float* sS, int dp,
                                               const float* __restrict__ rpe_table) {
    const int lp = g.Lk + 1, dp4 = ((g.d + 3) & ~3) >> 2;
    for (int e = threadIdx.x; e < g.Lq * g.Lk; e += blockDim.x) {
        const int i = e / g.Lk, j = e - i * g.Lk;
        float s = dot_rows(sQ + i * dp, sK + j * dp, dp4) * g.scale;
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

// out[i][c4] = mul * sum_j W[i*lw + j] * T[j][c4]  (i < Li, j < Lj) -> staged into dstTile rows (float4 per thread)
__device__ __forceinline__ void weighted_rows(const float* W, int lw, bool transposed, const float* T, float* dstTile, int Li, int Lj,
                                              int dp, int d4) {
    for (int e = threadIdx.x; e < Li * d4; e += blockDim.x) {
        const int i = e / d4, c = e - i * d4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < Lj; ++j) {
            const float w = transposed ? W[j * lw + i] : W[i * lw + j];
            const float4 t = reinterpret_cast<const float4*>(T + j * dp)[c];
            acc.x = fmaf(w, t.x, acc.x); acc.y = fmaf(w, t.y, acc.y); acc.z = fmaf(w, t.z, acc.z); acc.w = fmaf(w, t.w, acc.w);
        }
        reinterpret_cast<float4*>(dstTile + i * dp)[c] = acc;
    }
}

__global__ void __launch_bounds__(128) attn_fwd_kernel(const float* __restrict__ Q, long long ldq, const float* __restrict__ K,
                                                       long long ldk, const float* __restrict__ V, long long ldv,
                                                       float* __restrict__ O, long long ldo, const float* __restrict__ rpe_table,
                                                       const AttnGeom g, int batches) {
    extern __shared__ __align__(16) float sm[];
    const int dp = tile_pitch(g.d), lp = g.Lk + 1, d4 = ((g.d + 3) & ~3) >> 2;
    float* sQ = sm;                       // also the output staging tile
    float* sK = sQ + g.Lq * dp;
    float* sV = sK + g.Lk * dp;
    float* sS = sV + g.Lk * dp;
    const int h = blockIdx.y;
    const int col0 = h * g.d;
    for (int b = blockIdx.x; b < batches; b += gridDim.x) {
        load_rows(sQ, Q, ldq, g, b, g.Lq, true, col0, dp);
        load_rows(sK, K, ldk, g, b, g.Lk, false, col0, dp);
        load_rows(sV, V, ldv, g, b, g.Lk, false, col0, dp);
        __syncthreads();
        scores_softmax(g, h, sQ, sK, sS, dp, rpe_table);
        if (g.drop_p > 0.f) {
            for (int e = threadIdx.x; e < g.Lq * g.Lk; e += blockDim.x) {
                const int i = e / g.Lk, j = e - i * g.Lk;
                sS[i * lp + j] *= prob_drop(g, b, h, i, j);
            }
            __syncthreads();
        }
        weighted_rows(sS, lp, false, sV, sQ, g.Lq, g.Lk, dp, d4);      // O = P V, staged over the Q tile
        __syncthreads();
        store_rows(sQ, O, ldo, g, b, g.Lq, true, col0, dp, 1.f);
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
    extern __shared__ __align__(16) float sm[];
    const int dp = tile_pitch(g.d), lp = g.Lk + 1, d4 = ((g.d + 3) & ~3) >> 2, dp4 = d4;
    float* sQ = sm;
    float* sK = sQ + g.Lq * dp;
    float* sV = sK + g.Lk * dp;
    float* sdO = sV + g.Lk * dp;
    float* sOut = sdO + g.Lq * dp;                 // staging tile for dV / dK / dQ: max(Lq, Lk) rows
    float* sP = sOut + max(g.Lq, g.Lk) * dp;       // P (undropped), later PD = P * keep-scale
    float* sdS = sP + g.Lq * lp;
    float* sdB = sdS + g.Lq * lp;                  // [Lq][Lk] bias-gradient accumulator
    const int h = blockIdx.y;
    const int col0 = h * g.d;
    if (d_rpe_table)
        for (int e = threadIdx.x; e < g.Lq * g.Lk; e += blockDim.x) sdB[e] = 0.f;
    for (int b = blockIdx.x; b < batches; b += gridDim.x) {
        load_rows(sQ, Q, ldq, g, b, g.Lq, true, col0, dp);
        load_rows(sdO, dO, ldo, g, b, g.Lq, true, col0, dp);
        load_rows(sK, K, ldk, g, b, g.Lk, false, col0, dp);
        load_rows(sV, V, ldv, g, b, g.Lk, false, col0, dp);
        __syncthreads();
        scores_softmax(g, h, sQ, sK, sP, dp, rpe_table);
        // dP = (dO V^T) * keep-scale   (gradient w.r.t. the undropped probabilities); PD = P * keep-scale
        const bool drop = g.drop_p > 0.f;
        for (int e = threadIdx.x; e < g.Lq * g.Lk; e += blockDim.x) {
            const int i = e / g.Lk, j = e - i * g.Lk;
            float s = dot_rows(sdO + i * dp, sV + j * dp, dp4);
            if (drop) s *= prob_drop(g, b, h, i, j);
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
                    const float pj = sP[i * lp + j];
                    const float ds = pj * (sdS[i * lp + j] - t);
                    sdS[i * lp + j] = ds;
                    if (drop) sP[i * lp + j] = pj * prob_drop(g, b, h, i, j);
                    if (d_rpe_table) sdB[i * g.Lk + j] += ds;
                }
            }
        }
        __syncthreads();
        weighted_rows(sP, lp, true, sdO, sOut, g.Lk, g.Lq, dp, d4);     // dV = PD^T dO
        __syncthreads();
        store_rows(sOut, dV, lddv, g, b, g.Lk, false, col0, dp, 1.f);
        __syncthreads();
        weighted_rows(sdS, lp, true, sQ, sOut, g.Lk, g.Lq, dp, d4);     // dK = scale * dS^T Q
        __syncthreads();
        store_rows(sOut, dK, lddk, g, b, g.Lk, false, col0, dp, g.scale);
        __syncthreads();
        weighted_rows(sdS, lp, false, sK, sOut, g.Lq, g.Lk, dp, d4);    // dQ = scale * dS K
        __syncthreads();
        store_rows(sOut, dQ, lddq, g, b, g.Lq, true, col0, dp, g.scale);
        __syncthreads();
    }
    if (d_rpe_table) {
        for (int e = threadIdx.x; e < g.Lq * g.Lk; e += blockDim.x) {
            const int i = e / g.Lk, j = e - i * g.Lk;
            atomicAdd(d_rpe_table + rel_pos_index(g.ws, i, j) * g.nhead + h, sdB[e]);
        }
    }
}

// =============================================================================================
// Main path: ONE CTA PER BATCH ENTRY, ALL HEADS.  The Lq (Lk) token rows of a window / pixel sequence are whole C-float rows
// in memory, so the CTA stages them with fully coalesced float4 loads (Q, K, V [, dO] tiles of [L][C+2] floats), then each
// warp owns one head: scores -> softmax -> P.V (and the backward products) entirely out of shared memory, with no block
// barrier between the load and the store phases.  Outputs are written back over dead tiles and stored as whole rows.
// Row pitch C+2 floats makes the 16 float2 accesses of a half-warp hit 32 distinct banks.
struct AhSmem {
    int Cp;        // row pitch (floats)
    float* q; float* k; float* v; float* go;   // tiles
    float* s;      // [nhead][Lq][Lk+1] probabilities
    float* ds;     // [nhead][Lq][Lk+1] (backward)
    float* db;     // [nhead][Lq][Lk]   (backward, bias-gradient accumulator)
    long long* rq; long long* rk;              // row indices of this batch entry
};

__device__ __forceinline__ void ah_load_tile(float* tile, int Cp, const float* __restrict__ src, long long ld, const long long* rows, int L,
                                             int C4) {
    for (int e = threadIdx.x; e < L * C4; e += blockDim.x) {
        const int l = e / C4, c = e - l * C4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + rows[l] * ld) + c);
        float2* d = reinterpret_cast<float2*>(tile + l * Cp + 4 * c);
        d[0] = make_float2(v.x, v.y);
        d[1] = make_float2(v.z, v.w);
    }
}
__device__ __forceinline__ void ah_store_tile(const float* tile, int Cp, float* __restrict__ dst, long long ld, const long long* rows, int L,
                                              int C4, float mul, int round_tf32) {
    for (int e = threadIdx.x; e < L * C4; e += blockDim.x) {
        const int l = e / C4, c = e - l * C4;
        const float2* sp = reinterpret_cast<const float2*>(tile + l * Cp + 4 * c);
        const float2 a = sp[0], b = sp[1];
        float4 v = make_float4(a.x * mul, a.y * mul, b.x * mul, b.y * mul);
        if (round_tf32) { v.x = vptr_round_tf32(v.x); v.y = vptr_round_tf32(v.y); v.z = vptr_round_tf32(v.z); v.w = vptr_round_tf32(v.w); }
        reinterpret_cast<float4*>(dst + rows[l] * ld)[c] = v;
    }
}
// warp-level: out[i*lp + j] = sum_c A[i][col+c] * B[j][col+c]  (i < Li, j < Lj).  Work item = (key row j, group of 4 query
// rows): the key element is loaded once for four independent accumulator chains (the plain one-pair-per-lane loop is latency
// bound: two dependent shared loads per FMA pair with only two warps per scheduler to hide them).
__device__ __forceinline__ void ah_dots(const float* A, const float* B, int Cp, int col, int Li, int Lj, int d2, float* out, int lp, int lane) {
    const int groups = (Li + 3) >> 2;
    for (int e = lane; e < Lj * groups; e += 32) {
        const int ig = e / Lj, j = e - ig * Lj;
        const int i0 = ig * 4;
        const float2* b = reinterpret_cast<const float2*>(B + j * Cp + col);
        const float2* a0 = reinterpret_cast<const float2*>(A + min(i0, Li - 1) * Cp + col);
        const float2* a1 = reinterpret_cast<const float2*>(A + min(i0 + 1, Li - 1) * Cp + col);
        const float2* a2 = reinterpret_cast<const float2*>(A + min(i0 + 2, Li - 1) * Cp + col);
        const float2* a3 = reinterpret_cast<const float2*>(A + min(i0 + 3, Li - 1) * Cp + col);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll 3
        for (int c = 0; c < d2; ++c) {
            const float2 y = b[c];
            const float2 x0 = a0[c], x1 = a1[c], x2 = a2[c], x3 = a3[c];
            s0 = fmaf(x0.x, y.x, s0); t0 = fmaf(x0.y, y.y, t0);
            s1 = fmaf(x1.x, y.x, s1); t1 = fmaf(x1.y, y.y, t1);
            s2 = fmaf(x2.x, y.x, s2); t2 = fmaf(x2.y, y.y, t2);
            s3 = fmaf(x3.x, y.x, s3); t3 = fmaf(x3.y, y.y, t3);
        }
        out[i0 * lp + j] = s0 + t0;
        if (i0 + 1 < Li) out[(i0 + 1) * lp + j] = s1 + t1;
        if (i0 + 2 < Li) out[(i0 + 2) * lp + j] = s2 + t2;
        if (i0 + 3 < Li) out[(i0 + 3) * lp + j] = s3 + t3;
    }
    __syncwarp();
}
// warp-level: S[i][j] = softmax_j(scale * q_i.k_j + bias) for head h (probabilities, undropped)
__device__ __forceinline__ void ah_scores_softmax(const AttnGeom& g, int h, const float* sq, const float* sk, int Cp, float* S,
                                                  const float* __restrict__ rpe_table, int lane) {
    const int lp = g.Lk + 1, d2 = g.d >> 1, col = h * g.d;
    ah_dots(sq, sk, Cp, col, g.Lq, g.Lk, d2, S, lp, lane);
    for (int e = lane; e < g.Lq * g.Lk; e += 32) {
        const int i = e / g.Lk, j = e - i * g.Lk;
        float s = S[i * lp + j] * g.scale;
        if (rpe_table) s += __ldg(rpe_table + rel_pos_index(g.ws, i, j) * g.nhead + h);
        if (g.causal && j > i) s = -INFINITY;
        S[i * lp + j] = s;
    }
    __syncwarp();
    for (int i = lane; i < g.Lq; i += 32) {          // one lane per row: rows are short (<= 64)
        float m = -INFINITY;
        for (int j = 0; j < g.Lk; ++j) m = fmaxf(m, S[i * lp + j]);
        float sum = 0.f;
        for (int j = 0; j < g.Lk; ++j) { const float p = __expf(S[i * lp + j] - m); S[i * lp + j] = p; sum += p; }
        const float inv = 1.f / sum;
        for (int j = 0; j < g.Lk; ++j) S[i * lp + j] *= inv;
    }
    __syncwarp();
}
// warp-level: out[i][col + c] = mul-free sum_j W(i,j) * T[j][col + c], i < Li, j < Lj; lane = float2 column, rows in registers chunks
template <bool TRANSPOSED>
__device__ __forceinline__ void ah_weighted(const float* W, int lw, const float* T, float* out, int Cp, int col, int Li, int Lj, int d2,
                                            int lane) {
    // work item = (float2 column c, group of 4 output rows): d2 = 33 columns do not divide the warp, so items (not columns) are
    // spread over the lanes -- with one column per lane, lane 0 alone ran a second pass and doubled the time
    const int groups = (Li + 3) >> 2;
    for (int e = lane; e < d2 * groups; e += 32) {
        const int rg = e / d2, c = e - rg * d2;
        const int i0 = rg * 4;
        float2 acc[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll 2
        for (int j = 0; j < Lj; ++j) {
            const float2 t = *reinterpret_cast<const float2*>(T + j * Cp + col + 2 * c);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = min(i0 + r, Li - 1);
                const float w = TRANSPOSED ? W[j * lw + i] : W[i * lw + j];
                acc[r].x = fmaf(w, t.x, acc[r].x);
                acc[r].y = fmaf(w, t.y, acc[r].y);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (i0 + r < Li) *reinterpret_cast<float2*>(out + (i0 + r) * Cp + col + 2 * c) = acc[r];
    }
}
__device__ __forceinline__ void ah_rows(const AttnGeom& g, int b, long long* rq, long long* rk) {
    for (int l = threadIdx.x; l < g.Lq; l += blockDim.x) rq[l] = q_row(g, b, l);
    for (int l = threadIdx.x; l < g.Lk; l += blockDim.x) rk[l] = k_row(g, b, l);
}

__global__ void __launch_bounds__(256) attn_fwd_allheads_kernel(const float* __restrict__ Q, long long ldq, const float* __restrict__ K,
                                                                long long ldk, const float* __restrict__ V, long long ldv,
                                                                float* __restrict__ O, long long ldo,
                                                                const float* __restrict__ rpe_table, const AttnGeom g, int batches) {
    extern __shared__ __align__(16) float sm[];
    const int C = g.nhead * g.d, C4 = C >> 2, Cp = C + 2, lp = g.Lk + 1, d2 = g.d >> 1;
    float* sq = sm;
    float* sk = sq + g.Lq * Cp;
    float* sv = sk + g.Lk * Cp;
    float* sS = sv + g.Lk * Cp;                                   // [nhead][Lq][lp]
    long long* rq = reinterpret_cast<long long*>(sS + ((g.nhead * g.Lq * lp + 1) & ~1));
    long long* rk = rq + g.Lq;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int b = blockIdx.x; b < batches; b += gridDim.x) {
        ah_rows(g, b, rq, rk);
        __syncthreads();
        ah_load_tile(sq, Cp, Q, ldq, rq, g.Lq, C4);
        ah_load_tile(sk, Cp, K, ldk, rk, g.Lk, C4);
        ah_load_tile(sv, Cp, V, ldv, rk, g.Lk, C4);
        __syncthreads();
        for (int h = warp; h < g.nhead; h += nwarps) {
            float* S = sS + h * g.Lq * lp;
            ah_scores_softmax(g, h, sq, sk, Cp, S, rpe_table, lane);
            if (g.drop_p > 0.f) {
                for (int e = lane; e < g.Lq * g.Lk; e += 32) {
                    const int i = e / g.Lk, j = e - i * g.Lk;
                    S[i * lp + j] *= prob_drop(g, b, h, i, j);
                }
                __syncwarp();
            }
            ah_weighted<false>(S, lp, sv, sq, Cp, h * g.d, g.Lq, g.Lk, d2, lane);     // O = P V over this head's Q columns
        }
        __syncthreads();
        ah_store_tile(sq, Cp, O, ldo, rq, g.Lq, C4, 1.f, g.round_tf32);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) attn_bwd_allheads_kernel(const float* __restrict__ Q, long long ldq, const float* __restrict__ K,
                                                                long long ldk, const float* __restrict__ V, long long ldv,
                                                                const float* __restrict__ dO, long long ldo, float* __restrict__ dQ,
                                                                long long lddq, float* __restrict__ dK, long long lddk,
                                                                float* __restrict__ dV, long long lddv,
                                                                const float* __restrict__ rpe_table, float* __restrict__ d_rpe_table,
                                                                const AttnGeom g, int batches) {
    extern __shared__ __align__(16) float sm[];
    const int C = g.nhead * g.d, C4 = C >> 2, Cp = C + 2, lp = g.Lk + 1, d2 = g.d >> 1;
    float* sq = sm;
    float* sgo = sq + g.Lq * Cp;
    float* sk = sgo + g.Lq * Cp;
    float* sv = sk + g.Lk * Cp;
    float* sP = sv + g.Lk * Cp;                                   // [nhead][Lq][lp]
    float* sdS = sP + g.nhead * g.Lq * lp;
    float* sdB = sdS + g.nhead * g.Lq * lp;                       // [nhead][Lq][Lk]
    long long* rq = reinterpret_cast<long long*>(sdB + ((g.nhead * g.Lq * g.Lk + 1) & ~1));
    long long* rk = rq + g.Lq;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    if (d_rpe_table)
        for (int e = threadIdx.x; e < g.nhead * g.Lq * g.Lk; e += blockDim.x) sdB[e] = 0.f;
    for (int b = blockIdx.x; b < batches; b += gridDim.x) {
        ah_rows(g, b, rq, rk);
        __syncthreads();
        ah_load_tile(sq, Cp, Q, ldq, rq, g.Lq, C4);
        ah_load_tile(sgo, Cp, dO, ldo, rq, g.Lq, C4);
        ah_load_tile(sk, Cp, K, ldk, rk, g.Lk, C4);
        ah_load_tile(sv, Cp, V, ldv, rk, g.Lk, C4);
        __syncthreads();
        for (int h = warp; h < g.nhead; h += nwarps) {
            float* P = sP + h * g.Lq * lp;
            float* dS = sdS + h * g.Lq * lp;
            const int col = h * g.d;
            ah_scores_softmax(g, h, sq, sk, Cp, P, rpe_table, lane);
            const bool drop = g.drop_p > 0.f;
            // dP = (dO V^T) * keep-scale
            ah_dots(sgo, sv, Cp, col, g.Lq, g.Lk, d2, dS, lp, lane);
            if (drop) {
                for (int e = lane; e < g.Lq * g.Lk; e += 32) {
                    const int i = e / g.Lk, j = e - i * g.Lk;
                    dS[i * lp + j] *= prob_drop(g, b, h, i, j);
                }
                __syncwarp();
            }
            // dS = P * (dP - rowsum(P*dP)); PD = P * keep-scale; bias-gradient accumulation
            for (int i = lane; i < g.Lq; i += 32) {
                float t = 0.f;
                for (int j = 0; j < g.Lk; ++j) t = fmaf(P[i * lp + j], dS[i * lp + j], t);
                for (int j = 0; j < g.Lk; ++j) {
                    const float pj = P[i * lp + j];
                    const float ds = pj * (dS[i * lp + j] - t);
                    dS[i * lp + j] = ds;
                    if (drop) P[i * lp + j] = pj * prob_drop(g, b, h, i, j);
                    if (d_rpe_table) sdB[(h * g.Lq + i) * g.Lk + j] += ds;
                }
            }
            __syncwarp();
            ah_weighted<true>(P, lp, sgo, sv, Cp, col, g.Lk, g.Lq, d2, lane);      // dV = PD^T dO   -> over V (dead for this head)
            __syncwarp();
            ah_weighted<false>(dS, lp, sk, sgo, Cp, col, g.Lq, g.Lk, d2, lane);    // dQ = dS K      -> over dO (dead)
            __syncwarp();
            ah_weighted<true>(dS, lp, sq, sk, Cp, col, g.Lk, g.Lq, d2, lane);      // dK = dS^T Q    -> over K (dead)
        }
        __syncthreads();
        ah_store_tile(sv, Cp, dV, lddv, rk, g.Lk, C4, 1.f, g.round_tf32);
        ah_store_tile(sgo, Cp, dQ, lddq, rq, g.Lq, C4, g.scale, g.round_tf32);
        ah_store_tile(sk, Cp, dK, lddk, rk, g.Lk, C4, g.scale, g.round_tf32);
        __syncthreads();
    }
    if (d_rpe_table) {
        for (int e = threadIdx.x; e < g.nhead * g.Lq * g.Lk; e += blockDim.x) {
            const int h = e / (g.Lq * g.Lk), r = e - h * g.Lq * g.Lk;
            atomicAdd(d_rpe_table + rel_pos_index(g.ws, r / g.Lk, r % g.Lk) * g.nhead + h, sdB[e]);
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

// =============================================================================================
// Tensor-core path (Lq, Lk <= 32, head_dim 66): warp-level mma.sync m16n8k8 TF32 with the 3xTF32 split
// (x = hi + lo, d += a_lo*b_hi + a_hi*b_lo + a_hi*b_hi), which keeps fp32-level accuracy -- the attention core is HBM-bound
// (AI ~ 4 FLOP/B), so the extra MMAs are free, while the scalar kernels above are shared-memory-latency bound at 4-7x the HBM
// floor.  One CTA = HPC heads of one batch entry (window / pixel sequence): the HPC*d-float row segments of Q, K, V (and dO) are
// staged with coalesced float4 loads, then every warp owns one head and keeps S / P / dS and all accumulators in registers:
//   S = Q K^T (A = Q rows, B = K rows, contraction over d),  P = softmax(scale*S + rpe),  O = P V,
//   dP = dO V^T,  dS = P o (dP - rowsum(P o dP)),  dV = PD^T dO,  dQ = dS K,  dK = dS^T Q.
// Contractions over the key index reuse the S accumulator registers directly as the A operand: a k-slot permutation
// (slot t <-> key 2t, slot t+4 <-> key 2t+1, applied to the B rows as well) makes the m16n8 accumulator layout coincide with
// the m16k8 operand layout, so P and dS never leave registers for O and dQ.  The transposed products (dV, dK) read P / dS back
// from a small per-warp shared tile.  Tile row pitches are chosen so every fragment load is bank-conflict free.
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// x = hi + lo exactly: hi = x truncated to tf32 (one LOP3), lo = x - hi (<= 13 significant bits; the tensor core reads its top
// 11, so the split is good to ~2^-21).  cvt.rna.tf32.f32 is EMULATED on sm_100 (VIADD + FSETP + SEL + LOP3): the rounding split
// cost 9 ALU instructions per fragment element and dominated the kernel (ncu source page); truncation costs 2.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
struct Frag4 { uint32_t hi[4], lo[4]; };
struct Frag2 { uint32_t hi[2], lo[2]; };
__device__ __forceinline__ void mma3(float (&d)[4], const Frag4& a, const Frag2& b) {
    mma_tf32_16x8x8(d, a.lo, b.hi);
    mma_tf32_16x8x8(d, a.hi, b.lo);
    mma_tf32_16x8x8(d, a.hi, b.hi);
}
// same, with the small cross terms in their own accumulator (two independent dependency chains instead of one)
__device__ __forceinline__ void mma3_split(float (&d)[4], float (&dc)[4], const Frag4& a, const Frag2& b) {
    mma_tf32_16x8x8(dc, a.lo, b.hi);
    mma_tf32_16x8x8(d, a.hi, b.hi);
    mma_tf32_16x8x8(dc, a.hi, b.lo);
}

template <int MT, int NT, int HPC>
struct MmaCfg {
    static constexpr int DT = 9;                       // head_dim padded to 72 = 9 k-steps / n-tiles of 8
    static constexpr int LQP = 16 * MT, LKP = 8 * NT;  // padded tile rows
    static constexpr int MTK = (LKP + 15) / 16;        // m-tiles over the key index (dV, dK)
    static constexpr int LP = NT > 2 ? 36 : 20;        // pitch of the per-warp P / dS tiles ([LQP][LP])
};
constexpr int MMA_D = 66;   // head dim of the tensor-core path (d_model 528 / 8 heads); compile time so fragment addresses fold
__host__ __device__ constexpr int mma_pitch(int hpc, int d) {   // tile row pitch (floats): == 4 or 12 (mod 32), multiple of 4
    int w = (hpc * d + 3) & ~3;
    while ((w & 31) != 4 && (w & 31) != 12) w += 4;
    return w;
}

// C[i][j] += sum_c A[i][c0 + c] * B[j][c0 + c], c < d (contraction over the head dim; columns >= d are masked on the A side)
template <int MT, int NT>
__device__ __forceinline__ void mma_rows_dot(float (&acc)[MT][NT][4], const float* sA, const float* sB, int Cp, int c0, int d, int g, int t) {
    float corr[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) corr[mt][nt][0] = corr[mt][nt][1] = corr[mt][nt][2] = corr[mt][nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 9; ++ks) {
        const int c = ks * 8 + t;
        const bool ok0 = c < d, ok1 = c + 4 < d;
        const int o0 = ok0 ? 0 : d - 1 - c, o1 = ok1 ? 4 : d - 1 - c;     // offsets from column c, clamped to the head's last column
        Frag4 a[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const float* r0 = sA + (mt * 16 + g) * Cp + c0 + c;
            const float* r1 = r0 + 8 * Cp;
            // masked columns are read from a clamped (in-head) address and discarded: no access ever leaves this head's columns
            split_tf32(ok0 ? r0[o0] : 0.f, a[mt].hi[0], a[mt].lo[0]);
            split_tf32(ok0 ? r1[o0] : 0.f, a[mt].hi[1], a[mt].lo[1]);
            split_tf32(ok1 ? r0[o1] : 0.f, a[mt].hi[2], a[mt].lo[2]);
            split_tf32(ok1 ? r1[o1] : 0.f, a[mt].hi[3], a[mt].lo[3]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float* rb = sB + (nt * 8 + g) * Cp + c0 + c;
            Frag2 b;   // columns >= d belong to the neighbouring head (whose warp may be overwriting them): never read
            split_tf32(ok0 ? rb[o0] : 0.f, b.hi[0], b.lo[0]);
            split_tf32(ok1 ? rb[o1] : 0.f, b.hi[1], b.lo[1]);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) mma3_split(acc[mt][nt], corr[mt][nt], a[mt], b);
        }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[mt][nt][r] += corr[mt][nt][r];
}
// out[i][c0 + c] = mul * sum_j W[i][j] * T[j][c0 + c] with W in accumulator registers (k-slot permutation), T rows in shared
// memory; the result is written to dst rows (this head's columns, c < d) as float2.
template <int MT, int NT>
__device__ __forceinline__ void mma_regs_times_rows(const float (&W)[MT][NT][4], const float* sT, float* dst, int Cp, int c0, int d, float mul,
                                                    int g, int t) {
    Frag4 a[MT][NT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int kt = 0; kt < NT; ++kt) {
            split_tf32(W[mt][kt][0], a[mt][kt].hi[0], a[mt][kt].lo[0]);   // (row g,   key 2t)   -> slot t
            split_tf32(W[mt][kt][2], a[mt][kt].hi[1], a[mt][kt].lo[1]);   // (row g+8, key 2t)
            split_tf32(W[mt][kt][1], a[mt][kt].hi[2], a[mt][kt].lo[2]);   // (row g,   key 2t+1) -> slot t+4
            split_tf32(W[mt][kt][3], a[mt][kt].hi[3], a[mt][kt].lo[3]);   // (row g+8, key 2t+1)
        }
#pragma unroll 3
    for (int n9 = 0; n9 < 9; ++n9) {
        float o[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < NT; ++kt) {
            const float* rb = sT + (kt * 8 + 2 * t) * Cp + c0 + n9 * 8 + g;
            const bool okn = n9 * 8 + g < d;      // output columns >= d are never stored; their inputs are another head's
            const int on = okn ? 0 : d - 1 - (n9 * 8 + g);                     // clamped to the head's last column
            Frag2 b;
            split_tf32(okn ? rb[on] : 0.f, b.hi[0], b.lo[0]);
            split_tf32(okn ? rb[Cp + on] : 0.f, b.hi[1], b.lo[1]);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) mma3(o[mt], a[mt][kt], b);
        }
        const int c = n9 * 8 + 2 * t;
        if (c < d) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                *reinterpret_cast<float2*>(dst + (mt * 16 + g) * Cp + c0 + c) = make_float2(o[mt][0] * mul, o[mt][1] * mul);
                *reinterpret_cast<float2*>(dst + (mt * 16 + g + 8) * Cp + c0 + c) = make_float2(o[mt][2] * mul, o[mt][3] * mul);
            }
        }
    }
}
// out[j][c0 + c] = mul * sum_i Wt[i][j] * T[i][c0 + c]: Wt = per-warp shared tile [LQP][LP] read transposed, T rows in shared
// memory (k index = query i, same slot permutation so the T-row loads stay conflict free).  MTK m-tiles over j, KS k-steps over i.
template <int MTK, int KS>
__device__ __forceinline__ void mma_smemT_times_rows(const float* Wt, int lp, const float* sT, float* dst, int Cp, int c0, int d, float mul,
                                                     int g, int t) {
    Frag4 a[MTK][KS];
#pragma unroll
    for (int mk = 0; mk < MTK; ++mk)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const float* w0 = Wt + (ks * 8 + 2 * t) * lp + mk * 16 + g;
            split_tf32(w0[0], a[mk][ks].hi[0], a[mk][ks].lo[0]);
            split_tf32(w0[8], a[mk][ks].hi[1], a[mk][ks].lo[1]);
            split_tf32(w0[lp], a[mk][ks].hi[2], a[mk][ks].lo[2]);
            split_tf32(w0[lp + 8], a[mk][ks].hi[3], a[mk][ks].lo[3]);
        }
#pragma unroll 3
    for (int n9 = 0; n9 < 9; ++n9) {
        float o[MTK][4];
#pragma unroll
        for (int mk = 0; mk < MTK; ++mk) o[mk][0] = o[mk][1] = o[mk][2] = o[mk][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const float* rb = sT + (ks * 8 + 2 * t) * Cp + c0 + n9 * 8 + g;
            const bool okn = n9 * 8 + g < d;
            const int on = okn ? 0 : d - 1 - (n9 * 8 + g);
            Frag2 b;
            split_tf32(okn ? rb[on] : 0.f, b.hi[0], b.lo[0]);
            split_tf32(okn ? rb[Cp + on] : 0.f, b.hi[1], b.lo[1]);
#pragma unroll
            for (int mk = 0; mk < MTK; ++mk) mma3(o[mk], a[mk][ks], b);
        }
        const int c = n9 * 8 + 2 * t;
        if (c < d) {
#pragma unroll
            for (int mk = 0; mk < MTK; ++mk) {
                *reinterpret_cast<float2*>(dst + (mk * 16 + g) * Cp + c0 + c) = make_float2(o[mk][0] * mul, o[mk][1] * mul);
                *reinterpret_cast<float2*>(dst + (mk * 16 + g + 8) * Cp + c0 + c) = make_float2(o[mk][2] * mul, o[mk][3] * mul);
            }
        }
    }
}

// coalesced 16-byte staging of `L` row segments [col0, col0 + 4*W4) into a [.][Cp] tile with cp.async (no register staging,
// all of a thread's requests in flight at once), one warp per row; and the store back (optionally rounded to tf32)
template <int W4>
__device__ __forceinline__ void mma_load_tile(float* tile, int Cp, const float* __restrict__ src, long long ld, const long long* rows, int L,
                                              int col0) {
    // flat (row, float4) index with a compile-time row width: the division is a multiply-shift and every thread issues
    // ceil(L * W4 / blockDim) copies (the one-warp-per-row form spent 3 iterations, the last 2 lanes wide, on 66 float4)
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tile);
    for (int e = threadIdx.x; e < L * W4; e += blockDim.x) {
        const int l = e / W4, c = e - l * W4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + (uint32_t)(l * Cp + 4 * c) * 4u), "l"(src + rows[l] * ld + col0 + 4 * c)
                     : "memory");
    }
}
__device__ __forceinline__ void mma_load_wait() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
// scol != NULL: the column sums of what is stored accumulate in shared memory (the bias gradient of the projection whose output
// gradient this is -- in_proj_bias / q,k,v_proj.bias -- so no separate pass re-reads dq / dk / dv; flushed once per CTA)
template <int W4>
__device__ __forceinline__ void mma_store_tile(const float* tile, int Cp, float* __restrict__ dst, long long ld, const long long* rows, int L,
                                               int col0, int round_tf32, float* scol = nullptr) {
    for (int e = threadIdx.x; e < L * W4; e += blockDim.x) {
        const int l = e / W4, c = e - l * W4;
        float4 v = *reinterpret_cast<const float4*>(tile + l * Cp + 4 * c);
        if (round_tf32) { v.x = vptr_round_tf32(v.x); v.y = vptr_round_tf32(v.y); v.z = vptr_round_tf32(v.z); v.w = vptr_round_tf32(v.w); }
        *reinterpret_cast<float4*>(dst + rows[l] * ld + col0 + 4 * c) = v;
        if (scol) { atomicAdd(scol + 4 * c, v.x); atomicAdd(scol + 4 * c + 1, v.y); atomicAdd(scol + 4 * c + 2, v.z); atomicAdd(scol + 4 * c + 3, v.w); }
    }
}
// dropout keep-scales of probabilities (b, h, i, j) and (b, h, i, j + 1), j even: one hash when both fall into one group of four
__device__ __forceinline__ void prob_drop2(const AttnGeom& g, unsigned long long idx, float& k0, float& k1) {
    if (idx & 1) { k0 = vptr_drop_scale(g.drop_seed, idx, g.drop_p); k1 = vptr_drop_scale(g.drop_seed, idx + 1, g.drop_p); return; }
    const unsigned z = (unsigned)(vptr_hash4(g.drop_seed, idx >> 2) >> (16 * (unsigned)(idx & 3)));
    const unsigned thr = vptr_drop_threshold(g.drop_p);
    const float inv = 1.f / (1.f - g.drop_p);
    k0 = (z & 0xFFFFu) >= thr ? inv : 0.f;
    k1 = ((z >> 16) & 0xFFFFu) >= thr ? inv : 0.f;
}

template <int MT, int NT, int HPC, bool BWD>
__global__ void __launch_bounds__(HPC * 32) attn_mma_kernel(const float* __restrict__ Q, long long ldq, const float* __restrict__ K, long long ldk,
                                                            const float* __restrict__ V, long long ldv, const float* __restrict__ dO,
                                                            float* __restrict__ O_or_dQ, long long ldo, float* __restrict__ dK, long long lddk,
                                                            float* __restrict__ dV, long long lddv, long long lddo,
                                                            const float* __restrict__ rpe_table, float* __restrict__ d_rpe_table,
                                                            const AttnGeom g, int batches) {
    using Cfg = MmaCfg<MT, NT, HPC>;
    constexpr int LQP = Cfg::LQP, LKP = Cfg::LKP, LP = Cfg::LP;
    extern __shared__ __align__(16) float sm[];
    __shared__ float scol[3][HPC * MMA_D];            // column sums of dQ / dK / dV (bias gradients), backward only
    if (BWD) for (int e = threadIdx.x; e < 3 * HPC * MMA_D; e += blockDim.x) (&scol[0][0])[e] = 0.f;
    constexpr int Cp = mma_pitch(HPC, MMA_D);
    constexpr int W4 = HPC * MMA_D / 4;
    constexpr int D = MMA_D;
    float* sq = sm;                                   // [LQP][Cp]   (forward: O is staged over it)
    float* sk = sq + LQP * Cp;                        // [LKP][Cp]   (backward: dK over it)
    float* sv = sk + LKP * Cp;                        // [LKP][Cp]   (backward: dV over it)
    float* sgo = sv + LKP * Cp;                       // [LQP][Cp]   (backward only: dO, dQ over it)
    float* spw = sgo + (BWD ? LQP * Cp : 0);          // per-warp PD and dS tiles [HPC][2][LQP][LP] (backward only)
    float* sdb = spw + (BWD ? HPC * 2 * LQP * LP : 0);                   // [HPC][Lq][Lk] bias-gradient accumulator (backward, rpe)
    float* sbias = sdb + ((BWD && d_rpe_table) ? ((HPC * g.Lq * g.Lk + 1) & ~1) : 0);   // [HPC][LQP][LKP] rpe bias of this CTA's heads
    // + 8 zero floats: fragment loads of the last head's pad columns run a few floats past the last tile row
    long long* rq = reinterpret_cast<long long*>(sbias + (rpe_table ? HPC * LQP * LKP : 0) + 8);
    long long* rk = rq + LQP;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
    const int hgroups = g.nhead / HPC;
    // pad rows / columns are never loaded: zero them once so every product with them is finite (and zero)
    for (int e = threadIdx.x; e < (int)(reinterpret_cast<float*>(rq) - sm); e += blockDim.x) sm[e] = 0.f;
    __syncthreads();
    if (rpe_table) {   // gridDim.x is a multiple of hgroups: all items of this CTA share one head group
        const int hb = (blockIdx.x % hgroups) * HPC;
        for (int e = threadIdx.x; e < HPC * g.Lq * g.Lk; e += blockDim.x) {
            const int w = e / (g.Lq * g.Lk), r = e - w * g.Lq * g.Lk, i = r / g.Lk, j = r - i * g.Lk;
            sbias[(w * LQP + i) * LKP + j] = __ldg(rpe_table + rel_pos_index(g.ws, i, j) * g.nhead + hb + w);
        }
    }
    __syncthreads();
    for (int item = blockIdx.x; item < batches * hgroups; item += gridDim.x) {
        const int b = item / hgroups, h0 = (item - b * hgroups) * HPC;
        for (int l = threadIdx.x; l < g.Lq; l += blockDim.x) rq[l] = q_row(g, b, l);
        for (int l = threadIdx.x; l < g.Lk; l += blockDim.x) rk[l] = k_row(g, b, l);
        __syncthreads();
        mma_load_tile<W4>(sq, Cp, Q, ldq, rq, g.Lq, h0 * D);
        mma_load_tile<W4>(sk, Cp, K, ldk, rk, g.Lk, h0 * D);
        mma_load_tile<W4>(sv, Cp, V, ldv, rk, g.Lk, h0 * D);
        if (BWD) mma_load_tile<W4>(sgo, Cp, dO, lddo, rq, g.Lq, h0 * D);
        mma_load_wait();
        __syncthreads();
        {
            const int h = h0 + warp, c0 = warp * D;
            // ---- S = Q K^T, P = softmax(scale * S + bias) (undropped probabilities)
            float P[MT][NT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) P[mt][nt][0] = P[mt][nt][1] = P[mt][nt][2] = P[mt][nt][3] = 0.f;
            mma_rows_dot<MT, NT>(P, sq, sk, Cp, c0, D, gq, t);
            __syncwarp();   // all lanes' reads of this head's Q / K columns precede the stores that later reuse them (O, dK)
            float keep[MT][NT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                float mx[2] = {-INFINITY, -INFINITY};
                unsigned long long drop_row[2];      // index of probability (b, h, i, 0) for this lane's two rows
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
                    drop_row[hh] = (((unsigned long long)b * g.nhead + h) * g.Lq + (mt * 16 + gq + hh * 8)) * g.Lk;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int i = mt * 16 + gq + hh * 8, j = nt * 8 + 2 * t;
                        float2 bias = make_float2(0.f, 0.f);
                        if (rpe_table) bias = *reinterpret_cast<const float2*>(sbias + (warp * LQP + i) * LKP + j);   // pad entries are 0
                        float s0 = fmaf(P[mt][nt][2 * hh], g.scale, bias.x), s1 = fmaf(P[mt][nt][2 * hh + 1], g.scale, bias.y);
                        if (j >= g.Lk || (g.causal && j > i)) s0 = -INFINITY;
                        if (j + 1 >= g.Lk || (g.causal && j + 1 > i)) s1 = -INFINITY;
                        P[mt][nt][2 * hh] = s0;
                        P[mt][nt][2 * hh + 1] = s1;
                        mx[hh] = fmaxf(mx[hh], fmaxf(s0, s1));
                        float k0 = 1.f, k1 = 1.f;
                        if (g.drop_p > 0.f && i < g.Lq && j < g.Lk) {
                            prob_drop2(g, drop_row[hh] + j, k0, k1);
                            if (j + 1 >= g.Lk) k1 = 1.f;
                        }
                        keep[mt][nt][2 * hh] = k0;
                        keep[mt][nt][2 * hh + 1] = k1;
                    }
                float sum[2] = {0.f, 0.f};
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
                    mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float e = __expf(P[mt][nt][r] - mx[r >> 1]);
                        P[mt][nt][r] = e;
                        sum[r >> 1] += e;
                    }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    sum[hh] += __shfl_xor_sync(0xffffffffu, sum[hh], 1);
                    sum[hh] += __shfl_xor_sync(0xffffffffu, sum[hh], 2);
                    sum[hh] = 1.f / sum[hh];
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) P[mt][nt][r] *= sum[r >> 1];
            }
            if (!BWD) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r) P[mt][nt][r] *= keep[mt][nt][r];
                mma_regs_times_rows<MT, NT>(P, sv, sq, Cp, c0, D, 1.f, gq, t);        // O = PD V, staged over this head's Q columns
            } else {
                // ---- dP = (dO V^T) * keep ; dS = P * (dP - rowsum(P * dP)) ; PD = P * keep
                float dS[MT][NT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dS[mt][nt][0] = dS[mt][nt][1] = dS[mt][nt][2] = dS[mt][nt][3] = 0.f;
                mma_rows_dot<MT, NT>(dS, sgo, sv, Cp, c0, D, gq, t);
                __syncwarp();
                float* PDw = spw + warp * 2 * LQP * LP;
                float* dSw = PDw + LQP * LP;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    float tsum[2] = {0.f, 0.f};
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            dS[mt][nt][r] *= keep[mt][nt][r];
                            tsum[r >> 1] = fmaf(P[mt][nt][r], dS[mt][nt][r], tsum[r >> 1]);
                        }
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        tsum[hh] += __shfl_xor_sync(0xffffffffu, tsum[hh], 1);
                        tsum[hh] += __shfl_xor_sync(0xffffffffu, tsum[hh], 2);
                    }
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const float ds = P[mt][nt][r] * (dS[mt][nt][r] - tsum[r >> 1]);
                            dS[mt][nt][r] = ds;
                            P[mt][nt][r] *= keep[mt][nt][r];
                            if (d_rpe_table) {
                                const int i = mt * 16 + gq + (r >> 1) * 8, j = nt * 8 + 2 * t + (r & 1);
                                if (i < g.Lq && j < g.Lk) sdb[(warp * g.Lq + i) * g.Lk + j] += ds;
                            }
                        }
                        const int i0 = mt * 16 + gq, j0 = nt * 8 + 2 * t;
                        *reinterpret_cast<float2*>(PDw + i0 * LP + j0) = make_float2(P[mt][nt][0], P[mt][nt][1]);
                        *reinterpret_cast<float2*>(PDw + (i0 + 8) * LP + j0) = make_float2(P[mt][nt][2], P[mt][nt][3]);
                        *reinterpret_cast<float2*>(dSw + i0 * LP + j0) = make_float2(dS[mt][nt][0], dS[mt][nt][1]);
                        *reinterpret_cast<float2*>(dSw + (i0 + 8) * LP + j0) = make_float2(dS[mt][nt][2], dS[mt][nt][3]);
                    }
                }
                __syncwarp();
                mma_smemT_times_rows<Cfg::MTK, 2 * MT>(PDw, LP, sgo, sv, Cp, c0, D, 1.f, gq, t);      // dV = PD^T dO  -> over V (dead)
                __syncwarp();
                mma_regs_times_rows<MT, NT>(dS, sk, sgo, Cp, c0, D, g.scale, gq, t);                   // dQ = dS K     -> over dO (dead)
                __syncwarp();
                mma_smemT_times_rows<Cfg::MTK, 2 * MT>(dSw, LP, sq, sk, Cp, c0, D, g.scale, gq, t);   // dK = dS^T Q   -> over K (dead)
            }
        }
        __syncthreads();
        if (!BWD) {
            mma_store_tile<W4>(sq, Cp, O_or_dQ, ldo, rq, g.Lq, h0 * D, g.round_tf32);
        } else {
            mma_store_tile<W4>(sv, Cp, dV, lddv, rk, g.Lk, h0 * D, g.round_tf32, g.dbv ? scol[2] : nullptr);
            mma_store_tile<W4>(sgo, Cp, O_or_dQ, ldo, rq, g.Lq, h0 * D, g.round_tf32, g.dbq ? scol[0] : nullptr);
            mma_store_tile<W4>(sk, Cp, dK, lddk, rk, g.Lk, h0 * D, g.round_tf32, g.dbk ? scol[1] : nullptr);
        }
        __syncthreads();
    }
    if (BWD && (g.dbq || g.dbk || g.dbv)) {     // all items of this CTA share one head group: flush its bias-gradient columns once
        const int hc = (blockIdx.x % hgroups) * HPC * D;
        for (int e = threadIdx.x; e < HPC * D; e += blockDim.x) {
            if (g.dbq) atomicAdd(g.dbq + hc + e, scol[0][e]);
            if (g.dbk) atomicAdd(g.dbk + hc + e, scol[1][e]);
            if (g.dbv) atomicAdd(g.dbv + hc + e, scol[2][e]);
        }
    }
    if (BWD && d_rpe_table) {
        // the launch makes gridDim.x a multiple of hgroups, so all items of a CTA share one head group
        // (item % hgroups == blockIdx.x % hgroups) and the accumulator can be flushed once, here
        // reduce the (i, j) entries into the (2ws-1)^2 table bins in shared memory first (over the dead PD/dS tiles): 5x fewer,
        // far less contended global atomics than one per (i, j)
        const int h0 = (blockIdx.x % hgroups) * HPC;
        const int bins = (2 * g.ws - 1) * (2 * g.ws - 1);
        const bool binned = BWD && HPC * bins <= HPC * 2 * LQP * LP;
        __syncthreads();
        if (binned) {
            for (int e = threadIdx.x; e < HPC * bins; e += blockDim.x) spw[e] = 0.f;
            __syncthreads();
        }
        for (int e = threadIdx.x; e < HPC * g.Lq * g.Lk; e += blockDim.x) {
            const int w = e / (g.Lq * g.Lk), r = e - w * g.Lq * g.Lk;
            const int idx = rel_pos_index(g.ws, r / g.Lk, r % g.Lk);
            if (binned) atomicAdd(spw + w * bins + idx, sdb[e]);
            else atomicAdd(d_rpe_table + idx * g.nhead + h0 + w, sdb[e]);
        }
        if (binned) {
            __syncthreads();
            for (int e = threadIdx.x; e < HPC * bins; e += blockDim.x) {
                const int w = e / bins, idx = e - w * bins;
                atomicAdd(d_rpe_table + idx * g.nhead + h0 + w, spw[e]);
            }
        }
    }
}

template <int MT, int NT, int HPC, bool BWD>
int launch_attn_mma(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv, const float* dO, float* O_or_dQ,
                    long long ldo, float* dK, long long lddk, float* dV, long long lddv, long long lddo, const float* rpe_table,
                    float* d_rpe_table, const AttnGeom& g, int batches, cudaStream_t stream) {
    using Cfg = MmaCfg<MT, NT, HPC>;
    constexpr int Cp = mma_pitch(HPC, MMA_D);
    size_t floats = (size_t)(Cfg::LQP + 2 * Cfg::LKP) * Cp;
    if (BWD) floats += (size_t)Cfg::LQP * Cp + (size_t)HPC * 2 * Cfg::LQP * Cfg::LP + ((d_rpe_table ? (size_t)HPC * g.Lq * g.Lk + 1 : 0) & ~(size_t)1);
    if (rpe_table) floats += (size_t)HPC * Cfg::LQP * Cfg::LKP;
    const size_t smem = (floats + 8) * sizeof(float) + sizeof(long long) * (size_t)(Cfg::LQP + Cfg::LKP);
    auto kern = attn_mma_kernel<MT, NT, HPC, BWD>;
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        VPTR_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(attn_mma, smem=%zu): %s", smem, cudaGetErrorString(e));
        attr = smem;
    }
    const int hgroups = g.nhead / HPC;
    const long long items = (long long)batches * hgroups;
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long grid = 148LL * per_sm;
    grid -= grid % hgroups;                       // a CTA keeps one head group: item % hgroups == blockIdx.x % hgroups
    if (grid > items) grid = items;               // (items is a multiple of hgroups)
    kern<<<(int)grid, HPC * 32, smem, stream>>>(Q, ldq, K, ldk, V, ldv, dO, O_or_dQ, ldo, dK, lddk, dV, lddv, lddo, rpe_table, d_rpe_table, g,
                                                batches);
    return vptr_check_launch("attn_mma_kernel");
}


// ---------------------------------------------------------------------------------------------------------------------------------
// Wide tensor-core path: groups of 33..64 tokens (the 8x8 windows of the 16x16 grid, cfg4; temporal sequences of up to 64 frames).
// Same 3xTF32 mma.sync arithmetic and shared-memory tiles as attn_mma_kernel, but a head no longer fits one warp's registers
// (S alone is 64 x 64), so FOUR warps share a head: warp (hw, mt) owns the 16 query rows of m-tile mt -- S, softmax, dP, dS, O and
// dQ are row-local -- and, for the products that contract over the query index (dV = PD^T dO, dK = dS^T Q), the 16 key rows of
// slice mt, reading the whole head's PD / dS tiles from shared memory after a 128-thread named barrier.  Two heads per CTA.
// The relative-position bias is looked up through a [64][64] uint8 index table built once per CTA ((2 ws - 1)^2 <= 255 bins) and
// its gradient accumulates in registers across the CTA's items (the (i, j) ownership of a thread never changes).
// The scalar kernels this replaces ran the cfg4 window attention at 6.3 ms per backward call (30 % of the step).
constexpr int W64_L = 64, W64_HPC = 2, W64_LP = 68, W64_THREADS = W64_HPC * 4 * 32;
__device__ __forceinline__ void bar_head(int hw) { asm volatile("bar.sync %0, 128;" ::"r"(hw + 1) : "memory"); }

template <bool BWD>
__global__ void __launch_bounds__(W64_THREADS, 1) attn_mma64_kernel(const float* __restrict__ Q, long long ldq, const float* __restrict__ K, long long ldk,
                                                                    const float* __restrict__ V, long long ldv, const float* __restrict__ dO,
                                                                    float* __restrict__ O_or_dQ, long long ldo, float* __restrict__ dK, long long lddk,
                                                                    float* __restrict__ dV, long long lddv, long long lddo,
                                                                    const float* __restrict__ rpe_table, float* __restrict__ d_rpe_table,
                                                                    const AttnGeom g, int batches) {
    constexpr int L = W64_L, HPC = W64_HPC, LP = W64_LP, D = MMA_D, NT = 8;
    constexpr int Cp = mma_pitch(HPC, MMA_D);
    constexpr int W4 = HPC * MMA_D / 4;
    extern __shared__ __align__(16) float sm[];
    __shared__ float scol[3][HPC * MMA_D];            // column sums of dQ / dK / dV (bias gradients), backward only
    if (BWD) for (int e = threadIdx.x; e < 3 * HPC * MMA_D; e += blockDim.x) (&scol[0][0])[e] = 0.f;
    float* sq = sm;                                   // [64][Cp]   (forward: O staged over it)
    float* sk = sq + L * Cp;                          // [64][Cp]   (backward: dK over it)
    float* sv = sk + L * Cp;                          // [64][Cp]   (backward: dV over it)
    float* sgo = sv + L * Cp;                         // [64][Cp]   (backward only: dO, dQ over it)
    float* spw = sgo + (BWD ? L * Cp : 0);            // [HPC][2][64][LP] PD and dS tiles of each head (backward only)
    const int bins = (2 * g.ws - 1) * (2 * g.ws - 1);
    float* stab = spw + (BWD ? HPC * 2 * L * LP : 0);                    // [HPC][bins] bias table of this CTA's heads
    float* sbin = stab + (rpe_table ? HPC * bins : 0);                   // [HPC][bins] bias-gradient bins (flush)
    float* tail = sbin + ((BWD && d_rpe_table) ? HPC * bins : 0);
    tail += (4 - ((tail - sm) & 3)) & 3;
    long long* rq = reinterpret_cast<long long*>(tail + 8);             // + 8 zero floats (fragment loads run a little past the last row)
    long long* rk = rq + L;
    unsigned char* lut = reinterpret_cast<unsigned char*>(rk + L);      // [64][64] relative-position index of (i, j)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
    const int hw = warp >> 2, mt = warp & 3, row0 = mt * 16;
    const int hgroups = g.nhead / HPC;
    for (int e = threadIdx.x; e < (int)(reinterpret_cast<float*>(rq) - sm); e += blockDim.x) sm[e] = 0.f;
    __syncthreads();
    if (rpe_table) {   // gridDim.x is a multiple of hgroups: all items of this CTA share one head group
        const int hb = (blockIdx.x % hgroups) * HPC;
        for (int e = threadIdx.x; e < HPC * bins; e += blockDim.x) stab[e] = __ldg(rpe_table + (e % bins) * g.nhead + hb + e / bins);
        for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
            const int i = e >> 6, j = e & 63;
            lut[e] = (i < g.Lq && j < g.Lk) ? (unsigned char)rel_pos_index(g.ws, i, j) : 0;
        }
    }
    float dbacc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) dbacc[nt][0] = dbacc[nt][1] = dbacc[nt][2] = dbacc[nt][3] = 0.f;
    __syncthreads();
    for (int item = blockIdx.x; item < batches * hgroups; item += gridDim.x) {
        const int b = item / hgroups, h0 = (item - b * hgroups) * HPC;
        for (int l = threadIdx.x; l < g.Lq; l += blockDim.x) rq[l] = q_row(g, b, l);
        for (int l = threadIdx.x; l < g.Lk; l += blockDim.x) rk[l] = k_row(g, b, l);
        __syncthreads();
        mma_load_tile<W4>(sq, Cp, Q, ldq, rq, g.Lq, h0 * D);
        mma_load_tile<W4>(sk, Cp, K, ldk, rk, g.Lk, h0 * D);
        mma_load_tile<W4>(sv, Cp, V, ldv, rk, g.Lk, h0 * D);
        if (BWD) mma_load_tile<W4>(sgo, Cp, dO, lddo, rq, g.Lq, h0 * D);
        mma_load_wait();
        __syncthreads();
        {
            const int h = h0 + hw, c0 = hw * D;
            // ---- S = Q_mt K^T, P = softmax(scale * S + bias) for the 16 query rows of this warp
            float P[1][NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) P[0][nt][0] = P[0][nt][1] = P[0][nt][2] = P[0][nt][3] = 0.f;
            mma_rows_dot<1, NT>(P, sq + row0 * Cp, sk, Cp, c0, D, gq, t);
            uint32_t keepm = 0xffffffffu;                 // bit nt*4 + r: probability kept by dropout
            float mx[2] = {-INFINITY, -INFINITY};
            unsigned long long drop_row[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) drop_row[hh] = (((unsigned long long)b * g.nhead + h) * g.Lq + (row0 + gq + hh * 8)) * g.Lk;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int i = row0 + gq + hh * 8, j = nt * 8 + 2 * t;
                    float b0 = 0.f, b1 = 0.f;
                    if (rpe_table) { b0 = stab[hw * bins + lut[i * L + j]]; b1 = stab[hw * bins + lut[i * L + j + 1]]; }
                    float s0 = fmaf(P[0][nt][2 * hh], g.scale, b0), s1 = fmaf(P[0][nt][2 * hh + 1], g.scale, b1);
                    if (j >= g.Lk || (g.causal && j > i)) s0 = -INFINITY;
                    if (j + 1 >= g.Lk || (g.causal && j + 1 > i)) s1 = -INFINITY;
                    P[0][nt][2 * hh] = s0;
                    P[0][nt][2 * hh + 1] = s1;
                    mx[hh] = fmaxf(mx[hh], fmaxf(s0, s1));
                    if (g.drop_p > 0.f && i < g.Lq && j < g.Lk) {
                        float k0, k1;
                        prob_drop2(g, drop_row[hh] + j, k0, k1);
                        if (j + 1 >= g.Lk) k1 = 1.f;
                        if (k0 == 0.f) keepm &= ~(1u << (nt * 4 + 2 * hh));
                        if (k1 == 0.f) keepm &= ~(1u << (nt * 4 + 2 * hh + 1));
                    }
                }
            const float keep_scale = g.drop_p > 0.f ? 1.f / (1.f - g.drop_p) : 1.f;
            float sum[2] = {0.f, 0.f};
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
                mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float e = __expf(P[0][nt][r] - mx[r >> 1]);
                    P[0][nt][r] = e;
                    sum[r >> 1] += e;
                }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                sum[hh] += __shfl_xor_sync(0xffffffffu, sum[hh], 1);
                sum[hh] += __shfl_xor_sync(0xffffffffu, sum[hh], 2);
                sum[hh] = 1.f / sum[hh];
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int r = 0; r < 4; ++r) P[0][nt][r] *= sum[r >> 1];
            if (!BWD) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) P[0][nt][r] *= ((keepm >> (nt * 4 + r)) & 1u) ? keep_scale : 0.f;
                __syncwarp();
                mma_regs_times_rows<1, NT>(P, sv, sq + row0 * Cp, Cp, c0, D, 1.f, gq, t);      // O_mt = PD V, staged over this warp's own Q rows
            } else {
                // ---- dP = (dO_mt V^T) * keep ; dS = P o (dP - rowsum(P o dP)) ; PD = P o keep
                float dS[1][NT][4];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dS[0][nt][0] = dS[0][nt][1] = dS[0][nt][2] = dS[0][nt][3] = 0.f;
                mma_rows_dot<1, NT>(dS, sgo + row0 * Cp, sv, Cp, c0, D, gq, t);
                float* PDh = spw + hw * 2 * L * LP;
                float* dSh = PDh + L * LP;
                float tsum[2] = {0.f, 0.f};
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        dS[0][nt][r] *= ((keepm >> (nt * 4 + r)) & 1u) ? keep_scale : 0.f;
                        tsum[r >> 1] = fmaf(P[0][nt][r], dS[0][nt][r], tsum[r >> 1]);
                    }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    tsum[hh] += __shfl_xor_sync(0xffffffffu, tsum[hh], 1);
                    tsum[hh] += __shfl_xor_sync(0xffffffffu, tsum[hh], 2);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float ds = P[0][nt][r] * (dS[0][nt][r] - tsum[r >> 1]);
                        dS[0][nt][r] = ds;
                        P[0][nt][r] *= ((keepm >> (nt * 4 + r)) & 1u) ? keep_scale : 0.f;
                        dbacc[nt][r] += ds;                 // (entries outside Lq x Lk are exact zeros: P is 0 there)
                    }
                    const int i0 = row0 + gq, j0 = nt * 8 + 2 * t;
                    *reinterpret_cast<float2*>(PDh + i0 * LP + j0) = make_float2(P[0][nt][0], P[0][nt][1]);
                    *reinterpret_cast<float2*>(PDh + (i0 + 8) * LP + j0) = make_float2(P[0][nt][2], P[0][nt][3]);
                    *reinterpret_cast<float2*>(dSh + i0 * LP + j0) = make_float2(dS[0][nt][0], dS[0][nt][1]);
                    *reinterpret_cast<float2*>(dSh + (i0 + 8) * LP + j0) = make_float2(dS[0][nt][2], dS[0][nt][3]);
                }
                bar_head(hw);            // the head's PD / dS tiles are complete; every warp is done reading V
                mma_smemT_times_rows<1, 8>(PDh + row0, LP, sgo, sv + row0 * Cp, Cp, c0, D, 1.f, gq, t);       // dV[slice mt] = PD^T dO -> over V
                bar_head(hw);            // every warp is done reading dO
                mma_regs_times_rows<1, NT>(dS, sk, sgo + row0 * Cp, Cp, c0, D, g.scale, gq, t);              // dQ_mt = dS K -> over dO
                bar_head(hw);            // every warp is done reading K
                mma_smemT_times_rows<1, 8>(dSh + row0, LP, sq, sk + row0 * Cp, Cp, c0, D, g.scale, gq, t);   // dK[slice mt] = dS^T Q -> over K
            }
        }
        __syncthreads();
        if (!BWD) {
            mma_store_tile<W4>(sq, Cp, O_or_dQ, ldo, rq, g.Lq, h0 * D, g.round_tf32);
        } else {
            mma_store_tile<W4>(sv, Cp, dV, lddv, rk, g.Lk, h0 * D, g.round_tf32, g.dbv ? scol[2] : nullptr);
            mma_store_tile<W4>(sgo, Cp, O_or_dQ, ldo, rq, g.Lq, h0 * D, g.round_tf32, g.dbq ? scol[0] : nullptr);
            mma_store_tile<W4>(sk, Cp, dK, lddk, rk, g.Lk, h0 * D, g.round_tf32, g.dbk ? scol[1] : nullptr);
        }
        __syncthreads();
    }
    if (BWD && (g.dbq || g.dbk || g.dbv)) {     // all items of this CTA share one head group: flush its bias-gradient columns once
        const int hc = (blockIdx.x % hgroups) * HPC * D;
        for (int e = threadIdx.x; e < HPC * D; e += blockDim.x) {
            if (g.dbq) atomicAdd(g.dbq + hc + e, scol[0][e]);
            if (g.dbk) atomicAdd(g.dbk + hc + e, scol[1][e]);
            if (g.dbv) atomicAdd(g.dbv + hc + e, scol[2][e]);
        }
    }
    if (BWD && d_rpe_table) {
        for (int e = threadIdx.x; e < HPC * bins; e += blockDim.x) sbin[e] = 0.f;
        __syncthreads();
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = row0 + gq + (r >> 1) * 8, j = nt * 8 + 2 * t + (r & 1);
                if (i < g.Lq && j < g.Lk) atomicAdd(sbin + hw * bins + lut[i * L + j], dbacc[nt][r]);
            }
        __syncthreads();
        const int hb = (blockIdx.x % hgroups) * HPC;
        for (int e = threadIdx.x; e < HPC * bins; e += blockDim.x) atomicAdd(d_rpe_table + (e % bins) * g.nhead + hb + e / bins, sbin[e]);
    }
}

template <bool BWD>
int launch_attn_mma64(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv, const float* dO, float* O_or_dQ,
                      long long ldo, float* dK, long long lddk, float* dV, long long lddv, long long lddo, const float* rpe_table,
                      float* d_rpe_table, const AttnGeom& g, int batches, cudaStream_t stream) {
    constexpr int Cp = mma_pitch(W64_HPC, MMA_D);
    const int bins = (2 * g.ws - 1) * (2 * g.ws - 1);
    size_t floats = (size_t)(BWD ? 4 : 3) * W64_L * Cp + (BWD ? (size_t)W64_HPC * 2 * W64_L * W64_LP : 0);
    if (rpe_table) floats += (size_t)W64_HPC * bins;
    if (BWD && d_rpe_table) floats += (size_t)W64_HPC * bins;
    const size_t smem = (floats + 16) * sizeof(float) + sizeof(long long) * 2 * W64_L + (size_t)W64_L * W64_L;
    VPTR_REQUIRE(smem + 2048 <= 227 * 1024, VPTR_ERR_UNSUPPORTED, "attn_mma64: %zu bytes of shared memory", smem);
    auto kern = attn_mma64_kernel<BWD>;
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        VPTR_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(attn_mma64, smem=%zu): %s", smem, cudaGetErrorString(e));
        attr = smem;
    }
    const int hgroups = g.nhead / W64_HPC;
    const long long items = (long long)batches * hgroups;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    long long grid = 148LL * per_sm;
    grid -= grid % hgroups;                       // a CTA keeps one head group: item % hgroups == blockIdx.x % hgroups
    if (grid > items) grid = items;
    kern<<<(int)grid, W64_THREADS, smem, stream>>>(Q, ldq, K, ldk, V, ldv, dO, O_or_dQ, ldo, dK, lddk, dV, lddv, lddo, rpe_table, d_rpe_table, g, batches);
    return vptr_check_launch("attn_mma64_kernel");
}
// shapes of the wide path: some side in 33..64, head_dim 66, an even number of heads, (2 ws - 1)^2 <= 255 relative positions
bool attn_mma64_ok(const AttnGeom& g, int nhead, int d, bool rpe) {
    static const bool off = [] { const char* e = getenv("VPTR_ATTN_NO_MMA64"); return e && e[0] == '1'; }();
    return !off && g.mode != 2 && g.Lq <= 64 && g.Lk <= 64 && (g.Lq > 32 || g.Lk > 32) && d == MMA_D && nhead % 2 == 0 &&
           (!rpe || (g.mode == 0 && (2 * g.ws - 1) * (2 * g.ws - 1) <= 255));
}

bool attn_mma_disabled() {
    static const bool v = [] { const char* e = getenv("VPTR_ATTN_NO_MMA"); return e && e[0] == '1'; }();
    return v;
}
// shapes the tensor-core path covers: Lq, Lk <= 32, head_dim 66, an even number of heads, 16-byte aligned rows
bool attn_mma_ok(const AttnGeom& g, int nhead, int d) {
    return !attn_mma_disabled() && g.mode != 2 && g.Lq <= 32 && g.Lk <= 32 && d == MMA_D && nhead % 2 == 0;
}
template <bool BWD>
int dispatch_attn_mma(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv, const float* dO, float* O_or_dQ,
                      long long ldo, float* dK, long long lddk, float* dV, long long lddv, long long lddo, const float* rpe_table,
                      float* d_rpe_table, const AttnGeom& g, int batches, cudaStream_t stream) {
    static const int hpc_env = [] { const char* e = getenv("VPTR_ATTN_HPC"); return e ? atoi(e) : 0; }();
    if (g.Lq <= 16 && g.Lk <= 16) {
        if (g.nhead % 4 == 0 && hpc_env != 2)
            return launch_attn_mma<1, 2, 4, BWD>(Q, ldq, K, ldk, V, ldv, dO, O_or_dQ, ldo, dK, lddk, dV, lddv, lddo, rpe_table, d_rpe_table, g, batches, stream);
        return launch_attn_mma<1, 2, 2, BWD>(Q, ldq, K, ldk, V, ldv, dO, O_or_dQ, ldo, dK, lddk, dV, lddv, lddo, rpe_table, d_rpe_table, g, batches, stream);
    }
    return launch_attn_mma<2, 4, 2, BWD>(Q, ldq, K, ldk, V, ldv, dO, O_or_dQ, ldo, dK, lddk, dV, lddv, lddo, rpe_table, d_rpe_table, g, batches, stream);
}

bool attn_no_allheads() {
    static const bool v = [] { const char* e = getenv("VPTR_ATTN_PERHEAD"); return e && e[0] == '1'; }();
    return v;
}

int fill_geom(AttnGeom& g, int mode, int F_or_N, int H, int W, int ws, int Tq, int Tk, int nhead, int d, int causal, float scale, int* batches) {
    g = AttnGeom{};
    g.mode = mode; g.nhead = nhead; g.d = d; g.causal = causal; g.scale = scale;
    if (mode == 0) {
        VPTR_REQUIRE(ws > 0 && H % ws == 0 && W % ws == 0, VPTR_ERR_SHAPE, "window attention: H=%d W=%d not multiples of ws=%d (pad first)", H, W, ws);
        g.H = H; g.W = W; g.ws = ws; g.nwh = H / ws; g.nww = W / ws; g.Lq = g.Lk = ws * ws;
        *batches = F_or_N * g.nwh * g.nww;
    } else if (mode == 2) {
        VPTR_REQUIRE(ws > 0 && H % ws == 0 && W % ws == 0 && Tq > 0 && Tk > 0 && !causal, VPTR_ERR_SHAPE,
                     "temporal-spatial window attention: H=%d W=%d ws=%d Tq=%d Tk=%d (grid must be a multiple of the window)", H, W, ws, Tq, Tk);
        g.H = H; g.W = W; g.ws = ws; g.nwh = H / ws; g.nww = W / ws; g.Tq = Tq; g.Tk = Tk; g.HW = H * W;
        g.Lq = ws * ws; g.Lk = Tk * ws * ws;
        *batches = F_or_N * Tq * g.nwh * g.nww;
    } else {
        VPTR_REQUIRE(Tq > 0 && Tk > 0, VPTR_ERR_SHAPE, "temporal attention: Tq=%d Tk=%d", Tq, Tk);
        VPTR_REQUIRE(!causal || Tq == Tk, VPTR_ERR_SHAPE, "causal temporal attention needs Tq == Tk");
        g.Tq = Tq; g.Tk = Tk; g.HW = H * W; g.Lq = Tq; g.Lk = Tk;
        *batches = F_or_N * H * W;
    }
    return VPTR_OK;
}

}  // namespace

// tcgen05 / TMA forward (attn_tcgen05.cu); VPTR_ERR_UNSUPPORTED when the shape is outside its domain
extern "C" int vptr_attn_fwd_tcgen05(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv, float* O, long long ldo,
                          const float* rpe_table, int mode, int F_or_N, int H, int W, int ws, int Tq, int Tk, int nhead, int d, int causal,
                          float scale, int round_tf32, unsigned long long drop_seed, float drop_p, cudaStream_t stream);
static bool attn_tc_enabled() {
    static const bool v = [] { const char* e = getenv("VPTR_ATTN_TC"); return e && e[0] == '1'; }();
    return v;
}

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
    if (attn_tc_enabled()) {   // tcgen05 + TMA + TMEM forward
        rc = vptr_attn_fwd_tcgen05(Q, ldq, K, ldk, V, ldv, O, ldo, rpe_table, mode, F_or_N, H, W, ws, Tq, Tk, nhead, d, causal, scale, round_tf32,
                                   drop_seed, drop_p, stream);
        if (rc != VPTR_ERR_UNSUPPORTED) return rc;
    }
    if (attn_mma_ok(g, nhead, d) && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 && ((uintptr_t)Q % 16 == 0) &&
        ((uintptr_t)K % 16 == 0) && ((uintptr_t)V % 16 == 0) && ((uintptr_t)O % 16 == 0))   // tensor-core path (mma.sync, 3xTF32)
        return dispatch_attn_mma<false>(Q, ldq, K, ldk, V, ldv, nullptr, O, ldo, nullptr, 0, nullptr, 0, 0, rpe_table, nullptr, g, batches, stream);
    if (!attn_mma_disabled() && attn_mma64_ok(g, nhead, d, rpe_table != nullptr) && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 &&
        ((uintptr_t)Q % 16 == 0) && ((uintptr_t)K % 16 == 0) && ((uintptr_t)V % 16 == 0) && ((uintptr_t)O % 16 == 0))   // wide tensor-core path
        return launch_attn_mma64<false>(Q, ldq, K, ldk, V, ldv, nullptr, O, ldo, nullptr, 0, nullptr, 0, 0, rpe_table, nullptr, g, batches, stream);
    {   // one CTA per batch entry, all heads (scalar)
        const int C = nhead * d;
        const size_t ah = sizeof(float) * ((size_t)(g.Lq + 2 * g.Lk) * (C + 2) + (((size_t)nhead * g.Lq * (g.Lk + 1) + 1) & ~(size_t)1)) +
                          sizeof(long long) * (size_t)(g.Lq + g.Lk);
        if (!attn_no_allheads() && g.mode != 2 && d % 2 == 0 && C % 4 == 0 && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 && ah <= 220 * 1024 &&
            ((uintptr_t)Q % 16 == 0) && ((uintptr_t)K % 16 == 0) && ((uintptr_t)V % 16 == 0) && ((uintptr_t)O % 16 == 0)) {
            if (ah > 48 * 1024) cudaFuncSetAttribute(attn_fwd_allheads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ah);
            const int gx = batches < 148 * 8 ? batches : 148 * 8;
            attn_fwd_allheads_kernel<<<gx, 256, ah, stream>>>(Q, ldq, K, ldk, V, ldv, O, ldo, rpe_table, g, batches);
            return vptr_check_launch("attn_fwd_allheads_kernel");
        }
    }
    VPTR_REQUIRE(d % 2 == 0 && ldq % 2 == 0 && ldk % 2 == 0 && ldv % 2 == 0 && ldo % 2 == 0 && ((uintptr_t)Q % 8 == 0) &&
                     ((uintptr_t)K % 8 == 0) && ((uintptr_t)V % 8 == 0) && ((uintptr_t)O % 8 == 0),
                 VPTR_ERR_ALIGN, "vptr_attn_fwd: head_dim and pitches must be even, pointers 8-byte aligned");
    const int dpitch = ((d + 3) & ~3) + 4;
    size_t smem = sizeof(float) * ((size_t)(g.Lq + 2 * g.Lk) * dpitch + (size_t)g.Lq * (g.Lk + 1));
    VPTR_REQUIRE(smem <= 200 * 1024, VPTR_ERR_UNSUPPORTED, "vptr_attn_fwd: tile too large (%zu B of shared memory)", smem);
    if (smem > 48 * 1024) cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(batches < 65535 * 8 ? batches : 65535 * 8, nhead);
    attn_fwd_kernel<<<grid, 128, smem, stream>>>(Q, ldq, K, ldk, V, ldv, O, ldo, rpe_table, g, batches);
    return vptr_check_launch("attn_fwd_kernel");
}

extern "C" int vptr_attn_bwd_bias(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv,
                                  const float* dO, long long ldo, float* dQ, long long lddq, float* dK, long long lddk, float* dV,
                                  long long lddv, const float* rpe_table, float* d_rpe_table, int mode, int F_or_N, int H, int W, int ws,
                                  int Tq, int Tk, int nhead, int d, int causal, float scale, int round_tf32, unsigned long long drop_seed,
                                  float drop_p, float* dbq, float* dbk, float* dbv, cudaStream_t stream);
extern "C" int vptr_attn_bwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv,
                             const float* dO, long long ldo, float* dQ, long long lddq, float* dK, long long lddk, float* dV,
                             long long lddv, const float* rpe_table, float* d_rpe_table, int mode, int F_or_N, int H, int W, int ws,
                             int Tq, int Tk, int nhead, int d, int causal, float scale, int round_tf32, unsigned long long drop_seed,
                             float drop_p, cudaStream_t stream) {
    return vptr_attn_bwd_bias(Q, ldq, K, ldk, V, ldv, dO, ldo, dQ, lddq, dK, lddk, dV, lddv, rpe_table, d_rpe_table, mode, F_or_N, H, W, ws, Tq, Tk,
                              nhead, d, causal, scale, round_tf32, drop_seed, drop_p, nullptr, nullptr, nullptr, stream);
}
extern "C" int vptr_colsum(const float* x, float* out, long long rows, int C, long long ld, cudaStream_t stream);
// Same, and dbq / dbk / dbv ([nhead*d] each, any may be NULL) += the column sums of dQ / dK / dV -- the bias gradients of the q / k / v
// projections (in_proj_bias of nn.MultiheadAttention, q/k/v_proj.bias of MultiheadAttentionRPE).  The tensor-core kernels produce them
// while storing the tiles; the scalar fallbacks run the stand-alone column-sum kernel afterwards.
extern "C" int vptr_attn_bwd_bias(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv,
                                  const float* dO, long long ldo, float* dQ, long long lddq, float* dK, long long lddk, float* dV,
                                  long long lddv, const float* rpe_table, float* d_rpe_table, int mode, int F_or_N, int H, int W, int ws,
                                  int Tq, int Tk, int nhead, int d, int causal, float scale, int round_tf32, unsigned long long drop_seed,
                                  float drop_p, float* dbq, float* dbk, float* dbv, cudaStream_t stream) {
    AttnGeom g;
    int batches = 0;
    int rc = fill_geom(g, mode, F_or_N, H, W, ws, Tq, Tk, nhead, d, causal, scale, &batches);
    if (rc) return rc;
    g.round_tf32 = round_tf32;
    g.drop_seed = drop_seed;
    g.drop_p = drop_p;
    g.dbq = dbq; g.dbk = dbk; g.dbv = dbv;
    struct AfterSums {   // scalar fallbacks: bias gradients by the separate column-sum pass
        float *dbq, *dbk, *dbv; const float *dQ, *dK, *dV; long long lq, lk, lv, rq, rk; int C; cudaStream_t st;
        int run() const {
            int r = 0;
            if (dbq && !r) r = vptr_colsum(dQ, dbq, rq, C, lq, st);
            if (dbk && !r) r = vptr_colsum(dK, dbk, rk, C, lk, st);
            if (dbv && !r) r = vptr_colsum(dV, dbv, rk, C, lv, st);
            return r;
        }
    } after{dbq, dbk, dbv, dQ, dK, dV, lddq, lddk, lddv, (long long)batches * g.Lq, (long long)batches * g.Lk, nhead * d, stream};
    VPTR_REQUIRE(batches > 0 && nhead > 0 && d > 0, VPTR_ERR_SHAPE, "vptr_attn_bwd: empty problem");
    if (attn_mma_ok(g, nhead, d) && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 && lddq % 4 == 0 && lddk % 4 == 0 &&
        lddv % 4 == 0 && ((uintptr_t)Q % 16 == 0) && ((uintptr_t)K % 16 == 0) && ((uintptr_t)V % 16 == 0) && ((uintptr_t)dO % 16 == 0) &&
        ((uintptr_t)dQ % 16 == 0) && ((uintptr_t)dK % 16 == 0) && ((uintptr_t)dV % 16 == 0))   // tensor-core path (mma.sync, 3xTF32)
        return dispatch_attn_mma<true>(Q, ldq, K, ldk, V, ldv, dO, dQ, lddq, dK, lddk, dV, lddv, ldo, rpe_table, d_rpe_table, g, batches, stream);
    if (!attn_mma_disabled() && attn_mma64_ok(g, nhead, d, rpe_table != nullptr) && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 &&
        lddq % 4 == 0 && lddk % 4 == 0 && lddv % 4 == 0 && ((uintptr_t)Q % 16 == 0) && ((uintptr_t)K % 16 == 0) && ((uintptr_t)V % 16 == 0) &&
        ((uintptr_t)dO % 16 == 0) && ((uintptr_t)dQ % 16 == 0) && ((uintptr_t)dK % 16 == 0) && ((uintptr_t)dV % 16 == 0))   // wide tensor-core path
        return launch_attn_mma64<true>(Q, ldq, K, ldk, V, ldv, dO, dQ, lddq, dK, lddk, dV, lddv, ldo, rpe_table, d_rpe_table, g, batches, stream);
    {   // one CTA per batch entry, all heads (scalar)
        const int C = nhead * d;
        const size_t ah = sizeof(float) * ((size_t)(2 * g.Lq + 2 * g.Lk) * (C + 2) + (size_t)2 * nhead * g.Lq * (g.Lk + 1) +
                                           (((size_t)nhead * g.Lq * g.Lk + 1) & ~(size_t)1)) + sizeof(long long) * (size_t)(g.Lq + g.Lk);
        if (!attn_no_allheads() && g.mode != 2 && d % 2 == 0 && C % 4 == 0 && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 && lddq % 4 == 0 &&
            lddk % 4 == 0 && lddv % 4 == 0 && ah <= 220 * 1024 && ((uintptr_t)Q % 16 == 0) && ((uintptr_t)K % 16 == 0) &&
            ((uintptr_t)V % 16 == 0) && ((uintptr_t)dO % 16 == 0) && ((uintptr_t)dQ % 16 == 0) && ((uintptr_t)dK % 16 == 0) &&
            ((uintptr_t)dV % 16 == 0)) {
            if (ah > 48 * 1024) cudaFuncSetAttribute(attn_bwd_allheads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ah);
            int gx = batches < 148 * 4 ? batches : 148 * 4;
            attn_bwd_allheads_kernel<<<gx, 256, ah, stream>>>(Q, ldq, K, ldk, V, ldv, dO, ldo, dQ, lddq, dK, lddk, dV, lddv, rpe_table,
                                                              d_rpe_table, g, batches);
            rc = vptr_check_launch("attn_bwd_allheads_kernel");
            return rc ? rc : after.run();
        }
    }
    VPTR_REQUIRE(d % 2 == 0 && ldq % 2 == 0 && ldk % 2 == 0 && ldv % 2 == 0 && ldo % 2 == 0 && lddq % 2 == 0 && lddk % 2 == 0 && lddv % 2 == 0 &&
                     ((uintptr_t)Q % 8 == 0) && ((uintptr_t)K % 8 == 0) && ((uintptr_t)V % 8 == 0) && ((uintptr_t)dO % 8 == 0) &&
                     ((uintptr_t)dQ % 8 == 0) && ((uintptr_t)dK % 8 == 0) && ((uintptr_t)dV % 8 == 0),
                 VPTR_ERR_ALIGN, "vptr_attn_bwd: head_dim and pitches must be even, pointers 8-byte aligned");
    const int dpitch = ((d + 3) & ~3) + 4;
    const int lmax = g.Lq > g.Lk ? g.Lq : g.Lk;
    size_t smem = sizeof(float) * ((size_t)(2 * g.Lq + 2 * g.Lk + lmax) * dpitch + (size_t)2 * g.Lq * (g.Lk + 1) + (size_t)g.Lq * g.Lk);
    VPTR_REQUIRE(smem <= 200 * 1024, VPTR_ERR_UNSUPPORTED, "vptr_attn_bwd: tile too large (%zu B of shared memory)", smem);
    if (smem > 48 * 1024) cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int gx = batches;
    if (d_rpe_table && gx > 148 * 4) gx = 148 * 4;  // bound the number of bias-gradient flushes
    dim3 grid(gx, nhead);
    attn_bwd_kernel<<<grid, 128, smem, stream>>>(Q, ldq, K, ldk, V, ldv, dO, ldo, dQ, lddq, dK, lddk, dV, lddv, rpe_table,
                                                 d_rpe_table, g, batches);
    rc = vptr_check_launch("attn_bwd_kernel");
    return rc ? rc : after.run();
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

VPTR_RNG_EPOCH_ACCESSOR(attn)
