// Normalisation kernels of the VidHRFormer blocks (token-major fp32 activations [rows][C]):
//   * LayerNorm over C (pre-LN of every sub-block, VidHRFormer_modules.py:44-56,137-161) with the
//     positional add fused as a second output (q/k source = LN(x)+pos, v source = LN(x)).
//   * the MlpDWBN norms (VidHRFormer_modules.py:397-400,424-442): BatchNorm2d(ch) (NAR encoder) or
//     LayerNorm((ch,H,W)) per frame (FAR, NAR decoder), each followed by exact GELU; norm3 also
//     carries the residual add.  Statistics are reduced in fp64.
#include "common.cuh"

namespace {

// branch regularisation around the MlpDWBN activations: y = rowscale[row / rows_per_group] * dropout(GELU(norm(x)))
// (nn.Dropout after act2 / act3, VidHRFormer_modules.py:436-441; DropPath on the branch, :68-71,563-575)
struct DropArgs {
    const float* rowscale;
    int rows_per_group;
    unsigned long long seed;
    float p;
};
// the four consecutive elements e .. e+3 (e % 4 == 0, one row): one hash
__device__ __forceinline__ float4 drop_factor4(const DropArgs& d, long long row, long long e) {
    float4 f = vptr_drop_scale4(d.seed, (unsigned long long)e >> 2, d.p);
    if (d.rowscale) {
        const float rs = __ldg(d.rowscale + row / d.rows_per_group);
        f.x *= rs; f.y *= rs; f.z *= rs; f.w *= rs;
    }
    return f;
}

// ------------------------------------------------------------------ LayerNorm(C) forward
// one warp per row; y = LN(x)*g+b ; y2 = y + add[((row / add_div) % add_mod)]  (both optional)
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y,
                                                            float* __restrict__ y2, const float* __restrict__ add,
                                                            int add_div, int add_mod, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, long long rows, int C, float eps,
                                                            int relu, int round_tf32) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * C;
    float s = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(xr + c);
        s += v.x + v.y + v.z + v.w;
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(xr + c);
        float a = v.x - mean, b = v.y - mean, d = v.z - mean, e = v.w - mean;
        q += a * a + b * b + d * d + e * e;
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
    const float* ar = add ? add + (long long)((row / add_div) % add_mod) * C : nullptr;
    for (int c = lane * 4; c < C; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(xr + c);
        float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
        float4 o;
        o.x = (v.x - mean) * rstd * g.x + b.x;
        o.y = (v.y - mean) * rstd * g.y + b.y;
        o.z = (v.z - mean) * rstd * g.z + b.z;
        o.w = (v.w - mean) * rstd * g.w + b.w;
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (y) {
            float4 w = o;
            if (round_tf32) { w.x = vptr_round_tf32(w.x); w.y = vptr_round_tf32(w.y); w.z = vptr_round_tf32(w.z); w.w = vptr_round_tf32(w.w); }
            *reinterpret_cast<float4*>(y + row * C + c) = w;
        }
        if (y2) {
            float4 p = __ldg(reinterpret_cast<const float4*>(ar + c));
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
            if (round_tf32) { o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w); }
            *reinterpret_cast<float4*>(y2 + row * C + c) = o;
        }
    }
}

// ------------------------------------------------------------------ LayerNorm(C) backward (dx)
// g = (dy1 + dy2) [* (y>0) when relu] ; dx = dres + rstd*(g*gamma - mean(g*gamma) - xhat*mean(g*gamma*xhat))
__global__ void __launch_bounds__(256) layernorm_bwd_dx_kernel(const float* __restrict__ dy1, const float* __restrict__ dy2,
                                                               const float* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const float* __restrict__ mean_in,
                                                               const float* __restrict__ rstd_in, const float* __restrict__ dres,
                                                               float* __restrict__ dx, long long rows, int C, int relu) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float mean = mean_in[row], rstd = rstd_in[row];
    const float* xr = x + row * C;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
        float g = dy1[row * C + c];
        if (dy2) g += dy2[row * C + c];
        float xh = (xr[c] - mean) * rstd;
        if (relu && xh * gamma[c] + beta[c] <= 0.f) g = 0.f;
        g *= gamma[c];
        s1 += g;
        s2 += g * xh;
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    for (int c = lane; c < C; c += 32) {
        float g = dy1[row * C + c];
        if (dy2) g += dy2[row * C + c];
        float xh = (xr[c] - mean) * rstd;
        if (relu && xh * gamma[c] + beta[c] <= 0.f) g = 0.f;
        g *= gamma[c];
        float o = rstd * (g - s1 - xh * s2);
        if (dres) o += dres[row * C + c];
        dx[row * C + c] = o;
    }
}

// dgamma[c] += sum_r g*xhat ; dbeta[c] += sum_r g   (thread per column, row chunk per blockIdx.y)
__global__ void __launch_bounds__(128) layernorm_bwd_affine_kernel(const float* __restrict__ dy1, const float* __restrict__ dy2,
                                                                   const float* __restrict__ x, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, const float* __restrict__ mean_in,
                                                                   const float* __restrict__ rstd_in, float* __restrict__ dgamma,
                                                                   float* __restrict__ dbeta, long long rows, int C, int rows_per_block,
                                                                   int relu) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, rows);
    const float gm = gamma[c], bt = beta[c];
    float ag = 0.f, ab = 0.f;
    for (long long r = r0; r < r1; ++r) {
        float g = dy1[r * C + c];
        if (dy2) g += dy2[r * C + c];
        float xh = (x[r * C + c] - mean_in[r]) * rstd_in[r];
        if (relu && xh * gm + bt <= 0.f) g = 0.f;
        ag += g * xh;
        ab += g;
    }
    atomicAdd(dgamma + c, ag);
    atomicAdd(dbeta + c, ab);
}

// ------------------------------------------------------------------ LayerNorm(C) backward, fused (C <= 1024, C % 4 == 0)
// One pass: a warp walks rows; the row's g = dy1 (+ dy2) and x live in registers (float4 per lane, up to 8), so dx needs no
// second read, and every lane accumulates dgamma / dbeta of its own columns over all the warp's rows in registers -- the
// separate affine kernel (which re-read dy1, dy2 and x) disappears.  Per-block partial sums are combined through shared memory
// and flushed with one atomic per column and block.
template <int NV>   // float4 per lane (C <= 128 * NV)
__global__ void __launch_bounds__(256, 2) layernorm_bwd_fused_kernel(const float* __restrict__ dy1, const float* __restrict__ dy2,
                                                                     const float* __restrict__ x, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, const float* __restrict__ mean_in,
                                                                     const float* __restrict__ rstd_in, const float* __restrict__ dres,
                                                                     float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                     long long rows, int C, int relu, int rows_per_block) {
    extern __shared__ float red[];   // [2][C] block accumulators
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int C4 = C >> 2;
    for (int e = threadIdx.x; e < 2 * C; e += blockDim.x) red[e] = 0.f;
    __syncthreads();
    float4 ag[NV], ab[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) ag[v] = ab[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* gm4 = reinterpret_cast<const float4*>(gamma);   // gamma / beta stay L1-resident: re-read per row, not held in registers
    const float4* bt4 = reinterpret_cast<const float4*>(beta);
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, rows);
    const float invC = 1.f / (float)C;
    for (long long row = r0 + warp; row < r1; row += nw) {
        const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
        float4 g[NV], xh[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {   // all of the row's loads first (independent, in flight together)
            const int c4 = lane + 32 * v;
            g[v] = xh[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c4 < C4) {
                g[v] = reinterpret_cast<const float4*>(dy1 + row * C)[c4];
                xh[v] = reinterpret_cast<const float4*>(x + row * C)[c4];
            }
        }
        if (dy2) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c4 = lane + 32 * v;
                if (c4 < C4) { const float4 t = reinterpret_cast<const float4*>(dy2 + row * C)[c4]; g[v].x += t.x; g[v].y += t.y; g[v].z += t.z; g[v].w += t.w; }
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c4 = lane + 32 * v;
            if (c4 < C4) {
                const float4 gm = __ldg(gm4 + c4);
                xh[v] = make_float4((xh[v].x - mean) * rstd, (xh[v].y - mean) * rstd, (xh[v].z - mean) * rstd, (xh[v].w - mean) * rstd);
                if (relu) {
                    const float4 bt = __ldg(bt4 + c4);
                    if (xh[v].x * gm.x + bt.x <= 0.f) g[v].x = 0.f;
                    if (xh[v].y * gm.y + bt.y <= 0.f) g[v].y = 0.f;
                    if (xh[v].z * gm.z + bt.z <= 0.f) g[v].z = 0.f;
                    if (xh[v].w * gm.w + bt.w <= 0.f) g[v].w = 0.f;
                }
                ab[v].x += g[v].x; ab[v].y += g[v].y; ab[v].z += g[v].z; ab[v].w += g[v].w;
                ag[v].x = fmaf(g[v].x, xh[v].x, ag[v].x); ag[v].y = fmaf(g[v].y, xh[v].y, ag[v].y);
                ag[v].z = fmaf(g[v].z, xh[v].z, ag[v].z); ag[v].w = fmaf(g[v].w, xh[v].w, ag[v].w);
                g[v].x *= gm.x; g[v].y *= gm.y; g[v].z *= gm.z; g[v].w *= gm.w;
                s1 += (g[v].x + g[v].y) + (g[v].z + g[v].w);
                s2 += (g[v].x * xh[v].x + g[v].y * xh[v].y) + (g[v].z * xh[v].z + g[v].w * xh[v].w);
            }
        }
        if (dx) {
            s1 = warp_sum(s1) * invC;
            s2 = warp_sum(s2) * invC;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c4 = lane + 32 * v;
                if (c4 < C4) {
                    float4 o = make_float4(rstd * (g[v].x - s1 - xh[v].x * s2), rstd * (g[v].y - s1 - xh[v].y * s2),
                                           rstd * (g[v].z - s1 - xh[v].z * s2), rstd * (g[v].w - s1 - xh[v].w * s2));
                    if (dres) { const float4 t = reinterpret_cast<const float4*>(dres + row * C)[c4]; o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
                    reinterpret_cast<float4*>(dx + row * C)[c4] = o;
                }
            }
        }
    }
    if (dgamma) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = (lane + 32 * v) * 4;
            if (c < C) {
                atomicAdd(red + c, ag[v].x); atomicAdd(red + c + 1, ag[v].y); atomicAdd(red + c + 2, ag[v].z); atomicAdd(red + c + 3, ag[v].w);
                atomicAdd(red + C + c, ab[v].x); atomicAdd(red + C + c + 1, ab[v].y); atomicAdd(red + C + c + 2, ab[v].z); atomicAdd(red + C + c + 3, ab[v].w);
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            atomicAdd(dgamma + c, red[c]);
            atomicAdd(dbeta + c, red[C + c]);
        }
    }
}

// ------------------------------------------------------------------ statistics
// per-column sum / sum of squares over rows (BatchNorm batch statistics), fp64 accumulators
__global__ void __launch_bounds__(128) colstats_kernel(const float* __restrict__ x, double* __restrict__ sum, double* __restrict__ sumsq,
                                                       long long rows, int ch, int rows_per_block) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ch) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, rows);
    float s = 0.f, q = 0.f;
    for (long long r = r0; r < r1; ++r) {
        float v = x[r * ch + c];
        s += v;
        q = fmaf(v, v, q);
    }
    atomicAdd(sum + c, (double)s);
    atomicAdd(sumsq + c, (double)q);
}

// mean/rstd from the fp64 sums; updates running stats like torch BatchNorm2d (momentum, unbiased var)
__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* running_mean, float* running_var, long long n, int ch,
                                   float eps, float momentum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ch) return;
    double m = sum[c] / (double)n;
    double var = sumsq[c] / (double)n - m * m;
    if (var < 0) var = 0;
    mean[c] = (float)m;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
}

// eval-mode BatchNorm: mean = running_mean, rstd = 1/sqrt(running_var + eps)
__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                     float* __restrict__ mean, float* __restrict__ rstd, int ch, float eps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ch) return;
    mean[c] = running_mean[c];
    rstd[c] = rsqrtf(running_var[c] + eps);
}

// per-group (frame) mean / rstd over `gsize` contiguous elements: one block per group
__global__ void __launch_bounds__(512) groupstats_kernel(const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ rstd,
                                                         long long gsize, float eps) {
    __shared__ double red[2][16];
    const float* xg = x + (long long)blockIdx.x * gsize;
    float s = 0.f, q = 0.f;
    for (long long i = threadIdx.x * 4; i < gsize; i += blockDim.x * 4) {
        float4 v = *reinterpret_cast<const float4*>(xg + i);
        s += v.x + v.y + v.z + v.w;
        q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    double ds = warp_sum(s), dq = warp_sum(q);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { red[0][w] = ds; red[1][w] = dq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0, tq = 0;
        for (int i = 0; i < (blockDim.x >> 5); ++i) { ts += red[0][i]; tq += red[1][i]; }
        double m = ts / (double)gsize;
        double var = tq / (double)gsize - m * m;
        if (var < 0) var = 0;
        mean[blockIdx.x] = (float)m;
        rstd[blockIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// mean / rstd of `groups` groups from fp64 (sum, sum of squares) pairs accumulated by a producer (vptr_dwconv3x3_stats)
__global__ void groupstats_finalize_kernel(const double* __restrict__ sums, int groups, double gsize, float* __restrict__ mean,
                                           float* __restrict__ rstd, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups) return;
    const double m = sums[i] / gsize;
    double var = sums[groups + i] / gsize - m * m;
    if (var < 0) var = 0;
    mean[i] = (float)m;
    rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

// ------------------------------------------------------------------ norm + GELU (+ residual) forward
// mode 0 (BatchNorm): stats and affine indexed by channel c.
// mode 1 (frame LayerNorm): stats indexed by frame = row / hw, affine indexed by (row % hw)*ch + c.
// Grid = (chunks of a frame, frames): every index below is 32-bit and relative to the frame -- the flat-index version spent two
// 64-bit divisions (e / ch, row / hw) plus one more for DropPath per float4 and was ALU-bound at ~2x its HBM time.
template <int MODE>
__global__ void __launch_bounds__(256) norm_act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           const float* __restrict__ res, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, int frame4, int C4, int hw,
                                                           int round_tf32, const DropArgs da) {
    const int f = blockIdx.y;
    const long long fbase = (long long)f * frame4;                    // float4 index of the frame's first element
    float mm = 0.f, rr = 0.f;
    if (MODE == 1) { mm = __ldg(mean + f); rr = __ldg(rstd + f); }
    const bool drop = da.p > 0.f;
    float rs = 1.f;                                                   // DropPath: clips are whole frames (rows_per_group % hw == 0)
    if (da.rowscale) rs = __ldg(da.rowscale + ((long long)f * hw) / da.rows_per_group);
    // each block streams one contiguous chunk of the frame (DRAM-page friendly), 256 float4 per step
    const int per = ((frame4 + (int)gridDim.x - 1) / (int)gridDim.x + 255) & ~255;
    const int i_end = min((int)(blockIdx.x + 1) * per, frame4);
#pragma unroll 2
    for (int i = blockIdx.x * per + threadIdx.x; i < i_end; i += 256) {
        const float4 v = reinterpret_cast<const float4*>(x)[fbase + i];
        float4 m, r, g, b;
        if (MODE == 0) {
            const int c4 = i % C4;
            m = __ldg(reinterpret_cast<const float4*>(mean) + c4);
            r = __ldg(reinterpret_cast<const float4*>(rstd) + c4);
            g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
            b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
        } else {
            m = make_float4(mm, mm, mm, mm);
            r = make_float4(rr, rr, rr, rr);
            g = __ldg(reinterpret_cast<const float4*>(gamma) + i);
            b = __ldg(reinterpret_cast<const float4*>(beta) + i);
        }
        float4 o;
        o.x = vptr_gelu((v.x - m.x) * r.x * g.x + b.x);
        o.y = vptr_gelu((v.y - m.y) * r.y * g.y + b.y);
        o.z = vptr_gelu((v.z - m.z) * r.z * g.z + b.z);
        o.w = vptr_gelu((v.w - m.w) * r.w * g.w + b.w);
        if (drop) {
            const float4 k = vptr_drop_scale4(da.seed, (unsigned long long)(fbase + i), da.p);
            o.x *= k.x; o.y *= k.y; o.z *= k.z; o.w *= k.w;
        }
        o.x *= rs; o.y *= rs; o.z *= rs; o.w *= rs;
        if (res) {
            const float4 q = reinterpret_cast<const float4*>(res)[fbase + i];
            o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
        }
        if (round_tf32) { o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w); }
        reinterpret_cast<float4*>(y)[fbase + i] = o;
    }
}

// ------------------------------------------------------------------ norm + GELU backward
// Two-stage backward.  Pass A evaluates g0 = dy * dropfactor * GELU'(z) ONCE (the erf/exp are the expensive part), stores it in
// the dx buffer (which may alias dy) and reduces what the normalisation's backward needs; the later passes read g0.
//
// BatchNorm, pass A: per channel S1 = sum g0, S2 = sum g0*xhat (dbeta = S1, dgamma = S2).  4 channels per thread.
__global__ void __launch_bounds__(128) bn_act_bwd_pass_a_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ g0, float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta, float* __restrict__ s1, float* __restrict__ s2,
                                                                long long rows, int ch, int rows_per_block, const DropArgs da) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (c >= ch) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, rows);
    const float4 m = *reinterpret_cast<const float4*>(mean + c), r = *reinterpret_cast<const float4*>(rstd + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
    const bool drop = da.p > 0.f || da.rowscale;
    for (long long row = r0; row < r1; ++row) {
        const long long e = row * ch + c;
        const float4 xv = *reinterpret_cast<const float4*>(x + e);
        float4 d = *reinterpret_cast<const float4*>(dy + e);
        if (drop) { const float4 k = drop_factor4(da, row, e); d.x *= k.x; d.y *= k.y; d.z *= k.z; d.w *= k.w; }
        float4 xh = make_float4((xv.x - m.x) * r.x, (xv.y - m.y) * r.y, (xv.z - m.z) * r.z, (xv.w - m.w) * r.w);
        d.x *= vptr_gelu_grad(xh.x * g.x + b.x); d.y *= vptr_gelu_grad(xh.y * g.y + b.y);
        d.z *= vptr_gelu_grad(xh.z * g.z + b.z); d.w *= vptr_gelu_grad(xh.w * g.w + b.w);
        *reinterpret_cast<float4*>(g0 + e) = d;
        a1.x += d.x; a1.y += d.y; a1.z += d.z; a1.w += d.w;
        a2.x = fmaf(d.x, xh.x, a2.x); a2.y = fmaf(d.y, xh.y, a2.y); a2.z = fmaf(d.z, xh.z, a2.z); a2.w = fmaf(d.w, xh.w, a2.w);
    }
    atomicAdd(s1 + c, a1.x); atomicAdd(s1 + c + 1, a1.y); atomicAdd(s1 + c + 2, a1.z); atomicAdd(s1 + c + 3, a1.w);
    atomicAdd(s2 + c, a2.x); atomicAdd(s2 + c + 1, a2.y); atomicAdd(s2 + c + 2, a2.z); atomicAdd(s2 + c + 3, a2.w);
    atomicAdd(dbeta + c, a1.x); atomicAdd(dbeta + c + 1, a1.y); atomicAdd(dbeta + c + 2, a1.z); atomicAdd(dbeta + c + 3, a1.w);
    atomicAdd(dgamma + c, a2.x); atomicAdd(dgamma + c + 1, a2.y); atomicAdd(dgamma + c + 2, a2.z); atomicAdd(dgamma + c + 3, a2.w);
}

// frame LayerNorm, passes A + affine merged: a thread owns four (hw, ch) positions and walks a chunk of frames, so the affine
// gradients dgamma / dbeta (sums over frames at fixed position) accumulate in registers and gamma / beta are read once, while
// the per-frame sums P1 = sum g0*gamma, P2 = sum g0*gamma*xhat (needed by the dx pass) go warp-reduce -> shared -> one atomic
// per frame and block.  Saves the separate affine kernel's re-read of g0 and x (692 MB per 2112-channel call at cfg1).
constexpr int LN3_FPB = 64;   // frames per block
__global__ void __launch_bounds__(256) ln3_act_bwd_pass_ab_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  float* __restrict__ g0, float* __restrict__ p1, float* __restrict__ p2,
                                                                  float* __restrict__ dgamma, float* __restrict__ dbeta, long long gsize, int ch,
                                                                  int frames, const DropArgs da) {
    __shared__ float part[LN3_FPB][8][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long a = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const bool live = a < gsize;
    const int f0 = blockIdx.y * LN3_FPB;
    const int f1 = min(f0 + LN3_FPB, frames);
    const bool drop = da.p > 0.f || da.rowscale;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f), b = g, ag = g, ab = g;
    if (live) { g = __ldg(reinterpret_cast<const float4*>(gamma + a)); b = __ldg(reinterpret_cast<const float4*>(beta + a)); }
#pragma unroll 2
    for (int f = f0; f < f1; ++f) {
        float a1 = 0.f, a2 = 0.f;
        if (live) {
            const float m = __ldg(mean + f), r = __ldg(rstd + f);
            const long long e = (long long)f * gsize + a;
            const float4 xv = *reinterpret_cast<const float4*>(x + e);
            float4 d = *reinterpret_cast<const float4*>(dy + e);
            if (drop) {
                const float4 k = drop_factor4(da, (long long)f * (int)(gsize / ch) + (int)a / ch, e);   // 32-bit in-frame row; 4 elements share a row
                d.x *= k.x; d.y *= k.y; d.z *= k.z; d.w *= k.w;
            }
            const float4 xh = make_float4((xv.x - m) * r, (xv.y - m) * r, (xv.z - m) * r, (xv.w - m) * r);
            d.x *= vptr_gelu_grad(xh.x * g.x + b.x); d.y *= vptr_gelu_grad(xh.y * g.y + b.y);
            d.z *= vptr_gelu_grad(xh.z * g.z + b.z); d.w *= vptr_gelu_grad(xh.w * g.w + b.w);
            *reinterpret_cast<float4*>(g0 + e) = d;
            ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
            ag.x = fmaf(d.x, xh.x, ag.x); ag.y = fmaf(d.y, xh.y, ag.y); ag.z = fmaf(d.z, xh.z, ag.z); ag.w = fmaf(d.w, xh.w, ag.w);
            const float4 t = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
            a1 = (t.x + t.y) + (t.z + t.w);
            a2 = (t.x * xh.x + t.y * xh.y) + (t.z * xh.z + t.w * xh.w);
        }
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (lane == 0) { part[f - f0][warp][0] = a1; part[f - f0][warp][1] = a2; }
    }
    if (live) {
        atomicAdd(dgamma + a, ag.x); atomicAdd(dgamma + a + 1, ag.y); atomicAdd(dgamma + a + 2, ag.z); atomicAdd(dgamma + a + 3, ag.w);
        atomicAdd(dbeta + a, ab.x); atomicAdd(dbeta + a + 1, ab.y); atomicAdd(dbeta + a + 2, ab.z); atomicAdd(dbeta + a + 3, ab.w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (f1 - f0) * 2; i += blockDim.x) {
        const int fi = i >> 1, w = i & 1;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += part[fi][k][w];
        atomicAdd((w ? p2 : p1) + f0 + fi, s);
    }
}

// Final pass (in place on the g0 buffer): dx from g0 and the reduced sums.
//  MODE 0: dx = gamma*rstd*(g0 - S1/n - xhat*S2/n)        (S indexed by channel, n = rows)
//  MODE 1: dx = rstd_f*(g0*gamma - P1_f/n - xhat*P2_f/n)   (P indexed by frame,  n = hw*ch)
//  MODE 2: eval BatchNorm (statistics are constants): dx = gamma*rstd*g0
template <int MODE>
__global__ void __launch_bounds__(256) norm_act_bwd_dx_kernel(float* __restrict__ g0dx, const float* __restrict__ x,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma, const float* __restrict__ s1,
                                                              const float* __restrict__ s2, int frame4, int C4, float inv_n, int round_tf32) {
    const int f = blockIdx.y;                                          // (chunk, frame) grid, 32-bit in-frame indices (see forward)
    const long long fbase = (long long)f * frame4;
    float mm = 0.f, rr = 0.f, t1 = 0.f, t2 = 0.f;
    if (MODE == 1) { mm = __ldg(mean + f); rr = __ldg(rstd + f); t1 = __ldg(s1 + f) * inv_n; t2 = __ldg(s2 + f) * inv_n; }
    // each block streams one contiguous chunk of the frame (DRAM-page friendly), 256 float4 per step
    const int per = ((frame4 + (int)gridDim.x - 1) / (int)gridDim.x + 255) & ~255;
    const int i_end = min((int)(blockIdx.x + 1) * per, frame4);
#pragma unroll 2
    for (int i = blockIdx.x * per + threadIdx.x; i < i_end; i += 256) {
        const float4 d = reinterpret_cast<const float4*>(g0dx)[fbase + i];
        const float4 xv = reinterpret_cast<const float4*>(x)[fbase + i];
        float4 o;
        if (MODE == 1) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i);
            o.x = rr * (d.x * g.x - t1 - (xv.x - mm) * rr * t2); o.y = rr * (d.y * g.y - t1 - (xv.y - mm) * rr * t2);
            o.z = rr * (d.z * g.z - t1 - (xv.z - mm) * rr * t2); o.w = rr * (d.w * g.w - t1 - (xv.w - mm) * rr * t2);
        } else {
            const int c4 = i % C4;
            const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + c4), r = __ldg(reinterpret_cast<const float4*>(rstd) + c4);
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
            if (MODE == 0) {
                const float4 u1 = __ldg(reinterpret_cast<const float4*>(s1) + c4), u2 = __ldg(reinterpret_cast<const float4*>(s2) + c4);
                o.x = g.x * r.x * (d.x - u1.x * inv_n - (xv.x - m.x) * r.x * u2.x * inv_n);
                o.y = g.y * r.y * (d.y - u1.y * inv_n - (xv.y - m.y) * r.y * u2.y * inv_n);
                o.z = g.z * r.z * (d.z - u1.z * inv_n - (xv.z - m.z) * r.z * u2.z * inv_n);
                o.w = g.w * r.w * (d.w - u1.w * inv_n - (xv.w - m.w) * r.w * u2.w * inv_n);
            } else {
                o.x = g.x * r.x * d.x; o.y = g.y * r.y * d.y; o.z = g.z * r.z * d.z; o.w = g.w * r.w * d.w;
            }
        }
        if (round_tf32) { o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w); }
        reinterpret_cast<float4*>(g0dx)[fbase + i] = o;
    }
}

// Column-slab form of the final pass (see elementwise.cu): a thread owns one float4 column and walks rows_per_block rows, so the
// column sums of dx -- the bias gradient of the 1x1 conv in front of the norm (fc1 / fc2 of MlpDWBN) -- come out of this pass.
template <int MODE>
__global__ void __launch_bounds__(1024) norm_act_bwd_dx_colsum_kernel(float* __restrict__ g0dx, const float* __restrict__ x,
                                                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                      const float* __restrict__ gamma, const float* __restrict__ s1,
                                                                      const float* __restrict__ s2, long long rows, int C4, int hw, float inv_n,
                                                                      int round_tf32, float* __restrict__ colsum, int rows_per_block) {
    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (c4 >= C4) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, rows);
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f), r = m, g = m, u1 = m, u2 = m;
    if (MODE != 1) {
        m = __ldg(reinterpret_cast<const float4*>(mean) + c4); r = __ldg(reinterpret_cast<const float4*>(rstd) + c4);
        g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        if (MODE == 0) { u1 = __ldg(reinterpret_cast<const float4*>(s1) + c4); u2 = __ldg(reinterpret_cast<const float4*>(s2) + c4); }
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int f = (int)(r0 / hw), p = (int)(r0 - (long long)f * hw);      // frame and in-frame position of the current row
#pragma unroll 4
    for (long long row = r0; row < r1; ++row) {
        const long long i = row * C4 + c4;
        const float4 d = reinterpret_cast<const float4*>(g0dx)[i];
        const float4 xv = reinterpret_cast<const float4*>(x)[i];
        float4 o;
        if (MODE == 1) {
            const float mm = __ldg(mean + f), rr = __ldg(rstd + f), t1 = __ldg(s1 + f) * inv_n, t2 = __ldg(s2 + f) * inv_n;
            const float4 gg = __ldg(reinterpret_cast<const float4*>(gamma) + (long long)p * C4 + c4);
            o.x = rr * (d.x * gg.x - t1 - (xv.x - mm) * rr * t2); o.y = rr * (d.y * gg.y - t1 - (xv.y - mm) * rr * t2);
            o.z = rr * (d.z * gg.z - t1 - (xv.z - mm) * rr * t2); o.w = rr * (d.w * gg.w - t1 - (xv.w - mm) * rr * t2);
            if (++p == hw) { p = 0; ++f; }
        } else if (MODE == 0) {
            o.x = g.x * r.x * (d.x - u1.x * inv_n - (xv.x - m.x) * r.x * u2.x * inv_n);
            o.y = g.y * r.y * (d.y - u1.y * inv_n - (xv.y - m.y) * r.y * u2.y * inv_n);
            o.z = g.z * r.z * (d.z - u1.z * inv_n - (xv.z - m.z) * r.z * u2.z * inv_n);
            o.w = g.w * r.w * (d.w - u1.w * inv_n - (xv.w - m.w) * r.w * u2.w * inv_n);
        } else {
            o.x = g.x * r.x * d.x; o.y = g.y * r.y * d.y; o.z = g.z * r.z * d.z; o.w = g.w * r.w * d.w;
        }
        if (round_tf32) { o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w); }
        reinterpret_cast<float4*>(g0dx)[i] = o;
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    atomicAdd(colsum + 4 * c4, acc.x); atomicAdd(colsum + 4 * c4 + 1, acc.y); atomicAdd(colsum + 4 * c4 + 2, acc.z); atomicAdd(colsum + 4 * c4 + 3, acc.w);
}

// blocks per frame for the (chunk, frame) grids: enough CTAs to fill the GPU a few times over, >= 2 float4 per thread
int norm_chunks(int frame4, int frames) {
    int chunks = (frame4 + 2 * 256 - 1) / (2 * 256);
    const int want = (148 * 16 + frames - 1) / frames;
    if (chunks > want) chunks = want;
    return chunks < 1 ? 1 : chunks;
}

int ew_grid(long long n, int block) {
    long long g = (n + block - 1) / block;
    long long cap = 148LL * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" int vptr_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* y2, const float* add,
                                  int add_div, int add_mod, float* mean, float* rstd, long long rows, int C, float eps, int relu,
                                  int round_tf32, cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && C > 0 && C % 4 == 0, VPTR_ERR_SHAPE, "vptr_layernorm_fwd: rows=%lld C=%d (C %% 4 == 0 required)", rows, C);
    VPTR_REQUIRE(y2 == nullptr || (add != nullptr && add_div > 0 && add_mod > 0), VPTR_ERR_SHAPE, "vptr_layernorm_fwd: y2 needs add/add_div/add_mod");
    layernorm_fwd_kernel<<<vptr_cdiv(rows, 8), 256, 0, stream>>>(x, gamma, beta, y, y2, add, add_div, add_mod, mean, rstd, rows, C, eps, relu, round_tf32);
    return vptr_check_launch("layernorm_fwd_kernel");
}

extern "C" int vptr_layernorm_bwd(const float* dy1, const float* dy2, const float* x, const float* gamma, const float* beta,
                                  const float* mean, const float* rstd, const float* dres, float* dx, float* dgamma, float* dbeta,
                                  long long rows, int C, int relu, cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && C > 0, VPTR_ERR_SHAPE, "vptr_layernorm_bwd: rows=%lld C=%d", rows, C);
    auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    if (C % 4 == 0 && C <= 1024 && al16(dy1) && al16(dy2) && al16(x) && al16(gamma) && al16(beta) && al16(dres) && al16(dx)) {
        // fused single pass (dx and the affine gradients from one read of dy / x)
        int blocks = 148 * 4;
        int rpb = (int)((rows + blocks - 1) / blocks);
        if (rpb < 8) rpb = 8;
        blocks = (int)((rows + rpb - 1) / rpb);
        const size_t smem = sizeof(float) * 2 * C;
        if (C <= 512) layernorm_bwd_fused_kernel<4><<<blocks, 256, smem, stream>>>(dy1, dy2, x, gamma, beta, mean, rstd, dres, dx, dgamma, dbeta, rows, C, relu, rpb);
        else if (C <= 640) layernorm_bwd_fused_kernel<5><<<blocks, 256, smem, stream>>>(dy1, dy2, x, gamma, beta, mean, rstd, dres, dx, dgamma, dbeta, rows, C, relu, rpb);
        else if (C <= 768) layernorm_bwd_fused_kernel<6><<<blocks, 256, smem, stream>>>(dy1, dy2, x, gamma, beta, mean, rstd, dres, dx, dgamma, dbeta, rows, C, relu, rpb);
        else layernorm_bwd_fused_kernel<8><<<blocks, 256, smem, stream>>>(dy1, dy2, x, gamma, beta, mean, rstd, dres, dx, dgamma, dbeta, rows, C, relu, rpb);
        return vptr_check_launch("layernorm_bwd_fused_kernel");
    }
    if (dx) {
        layernorm_bwd_dx_kernel<<<vptr_cdiv(rows, 8), 256, 0, stream>>>(dy1, dy2, x, gamma, beta, mean, rstd, dres, dx, rows, C, relu);
        int rc = vptr_check_launch("layernorm_bwd_dx_kernel");
        if (rc) return rc;
    }
    if (dgamma) {
        int rpb = 256;
        dim3 grid(vptr_cdiv(C, 128), vptr_cdiv(rows, rpb));
        layernorm_bwd_affine_kernel<<<grid, 128, 0, stream>>>(dy1, dy2, x, gamma, beta, mean, rstd, dgamma, dbeta, rows, C, rpb, relu);
        return vptr_check_launch("layernorm_bwd_affine_kernel");
    }
    return VPTR_OK;
}

// BatchNorm statistics over rows (training) -> mean/rstd [ch]; ws = 2*ch doubles of scratch.
extern "C" int vptr_bn_stats(const float* x, long long rows, int ch, float* mean, float* rstd, float* running_mean, float* running_var,
                             float eps, float momentum, double* ws, cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && ch > 0, VPTR_ERR_SHAPE, "vptr_bn_stats: rows=%lld ch=%d", rows, ch);
    cudaMemsetAsync(ws, 0, sizeof(double) * 2 * ch, stream);
    int rpb = 256;
    dim3 grid(vptr_cdiv(ch, 128), vptr_cdiv(rows, rpb));
    colstats_kernel<<<grid, 128, 0, stream>>>(x, ws, ws + ch, rows, ch, rpb);
    bn_finalize_kernel<<<vptr_cdiv(ch, 128), 128, 0, stream>>>(ws, ws + ch, mean, rstd, running_mean, running_var, rows, ch, eps, momentum);
    return vptr_check_launch("vptr_bn_stats");
}

extern "C" int vptr_bn_eval_stats(const float* running_mean, const float* running_var, float* mean, float* rstd, int ch, float eps,
                                  cudaStream_t stream) {
    bn_eval_stats_kernel<<<vptr_cdiv(ch, 128), 128, 0, stream>>>(running_mean, running_var, mean, rstd, ch, eps);
    return vptr_check_launch("bn_eval_stats_kernel");
}

extern "C" int vptr_group_stats(const float* x, int groups, long long gsize, float* mean, float* rstd, float eps, cudaStream_t stream) {
    VPTR_REQUIRE(groups > 0 && gsize > 0 && gsize % 4 == 0, VPTR_ERR_SHAPE, "vptr_group_stats: groups=%d gsize=%lld", groups, gsize);
    groupstats_kernel<<<groups, 512, 0, stream>>>(x, mean, rstd, gsize, eps);
    return vptr_check_launch("groupstats_kernel");
}

extern "C" int vptr_group_stats_finalize(const double* sums, int groups, long long gsize, float* mean, float* rstd, float eps,
                                         cudaStream_t stream) {
    VPTR_REQUIRE(sums != nullptr && groups > 0 && gsize > 0, VPTR_ERR_SHAPE, "vptr_group_stats_finalize: groups=%d gsize=%lld", groups, gsize);
    groupstats_finalize_kernel<<<vptr_cdiv(groups, 128), 128, 0, stream>>>(sums, groups, (double)gsize, mean, rstd, eps);
    return vptr_check_launch("groupstats_finalize_kernel");
}

// y = GELU(norm(x)) (+ res).  mode 0: BatchNorm (per-channel stats/affine); mode 1: LayerNorm((ch,H,W)) per frame with
// affine stored token-major [hw][ch].
extern "C" int vptr_norm_act_fwd(const float* x, float* y, const float* res, const float* mean, const float* rstd, const float* gamma,
                                 const float* beta, long long rows, int ch, int hw, int mode, int round_tf32, const float* rowscale,
                                 int rows_per_group, unsigned long long drop_seed, float drop_p, cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && ch > 0 && ch % 4 == 0, VPTR_ERR_SHAPE, "vptr_norm_act_fwd: rows=%lld ch=%d", rows, ch);
    const DropArgs da{rowscale, rows_per_group > 0 ? rows_per_group : 1, drop_seed, drop_p};
    VPTR_REQUIRE(hw > 0 && rows % hw == 0 && rows / hw < 65536, VPTR_ERR_SHAPE, "vptr_norm_act_fwd: rows=%lld not whole frames of hw=%d", rows, hw);
    VPTR_REQUIRE(rowscale == nullptr || rows_per_group % hw == 0, VPTR_ERR_SHAPE, "vptr_norm_act_fwd: DropPath groups must be whole frames");
    const int frames = (int)(rows / hw), frame4 = hw * (ch / 4);
    dim3 grid(norm_chunks(frame4, frames), frames);
    if (mode == 0) norm_act_fwd_kernel<0><<<grid, 256, 0, stream>>>(x, y, res, mean, rstd, gamma, beta, frame4, ch / 4, hw, round_tf32, da);
    else norm_act_fwd_kernel<1><<<grid, 256, 0, stream>>>(x, y, res, mean, rstd, gamma, beta, frame4, ch / 4, hw, round_tf32, da);
    return vptr_check_launch("norm_act_fwd_kernel");
}

// Backward of y = rowscale*dropout(GELU(norm(x))).  mode 0: train BatchNorm, 1: frame LayerNorm, 2: eval BatchNorm.
// ws: mode 0/2 -> 2*ch floats; mode 1 -> 2*frames floats.  dgamma/dbeta are accumulated (+=).  dx may alias dy (in place).
extern "C" int vptr_norm_act_bwd_colsum(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                        const float* beta, float* dx, float* dgamma, float* dbeta, long long rows, int ch, int hw, int mode,
                                        float* ws, int round_tf32, const float* rowscale, int rows_per_group, unsigned long long drop_seed,
                                        float drop_p, float* colsum, cudaStream_t stream);
extern "C" int vptr_norm_act_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                 const float* beta, float* dx, float* dgamma, float* dbeta, long long rows, int ch, int hw, int mode,
                                 float* ws, int round_tf32, const float* rowscale, int rows_per_group, unsigned long long drop_seed,
                                 float drop_p, cudaStream_t stream) {
    return vptr_norm_act_bwd_colsum(dy, x, mean, rstd, gamma, beta, dx, dgamma, dbeta, rows, ch, hw, mode, ws, round_tf32, rowscale, rows_per_group,
                                    drop_seed, drop_p, nullptr, stream);
}
// colsum != NULL: colsum[c] += sum over rows of dx[.][c] (the bias gradient of the 1x1 conv that produced x), out of the final pass
extern "C" int vptr_norm_act_bwd_colsum(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                        const float* beta, float* dx, float* dgamma, float* dbeta, long long rows, int ch, int hw, int mode,
                                        float* ws, int round_tf32, const float* rowscale, int rows_per_group, unsigned long long drop_seed,
                                        float drop_p, float* colsum, cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && ch > 0 && ch % 4 == 0, VPTR_ERR_SHAPE, "vptr_norm_act_bwd: rows=%lld ch=%d (ch %% 4 == 0 required)", rows, ch);
    const DropArgs da{rowscale, rows_per_group > 0 ? rows_per_group : 1, drop_seed, drop_p};
    VPTR_REQUIRE(hw > 0 && rows % hw == 0 && rows / hw < 65536, VPTR_ERR_SHAPE, "vptr_norm_act_bwd: rows=%lld not whole frames of hw=%d", rows, hw);
    const int frame4 = hw * (ch / 4);
    const dim3 gdx(norm_chunks(frame4, (int)(rows / hw)), (unsigned)(rows / hw));
    if (mode == 0 || mode == 2) {
        cudaMemsetAsync(ws, 0, sizeof(float) * 2 * ch, stream);
        int rpb = 128;
        dim3 g2(vptr_cdiv(ch / 4, 128), vptr_cdiv(rows, rpb));
        bn_act_bwd_pass_a_kernel<<<g2, 128, 0, stream>>>(dy, x, mean, rstd, gamma, beta, dx, dgamma, dbeta, ws, ws + ch, rows, ch, rpb, da);
        if (colsum) {
            const int C4 = ch / 4, tx = C4 < 1024 ? ((C4 + 31) & ~31) : 1024, xb = vptr_cdiv(C4, tx);
            int rpbc = (int)((rows + (148 * 4 + xb - 1) / xb - 1) / ((148 * 4 + xb - 1) / xb));
            if (rpbc < 16) rpbc = 16;
            const dim3 gc(xb, vptr_cdiv(rows, rpbc));
            if (mode == 0)
                norm_act_bwd_dx_colsum_kernel<0><<<gc, tx, 0, stream>>>(dx, x, mean, rstd, gamma, ws, ws + ch, rows, C4, hw, 1.0f / (float)rows, round_tf32, colsum, rpbc);
            else
                norm_act_bwd_dx_colsum_kernel<2><<<gc, tx, 0, stream>>>(dx, x, mean, rstd, gamma, ws, ws + ch, rows, C4, hw, 0.f, round_tf32, colsum, rpbc);
        } else if (mode == 0)
            norm_act_bwd_dx_kernel<0><<<gdx, 256, 0, stream>>>(dx, x, mean, rstd, gamma, ws, ws + ch, frame4, ch / 4, 1.0f / (float)rows, round_tf32);
        else
            norm_act_bwd_dx_kernel<2><<<gdx, 256, 0, stream>>>(dx, x, mean, rstd, gamma, ws, ws + ch, frame4, ch / 4, 0.f, round_tf32);
    } else {
        const long long gsize = (long long)hw * ch;
        const int frames = (int)(rows / hw);
        cudaMemsetAsync(ws, 0, sizeof(float) * 2 * frames, stream);
        dim3 g2(vptr_cdiv(gsize / 4, 256), vptr_cdiv(frames, LN3_FPB));
        ln3_act_bwd_pass_ab_kernel<<<g2, 256, 0, stream>>>(dy, x, mean, rstd, gamma, beta, dx, ws, ws + frames, dgamma, dbeta, gsize, ch, frames, da);
        if (colsum) {
            const int C4 = ch / 4, tx = C4 < 1024 ? ((C4 + 31) & ~31) : 1024, xb = vptr_cdiv(C4, tx);
            int rpbc = (int)((rows + (148 * 4 + xb - 1) / xb - 1) / ((148 * 4 + xb - 1) / xb));
            if (rpbc < 16) rpbc = 16;
            const dim3 gc(xb, vptr_cdiv(rows, rpbc));
            norm_act_bwd_dx_colsum_kernel<1><<<gc, tx, 0, stream>>>(dx, x, mean, rstd, gamma, ws, ws + frames, rows, C4, hw, 1.0f / (float)gsize, round_tf32, colsum, rpbc);
        } else
            norm_act_bwd_dx_kernel<1><<<gdx, 256, 0, stream>>>(dx, x, mean, rstd, gamma, ws, ws + frames, frame4, ch / 4, 1.0f / (float)gsize, round_tf32);
    }
    return vptr_check_launch("vptr_norm_act_bwd");
}

VPTR_RNG_EPOCH_ACCESSOR(norm)
