// Tail of the training iteration: the image-space losses of cal_lossT, the global-norm gradient clip and the
// AdamW update, as a handful of HBM-bound kernels instead of the reference's ~60 PyTorch launches.
//   reference: model/criterion.py:105-204 (MSELoss, GDL), train_NAR.py:33-47,83-86 (cal_lossT, clip_grad_norm_, optimizer_T.step()),
//              torch.optim.AdamW (decoupled weight decay, bias correction), torch.nn.utils.clip_grad_norm_.
// Every kernel is a single pass over its operands (8-28 B / element); nothing here is GEMM-shaped.
#include "common.cuh"

namespace {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------- MSE + GDL (alpha = 1)
// planes = N*T*C image planes of H x W.  sums[0] += sum (p-g)^2, sums[1] += sum | |g[i+1]-g[i]| - |p[i+1]-p[i]| | (vertical,
// (H-1) x W terms per plane), sums[2] += the horizontal one (H x (W-1) terms).  criterion.py:160-181.
__global__ void __launch_bounds__(256) mse_gdl_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long long planes,
                                                          int H, int W, double* __restrict__ sums) {
    const long long total = planes * H * W;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        const int h = (int)((i / W) % H);
        const float p = pred[i], g = gt[i];
        const float d = p - g;
        s0 = fmaf(d, d, s0);
        if (h + 1 < H) s1 += fabsf(fabsf(gt[i + W] - g) - fabsf(pred[i + W] - p));
        if (w + 1 < W) s2 += fabsf(fabsf(g - gt[i + 1]) - fabsf(p - pred[i + 1]));
    }
    __shared__ double red[3][8];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    double a0 = warp_sum_d((double)s0), a1 = warp_sum_d((double)s1), a2 = warp_sum_d((double)s2);
    if (lane == 0) { red[0][wp] = a0; red[1][wp] = a1; red[2][wp] = a2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += red[threadIdx.x][k];
        atomicAdd(sums + threadIdx.x, t);
    }
}
// loss = mean(se) + mean(gdl1) + mean(gdl2)
__global__ void mse_gdl_loss_kernel(const double* __restrict__ sums, long long planes, int H, int W, float* __restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double n0 = (double)planes * H * W, n1 = (double)planes * (H - 1) * W, n2 = (double)planes * H * (W - 1);
        loss[0] = (float)(sums[0] / n0 + (n1 > 0 ? sums[1] / n1 : 0.0) + (n2 > 0 ? sums[2] / n2 : 0.0));
        loss[1] = (float)(sums[0] / n0);
        loss[2] = (float)((n1 > 0 ? sums[1] / n1 : 0.0) + (n2 > 0 ? sums[2] / n2 : 0.0));
    }
}
__device__ __forceinline__ float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }   // torch's abs backward: sign(0) = 0
// d loss / d pred, gathered per pixel: each of the (up to) four finite differences a pixel takes part in contributes
// -+ sign(|dg| - |dp|) * sign(dp) / count.
__global__ void __launch_bounds__(256) mse_gdl_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                          const float* __restrict__ dloss, float* __restrict__ dpred, long long planes,
                                                          int H, int W) {
    const long long total = planes * H * W;
    const float up = dloss ? dloss[0] : 1.f;
    const float c0 = up * 2.f / (float)((double)planes * H * W);
    const float c1 = H > 1 ? up / (float)((double)planes * (H - 1) * W) : 0.f;
    const float c2 = W > 1 ? up / (float)((double)planes * H * (W - 1)) : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        const int h = (int)((i / W) % H);
        const float p = pred[i], g = gt[i];
        float acc = c0 * (p - g);
        if (h + 1 < H) {   // a = p[h+1] - p[h]; e = |dg| - |a|; d|e|/dp[h] = -sign(e) * sign(a) * (-1)
            const float a = pred[i + W] - p, e = fabsf(gt[i + W] - g) - fabsf(a);
            acc += c1 * sgn(e) * sgn(a);
        }
        if (h > 0) {
            const float a = p - pred[i - W], e = fabsf(g - gt[i - W]) - fabsf(a);
            acc -= c1 * sgn(e) * sgn(a);
        }
        if (w + 1 < W) {   // a = p[w] - p[w+1]
            const float a = p - pred[i + 1], e = fabsf(g - gt[i + 1]) - fabsf(a);
            acc -= c2 * sgn(e) * sgn(a);
        }
        if (w > 0) {
            const float a = pred[i - 1] - p, e = fabsf(gt[i - 1] - g) - fabsf(a);
            acc += c2 * sgn(e) * sgn(a);
        }
        dpred[i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------- multi-tensor helpers
// table (int64, device): [n] pointer columns followed by ends[n] = cumulative unit counts (unit = float4 when VEC else float).
__device__ __forceinline__ int find_tensor(const long long* __restrict__ ends, int n, long long u) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (u < ends[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

template <bool VEC>
__global__ void __launch_bounds__(256) sqnorm_multi_kernel(const long long* __restrict__ table, int n, long long total, double* __restrict__ out) {
    const long long* ends = table + n;
    float s = 0.f;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (long long)gridDim.x * blockDim.x) {
        const int t = find_tensor(ends, n, u);
        const long long off = u - (t ? ends[t - 1] : 0);
        if (VEC) {
            const float4 v = reinterpret_cast<const float4*>(table[t])[off];
            s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
        } else {
            const float v = reinterpret_cast<const float*>(table[t])[off];
            s = fmaf(v, v, s);
        }
    }
    __shared__ float red[32];
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(out, (double)s);
}

struct AdamWArgs {
    float lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2_sqrt, max_norm;
};
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWArgs& a, float clip) {
    g *= clip;
    p *= 1.f - a.lr * a.weight_decay;                       // decoupled weight decay (torch.optim.AdamW)
    m = fmaf(a.beta1, m, (1.f - a.beta1) * g);              // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(a.beta2, v, (1.f - a.beta2) * g * g);          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / a.bias_corr2_sqrt + a.eps;
    p -= (a.lr / a.bias_corr1) * (m / denom);
}
// table columns: param, grad, exp_avg, exp_avg_sq.  sqnorm (may be null): sum of squares of ALL gradients -> the
// clip_grad_norm_ coefficient min(1, max_norm / (norm + 1e-6)) is applied to the gradient on the fly (one pass less).
template <bool VEC>
__global__ void __launch_bounds__(256) adamw_multi_kernel(const long long* __restrict__ table, int n, long long total, AdamWArgs a,
                                                          const double* __restrict__ sqnorm, const long long* __restrict__ step_dev) {
    const long long* ends = table + 4 * n;
    if (step_dev) {      // step count lives on the device (captured graphs: the host-side value would be frozen into the launch)
        const double st = (double)*step_dev;
        a.bias_corr1 = (float)(1.0 - pow((double)a.beta1, st));
        a.bias_corr2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, st));
    }
    float clip = 1.f;
    if (sqnorm) clip = fminf(1.f, a.max_norm / ((float)sqrt(*sqnorm) + 1e-6f));
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (long long)gridDim.x * blockDim.x) {
        const int t = find_tensor(ends, n, u);
        const long long off = u - (t ? ends[t - 1] : 0);
        if (VEC) {
            float4* pp = reinterpret_cast<float4*>(table[t]) + off;
            const float4 g = reinterpret_cast<const float4*>(table[n + t])[off];
            float4* mp = reinterpret_cast<float4*>(table[2 * n + t]) + off;
            float4* vp = reinterpret_cast<float4*>(table[3 * n + t]) + off;
            float4 p = *pp, m = *mp, v = *vp;
            adamw_one(p.x, g.x, m.x, v.x, a, clip);
            adamw_one(p.y, g.y, m.y, v.y, a, clip);
            adamw_one(p.z, g.z, m.z, v.z, a, clip);
            adamw_one(p.w, g.w, m.w, v.w, a, clip);
            *pp = p; *mp = m; *vp = v;
        } else {
            float* pp = reinterpret_cast<float*>(table[t]) + off;
            const float g = reinterpret_cast<const float*>(table[n + t])[off];
            float* mp = reinterpret_cast<float*>(table[2 * n + t]) + off;
            float* vp = reinterpret_cast<float*>(table[3 * n + t]) + off;
            float p = *pp, m = *mp, v = *vp;
            adamw_one(p, g, m, v, a, clip);
            *pp = p; *mp = m; *vp = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------- BiPatchNCE (fused)
// Bidirectional patch-wise contrastive loss of cal_lossT (model/criterion.py:206-259, train_NAR.py:36 with F.normalize over the
// channel dim of both operands) for one frame per CTA: A = gt features [L][C], B = predicted features [L][C], L = h*w <= 64.
//   a^ = a / max(|a|, 1e-12), b^ likewise;  S^_ij = a^_i . b^_j;  score1 = S^ / tau (rows: gt -> pred), score2 = score1^T;
//   loss = 0.5 * (mean_i CE(score1_i, i) + mean_j CE(score2_j, j)) over all F*L rows.
// Stop-gradients of the reference (negatives detached on the "other" side): with G1 = softmax_row(score1) - I and
// H_ij = softmax_col(score1)_ij - I (both scaled by 0.5 / (F L tau)),
//   d a^_i = sum_j G1_ij b^_j + H_ii b^_i ,   d b^_j = sum_i H_ij a^_i + G1_jj a^_j .
// Forward keeps S^ (16 KB per frame) and the row / column statistics; the backward rebuilds the coefficient matrices from them, so
// the only big tensors it touches are A, B (read) and dA, dB (written): 4 passes over the features instead of the reference's
// normalize / rearrange / 4 bmm / mask / cross-entropy chain and its autograd mirror (~60 launches).
constexpr int NCE_L = 64, NCE_CK = 32, NCE_CKB = 24, NCE_SP = 65;   // (backward: 24-column chunks keep static smem < 48 KB)

__global__ void __launch_bounds__(256) bipatch_nce_fwd_kernel(const float* __restrict__ A, const float* __restrict__ B, int L, int C, float inv_tau,
                                                              float* __restrict__ S_out, float* __restrict__ stats, double* __restrict__ loss_sum) {
    __shared__ float sA[NCE_L][NCE_CK + 1], sB[NCE_L][NCE_CK + 1], sS[NCE_L][NCE_SP];
    __shared__ float sn[2][NCE_L], red[8];
    const int f = blockIdx.x, tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const float* Af = A + (long long)f * L * C;
    const float* Bf = B + (long long)f * L * C;
    float acc[4][4] = {};
    float nrm = 0.f;                                   // threads 0..63: |a_tid|^2 ; 64..127: |b_(tid-64)|^2
    for (int c0 = 0; c0 < C; c0 += NCE_CK) {
        const int ck = min(NCE_CK, C - c0);
        for (int e = tid; e < NCE_L * NCE_CK; e += 256) {
            const int r = e / NCE_CK, c = e - r * NCE_CK;
            const bool ok = r < L && c < ck;
            sA[r][c] = ok ? Af[(long long)r * C + c0 + c] : 0.f;
            sB[r][c] = ok ? Bf[(long long)r * C + c0 + c] : 0.f;
        }
        __syncthreads();
        if (tid < 64) { for (int c = 0; c < NCE_CK; ++c) nrm = fmaf(sA[tid][c], sA[tid][c], nrm); }
        else if (tid < 128) { for (int c = 0; c < NCE_CK; ++c) nrm = fmaf(sB[tid - 64][c], sB[tid - 64][c], nrm); }
#pragma unroll 4
        for (int c = 0; c < NCE_CK; ++c) {
            float a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { a[k] = sA[ty * 4 + k][c]; b[k] = sB[tx * 4 + k][c]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    if (tid < 128) sn[tid >> 6][tid & 63] = fmaxf(sqrtf(nrm), 1e-12f);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sS[ty * 4 + i][tx * 4 + j] = acc[i][j] / (sn[0][ty * 4 + i] * sn[1][tx * 4 + j]);
    __syncthreads();
    float part = 0.f, lse = 0.f;
    if (tid < 128) {
        const int r = tid & 63;
        const bool col = tid >= 64;
        if (r < L) {
            float mx = -INFINITY;
            for (int k = 0; k < L; ++k) mx = fmaxf(mx, (col ? sS[k][r] : sS[r][k]) * inv_tau);
            float se = 0.f;
            for (int k = 0; k < L; ++k) se += __expf((col ? sS[k][r] : sS[r][k]) * inv_tau - mx);
            lse = mx + __logf(se);
            part = lse - sS[r][r] * inv_tau;
        }
    }
    float* st = stats + (long long)f * 4 * NCE_L;      // na | nb | lse1 | lse2
    if (tid < 128) { st[tid] = sn[tid >> 6][tid & 63]; st[128 + tid] = lse; }
    for (int e = tid; e < NCE_L * NCE_L; e += 256) S_out[(long long)f * NCE_L * NCE_L + e] = sS[e >> 6][e & 63];
    part = block_sum(part, red);
    if (tid == 0) atomicAdd(loss_sum, (double)part);
}
__global__ void bipatch_nce_loss_kernel(const double* __restrict__ loss_sum, double rows, float* __restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) loss[0] = (float)(0.5 * loss_sum[0] / rows);
}
__global__ void __launch_bounds__(256) bipatch_nce_bwd_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ S_in,
                                                              const float* __restrict__ stats, const float* __restrict__ dloss, int L, int C,
                                                              float inv_tau, float coef, float* __restrict__ dA, float* __restrict__ dB) {
    __shared__ float sM[NCE_L][NCE_SP], sMt[NCE_L][NCE_SP];   // M_ij (for dA), M'_ij (for dB)
    __shared__ float sA[NCE_L][NCE_CKB + 1], sB[NCE_L][NCE_CKB + 1];
    __shared__ float sna[NCE_L], snb[NCE_L], sr[2][NCE_L];
    const int f = blockIdx.x, tid = threadIdx.x;
    const float* st = stats + (long long)f * 4 * NCE_L;
    const float up = (dloss ? dloss[0] : 1.f) * coef * inv_tau;        // coef = 0.5 / (F L)
    if (tid < 64) { sna[tid] = st[tid]; snb[tid] = st[64 + tid]; }
    for (int e = tid; e < NCE_L * NCE_L; e += 256) {
        const int i = e >> 6, j = e & 63;
        float m = 0.f, mt = 0.f;
        if (i < L && j < L) {
            const float sc = S_in[(long long)f * NCE_L * NCE_L + e] * inv_tau;
            const float g1 = __expf(sc - st[128 + i]) - (i == j ? 1.f : 0.f);        // softmax over the row i
            const float h = __expf(sc - st[192 + j]) - (i == j ? 1.f : 0.f);         // softmax over the column j
            m = g1 + (i == j ? h : 0.f);
            mt = h + (i == j ? g1 : 0.f);
        }
        sM[i][j] = m * up;
        sMt[i][j] = mt * up;
    }
    __syncthreads();
    if (tid < 128) {       // r_i = sum_j M_ij S^_ij ; r'_j = sum_i M'_ij S^_ij   (= x^ . dx^ of the normalisation's backward)
        const int r = tid & 63;
        const bool col = tid >= 64;
        float acc = 0.f;
        const float* Sf = S_in + (long long)f * NCE_L * NCE_L;
        if (r < L)
            for (int k = 0; k < L; ++k) acc = fmaf(col ? sMt[k][r] : sM[r][k], col ? Sf[k * NCE_L + r] : Sf[r * NCE_L + k], acc);
        sr[col ? 1 : 0][r] = acc;
    }
    __syncthreads();
    const float* Af = A + (long long)f * L * C;
    const float* Bf = B + (long long)f * L * C;
    float* dAf = dA + (long long)f * L * C;
    float* dBf = dB + (long long)f * L * C;
    constexpr int CPT = NCE_CKB / 4;
    const int r = tid >> 2, cg = (tid & 3) * CPT;
    for (int c0 = 0; c0 < C; c0 += NCE_CKB) {
        const int ck = min(NCE_CKB, C - c0);
        for (int e = tid; e < NCE_L * NCE_CKB; e += 256) {
            const int rr = e / NCE_CKB, c = e - rr * NCE_CKB;
            const bool ok = rr < L && c < ck;
            sA[rr][c] = ok ? Af[(long long)rr * C + c0 + c] / sna[rr] : 0.f;           // normalised rows
            sB[rr][c] = ok ? Bf[(long long)rr * C + c0 + c] / snb[rr] : 0.f;
        }
        __syncthreads();
        float oa[CPT] = {}, ob[CPT] = {};
        for (int k = 0; k < L; ++k) {
            const float m = sM[r][k], mt = sMt[k][r];
#pragma unroll
            for (int c = 0; c < CPT; ++c) { oa[c] = fmaf(m, sB[k][cg + c], oa[c]); ob[c] = fmaf(mt, sA[k][cg + c], ob[c]); }
        }
        if (r < L) {
            const float ra = sr[0][r], rb = sr[1][r], ia = 1.f / sna[r], ib = 1.f / snb[r];
#pragma unroll
            for (int c = 0; c < CPT; ++c)
                if (cg + c < ck) {
                    dAf[(long long)r * C + c0 + cg + c] = (oa[c] - sA[r][cg + c] * ra) * ia;
                    dBf[(long long)r * C + c0 + cg + c] = (ob[c] - sB[r][cg + c] * rb) * ib;
                }
        }
        __syncthreads();
    }
}

inline int grid_for(long long units) {
    long long b = (units + 255) / 256;
    const long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

// sums: 3 device doubles, zeroed by the caller.
extern "C" int vptr_mse_gdl_fwd(const float* pred, const float* target, long long planes, int H, int W, double* sums, float* loss3,
                                cudaStream_t stream) {
    VPTR_REQUIRE(planes > 0 && H > 0 && W > 0, VPTR_ERR_SHAPE, "vptr_mse_gdl_fwd: planes=%lld H=%d W=%d", planes, H, W);
    mse_gdl_fwd_kernel<<<grid_for(planes * H * W), 256, 0, stream>>>(pred, target, planes, H, W, sums);
    int rc = vptr_check_launch("mse_gdl_fwd_kernel");
    if (rc) return rc;
    mse_gdl_loss_kernel<<<1, 32, 0, stream>>>(sums, planes, H, W, loss3);
    return vptr_check_launch("mse_gdl_loss_kernel");
}
// dloss: device scalar (upstream gradient of the loss) or NULL (= 1).
extern "C" int vptr_mse_gdl_bwd(const float* pred, const float* target, const float* dloss, float* dpred, long long planes, int H, int W,
                                cudaStream_t stream) {
    VPTR_REQUIRE(planes > 0 && H > 0 && W > 0, VPTR_ERR_SHAPE, "vptr_mse_gdl_bwd: planes=%lld H=%d W=%d", planes, H, W);
    mse_gdl_bwd_kernel<<<grid_for(planes * H * W), 256, 0, stream>>>(pred, target, dloss, dpred, planes, H, W);
    return vptr_check_launch("mse_gdl_bwd_kernel");
}
// table: device int64 [n pointers][n cumulative unit ends]; vec != 0: units are float4 (every tensor 16-byte aligned, numel % 4 == 0).
// out (device double) += sum of squares; caller zeroes it.
extern "C" int vptr_sqnorm_multi(const long long* table, int n, long long total_units, int vec, double* out, cudaStream_t stream) {
    if (n <= 0 || total_units <= 0) return VPTR_OK;
    if (vec) sqnorm_multi_kernel<true><<<grid_for(total_units), 256, 0, stream>>>(table, n, total_units, out);
    else sqnorm_multi_kernel<false><<<grid_for(total_units), 256, 0, stream>>>(table, n, total_units, out);
    return vptr_check_launch("sqnorm_multi_kernel");
}
// One AdamW step over n tensors (table: device int64 [param][grad][exp_avg][exp_avg_sq][ends], n entries each).
// step >= 1 is the update count AFTER this step (bias corrections 1 - beta^step).  sqnorm != NULL: fused clip_grad_norm_(max_norm).
extern "C" int vptr_adamw_multi_dev(const long long* table, int n, long long total_units, int vec, float lr, float beta1, float beta2, float eps,
                                    float weight_decay, long long step, const long long* step_dev, const double* sqnorm, float max_norm,
                                    cudaStream_t stream);
extern "C" int vptr_adamw_multi(const long long* table, int n, long long total_units, int vec, float lr, float beta1, float beta2, float eps,
                                float weight_decay, long long step, const double* sqnorm, float max_norm, cudaStream_t stream) {
    return vptr_adamw_multi_dev(table, n, total_units, vec, lr, beta1, beta2, eps, weight_decay, step, nullptr, sqnorm, max_norm, stream);
}
// step_dev != NULL: the update count (AFTER this step) is read from device memory instead of `step`
extern "C" int vptr_adamw_multi_dev(const long long* table, int n, long long total_units, int vec, float lr, float beta1, float beta2, float eps,
                                    float weight_decay, long long step, const long long* step_dev, const double* sqnorm, float max_norm,
                                    cudaStream_t stream) {
    if (n <= 0 || total_units <= 0) return VPTR_OK;
    VPTR_REQUIRE(step >= 1 || step_dev != nullptr, VPTR_ERR_SHAPE, "vptr_adamw_multi: step=%lld must be >= 1", step);
    if (step < 1) step = 1;
    AdamWArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm;
    a.bias_corr1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bias_corr2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    if (vec) adamw_multi_kernel<true><<<grid_for(total_units), 256, 0, stream>>>(table, n, total_units, a, sqnorm, step_dev);
    else adamw_multi_kernel<false><<<grid_for(total_units), 256, 0, stream>>>(table, n, total_units, a, sqnorm, step_dev);
    return vptr_check_launch("adamw_multi_kernel");
}

// gt / pred: [F][L][C] channel-last feature rows of F = N*T frames (L = h*w <= 64).  S_save: F*64*64 floats, stats_save: F*256 floats
// (kept for the backward); loss_sum: device double zeroed by the caller; loss_out: device float = the BiPatchNCE value.
extern "C" int vptr_bipatch_nce_fwd(const float* gt, const float* pred, int F, int L, int C, float temperature, float* S_save, float* stats_save,
                                    double* loss_sum, float* loss_out, cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && L > 0 && L <= NCE_L && C > 0 && temperature > 0.f, VPTR_ERR_UNSUPPORTED,
                 "vptr_bipatch_nce_fwd: F=%d L=%d (<= 64 supported) C=%d temperature=%g", F, L, C, (double)temperature);
    bipatch_nce_fwd_kernel<<<F, 256, 0, stream>>>(gt, pred, L, C, 1.f / temperature, S_save, stats_save, loss_sum);
    int rc = vptr_check_launch("bipatch_nce_fwd_kernel");
    if (rc) return rc;
    bipatch_nce_loss_kernel<<<1, 32, 0, stream>>>(loss_sum, (double)F * L, loss_out);
    return vptr_check_launch("bipatch_nce_loss_kernel");
}
// d(gt), d(pred) = dloss * d(BiPatchNCE)/d(.) through the reference's stop-gradients and the channel normalisation
extern "C" int vptr_bipatch_nce_bwd(const float* gt, const float* pred, const float* S_save, const float* stats_save, const float* dloss, int F,
                                    int L, int C, float temperature, float* dgt, float* dpred, cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && L > 0 && L <= NCE_L && C > 0 && temperature > 0.f, VPTR_ERR_UNSUPPORTED,
                 "vptr_bipatch_nce_bwd: F=%d L=%d (<= 64 supported) C=%d", F, L, C);
    bipatch_nce_bwd_kernel<<<F, 256, 0, stream>>>(gt, pred, S_save, stats_save, dloss, L, C, 1.f / temperature, (float)(0.5 / ((double)F * L)),
                                                  dgt, dpred);
    return vptr_check_launch("bipatch_nce_bwd_kernel");
}
