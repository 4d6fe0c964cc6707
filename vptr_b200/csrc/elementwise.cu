// Small HBM-bound helpers: residual adds, GELU/ReLU passes, bias-gradient column sums, layout transposes,
// centre zero-pad / crop for non-divisible window grids (PadBlock, VidHRFormer_modules.py:527-561),
// squared-norm + scale for gradient clipping.
#include "common.cuh"

namespace {

int ew_grid(long long n, int block) {
    long long g = (n + block - 1) / block;
    long long cap = 148LL * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                                  long long n4, float alpha, float beta) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 x = reinterpret_cast<const float4*>(a)[i];
        float4 y = reinterpret_cast<const float4*>(b)[i];
        float4 o = make_float4(alpha * x.x + beta * y.x, alpha * x.y + beta * y.y, alpha * x.z + beta * y.z, alpha * x.w + beta * y.w);
        reinterpret_cast<float4*>(out)[i] = o;
    }
}

// out[row] = x[row] + add[(row / div) % mod]
__global__ void __launch_bounds__(256) add_rows_kernel(const float* __restrict__ x, const float* __restrict__ add, float* __restrict__ out,
                                                       long long total4, int C4, int div, int mod, int round_tf32) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / C4;
        const int c = (int)(i - row * C4);
        float4 v = reinterpret_cast<const float4*>(x)[i];
        float4 p = __ldg(reinterpret_cast<const float4*>(add) + (long long)((row / div) % mod) * C4 + c);
        float4 o = make_float4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
        if (round_tf32) { o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w); }
        reinterpret_cast<float4*>(out)[i] = o;
    }
}

// y = [round_tf32]( x * rowscale[i / group_elems] * dropmask(seed, i) )
__global__ void __launch_bounds__(256) round_copy_kernel(const float* __restrict__ x, float* __restrict__ y, long long n4, int do_round,
                                                         const float* __restrict__ rowscale, long long group_elems,
                                                         unsigned long long seed, float p) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        if (p > 0.f) {
            const float4 k = vptr_drop_scale4(seed, (unsigned long long)i, p);
            v.x *= k.x; v.y *= k.y; v.z *= k.z; v.w *= k.w;
        }
        if (rowscale) {
            const float rs = __ldg(rowscale + (i * 4) / group_elems);
            v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
        }
        if (do_round) { v.x = vptr_round_tf32(v.x); v.y = vptr_round_tf32(v.y); v.z = vptr_round_tf32(v.z); v.w = vptr_round_tf32(v.w); }
        reinterpret_cast<float4*>(y)[i] = v;
    }
}

// Multi-tensor rounded copy: n source tensors (device table: src pointer, then the END offset of each tensor in the flat
// destination, offsets in float4 units) -> one flat buffer.  One launch replaces the per-weight round_copy launches of a
// forward / backward pass (372 launches, 4 ms per cfg1 step, all launch-latency).
__global__ void __launch_bounds__(256) round_copy_multi_kernel(const long long* __restrict__ table, int n, float* __restrict__ dst,
                                                               long long total4) {
    const long long* ends = table + n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = n - 1;                       // first tensor whose end offset is > i
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(ends + mid) > i) hi = mid; else lo = mid + 1;
        }
        const long long begin = lo ? __ldg(ends + lo - 1) : 0;
        const float4* src = reinterpret_cast<const float4*>(__ldg(table + lo));
        float4 v = src[i - begin];
        v.x = vptr_round_tf32(v.x); v.y = vptr_round_tf32(v.y); v.z = vptr_round_tf32(v.z); v.w = vptr_round_tf32(v.w);
        reinterpret_cast<float4*>(dst)[i] = v;
    }
}

// out[r][0:K] = rna_tf32(w[r][:]) ; out[r][K:2K] = rna_tf32(w[r][:] - hi)   ("2xTF32" split of a weight matrix, rows of K)
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ w, float* __restrict__ out, long long rows, long long K) {
    const long long total = rows * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / K, c = i - r * K;
        const float v = w[i];
        const float hi = vptr_round_tf32(v);
        out[r * 2 * K + c] = hi;
        out[r * 2 * K + K + c] = vptr_round_tf32(v - hi);
    }
}

// DropPath keep-scales per sample: 0 with probability p else 1/(1-p) (drop_path, VidHRFormer_modules.py:563-575)
__global__ void droppath_scales_kernel(float* __restrict__ out, int n, unsigned long long seed, float p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = vptr_drop_scale(seed, (unsigned long long)i, p);
}

// out[g][c] += sum_n dy[(n*mod + g)][c]   (gradient of a per-(t,h,w) learned query broadcast over clips)
__global__ void __launch_bounds__(256) rowgroup_sum_kernel(const float* __restrict__ dy, float* __restrict__ out, long long group_elems,
                                                           int reps) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < group_elems; i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int n = 0; n < reps; ++n) s += dy[(long long)n * group_elems + i];
        out[i] += s;
    }
}

__global__ void __launch_bounds__(256) gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n4, int round_tf32,
                                                       unsigned long long seed, float p) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        float4 o = make_float4(vptr_gelu(v.x), vptr_gelu(v.y), vptr_gelu(v.z), vptr_gelu(v.w));
        if (p > 0.f) {
            const float4 k = vptr_drop_scale4(seed, (unsigned long long)i, p);
            o.x *= k.x; o.y *= k.y; o.z *= k.z; o.w *= k.w;
        }
        if (round_tf32) { o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w); }
        reinterpret_cast<float4*>(y)[i] = o;
    }
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx,
                                                       long long n4, int round_tf32, unsigned long long seed, float p) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 g = reinterpret_cast<const float4*>(dy)[i];
        float4 v = reinterpret_cast<const float4*>(x)[i];
        float4 o = make_float4(g.x * vptr_gelu_grad(v.x), g.y * vptr_gelu_grad(v.y), g.z * vptr_gelu_grad(v.z), g.w * vptr_gelu_grad(v.w));
        if (p > 0.f) {
            const float4 k = vptr_drop_scale4(seed, (unsigned long long)i, p);
            o.x *= k.x; o.y *= k.y; o.z *= k.z; o.w *= k.w;
        }
        if (round_tf32) { o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w); }
        reinterpret_cast<float4*>(dx)[i] = o;
    }
}
// dx = dy * (y > 0)
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                                       long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 g = reinterpret_cast<const float4*>(dy)[i];
        float4 v = reinterpret_cast<const float4*>(y)[i];
        reinterpret_cast<float4*>(dx)[i] = make_float4(v.x > 0.f ? g.x : 0.f, v.y > 0.f ? g.y : 0.f, v.z > 0.f ? g.z : 0.f, v.w > 0.f ? g.w : 0.f);
    }
}
__global__ void __launch_bounds__(256) relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        reinterpret_cast<float4*>(y)[i] = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    }
}

// out[c] += sum_r x[r][c]
// ---- column-slab variants: a thread owns ONE float4 column and walks a block of rows, so the column sums of what it writes (the
// bias gradient of the Linear / 1x1 conv that consumes the tensor: db = sum over tokens of dY) accumulate in registers and cost four
// atomics per thread instead of a separate pass over dY (the stand-alone colsum kernel re-read every dY of the backward: 164
// launches, 6 ms per cfg1 step).  A block reads whole rows (blockDim.x * 16 B contiguous per row), rows_per_block consecutive rows.
__global__ void __launch_bounds__(1024) round_copy_colsum_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int C4,
                                                                 int do_round, const float* __restrict__ rowscale, int group_rows,
                                                                 unsigned long long seed, float p, float* __restrict__ colsum, int rows_per_block) {
    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (c4 >= C4) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, rows);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) {
        const long long i = r * C4 + c4;
        float4 v = reinterpret_cast<const float4*>(x)[i];
        if (p > 0.f) {
            const float4 k = vptr_drop_scale4(seed, (unsigned long long)i, p);
            v.x *= k.x; v.y *= k.y; v.z *= k.z; v.w *= k.w;
        }
        if (rowscale) {
            const float rs = __ldg(rowscale + r / group_rows);
            v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
        }
        if (do_round) { v.x = vptr_round_tf32(v.x); v.y = vptr_round_tf32(v.y); v.z = vptr_round_tf32(v.z); v.w = vptr_round_tf32(v.w); }
        reinterpret_cast<float4*>(y)[i] = v;
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    atomicAdd(colsum + 4 * c4, acc.x); atomicAdd(colsum + 4 * c4 + 1, acc.y); atomicAdd(colsum + 4 * c4 + 2, acc.z); atomicAdd(colsum + 4 * c4 + 3, acc.w);
}
__global__ void __launch_bounds__(1024) gelu_bwd_colsum_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx,
                                                               long long rows, int C4, int round_tf32, unsigned long long seed, float p,
                                                               float* __restrict__ colsum, int rows_per_block) {
    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (c4 >= C4) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, rows);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) {
        const long long i = r * C4 + c4;
        const float4 g = reinterpret_cast<const float4*>(dy)[i];
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        float4 o = make_float4(g.x * vptr_gelu_grad(v.x), g.y * vptr_gelu_grad(v.y), g.z * vptr_gelu_grad(v.z), g.w * vptr_gelu_grad(v.w));
        if (p > 0.f) {
            const float4 k = vptr_drop_scale4(seed, (unsigned long long)i, p);
            o.x *= k.x; o.y *= k.y; o.z *= k.z; o.w *= k.w;
        }
        if (round_tf32) { o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w); }
        reinterpret_cast<float4*>(dx)[i] = o;
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    atomicAdd(colsum + 4 * c4, acc.x); atomicAdd(colsum + 4 * c4 + 1, acc.y); atomicAdd(colsum + 4 * c4 + 2, acc.z); atomicAdd(colsum + 4 * c4 + 3, acc.w);
}

// (Measured against two float4 forms -- eight loads in flight per thread with 64-row blocks, and ~2 blocks per SM with a
// shared-memory reduction over row slices: 24.4 us for 40960 x 528 here against 29.4 us and worse under ncu; 800 small blocks of
// four warps stream better than few fat ones, and 86.5 MB cannot be read in much under 20 us once launch ramp and tail are paid.)
__global__ void __launch_bounds__(128) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int C,
                                                     long long ld, int rows_per_block) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, rows);
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += x[r * ld + c];
    atomicAdd(out + c, s);
}

// batched 2-D transpose: in [B][R][C] -> out [B][C][R]; accumulate != 0 -> out += in^T
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C, int accumulate) {
    __shared__ float tile[32][33];
    const long long boff = (long long)blockIdx.z * R * C;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        int r = r0 + j, c = c0 + tx;
        tile[j][tx] = (r < R && c < C) ? in[boff + (long long)r * C + c] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int c = c0 + j, r = r0 + tx;
        if (r < R && c < C) {
            float* o = out + boff + (long long)c * R + r;
            if (accumulate) *o += tile[tx][j]; else *o = tile[tx][j];
        }
    }
}

// Several 2-D transposes in ONE launch (the frame-LayerNorm affine weights (ch,H,W) <-> the engine's [hw][ch] layout and the
// depthwise 3x3 weights, forward; their gradients back, accumulate mode): 232 separate 8-10 us launches per cfg1 step were
// 2.2 ms of launch-latency-sized kernels.  table: n entries of 5 int64 {src, dst, R, C, cumulative 32x32-tile count}.
__global__ void __launch_bounds__(256) transpose_multi_kernel(const long long* __restrict__ table, int n, int accumulate) {
    __shared__ float tile[32][33];
    int e = 0;
    while (e + 1 < n && (long long)blockIdx.x >= table[e * 5 + 4]) ++e;
    const float* in = reinterpret_cast<const float*>(table[e * 5]);
    float* out = reinterpret_cast<float*>(table[e * 5 + 1]);
    const int R = (int)table[e * 5 + 2], C = (int)table[e * 5 + 3];
    const int local = (int)(blockIdx.x - (e ? table[(e - 1) * 5 + 4] : 0));
    const int ctiles = (C + 31) / 32;
    const int c0 = (local % ctiles) * 32, r0 = (local / ctiles) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + tx;
        tile[j][tx] = (r < R && c < C) ? in[(long long)r * C + c] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, r = r0 + tx;
        if (r < R && c < C) {
            float* o = out + (long long)c * R + r;
            if (accumulate) *o += tile[tx][j]; else *o = tile[tx][j];
        }
    }
}

// centre zero-pad [F][H][W][C] -> [F][Hp][Wp][C] (dir 0) or crop back (dir 1). They are each other's adjoint.
__global__ void __launch_bounds__(256) pad_crop_kernel(const float* __restrict__ in, float* __restrict__ out, int F, int H, int W, int Hp,
                                                       int Wp, int ph0, int pw0, int C4, int dir) {
    const long long total = dir == 0 ? (long long)F * Hp * Wp * C4 : (long long)F * H * W * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        long long t = i / C4;
        if (dir == 0) {
            const int w = (int)(t % Wp); t /= Wp;
            const int h = (int)(t % Hp); const long long f = t / Hp;
            const int hs = h - ph0, wsrc = w - pw0;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hs >= 0 && hs < H && wsrc >= 0 && wsrc < W) v = reinterpret_cast<const float4*>(in)[((f * H + hs) * W + wsrc) * C4 + c];
            reinterpret_cast<float4*>(out)[i] = v;
        } else {
            const int w = (int)(t % W); t /= W;
            const int h = (int)(t % H); const long long f = t / H;
            reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(in)[((f * Hp + h + ph0) * Wp + w + pw0) * C4 + c];
        }
    }
}

__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ x, long long n, double* __restrict__ out) {
    __shared__ float red[32];
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s = fmaf(x[i], x[i], s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(out, (double)s);
}
// x *= min(1, max_norm / (sqrt(sqnorm) + 1e-6))   (torch.nn.utils.clip_grad_norm_ semantics)
__global__ void __launch_bounds__(256) clip_scale_kernel(float* __restrict__ x, long long n, const double* __restrict__ sqnorm, float max_norm) {
    const float coef = fminf(1.f, max_norm / ((float)sqrt(*sqnorm) + 1e-6f));
    if (coef >= 1.f) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= coef;
}

}  // namespace

#define REQ4(n, name) VPTR_REQUIRE((n) >= 0 && (n) % 4 == 0, VPTR_ERR_SHAPE, name ": element count %lld must be a multiple of 4", (long long)(n))

extern "C" int vptr_axpby(const float* a, const float* b, float* out, long long n, float alpha, float beta, cudaStream_t stream) {
    REQ4(n, "vptr_axpby");
    if (n == 0) return VPTR_OK;
    add_kernel<<<ew_grid(n / 4, 256), 256, 0, stream>>>(a, b, out, n / 4, alpha, beta);
    return vptr_check_launch("add_kernel");
}
extern "C" int vptr_add_rows(const float* x, const float* add, float* out, long long rows, int C, int div, int mod, int round_tf32,
                             cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && div > 0 && mod > 0, VPTR_ERR_SHAPE, "vptr_add_rows: rows=%lld C=%d div=%d mod=%d", rows, C, div, mod);
    add_rows_kernel<<<ew_grid(rows * C / 4, 256), 256, 0, stream>>>(x, add, out, rows * C / 4, C / 4, div, mod, round_tf32);
    return vptr_check_launch("add_rows_kernel");
}
extern "C" int vptr_rowgroup_sum(const float* dy, float* out, long long group_elems, int reps, cudaStream_t stream) {
    VPTR_REQUIRE(group_elems > 0 && reps > 0, VPTR_ERR_SHAPE, "vptr_rowgroup_sum: group_elems=%lld reps=%d", group_elems, reps);
    rowgroup_sum_kernel<<<ew_grid(group_elems, 256), 256, 0, stream>>>(dy, out, group_elems, reps);
    return vptr_check_launch("rowgroup_sum_kernel");
}
extern "C" int vptr_gelu_fwd(const float* x, float* y, long long n, int round_tf32, unsigned long long drop_seed, float drop_p,
                             cudaStream_t stream) {
    REQ4(n, "vptr_gelu_fwd");
    gelu_fwd_kernel<<<ew_grid(n / 4, 256), 256, 0, stream>>>(x, y, n / 4, round_tf32, drop_seed, drop_p);
    return vptr_check_launch("gelu_fwd_kernel");
}
extern "C" int vptr_gelu_bwd(const float* dy, const float* x, float* dx, long long n, int round_tf32, unsigned long long drop_seed,
                             float drop_p, cudaStream_t stream) {
    REQ4(n, "vptr_gelu_bwd");
    gelu_bwd_kernel<<<ew_grid(n / 4, 256), 256, 0, stream>>>(dy, x, dx, n / 4, round_tf32, drop_seed, drop_p);
    return vptr_check_launch("gelu_bwd_kernel");
}
// y = round-to-nearest tf32 of x (element count need not be a multiple of 4: the tail is handled by padding rules of the caller)
extern "C" int vptr_round_copy(const float* x, float* y, long long n, int do_round, const float* rowscale, long long group_elems,
                               unsigned long long drop_seed, float drop_p, cudaStream_t stream) {
    REQ4(n, "vptr_round_copy");
    if (n == 0) return VPTR_OK;
    VPTR_REQUIRE(rowscale == nullptr || (group_elems > 0 && group_elems % 4 == 0), VPTR_ERR_SHAPE, "vptr_round_copy: group_elems=%lld", group_elems);
    round_copy_kernel<<<ew_grid(n / 4, 256), 256, 0, stream>>>(x, y, n / 4, do_round, rowscale, group_elems, drop_seed, drop_p);
    return vptr_check_launch("round_copy_kernel");
}
// table: device array of 2n int64 -- n source pointers (16-byte aligned, element counts multiples of 4) followed by the n
// cumulative end offsets (in float4) of the tensors inside dst; total4 = last end offset.
extern "C" int vptr_round_copy_multi(const long long* table, int n, float* dst, long long total4, cudaStream_t stream) {
    VPTR_REQUIRE(table != nullptr && n > 0 && dst != nullptr && total4 > 0, VPTR_ERR_SHAPE, "vptr_round_copy_multi: n=%d total4=%lld", n, total4);
    round_copy_multi_kernel<<<ew_grid(total4, 256), 256, 0, stream>>>(table, n, dst, total4);
    return vptr_check_launch("round_copy_multi_kernel");
}
extern "C" int vptr_split_tf32(const float* w, float* out, long long rows, long long K, cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && K > 0, VPTR_ERR_SHAPE, "vptr_split_tf32: rows=%lld K=%lld", rows, K);
    split_tf32_kernel<<<ew_grid(rows * K, 256), 256, 0, stream>>>(w, out, rows, K);
    return vptr_check_launch("split_tf32_kernel");
}
extern "C" int vptr_droppath_scales(float* out, int n, unsigned long long seed, float p, cudaStream_t stream) {
    VPTR_REQUIRE(n > 0 && p >= 0.f && p < 1.f, VPTR_ERR_SHAPE, "vptr_droppath_scales: n=%d p=%g", n, p);
    droppath_scales_kernel<<<vptr_cdiv(n, 128), 128, 0, stream>>>(out, n, seed, p);
    return vptr_check_launch("droppath_scales_kernel");
}
extern "C" int vptr_relu_fwd(const float* x, float* y, long long n, cudaStream_t stream) {
    REQ4(n, "vptr_relu_fwd");
    relu_fwd_kernel<<<ew_grid(n / 4, 256), 256, 0, stream>>>(x, y, n / 4);
    return vptr_check_launch("relu_fwd_kernel");
}
extern "C" int vptr_relu_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t stream) {
    REQ4(n, "vptr_relu_bwd");
    relu_bwd_kernel<<<ew_grid(n / 4, 256), 256, 0, stream>>>(dy, y, dx, n / 4);
    return vptr_check_launch("relu_bwd_kernel");
}
extern "C" int vptr_colsum(const float* x, float* out, long long rows, int C, long long ld, cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && C > 0, VPTR_ERR_SHAPE, "vptr_colsum: rows=%lld C=%d", rows, C);
    int rpb = 256;
    dim3 grid(vptr_cdiv(C, 128), vptr_cdiv(rows, rpb));
    colsum_kernel<<<grid, 128, 0, stream>>>(x, out, rows, C, ld, rpb);
    return vptr_check_launch("colsum_kernel");
}
extern "C" int vptr_transpose(const float* in, float* out, int batch, int R, int C, int accumulate, cudaStream_t stream) {
    VPTR_REQUIRE(batch > 0 && R > 0 && C > 0 && batch < 65536, VPTR_ERR_SHAPE, "vptr_transpose: batch=%d R=%d C=%d", batch, R, C);
    dim3 grid(vptr_cdiv(C, 32), vptr_cdiv(R, 32), batch);
    transpose_kernel<<<grid, 256, 0, stream>>>(in, out, R, C, accumulate);
    return vptr_check_launch("transpose_kernel");
}
extern "C" int vptr_transpose_multi(const long long* table, int n, int total_tiles, int accumulate, cudaStream_t stream) {
    VPTR_REQUIRE(table != nullptr && n > 0 && total_tiles > 0, VPTR_ERR_SHAPE, "vptr_transpose_multi: n=%d tiles=%d", n, total_tiles);
    transpose_multi_kernel<<<total_tiles, 256, 0, stream>>>(table, n, accumulate);
    return vptr_check_launch("transpose_multi_kernel");
}
extern "C" int vptr_pad_crop(const float* in, float* out, int F, int H, int W, int Hp, int Wp, int ph0, int pw0, int C, int dir,
                             cudaStream_t stream) {
    VPTR_REQUIRE(C % 4 == 0 && Hp >= H && Wp >= W && ph0 >= 0 && pw0 >= 0 && ph0 + H <= Hp && pw0 + W <= Wp, VPTR_ERR_SHAPE,
                 "vptr_pad_crop: bad geometry H=%d W=%d Hp=%d Wp=%d ph0=%d pw0=%d C=%d", H, W, Hp, Wp, ph0, pw0, C);
    long long total = (dir == 0 ? (long long)F * Hp * Wp : (long long)F * H * W) * (C / 4);
    pad_crop_kernel<<<ew_grid(total, 256), 256, 0, stream>>>(in, out, F, H, W, Hp, Wp, ph0, pw0, C / 4, dir);
    return vptr_check_launch("pad_crop_kernel");
}
// sqnorm_out (device double) += sum x^2 ; caller zeroes it.
extern "C" int vptr_sqnorm_accumulate(const float* x, long long n, double* sqnorm_out, cudaStream_t stream) {
    if (n <= 0) return VPTR_OK;
    sqnorm_kernel<<<ew_grid(n, 256), 256, 0, stream>>>(x, n, sqnorm_out);
    return vptr_check_launch("sqnorm_kernel");
}
extern "C" int vptr_clip_scale(float* x, long long n, const double* sqnorm, float max_norm, cudaStream_t stream) {
    if (n <= 0) return VPTR_OK;
    clip_scale_kernel<<<ew_grid(n, 256), 256, 0, stream>>>(x, n, sqnorm, max_norm);
    return vptr_check_launch("clip_scale_kernel");
}

// launch geometry of the column-slab kernels: blockDim.x = float4 columns (<= 1024, 32-aligned), gridDim.y row blocks sized so that
// ~4 blocks per SM exist and a thread keeps >= 16 rows (its four atomics amortised)
static void slab_geometry(long long rows, int C4, int& tx, dim3& grid, int& rpb) {
    tx = C4 < 1024 ? ((C4 + 31) & ~31) : 1024;
    const int xb = vptr_cdiv(C4, tx);
    long long want = (148LL * 4 + xb - 1) / xb;
    rpb = (int)((rows + want - 1) / want);
    if (rpb < 16) rpb = 16;
    grid = dim3(xb, vptr_cdiv(rows, rpb));
}
// y = [round](x * dropmask * rowscale[row / group_rows]) over a [rows][C] matrix, and colsum[c] += sum_r y[r][c]
extern "C" int vptr_round_copy_colsum(const float* x, float* y, long long rows, int C, int do_round, const float* rowscale, int group_rows,
                                      unsigned long long drop_seed, float drop_p, float* colsum, cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && colsum != nullptr, VPTR_ERR_SHAPE, "vptr_round_copy_colsum: rows=%lld C=%d", rows, C);
    VPTR_REQUIRE(rowscale == nullptr || group_rows > 0, VPTR_ERR_SHAPE, "vptr_round_copy_colsum: group_rows=%d", group_rows);
    int tx, rpb; dim3 grid;
    slab_geometry(rows, C / 4, tx, grid, rpb);
    round_copy_colsum_kernel<<<grid, tx, 0, stream>>>(x, y, rows, C / 4, do_round, rowscale, group_rows > 0 ? group_rows : 1, drop_seed, drop_p, colsum, rpb);
    return vptr_check_launch("round_copy_colsum_kernel");
}
// dx = [round](dy * GELU'(x) * dropmask) over a [rows][C] matrix (dx may alias dy), and colsum[c] += sum_r dx[r][c]
extern "C" int vptr_gelu_bwd_colsum(const float* dy, const float* x, float* dx, long long rows, int C, int round_tf32, unsigned long long drop_seed,
                                    float drop_p, float* colsum, cudaStream_t stream) {
    VPTR_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && colsum != nullptr, VPTR_ERR_SHAPE, "vptr_gelu_bwd_colsum: rows=%lld C=%d", rows, C);
    int tx, rpb; dim3 grid;
    slab_geometry(rows, C / 4, tx, grid, rpb);
    gelu_bwd_colsum_kernel<<<grid, tx, 0, stream>>>(dy, x, dx, rows, C / 4, round_tf32, drop_seed, drop_p, colsum, rpb);
    return vptr_check_launch("gelu_bwd_colsum_kernel");
}

VPTR_RNG_EPOCH_ACCESSOR(elementwise)
