// TF32 tensor-core GEMM for sm_100a: tcgen05.mma (kind::tf32) with TMEM accumulators, operands
// staged by TMA (128B swizzle) through a multi-stage mbarrier ring, persistent CTAs (one per SM),
// double-buffered accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
//
//   D[M,N] (+)= act(alpha * sum_k Aop[m,k] * Bop[n,k] + bias[n]) + residual[m,n]
//
// Operand storage ("major"):
//   a_mn = 0 : A stored [M][K] row-major (K contiguous)  -> forward / dgrad left operand
//   a_mn = 1 : A stored [K][M] row-major (M contiguous)  -> wgrad  (dY^T, contraction over tokens)
//   b_mn = 0 : B stored [N][K] row-major (K contiguous)  -> forward (nn.Linear weight [out][in])
//   b_mn = 1 : B stored [K][N] row-major (N contiguous)  -> dgrad (weight read as is) / wgrad (X)
//
// This one kernel stands in for every cuBLAS sgemm/addmm the reference issues on the hot path
// (q/k/v/out projections MultiHeadAttentionRPE.py:543-545,688; nn.MultiheadAttention in/out
// projections; 1x1 convs VidHRFormer_modules.py:424-442; linear1/linear2 :87-89) and, fed with
// an im2col / padded-NHWC view, for the ResNet 3x3 convolutions (ResNetAutoEncoder.py:26-47).
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;   // fp32 elements per k-chunk == one 128-byte swizzle row
constexpr int UMMA_K = 8;     // kind::tf32: 32 bytes of K per instruction
constexpr int NUM_THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue

struct GemmParams {
    int M, N, K;
    // Work item `tile` = (ks, mn) with mn fastest: the (m, n) tiles of one K-slice run concurrently on neighbouring CTAs (pairs), so an
    // operand slab shared by several of them is fetched from HBM once and re-read from L2.  With ks fastest (round 1) the second
    // round of the persistent grid re-read every B slab from HBM: ncu showed 445 MB per weight-gradient launch against 269 MB
    // algorithmic, at 4.7 TB/s -- those GEMMs were HBM-bound by their own re-reads (profiles/r02_launches_step_cfg1_summary.txt).
    int m_tiles, n_tiles, k_splits, chunks_per_split, total_chunks;
    float* D;
    long long ldd;
    const float* bias;
    const float* residual;
    long long ldr;
    float alpha;
    int act;    // 0 none, 1 gelu, 2 relu
    int flags;  // bit0: atomic accumulate into D, bit1: round stored values to tf32 (rna)
    // branch regularisation fused into the epilogue: out = rowscale[m / rows_per_group] * dropout(act(...)) + residual
    const float* rowscale;  // DropPath keep-scales per clip (VidHRFormer_modules.py:563-575) or NULL
    int rows_per_group;
    unsigned long long drop_seed;  // elementwise nn.Dropout on the branch (drop1/drop3, VidHRFormer_modules.py:53-55)
    float drop_p;
    // implicit-GEMM 3x3 convolution (A = 4-D TMA view of the padded NHWC activation): 0 = plain GEMM
    int conv_taps, conv_cpt, conv_kw, conv_bh, conv_tiles_per_frame, conv_bf, conv_C;
    int conv_w8;       // W == 8 implicit conv (conv3x3_w8_kernel): tile rows are ordered (oh, frame, ow) -- see epi_row()
    int conv_qw, conv_qh;   // > 0: the "frames" of the W == 8 kernel are 8x8 quadrants of larger frames (conv_qw x conv_qh per frame)
    int conv_planes;
    int epi_tma;     // 2-CTA kernel: the output tile leaves through TMA bulk stores (epilogue_tile_tma) instead of per-lane stores
    long long* dbg;  // optional timeline buffer (tools/bench_gemm.py): block 0 records clock64() per tile
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > (1u << 22)) {  // ~seconds: turn a protocol bug into a trap instead of a hung GPU
            printf("vptr gemm: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// One lane of a CONVERGED warp.  The role loops below are executed by all 32 lanes with warp-uniform values and only the
// TMA / MMA / commit instructions are predicated on the elected lane: issued from a divergent `if (lane == 0)` region instead,
// every tcgen05.mma costs ~110-140 cycles of uniform-register marshalling (R2UR + BRA.U.ANY loops in the SASS), which made the
// single issuing thread -- not the tensor core -- the bottleneck (measured with tools/gemm_timeline.py).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Shared-memory matrix descriptor (sm_100 "version 1").
//   K-major : SWIZZLE_128B (layout 2): rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused (canonical value 1).
//   MN-major: 32-bit operands only exist as SWIZZLE_128B_BASE32B (layout 1; 32-byte chunks permuted inside each
//             128-byte row, period 4 rows; TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 32-float MN groups
//             `lbo_bytes` apart, 4-k-row groups `sbo_bytes` (512 B) apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;
    return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One NCOLS-wide chunk of a warp's 32 accumulator rows.  tcgen05.ld hands every lane one ROW (NCOLS consecutive columns);
// storing that directly makes each warp instruction touch 32 different rows (32 half-used sectors).  The chunk is therefore
// transposed through a warp-private shared-memory tile (row pitch 36 floats: 16-byte aligned, conflict-free for the
// quarter-warp float4 accesses) so that global loads (residual) and stores (or split-K reductions) are whole 128-byte row
// segments: NCOLS/4 lanes per row, float4 per lane, all iterations independent (loads in flight together).
constexpr int EPI_PITCH = 36;
// lane -> (row-in-iteration, 4-column group) mapping of the coalesced phase for an NCOLS-wide chunk
template <int NCOLS>
struct ChunkMap {
    static constexpr int LPR = NCOLS / 4;   // lanes per row
    static constexpr int RPI = 32 / LPR;    // rows per iteration
    static constexpr int ITERS = 32 / RPI;
};
// Global loads of the epilogue (bias, residual) have 1-3k cycles of latency while TMA keeps the memory system busy, and the
// chunk loop has nothing to overlap them with: issued inside the chunk they made every 32-column chunk cost ~3.3k cycles
// (tools/gemm_timeline.py).  They are therefore issued early: bias for the whole tile and the residual of chunk 0 before the
// accumulator-ready wait, the residual of chunk c+1 while chunk c is processed.
// Rarely-used epilogue options (activation, dropout, DropPath scale) live out of line: inlined into the unrolled store loop they
// made the epilogue several thousand instructions long and instruction fetch -- not memory -- bound it.
__device__ __noinline__ float4 epilogue_options(const GemmParams& p, float4 o, int m, int n) {
    if (p.act == 1) {
        o.x = vptr_gelu(o.x); o.y = vptr_gelu(o.y); o.z = vptr_gelu(o.z); o.w = vptr_gelu(o.w);
    } else if (p.act == 2) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    if (p.drop_p > 0.f) {
        const float4 k = vptr_drop_scale4(p.drop_seed, ((unsigned long long)m * p.N + n) >> 2, p.drop_p);   // N % 4 == 0, n % 4 == 0
        o.x *= k.x; o.y *= k.y; o.z *= k.z; o.w *= k.w;
    }
    if (p.rowscale) {
        const float rs = __ldg(p.rowscale + m / p.rows_per_group);
        o.x *= rs; o.y *= rs; o.z *= rs; o.w *= rs;
    }
    return o;
}
// Output row of accumulator row `rr` of the warp whose 32-row block starts at m_base.  Plain GEMM: m_base + rr.  W == 8 implicit
// conv: the 128 rows of a CTA tile are ordered (oh, frame-in-tile, ow) so that every 8-row UMMA group is one image row and the
// groups are a constant 10 padded pixels apart in the raw shared-memory tile (conv3x3_w8_kernel).
__device__ __forceinline__ int epi_row(const GemmParams& p, int m_base, int rr) {
    if (!p.conv_w8) return m_base + rr;
    const int L = (m_base & 127) + rr, g = L >> 3;
    if (!p.conv_qw) return (m_base & ~127) + (g & 1) * 64 + (g >> 1) * 8 + (L & 7);
    // quadrant mode: quadrant fq = (frame f, quadrant row qy, quadrant column qx) -> row of the full (conv_qh*8) x (conv_qw*8) frame
    const int fq = ((m_base >> 7) << 1) + (g & 1), nq = p.conv_qw * p.conv_qh;
    const int f = fq / nq, q = fq - f * nq, qy = q / p.conv_qw, qx = q - qy * p.conv_qw;
    return ((f * p.conv_qh + qy) * 8 + (g >> 1)) * (p.conv_qw * 8) + qx * 8 + (L & 7);
}
template <int NCOLS>
__device__ __forceinline__ float4 load_bias(const GemmParams& p, int lane, int n_base) {
    const int n = n_base + (lane % ChunkMap<NCOLS>::LPR) * 4;
    if (p.bias && n < p.N) return __ldg(reinterpret_cast<const float4*>(p.bias + n));
    return make_float4(0.f, 0.f, 0.f, 0.f);
}
template <int NCOLS, bool PERM>
__device__ __forceinline__ void load_residual(const GemmParams& p, int lane, int m_base, int n_base, float4 (&res)[8]) {
    using CM = ChunkMap<NCOLS>;
    const int r_in = lane / CM::LPR, n = n_base + (lane % CM::LPR) * 4;
#pragma unroll
    for (int it = 0; it < CM::ITERS; ++it) {
        const int m = PERM ? epi_row(p, m_base, it * CM::RPI + r_in) : m_base + it * CM::RPI + r_in;
        res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.residual && n < p.N && m < p.M) res[it] = *reinterpret_cast<const float4*>(p.residual + (long long)m * p.ldr + n);
    }
}
// MODE (compile time, so the unrolled store loop carries no per-element flag tests -- the epilogue warps run one per scheduler
// and are latency-bound, every branch costs): 0 plain store, 1 store rounded to tf32, 2 split-K reduction (red.add), 3 all
// options (activation / dropout / DropPath scale, out of line), 4 DropPath row scale (inline), 5 dropout (+ row scale) inline.
// Modes 4/5 exist because the out-of-line path made the regularised out-proj / linear2 GEMMs 3x slower than the plain ones
// (234-245 us vs 81 us at M=40960 N=K=528, tools/bench_conv.py); rounding to tf32 is a runtime flag in modes 3-5.
struct RowScale {   // DropPath keep-scales of a warp's 32 rows: at most two clips when rows_per_group >= 32
    float s0, s1;
    int boundary;   // first row of the second clip
};
__device__ __forceinline__ RowScale load_rowscale(const GemmParams& p, int m_base) {
    RowScale r{1.f, 1.f, 0x7fffffff};
    if (p.rowscale) {
        const int g0 = m_base / p.rows_per_group;
        const int last = (p.M - 1) / p.rows_per_group;
        r.s0 = __ldg(p.rowscale + min(g0, last));
        r.s1 = __ldg(p.rowscale + min(g0 + 1, last));
        r.boundary = (g0 + 1) * p.rows_per_group;
    }
    return r;
}
template <int NCOLS, int MODE, bool PERM>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const float* v, float* sT, int lane, int m_base, int n_base,
                                               const float4 bias, const float4 (&res)[8], const RowScale rsc) {
    using CM = ChunkMap<NCOLS>;
#pragma unroll
    for (int j = 0; j < NCOLS; j += 4)
        *reinterpret_cast<float4*>(sT + lane * EPI_PITCH + j) = make_float4(v[j] * p.alpha, v[j + 1] * p.alpha, v[j + 2] * p.alpha, v[j + 3] * p.alpha);
    __syncwarp();
    const int r_in = lane / CM::LPR, c = (lane % CM::LPR) * 4;
    const int n = n_base + c;
    // plain GEMM: rows are affine in `it` (pointer stepping, one bound); the permuted conv tile recomputes each row
    const int rows_ok = (n < p.N) ? p.M - m_base - r_in : 0;          // rows of this lane's column group that exist (N % 4 == 0)
    float* dcol = p.D + (long long)(m_base + r_in) * p.ldd + n;
    const long long dstep = (long long)CM::RPI * p.ldd;
#pragma unroll
    for (int it = 0; it < CM::ITERS; ++it) {
        const int m = PERM ? epi_row(p, m_base, it * CM::RPI + r_in) : m_base + it * CM::RPI + r_in;
        if (PERM ? (n < p.N && m < p.M) : (it * CM::RPI < rows_ok)) {
            float4 o = *reinterpret_cast<const float4*>(sT + (it * CM::RPI + r_in) * EPI_PITCH + c);
            o.x += bias.x; o.y += bias.y; o.z += bias.z; o.w += bias.w;
            if (MODE == 3) o = epilogue_options(p, o, m, n);
            if (MODE == 5) {   // (no activation in this mode)
                const float4 k = vptr_drop_scale4(p.drop_seed, ((unsigned long long)m * p.N + n) >> 2, p.drop_p);
                o.x *= k.x; o.y *= k.y; o.z *= k.z; o.w *= k.w;
            }
            if (MODE == 4 || MODE == 5) {
                if (p.act == 2) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                const float rs = m < rsc.boundary ? rsc.s0 : rsc.s1;
                o.x *= rs; o.y *= rs; o.z *= rs; o.w *= rs;
            }
            o.x += res[it].x; o.y += res[it].y; o.z += res[it].z; o.w += res[it].w;
            if (MODE == 1 || (MODE >= 3 && (p.flags & 2))) {
                o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w);
            }
            float* d = PERM ? p.D + (long long)m * p.ldd + n : dcol + it * dstep;
            if (MODE == 2)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
            else
                *reinterpret_cast<float4*>(d) = o;
        }
    }
    __syncwarp();
}

// Whole-tile epilogue of one warp (its 32 accumulator rows, BLOCK_N columns).  Global loads (bias, residual) have 1-3k cycles
// of latency while TMA keeps the memory system busy, so they are issued early: for chunk 0 before the accumulator-ready wait,
// for chunk c+1 while chunk c is stored.
template <int BLOCK_N, int MODE, bool PERM, class WaitFn>
__device__ __forceinline__ void epilogue_tile_mode(const GemmParams& p, uint32_t taddr, float* sT, int lane, int m_base, int n0,
                                                   WaitFn wait_full) {
    constexpr int NFULL = BLOCK_N / 32;
    constexpr bool TAIL = (BLOCK_N % 32) != 0;
    float4 res[8];
    float4 bias = load_bias<32>(p, lane, n0);
    load_residual<32, PERM>(p, lane, m_base, n0, res);
    const RowScale rsc = (MODE == 4 || MODE == 5) ? load_rowscale(p, m_base) : RowScale{1.f, 1.f, 0x7fffffff};
    wait_full();
#pragma unroll 1
    for (int c = 0; c < NFULL; ++c) {
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        const float4 bias_c = bias;
        float4 res_c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) res_c[i] = res[i];
        if (c + 1 < NFULL) {
            bias = load_bias<32>(p, lane, n0 + (c + 1) * 32);
            load_residual<32, PERM>(p, lane, m_base, n0 + (c + 1) * 32, res);
        } else if (TAIL) {
            bias = load_bias<16>(p, lane, n0 + (c + 1) * 32);
            load_residual<16, PERM>(p, lane, m_base, n0 + (c + 1) * 32, res);
        }
        epilogue_chunk<32, MODE, PERM>(p, v, sT, lane, m_base, n0 + c * 32, bias_c, res_c, rsc);
    }
    if (TAIL) {
        float v[16];
        tmem_ld16(taddr + NFULL * 32, v);
        tmem_ld_wait();
        epilogue_chunk<16, MODE, PERM>(p, v, sT, lane, m_base, n0 + NFULL * 32, bias, res, rsc);
    }
}
template <int BLOCK_N, bool PERM, class WaitFn>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t taddr, float* sT, int lane, int m_base, int n0, WaitFn wait_full) {
    const bool fancy = p.act != 0 || p.drop_p > 0.f || p.rowscale != nullptr;
    if (PERM) {   // W == 8 implicit conv: bias (+ ReLU) (+ residual) (+ tf32 rounding)
        if (fancy) epilogue_tile_mode<BLOCK_N, (PERM ? 4 : 0), PERM>(p, taddr, sT, lane, m_base, n0, wait_full);
        else if (p.flags & 2) epilogue_tile_mode<BLOCK_N, (PERM ? 1 : 0), PERM>(p, taddr, sT, lane, m_base, n0, wait_full);
        else epilogue_tile_mode<BLOCK_N, 0, PERM>(p, taddr, sT, lane, m_base, n0, wait_full);
        return;
    }
    const bool inline_ok = (p.act == 0 || p.act == 2) && (p.rowscale == nullptr || p.rows_per_group >= 32);
    if (p.flags & 1) epilogue_tile_mode<BLOCK_N, 2, false>(p, taddr, sT, lane, m_base, n0, wait_full);
    else if (fancy && inline_ok && p.drop_p > 0.f && p.act == 0) epilogue_tile_mode<BLOCK_N, 5, false>(p, taddr, sT, lane, m_base, n0, wait_full);
    else if (fancy && inline_ok && p.drop_p <= 0.f) epilogue_tile_mode<BLOCK_N, 4, false>(p, taddr, sT, lane, m_base, n0, wait_full);
    else if (fancy) epilogue_tile_mode<BLOCK_N, 3, false>(p, taddr, sT, lane, m_base, n0, wait_full);
    else if (p.flags & 2) epilogue_tile_mode<BLOCK_N, 1, false>(p, taddr, sT, lane, m_base, n0, wait_full);
    else epilogue_tile_mode<BLOCK_N, 0, false>(p, taddr, sT, lane, m_base, n0, wait_full);
}

// ---- TMA-store epilogue (2-CTA kernel, outputs without a residual operand or split-K reduction) --------------------------------
// The K = 528 GEMMs of the path are epilogue-bound: with per-lane stores a 128 x 176 tile costs ~14 k cycles (tcgen05.ld -> smem
// transpose -> ld.shared -> 128-byte row-segment STGs, each warp instruction touching 4-8 different lines) against ~9.7 k cycles
// of main loop at the sustained TF32 rate.  Here every lane keeps its accumulator ROW (what tcgen05.ld delivers), applies the
// epilogue math in registers, writes the 32-column chunk into a warp-private staging tile in the SWIZZLE_128B pattern (16-byte
// chunk j of row r at r*128 + ((j ^ (r & 7)) << 4): conflict-free) and ONE lane hands the 32 x 32 tile to the TMA engine
// (cp.async.bulk.tensor store; out-of-range rows / columns are clipped by the tensor map).  Two staging tiles per warp: the store
// of chunk c drains while chunk c+1 is computed; the tcgen05.ld of chunk c+1 is in flight during the math of chunk c.
constexpr int EPI_TMA_TILE = 4096;      // 32 rows x 128 B
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
template <int NCOLS, int MODE, bool RES>
__device__ __forceinline__ void epilogue_chunk_tma(const GemmParams& p, const CUtensorMap* map, const float* v, uint8_t* buf, int lane,
                                                   int m_base, int n_base, float rs) {
    const int m = min(m_base + lane, p.M - 1);      // rows past M are clipped by the store; keep their side look-ups in range
#pragma unroll
    for (int j = 0; j < NCOLS / 4; ++j) {
        const int n = n_base + 4 * j;
        float4 o = make_float4(v[4 * j] * p.alpha, v[4 * j + 1] * p.alpha, v[4 * j + 2] * p.alpha, v[4 * j + 3] * p.alpha);
        if (p.bias && n < p.N) {      // warp-uniform address: one broadcast transaction, L1-resident after the first tile
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
        if (MODE == 3) o = epilogue_options(p, o, m, n);
        if (MODE == 5) {
            const float4 k = vptr_drop_scale4(p.drop_seed, ((unsigned long long)m * p.N + n) >> 2, p.drop_p);
            o.x *= k.x; o.y *= k.y; o.z *= k.z; o.w *= k.w;
        }
        if (MODE == 4 || MODE == 5) {
            if (p.act == 2) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            o.x *= rs; o.y *= rs; o.z *= rs; o.w *= rs;
        }
        const uint32_t off = NCOLS == 32 ? lane * 128 + ((j ^ (lane & 7)) << 4) : lane * (NCOLS * 4) + (j << 4);
        if (RES) {      // the residual chunk was bulk-loaded into this very tile: add in place
            const float4 r = *reinterpret_cast<const float4*>(buf + off);
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (MODE == 1 || (MODE >= 3 && (p.flags & 2))) {
            o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w);
        }
        *reinterpret_cast<float4*>(buf + off) = o;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(map, buf, n_base, m_base);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}
template <int BLOCK_N, int MODE, class WaitFn>
__device__ __forceinline__ void epilogue_tile_tma_mode(const GemmParams& p, const CUtensorMap* map32, const CUtensorMap* map16, uint32_t taddr,
                                                       uint8_t* sbuf, uint32_t& nstore, int lane, int m_base, int n0, WaitFn wait_full) {
    constexpr int NFULL = BLOCK_N / 32;
    constexpr bool TAIL = (BLOCK_N % 32) != 0;
    const RowScale rsc = (MODE == 4 || MODE == 5) ? load_rowscale(p, m_base) : RowScale{1.f, 1.f, 0x7fffffff};
    const float rs = (m_base + lane) < rsc.boundary ? rsc.s0 : rsc.s1;
    const bool rows_live = m_base < p.M;            // warp-uniform
    wait_full();
    float v[32];
    tmem_ld32(taddr, v);
#pragma unroll 1
    for (int c = 0; c < NFULL; ++c) {
        float cur[32];
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) cur[i] = v[i];
        if (c + 1 < NFULL) tmem_ld32(taddr + (c + 1) * 32, v);
        else if (TAIL) tmem_ld16(taddr + NFULL * 32, v);
        const int n_base = n0 + c * 32;
        if (rows_live && n_base < p.N) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store issued two chunks ago has left its tile
            __syncwarp();
            epilogue_chunk_tma<32, MODE, false>(p, map32, cur, sbuf + (nstore & 1) * EPI_TMA_TILE, lane, m_base, n_base, rs);
            ++nstore;
        }
    }
    if (TAIL) {
        tmem_ld_wait();
        const int n_base = n0 + NFULL * 32;
        if (rows_live && n_base < p.N) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            epilogue_chunk_tma<16, MODE, false>(p, map16, v, sbuf + (nstore & 1) * EPI_TMA_TILE, lane, m_base, n_base, rs);
            ++nstore;
        }
    }
}
// Residual variant: the residual chunk is bulk-LOADED into the staging tile T-2 chunks ahead (T tiles per warp rotate: being
// loaded / being combined / being stored), the lane adds its accumulator row in place and the same tile is bulk-stored.  With
// per-lane residual loads the 528 -> 528 out-projections sat at 15-23 k cycles per tile (1-3 k cycles of load latency per chunk
// that nothing overlapped) against 9.4 k of main loop; the bulk loads queue behind the operand stages in the SM's TMA pipe, so
// one chunk of lookahead (T = 3) still left 11 k cycles per tile, two chunks (T = 4) 8.5 k (tools/gemm_timeline.py, RES=1).
struct ResPipe {
    uint8_t* buf;        // T staging tiles of this warp
    uint64_t* bar;       // T mbarriers (one per tile)
    uint32_t k;          // running chunk counter (selects the tile)
    uint32_t phases;     // bit b: parity the next wait on bar[b] expects
};
template <int BLOCK_N>
struct ResChunks {
    static constexpr int NFULL = BLOCK_N / 32;
    static constexpr bool TAIL = (BLOCK_N % 32) != 0;
    static constexpr int NCH = NFULL + (TAIL ? 1 : 0);
};
// residual chunk `ci` of the tile at (m_base, n0) -> staging tile (rp.k + ahead) % T; nothing is issued for chunks outside the matrix
template <int BLOCK_N, int T>
__device__ __forceinline__ void res_issue(const GemmParams& p, const CUtensorMap* r32, const CUtensorMap* r16, ResPipe& rp, uint32_t ahead,
                                          int m_base, int n0, int ci) {
    const int n_base = n0 + ci * 32;
    if (m_base >= p.M || n_base >= p.N) return;
    const uint32_t b = (rp.k + ahead) % (uint32_t)T;
    const bool tail = ResChunks<BLOCK_N>::TAIL && ci == ResChunks<BLOCK_N>::NFULL;
    mbar_expect_tx(&rp.bar[b], tail ? 32 * 64 : 32 * 128);
    tma_load_2d(tail ? r16 : r32, &rp.bar[b], rp.buf + b * EPI_TMA_TILE, n_base, m_base);
}
template <int BLOCK_N, int T, int MODE, class WaitFn>
__device__ __forceinline__ void epilogue_tile_tma_res_mode(const GemmParams& p, const CUtensorMap* d32, const CUtensorMap* d16,
                                                           const CUtensorMap* r32, const CUtensorMap* r16, uint32_t taddr, ResPipe& rp, int lane,
                                                           int m_base, int n0, bool has_next, int next_m_base, int next_n0, WaitFn wait_full) {
    using RC = ResChunks<BLOCK_N>;
    const RowScale rsc = (MODE == 4 || MODE == 5) ? load_rowscale(p, m_base) : RowScale{1.f, 1.f, 0x7fffffff};
    const float rs = (m_base + lane) < rsc.boundary ? rsc.s0 : rsc.s1;
    wait_full();
    float v[32];
    tmem_ld32(taddr, v);
#pragma unroll 1
    for (int c = 0; c < RC::NCH; ++c) {
        if (lane == 0) {   // staging tile (k + T-2) % T was last read by the store issued two chunks ago
            constexpr int AHEAD = T - 2;
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            if (c + AHEAD < RC::NCH) res_issue<BLOCK_N, T>(p, r32, r16, rp, AHEAD, m_base, n0, c + AHEAD);
            else if (has_next) res_issue<BLOCK_N, T>(p, r32, r16, rp, AHEAD, next_m_base, next_n0, c + AHEAD - RC::NCH);
        }
        float cur[32];
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) cur[i] = v[i];
        if (c + 1 < RC::NFULL) tmem_ld32(taddr + (c + 1) * 32, v);
        else if (RC::TAIL && c + 1 == RC::NFULL) tmem_ld16(taddr + RC::NFULL * 32, v);
        const int n_base = n0 + c * 32;
        if (m_base < p.M && n_base < p.N) {
            const uint32_t b = rp.k % (uint32_t)T;
            mbar_wait(&rp.bar[b], (rp.phases >> b) & 1u);
            rp.phases ^= 1u << b;
            if (RC::TAIL && c == RC::NFULL) epilogue_chunk_tma<16, MODE, true>(p, d16, cur, rp.buf + b * EPI_TMA_TILE, lane, m_base, n_base, rs);
            else epilogue_chunk_tma<32, MODE, true>(p, d32, cur, rp.buf + b * EPI_TMA_TILE, lane, m_base, n_base, rs);
        } else if (lane == 0) {
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");     // empty group: keeps "two chunks ago" == "two groups ago"
        }
        ++rp.k;
    }
}
template <int BLOCK_N, int T, class WaitFn>
__device__ __forceinline__ void epilogue_tile_tma_res(const GemmParams& p, const CUtensorMap* d32, const CUtensorMap* d16, const CUtensorMap* r32,
                                                      const CUtensorMap* r16, uint32_t taddr, ResPipe& rp, int lane, int m_base, int n0,
                                                      bool has_next, int next_m_base, int next_n0, WaitFn wait_full) {
    const bool fancy = p.act != 0 || p.drop_p > 0.f || p.rowscale != nullptr;
    const bool inline_ok = (p.act == 0 || p.act == 2) && (p.rowscale == nullptr || p.rows_per_group >= 32);
#define VPTR_EPI_RES(MODE) epilogue_tile_tma_res_mode<BLOCK_N, T, MODE>(p, d32, d16, r32, r16, taddr, rp, lane, m_base, n0, has_next, next_m_base, next_n0, wait_full)
    if (fancy && inline_ok && p.drop_p > 0.f && p.act == 0) VPTR_EPI_RES(5);
    else if (fancy && inline_ok && p.drop_p <= 0.f) VPTR_EPI_RES(4);
    else if (fancy) VPTR_EPI_RES(3);
    else if (p.flags & 2) VPTR_EPI_RES(1);
    else VPTR_EPI_RES(0);
#undef VPTR_EPI_RES
}

template <int BLOCK_N, class WaitFn>
__device__ __forceinline__ void epilogue_tile_tma(const GemmParams& p, const CUtensorMap* map32, const CUtensorMap* map16, uint32_t taddr,
                                                  uint8_t* sbuf, uint32_t& nstore, int lane, int m_base, int n0, WaitFn wait_full) {
    const bool fancy = p.act != 0 || p.drop_p > 0.f || p.rowscale != nullptr;
    const bool inline_ok = (p.act == 0 || p.act == 2) && (p.rowscale == nullptr || p.rows_per_group >= 32);
    if (fancy && inline_ok && p.drop_p > 0.f && p.act == 0) epilogue_tile_tma_mode<BLOCK_N, 5>(p, map32, map16, taddr, sbuf, nstore, lane, m_base, n0, wait_full);
    else if (fancy && inline_ok && p.drop_p <= 0.f) epilogue_tile_tma_mode<BLOCK_N, 4>(p, map32, map16, taddr, sbuf, nstore, lane, m_base, n0, wait_full);
    else if (fancy) epilogue_tile_tma_mode<BLOCK_N, 3>(p, map32, map16, taddr, sbuf, nstore, lane, m_base, n0, wait_full);
    else if (p.flags & 2) epilogue_tile_tma_mode<BLOCK_N, 1>(p, map32, map16, taddr, sbuf, nstore, lane, m_base, n0, wait_full);
    else epilogue_tile_tma_mode<BLOCK_N, 0>(p, map32, map16, taddr, sbuf, nstore, lane, m_base, n0, wait_full);
}

template <int BLOCK_N, int A_MN, int B_MN, int STAGES>
struct GemmCfg {
    static constexpr int A_BYTES = BLOCK_M * 128;
    static constexpr int B_GROUPS = (BLOCK_N + 31) / 32;
    static constexpr int B_BYTES = B_MN ? B_GROUPS * 4096 : BLOCK_N * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_BYTES = 4 * 32 * EPI_PITCH * 4;                      // per-epilogue-warp transpose tiles
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + EPI_BYTES + 1024;  // + barriers + manual 1024 B alignment slack
    static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N constraint for M=128");
    static_assert(B_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned for SWIZZLE_128B");
};

template <int BLOCK_N, int A_MN, int B_MN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N, A_MN, B_MN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* epi_tiles = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles * p.k_splits;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 512 columns = two BLOCK_N-wide fp32 accumulators (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {  // ===== TMA producer (whole warp converged; one elected lane issues) =====
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int ks = tile / (p.m_tiles * p.n_tiles);   // ks-major order: see GemmParams::k_splits
                const int mn = tile % (p.m_tiles * p.n_tiles);
                const int n_tile = mn % p.n_tiles;
                const int m_tile = mn / p.n_tiles;
                const int k0 = ks * p.chunks_per_split;
                const int k1 = min(k0 + p.chunks_per_split, p.total_chunks);
                for (int kc = k0; kc < k1; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sB = sA + Cfg::A_BYTES;
                    if (elect_one()) {
                    mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    if (!A_MN) {
                        tma_load_2d(&tma_a, &full_bar[stage], sA, kc * BLOCK_K, m_tile * BLOCK_M);
                    } else {
#pragma unroll
                        for (int g = 0; g < BLOCK_M / 32; ++g)
                            tma_load_2d(&tma_a, &full_bar[stage], sA + g * 4096, m_tile * BLOCK_M + g * 32, kc * BLOCK_K);
                    }
                    if (!B_MN) {
                        tma_load_2d(&tma_b, &full_bar[stage], sB, kc * BLOCK_K, n_tile * BLOCK_N);
                    } else {
#pragma unroll
                        for (int g = 0; g < Cfg::B_GROUPS; ++g)
                            tma_load_2d(&tma_b, &full_bar[stage], sB + g * 4096, n_tile * BLOCK_N + g * 32, kc * BLOCK_K);
                    }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        {  // ===== MMA issuer (whole warp converged; one elected lane issues) =====
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(A_MN) << 15) | (uint32_t(B_MN) << 16) |
                                       (uint32_t(BLOCK_N >> 3) << 17) | (uint32_t(BLOCK_M >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int ks = tile / (p.m_tiles * p.n_tiles);   // ks-major order: see GemmParams::k_splits
                const int k0 = ks * p.chunks_per_split;
                const int k1 = min(k0 + p.chunks_per_split, p.total_chunks);
                if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[(tile / gridDim.x) * 8 + 0] = clock64();
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[(tile / gridDim.x) * 8 + 1] = clock64();
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * BLOCK_N);
                for (int kc = k0; kc < k1; ++kc) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t b_base = a_base + Cfg::A_BYTES;
                    if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t da = A_MN ? make_smem_desc(a_base + k * 1024, 4096, 512, 1)
                                                 : make_smem_desc(a_base + k * 32, 16, 1024, 2);
                        const uint64_t db = B_MN ? make_smem_desc(b_base + k * 1024, 4096, 512, 1)
                                                 : make_smem_desc(b_base + k * 32, 16, 1024, 2);
                        umma_tf32(d_tmem, da, db, idesc, (kc > k0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
                __syncwarp();
                if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[(tile / gridDim.x) * 8 + 2] = clock64();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {  // ===== epilogue warps 2..5: TMEM -> registers -> global =====
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mn = tile % (p.m_tiles * p.n_tiles);
            const int n_tile = mn % p.n_tiles;
            const int m_tile = mn / p.n_tiles;
            const int m_base = m_tile * BLOCK_M + q * 32;
            float* sT = epi_tiles + (warp - 2) * 32 * EPI_PITCH;
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BLOCK_N);
            const bool stamp = p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0;
            if (stamp) p.dbg[(tile / gridDim.x) * 8 + 3] = clock64();
            epilogue_tile<BLOCK_N, false>(p, taddr, sT, lane, m_base, n_tile * BLOCK_N, [&] {
                mbar_wait(&tmem_full[acc], acc_phase);
                if (stamp) p.dbg[(tile / gridDim.x) * 8 + 4] = clock64();
                tcgen05_fence_after();
            });
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0) p.dbg[(tile / gridDim.x) * 8 + 5] = clock64();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// =============================================================================================
// 2-CTA variant (cta_group::2): a CTA pair on one TPC computes a 256 x BLOCK_N tile.  Each CTA stages its own 128 rows
// of A and HALF of the B tile; the pair's tensor cores read B halves from both shared memories, so per SM the operand
// traffic through shared memory drops from (128 + N) to (128 + N/2) rows per k-chunk.  That matters because the 1-CTA
// kernel is shared-memory-bandwidth bound: TMA writes + MMA operand reads of a 128x176 fp32 tile are 77.8 KB per
// k-chunk = 608 cycles at 128 B/clk against 352 cycles of MMA (measured 590, tools/gemm_timeline.py), and the saturated
// port also starves the epilogue's own loads/stores.
//   * leader CTA (cluster rank 0): single thread issues tcgen05.mma.cta_group::2; full barriers live in the leader
//     (one arrival: its own expect_tx for both CTAs' bytes; both CTAs' TMA complete_tx land there).
//   * tcgen05.commit multicasts to both CTAs' empty / tmem_full barriers.
//   * both epilogues arrive on the leader's tmem_empty barrier.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // arrive on CTA 0's copy of this barrier
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(bar)));
    // relaxed: the barrier only orders TMEM reads (already fenced by tcgen05.wait::ld + fence::before_thread_sync); a release
    // arrive would make the warp wait for all its outstanding global stores (ERRBAR, 5 % of the kernel's stall samples)
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    const uint32_t leader_bar = smem_u32(bar) & 0xFEFFFFFFu;   // same offset in the even (leader) CTA of the pair
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_2cta(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    const uint32_t leader_bar = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int BLOCK_N, int A_MN, int B_MN, int STAGES, int EPI_TILES = 2>
struct Gemm2Cfg {
    static constexpr int HALF_N = BLOCK_N / 2;
    static constexpr int A_BYTES = BLOCK_M * 128;
    static constexpr int B_GROUPS = HALF_N / 32;                       // MN-major only
    static constexpr int B_BYTES = B_MN ? B_GROUPS * 4096 : HALF_N * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_BYTES = 4 * EPI_TILES * EPI_TMA_TILE;   // TMA staging tiles per epilogue warp: 2, or 3 with a bulk-loaded
                                                                      // residual (the transpose tiles of the per-lane path alias them)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 256 + 1024;
    static_assert(EPI_BYTES >= 4 * 32 * EPI_PITCH * 4, "transpose tiles must fit");
    static_assert(BLOCK_N % 16 == 0 && BLOCK_N <= 256, "UMMA N constraint for M=256");
    static_assert(!B_MN || HALF_N % 32 == 0, "MN-major B halves must be whole 32-float groups");
    static_assert(B_MN || HALF_N % 8 == 0, "K-major B halves must be whole 8-row swizzle groups");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int BLOCK_N, int A_MN, int B_MN, int STAGES, int EPI_TILES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tf32_2cta_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                      const __grid_constant__ CUtensorMap tma_d, const __grid_constant__ CUtensorMap tma_d16,
                      const __grid_constant__ CUtensorMap tma_r, const __grid_constant__ CUtensorMap tma_r16, const GemmParams p) {
    using Cfg = Gemm2Cfg<BLOCK_N, A_MN, B_MN, STAGES, EPI_TILES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_base = smem + STAGES * Cfg::STAGE_BYTES;          // 1024-byte aligned (swizzled TMA staging tiles)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_base + Cfg::EPI_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint64_t* res_bar = tmem_empty + 3;                            // 4 warps x EPI_TILES residual-tile barriers (EPI_TILES > 2)
    static_assert((2 * STAGES + 4 + 1 + (EPI_TILES > 2 ? 4 * EPI_TILES : 0)) * 8 <= 256, "barrier block");
    float* epi_tiles = reinterpret_cast<float*>(epi_base);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int n_clusters = gridDim.x >> 1;
    // m_tiles here counts 256-row pair tiles
    const int total_tiles = p.m_tiles * p.n_tiles * p.k_splits;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
        if (p.epi_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_d) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);    // the leader's expect_tx arrive (used in the leader only; the peer's TMA bytes
                                           // can only land in the same phase because it waits on the multicast empty barrier)
            mbar_init(&empty_bar[s], 1);   // multicast tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);   // multicast tcgen05.commit
            mbar_init(&tmem_empty[a], 8);  // 4 epilogue warps x 2 CTAs (used in the leader only)
        }
        if (EPI_TILES > 2) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_r) : "memory");
            for (int i = 0; i < 4 * EPI_TILES; ++i) mbar_init(&res_bar[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {  // ===== TMA producer (both CTAs; whole warp converged, one elected lane issues) =====
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
                const int ks = tile / (p.m_tiles * p.n_tiles);   // ks-major order: see GemmParams::k_splits
                const int mn = tile % (p.m_tiles * p.n_tiles);
                const int n_tile = mn % p.n_tiles;
                const int m_pair = mn / p.n_tiles;
                const int k0 = ks * p.chunks_per_split;
                const int k1 = min(k0 + p.chunks_per_split, p.total_chunks);
                const int m0 = (m_pair * 2 + (int)rank) * BLOCK_M;
                const int n0 = n_tile * BLOCK_N + (int)rank * Cfg::HALF_N;
                for (int kc = k0; kc < k1; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sB = sA + Cfg::A_BYTES;
                    if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                    if (!A_MN && p.conv_taps) {
                        // implicit GEMM: k-chunk kc = (tap, 32-channel slice); the A tile is the tap-shifted window of the padded
                        // NHWC activation, fetched as one 4-D box {32 ch, W, bh rows, bf frames} = 128 output pixels
                        const int kk = kc % (p.conv_taps * p.conv_cpt);            // weight-split planes reuse the same A tiles
                        const int tap = kk / p.conv_cpt, cs = kk - tap * p.conv_cpt;
                        const int kh = tap / p.conv_kw, kw = tap - kh * p.conv_kw;
                        const int t = m_pair * 2 + (int)rank;                   // 128-pixel tile index
                        const int f0 = p.conv_bf > 1 ? t * p.conv_bf : t / p.conv_tiles_per_frame;
                        const int oh0 = p.conv_bf > 1 ? 0 : (t - f0 * p.conv_tiles_per_frame) * p.conv_bh;
                        tma_load_4d_2cta(&tma_a, &full_bar[stage], sA, cs * BLOCK_K, kw, oh0 + kh, f0);
                    } else if (!A_MN) {
                        tma_load_2d_2cta(&tma_a, &full_bar[stage], sA, kc * BLOCK_K, m0);
                    } else {
#pragma unroll
                        for (int g = 0; g < BLOCK_M / 32; ++g)
                            tma_load_2d_2cta(&tma_a, &full_bar[stage], sA + g * 4096, m0 + g * 32, kc * BLOCK_K);
                    }
                    if (!B_MN) {
                        int kcol = kc * BLOCK_K;
                        if (p.conv_taps) {
                            const int per_plane = p.conv_taps * p.conv_cpt;
                            const int plane = kc / per_plane, kk = kc - plane * per_plane;
                            kcol = plane * p.conv_taps * p.conv_C + (kk / p.conv_cpt) * p.conv_C + (kk % p.conv_cpt) * BLOCK_K;
                        }
                        tma_load_2d_2cta(&tma_b, &full_bar[stage], sB, kcol, n0);
                    } else {
#pragma unroll
                        for (int g = 0; g < Cfg::B_GROUPS; ++g)
                            tma_load_2d_2cta(&tma_b, &full_bar[stage], sB + g * 4096, n0 + g * 32, kc * BLOCK_K);
                    }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {  // ===== MMA issuer: the leader CTA's warp 1 (converged), one elected lane drives both tensor cores =====
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(A_MN) << 15) | (uint32_t(B_MN) << 16) |
                                       (uint32_t(BLOCK_N >> 3) << 17) | (uint32_t((2 * BLOCK_M) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
                const int ks = tile / (p.m_tiles * p.n_tiles);   // ks-major order: see GemmParams::k_splits
                const int k0 = ks * p.chunks_per_split;
                const int k1 = min(k0 + p.chunks_per_split, p.total_chunks);
                if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[(tile / n_clusters) * 8 + 0] = clock64();
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[(tile / n_clusters) * 8 + 1] = clock64();
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * BLOCK_N);
                for (int kc = k0; kc < k1; ++kc) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t b_base = a_base + Cfg::A_BYTES;
                    if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t da = A_MN ? make_smem_desc(a_base + k * 1024, 4096, 512, 1)
                                                 : make_smem_desc(a_base + k * 32, 16, 1024, 2);
                        const uint64_t db = B_MN ? make_smem_desc(b_base + k * 1024, 4096, 512, 1)
                                                 : make_smem_desc(b_base + k * 32, 16, 1024, 2);
                        umma_tf32_2cta(d_tmem, da, db, idesc, (kc > k0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit_2cta(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit_2cta(&tmem_full[acc]);
                __syncwarp();
                if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[(tile / n_clusters) * 8 + 2] = clock64();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {  // ===== epilogue warps 2..5 of both CTAs: own 128 rows =====
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t nstore = 0;
        uint8_t* sbuf = epi_base + (warp - 2) * EPI_TILES * EPI_TMA_TILE;
        ResPipe rp{sbuf, res_bar + (warp - 2) * EPI_TILES, 0u, 0u};
        auto tile_origin = [&](int tile, int& mb, int& nb) {
            const int mn = tile % (p.m_tiles * p.n_tiles);
            mb = ((mn / p.n_tiles) * 2 + (int)rank) * BLOCK_M + q * 32;
            nb = (mn % p.n_tiles) * BLOCK_N;
        };
        if (EPI_TILES > 2 && p.epi_tma == 2 && cluster_id < total_tiles && lane == 0) {   // residual of the very first chunk(s)
            int mb, nb;
            tile_origin(cluster_id, mb, nb);
            for (int a = 0; a < EPI_TILES - 2; ++a) res_issue<BLOCK_N, EPI_TILES>(p, &tma_r, &tma_r16, rp, a, mb, nb, a);
        }
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
            const int mn = tile % (p.m_tiles * p.n_tiles);
            const int n_tile = mn % p.n_tiles;
            const int m_pair = mn / p.n_tiles;
            const int m_base = (m_pair * 2 + (int)rank) * BLOCK_M + q * 32;
            float* sT = epi_tiles + (warp - 2) * 32 * EPI_PITCH;
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BLOCK_N);
            const bool stamp = p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0;
            if (stamp) p.dbg[(tile / n_clusters) * 8 + 3] = clock64();
            auto wait_full = [&] {
                mbar_wait(&tmem_full[acc], acc_phase);
                if (stamp) p.dbg[(tile / n_clusters) * 8 + 4] = clock64();
                tcgen05_fence_after();
            };
            if (EPI_TILES > 2 && p.epi_tma == 2) {
                const bool has_next = tile + n_clusters < total_tiles;
                int nmb = 0, nnb = 0;
                if (has_next) tile_origin(tile + n_clusters, nmb, nnb);
                epilogue_tile_tma_res<BLOCK_N, (EPI_TILES > 2 ? EPI_TILES : 3)>(p, &tma_d, &tma_d16, &tma_r, &tma_r16, taddr, rp, lane, m_base, n_tile * BLOCK_N, has_next, nmb, nnb,
                                               wait_full);
            } else if (p.epi_tma == 1) {
                epilogue_tile_tma<BLOCK_N>(p, &tma_d, &tma_d16, taddr, sbuf, nstore, lane, m_base, n_tile * BLOCK_N, wait_full);
            } else {
                epilogue_tile<BLOCK_N, false>(p, taddr, sT, lane, m_base, n_tile * BLOCK_N, wait_full);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
            if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0) p.dbg[(tile / n_clusters) * 8 + 5] = clock64();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (p.epi_tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all bulk stores of this warp complete
    }

    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}


// =============================================================================================
// Implicit-GEMM 3x3 stride-1 convolution for W == 8 feature grids (the 8x8 ResnetBlock grid of every 64x64 config).
// Feeding each (tap, 32-channel slice) k-chunk with its own 4-D TMA box (the generic path above) makes TMA the limit: boxes of
// 8-pixel runs cost ~6 cycles per 128-byte row, 943 cycles per k-chunk against 352 cycles of MMA (tools/bench_conv.py).  Here
// the RAW padded tile of a channel slice -- {32 ch, 10 w, 2 frames, 10 h} = 200 padded pixels x 128 B, stored [h][frame][w] --
// is loaded ONCE and all 9 taps (x weight planes) read it in place: the A descriptor of tap (kh,kw) starts kh*20+kw rows into
// the tile, every 8-row UMMA group is one image row (8 consecutive padded pixels), and consecutive groups (oh, frame) are a
// constant 10 rows = 1280 B apart (SBO).  Operand traffic through TMA drops from 9 x 16 KB to 25.6 KB per channel slice;
// only the weight chunks stream through the mbarrier ring.  Tile rows are therefore ordered (oh, frame, ow): epi_row().
constexpr int CW_BSTAGES = 10;
constexpr int CW_A_BYTES = 200 * 128;          // raw tile of one CTA (2 frames)
constexpr int CW_B_BYTES = 88 * 128;           // half of a 176-row weight chunk
constexpr int CW_SMEM_BYTES = 2 * CW_A_BYTES + CW_BSTAGES * CW_B_BYTES + 256 + 4 * 32 * EPI_PITCH * 4 + 1024;
static_assert(CW_A_BYTES % 1024 == 0 && CW_B_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned for SWIZZLE_128B");
static_assert(CW_SMEM_BYTES <= 227 * 1024, "shared memory budget");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv3x3_w8_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
    constexpr int BLOCK_N = 176;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sAraw = smem;
    uint8_t* sBring = smem + 2 * CW_A_BYTES;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(sBring + CW_BSTAGES * CW_B_BYTES);
    uint64_t* b_empty = b_full + CW_BSTAGES;
    uint64_t* a_full = b_empty + CW_BSTAGES;
    uint64_t* a_empty = a_full + 2;
    uint64_t* tmem_full = a_empty + 2;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* epi_tiles = reinterpret_cast<float*>(sBring + CW_BSTAGES * CW_B_BYTES + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int n_clusters = gridDim.x >> 1;
    const int total_tiles = p.m_tiles * p.n_tiles;
    const int chunks_per_slice = p.conv_planes * 9;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
        for (int s = 0; s < CW_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&a_full[a], 1);
            mbar_init(&a_empty[a], 1);
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {   // ===== TMA producer (both CTAs) =====
        int bs = 0, ab = 0;
        uint32_t bphase = 0, aphase = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
            const int n_tile = tile % p.n_tiles, m_pair = tile / p.n_tiles;
            const int f0 = (m_pair * 2 + (int)rank) * 2;                       // first of this CTA's two frames
            const int n0 = n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2);
            for (int cs = 0; cs < p.conv_cpt; ++cs) {
                mbar_wait(&a_empty[ab], aphase ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(&a_full[ab], 2 * CW_A_BYTES);
                    tma_load_4d_2cta(&tma_a, &a_full[ab], sAraw + ab * CW_A_BYTES, cs * BLOCK_K, 0, f0, 0);   // dims (c, w, frame, h)
                }
                __syncwarp();
                if (++ab == 2) { ab = 0; aphase ^= 1; }
                for (int j = 0; j < chunks_per_slice; ++j) {                   // j = plane * 9 + tap
                    mbar_wait(&b_empty[bs], bphase ^ 1);
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(&b_full[bs], 2 * CW_B_BYTES);
                        tma_load_2d_2cta(&tma_b, &b_full[bs], sBring + bs * CW_B_BYTES, j * p.conv_C + cs * BLOCK_K, n0);
                    }
                    __syncwarp();
                    if (++bs == CW_BSTAGES) { bs = 0; bphase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {   // ===== MMA issuer (leader CTA) =====
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BLOCK_N >> 3) << 17) | (uint32_t((2 * BLOCK_M) >> 4) << 24);
            int bs = 0, ab = 0, acc = 0;
            uint32_t bphase = 0, aphase = 0, acc_phase = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
                if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[(tile / n_clusters) * 8 + 0] = clock64();
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[(tile / n_clusters) * 8 + 1] = clock64();
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * BLOCK_N);
                for (int cs = 0; cs < p.conv_cpt; ++cs) {
                    mbar_wait(&a_full[ab], aphase);
                    tcgen05_fence_after();
                    const uint32_t a_base = smem_u32(sAraw + ab * CW_A_BYTES);
                    for (int j = 0; j < chunks_per_slice; ++j) {
                        const int tap = j % 9, kh = tap / 3, kw = tap - kh * 3;
                        const uint32_t a_tap = a_base + uint32_t(kh * 20 + kw) * 128u;        // [h][frame][w] rows of 128 B
                        mbar_wait(&b_full[bs], bphase);
                        tcgen05_fence_after();
                        const uint32_t b_base = smem_u32(sBring + bs * CW_B_BYTES);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                // the start is NOT on the 1024-byte swizzle repeat; the hardware swizzle is a function of the absolute
                                // shared-memory address (like TMA's), so the descriptor's base-offset field stays 0 (measured: setting it
                                // to (addr >> 7) & 7 gives wrong results, tools/bench_conv.py)
                                const uint64_t da = make_smem_desc(a_tap + k * 32, 16, 1280, 2);
                                const uint64_t db = make_smem_desc(b_base + k * 32, 16, 1024, 2);
                                umma_tf32_2cta(d_tmem, da, db, idesc, (cs > 0 || j > 0 || k > 0) ? 1u : 0u);
                            }
                            umma_commit_2cta(&b_empty[bs]);
                            if (j == chunks_per_slice - 1) umma_commit_2cta(&a_empty[ab]);
                        }
                        __syncwarp();
                        if (++bs == CW_BSTAGES) { bs = 0; bphase ^= 1; }
                    }
                    if (++ab == 2) { ab = 0; aphase ^= 1; }
                }
                if (elect_one()) umma_commit_2cta(&tmem_full[acc]);
                __syncwarp();
                if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[(tile / n_clusters) * 8 + 2] = clock64();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {   // ===== epilogue warps 2..5 of both CTAs =====
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
            const int n_tile = tile % p.n_tiles, m_pair = tile / p.n_tiles;
            const int m_base = (m_pair * 2 + (int)rank) * BLOCK_M + q * 32;
            float* sT = epi_tiles + (warp - 2) * 32 * EPI_PITCH;
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BLOCK_N);
            epilogue_tile<BLOCK_N, true>(p, taddr, sT, lane, m_base, n_tile * BLOCK_N, [&] {
                mbar_wait(&tmem_full[acc], acc_phase);
                tcgen05_fence_after();
            });
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// =============================================================================================
// The same raw-tile convolution with BOTH operands split into two bf16 planes (x = x_hi + x_lo, w = w_hi + w_lo, each rounded to
// nearest bf16) and three kind::f16 MMAs per product -- x_hi w_hi + x_lo w_hi + x_hi w_lo, fp32 accumulation; the dropped
// x_lo w_lo term is ~2^-16 relative.  Against the two-plane TF32 form above (x rounded to tf32, w = hi + lo in tf32) this
//   * removes the activation rounding (2^-11 -> 2^-16 per product: the frozen encoder's features land ~10x closer to fp32), and
//   * executes 3 bf16 passes = 1.5 TF32-pass equivalents instead of 2 (a kind::f16 instruction contracts 16 elements in the
//     time a kind::tf32 one contracts 8): 25 % less tensor-pipe time on a kernel that is 88-89 % tensor-bound.
// A 128-byte swizzle row holds 64 bf16 channels, so a channel slice is 64 wide and needs two raw tiles (hi, lo planes; the planes are
// stacked along the frame dimension of the padded buffer, [2][F][10][10][C] bf16).  Weight chunk j = tap * 2 + plane: the hi chunk
// serves two MMA groups (x_hi, x_lo), the lo chunk one (x_hi).
constexpr int CB_BSTAGES = 8;
constexpr int CB_SMEM_BYTES = 4 * CW_A_BYTES + CB_BSTAGES * CW_B_BYTES + 256 + 4 * 32 * EPI_PITCH * 4 + 1024;
static_assert(CB_SMEM_BYTES <= 227 * 1024, "shared memory budget");
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv3x3_w8_bf16x3_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p, int frames_total) {
    constexpr int BLOCK_N = 176;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sAraw = smem;                                  // [2 slots][hi, lo] raw tiles
    uint8_t* sBring = smem + 4 * CW_A_BYTES;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(sBring + CB_BSTAGES * CW_B_BYTES);
    uint64_t* b_empty = b_full + CB_BSTAGES;
    uint64_t* a_full = b_empty + CB_BSTAGES;
    uint64_t* a_empty = a_full + 2;
    uint64_t* tmem_full = a_empty + 2;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* epi_tiles = reinterpret_cast<float*>(sBring + CB_BSTAGES * CW_B_BYTES + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int n_clusters = gridDim.x >> 1;
    const int total_tiles = p.m_tiles * p.n_tiles;
    constexpr int SLICE = 64;                               // bf16 channels per 128-byte row

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
        for (int s = 0; s < CB_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&a_full[a], 1);
            mbar_init(&a_empty[a], 1);
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {   // ===== TMA producer (both CTAs) =====
        int bs = 0, ab = 0;
        uint32_t bphase = 0, aphase = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
            const int n_tile = tile % p.n_tiles, m_pair = tile / p.n_tiles;
            const int f0 = (m_pair * 2 + (int)rank) * 2;
            const int n0 = n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2);
            for (int cs = 0; cs < p.conv_cpt; ++cs) {
                mbar_wait(&a_empty[ab], aphase ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(&a_full[ab], 4 * CW_A_BYTES);      // hi + lo tiles of both CTAs
                    tma_load_4d_2cta(&tma_a, &a_full[ab], sAraw + (2 * ab) * CW_A_BYTES, cs * SLICE, 0, f0, 0);
                    tma_load_4d_2cta(&tma_a, &a_full[ab], sAraw + (2 * ab + 1) * CW_A_BYTES, cs * SLICE, 0, frames_total + f0, 0);
                }
                __syncwarp();
                if (++ab == 2) { ab = 0; aphase ^= 1; }
                for (int j = 0; j < 18; ++j) {                                       // j = tap * 2 + weight plane
                    mbar_wait(&b_empty[bs], bphase ^ 1);
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(&b_full[bs], 2 * CW_B_BYTES);
                        tma_load_2d_2cta(&tma_b, &b_full[bs], sBring + bs * CW_B_BYTES, ((j & 1) * 9 + (j >> 1)) * p.conv_C + cs * SLICE, n0);
                    }
                    __syncwarp();
                    if (++bs == CB_BSTAGES) { bs = 0; bphase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {   // ===== MMA issuer (leader CTA) =====
            // kind::f16 descriptor: fp32 accumulate (bit 4), A and B formats BF16 (1), K-major both, N, M = 256
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(BLOCK_N >> 3) << 17) | (uint32_t((2 * BLOCK_M) >> 4) << 24);
            int bs = 0, ab = 0, acc = 0;
            uint32_t bphase = 0, aphase = 0, acc_phase = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * BLOCK_N);
                for (int cs = 0; cs < p.conv_cpt; ++cs) {
                    mbar_wait(&a_full[ab], aphase);
                    tcgen05_fence_after();
                    const uint32_t a_hi = smem_u32(sAraw + (2 * ab) * CW_A_BYTES), a_lo = a_hi + CW_A_BYTES;
                    for (int j = 0; j < 18; ++j) {
                        const int tap = j >> 1, kh = tap / 3, kw = tap - kh * 3;
                        const uint32_t tap_off = uint32_t(kh * 20 + kw) * 128u;                 // [h][frame][w] rows of 128 B
                        mbar_wait(&b_full[bs], bphase);
                        tcgen05_fence_after();
                        const uint32_t b_base = smem_u32(sBring + bs * CW_B_BYTES);
                        if (elect_one()) {
                            const int groups = (j & 1) ? 1 : 2;       // w_hi: x_hi and x_lo; w_lo: x_hi only
                            for (int gq = 0; gq < groups; ++gq) {
                                const uint32_t a_tap = (gq ? a_lo : a_hi) + tap_off;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {         // 4 x 16 bf16 = one 128-byte row
                                    const uint64_t da = make_smem_desc(a_tap + k * 32, 16, 1280, 2);
                                    const uint64_t db = make_smem_desc(b_base + k * 32, 16, 1024, 2);
                                    umma_bf16_2cta(d_tmem, da, db, idesc, (cs > 0 || j > 0 || gq > 0 || k > 0) ? 1u : 0u);
                                }
                            }
                            umma_commit_2cta(&b_empty[bs]);
                            if (j == 17) umma_commit_2cta(&a_empty[ab]);
                        }
                        __syncwarp();
                        if (++bs == CB_BSTAGES) { bs = 0; bphase ^= 1; }
                    }
                    if (++ab == 2) { ab = 0; aphase ^= 1; }
                }
                if (elect_one()) umma_commit_2cta(&tmem_full[acc]);
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {   // ===== epilogue warps 2..5 of both CTAs =====
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
            const int n_tile = tile % p.n_tiles, m_pair = tile / p.n_tiles;
            const int m_base = (m_pair * 2 + (int)rank) * BLOCK_M + q * 32;
            float* sT = epi_tiles + (warp - 2) * 32 * EPI_PITCH;
            const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BLOCK_N);
            epilogue_tile<BLOCK_N, true>(p, taddr, sT, lane, m_base, n_tile * BLOCK_N, [&] {
                mbar_wait(&tmem_full[acc], acc_phase);
                tcgen05_fence_after();
            });
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 tensor map: `inner` contiguous elements, `outer` rows `pitch` elements apart.
int make_map_2d(CUtensorMap* map, const float* ptr, long long inner, long long outer, long long pitch, int box_inner, int box_outer,
                CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = get_encode_fn();
    VPTR_REQUIRE(enc != nullptr, VPTR_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPTR_REQUIRE(r == CUDA_SUCCESS, VPTR_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d): inner=%lld outer=%lld pitch=%lld box=%dx%d ptr=%p",
                 (int)r, inner, outer, pitch, box_inner, box_outer, (const void*)ptr);
    return VPTR_OK;
}

// 4-D fp32 tensor map over a padded NHWC activation [F][Hp][Wp][C]: box {32 channels, bw, bh, bf}
int make_map_nhwc(CUtensorMap* map, const float* ptr, long long F, long long Hp, long long Wp, long long C, int bw, int bh, int bf) {
    EncodeTiledFn enc = get_encode_fn();
    VPTR_REQUIRE(enc != nullptr, VPTR_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)F};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)Wp * C * 4, (cuuint64_t)Hp * Wp * C * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bf};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPTR_REQUIRE(r == CUDA_SUCCESS, VPTR_ERR_DRIVER, "cuTensorMapEncodeTiled(4d) failed (%d): F=%lld Hp=%lld Wp=%lld C=%lld box=%dx%dx%d", (int)r,
                 F, Hp, Wp, C, bw, bh, bf);
    return VPTR_OK;
}
// W == 8 raw-tile map: dims ordered (c, w, frame, h) so the box {32, 10, 2, 10} lands in shared memory as [h][frame][w][32 ch]
int make_map_nhwc_w8(CUtensorMap* map, const float* ptr, long long F, long long C) {
    EncodeTiledFn enc = get_encode_fn();
    VPTR_REQUIRE(enc != nullptr, VPTR_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, 10, (cuuint64_t)F, 10};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)100 * C * 4, (cuuint64_t)10 * C * 4};
    cuuint32_t box[4] = {32, 10, 2, 10};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPTR_REQUIRE(r == CUDA_SUCCESS, VPTR_ERR_DRIVER, "cuTensorMapEncodeTiled(w8) failed (%d): F=%lld C=%lld", (int)r, F, C);
    return VPTR_OK;
}

// bf16 raw-tile map: [2 planes * F][10][10][C] bf16 viewed as (c, w, frame, h); box {64, 10, 2, 10}
int make_map_nhwc_w8_bf16(CUtensorMap* map, const void* ptr, long long F2, long long C) {
    EncodeTiledFn enc = get_encode_fn();
    VPTR_REQUIRE(enc != nullptr, VPTR_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, 10, (cuuint64_t)F2, 10};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)100 * C * 2, (cuuint64_t)10 * C * 2};
    cuuint32_t box[4] = {64, 10, 2, 10};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPTR_REQUIRE(r == CUDA_SUCCESS, VPTR_ERR_DRIVER, "cuTensorMapEncodeTiled(w8 bf16) failed (%d): F2=%lld C=%lld", (int)r, F2, C);
    return VPTR_OK;
}
int make_map_2d_bf16(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long pitch, int box_inner, int box_outer) {
    EncodeTiledFn enc = get_encode_fn();
    VPTR_REQUIRE(enc != nullptr, VPTR_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VPTR_REQUIRE(r == CUDA_SUCCESS, VPTR_ERR_DRIVER, "cuTensorMapEncodeTiled(2d bf16) failed (%d): inner=%lld outer=%lld", (int)r, inner, outer);
    return VPTR_OK;
}

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int BLOCK_N, int A_MN, int B_MN, int STAGES>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<BLOCK_N, A_MN, B_MN, STAGES>;
    auto kern = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        VPTR_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
        attr_set = true;
    }
    int total = p.m_tiles * p.n_tiles * p.k_splits;
    int grid = total < num_sms() ? total : num_sms();
    kern<<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(ma, mb, p);
    return vptr_check_launch("gemm_tf32_kernel");
}


struct EpiMaps { CUtensorMap d, d16, r, r16; };
template <int BLOCK_N, int A_MN, int B_MN, int STAGES, int EPI_TILES = 2>
int launch_gemm_2cta(const CUtensorMap& ma, const CUtensorMap& mb, const EpiMaps& em, const GemmParams& p, cudaStream_t stream) {
    using Cfg = Gemm2Cfg<BLOCK_N, A_MN, B_MN, STAGES, EPI_TILES>;
    auto kern = gemm_tf32_2cta_kernel<BLOCK_N, A_MN, B_MN, STAGES, EPI_TILES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        VPTR_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
        attr_set = true;
    }
    int total = p.m_tiles * p.n_tiles * p.k_splits;
    int clusters = num_sms() / 2;
    if (total < clusters) clusters = total;
    kern<<<2 * clusters, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(ma, mb, em.d, em.d16, em.r, em.r16, p);
    return vptr_check_launch("gemm_tf32_2cta_kernel");
}

}  // namespace

static long long* g_gemm_dbg = nullptr;
// debug: device buffer (>= 8 * tiles-per-CTA int64) that block 0 fills with clock64() stamps; NULL disables
extern "C" int vptr_gemm_debug_buffer(long long* buf) { g_gemm_dbg = buf; return VPTR_OK; }

extern "C" int vptr_gemm_tf32(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, float* D,
                              long long ldd, int M, int N, int K, const float* bias, const float* residual, long long ldr,
                              float alpha, int act, int flags, int k_splits, const float* rowscale, int rows_per_group,
                              unsigned long long drop_seed, float drop_p, cudaStream_t stream) {
    VPTR_REQUIRE(M > 0 && N > 0 && K > 0, VPTR_ERR_SHAPE, "vptr_gemm_tf32: empty problem M=%d N=%d K=%d", M, N, K);
    VPTR_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, VPTR_ERR_ALIGN, "vptr_gemm_tf32: operand pitches must be multiples of 4 floats (TMA 16 B rule): lda=%lld ldb=%lld", lda, ldb);
    VPTR_REQUIRE(N % 4 == 0 && ldd % 4 == 0 && (residual == nullptr || ldr % 4 == 0), VPTR_ERR_ALIGN,
                 "vptr_gemm_tf32: N, ldd, ldr must be multiples of 4 (N=%d ldd=%lld ldr=%lld)", N, ldd, ldr);
    VPTR_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)D % 16 == 0) &&
                     ((uintptr_t)bias % 16 == 0) && ((uintptr_t)residual % 16 == 0),
                 VPTR_ERR_ALIGN, "vptr_gemm_tf32: pointers must be 16-byte aligned");
    VPTR_REQUIRE(!(flags & 1) || (bias == nullptr && residual == nullptr && act == 0 && rowscale == nullptr && drop_p <= 0.f),
                 VPTR_ERR_UNSUPPORTED, "vptr_gemm_tf32: atomic accumulate excludes bias/residual/activation/dropout");
    VPTR_REQUIRE(rowscale == nullptr || rows_per_group > 0, VPTR_ERR_SHAPE, "vptr_gemm_tf32: rowscale needs rows_per_group > 0");
    static const int mode_env = [] { const char* e = getenv("VPTR_GEMM_1CTA"); return (e && e[0] == '1') ? 1 : 0; }();
    const bool two_cta = !mode_env && M > BLOCK_M;          // CTA pairs (cta_group::2) unless the problem has a single M tile
    constexpr int BN1 = 176;                                 // 1-CTA N tile
    // wide outputs (fc1 / linear1, N = 2112): 256 x 256 pair tiles move 25 % fewer operand bytes per FLOP than 256 x 176 -- the TF32
    // main loop is operand-delivery bound -- when the column count tiles into 256 with <= 10 % waste (528 -> 2112: 181 -> 157 us)
    static const bool narrow_env = [] { const char* e = getenv("VPTR_GEMM_NARROW"); return e && e[0] == '1'; }();
    // (no gain measured for MN-major B, whose 192-column tiles already stage whole 32-column groups: dgrad 158 vs 160 us)
    const bool wide = !narrow_env && two_cta && !b_mn && N >= 1024 && (long long)vptr_cdiv(N, 256) * 256 * 10 <= (long long)N * 11;
    const int BN = two_cta ? (wide ? 256 : (b_mn ? 192 : 176)) : BN1;   // MN-major B halves: whole 32-column groups -> 192
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.m_tiles = vptr_cdiv(M, two_cta ? 2 * BLOCK_M : BLOCK_M);
    p.n_tiles = vptr_cdiv(N, BN);
    p.total_chunks = vptr_cdiv(K, BLOCK_K);
    if (!(flags & 1)) k_splits = 1;
    const int workers = two_cta ? num_sms() / 2 : num_sms();
    if (k_splits <= 0) {  // auto split-K for accumulate mode: fill exactly two rounds of the persistent grid
        // (rounding UP gave e.g. 9 tiles x 17 splits = 153 items on 74 CTA pairs = 2.07 waves, i.e. a third, nearly empty round:
        //  every weight-gradient GEMM of the path ran 3 rounds instead of 2)
        int tiles = p.m_tiles * p.n_tiles;
        k_splits = (2 * workers) / tiles;
        int max_splits = p.total_chunks / 8 > 0 ? p.total_chunks / 8 : 1;  // keep >= 8 chunks per split
        if (k_splits > max_splits) k_splits = max_splits;
        if (k_splits < 1) k_splits = 1;
    }
    if (k_splits > p.total_chunks) k_splits = p.total_chunks;
    p.chunks_per_split = vptr_cdiv(p.total_chunks, k_splits);
    p.k_splits = vptr_cdiv(p.total_chunks, p.chunks_per_split);  // no empty splits
    p.D = D; p.ldd = ldd; p.bias = bias; p.residual = residual; p.ldr = ldr;
    p.alpha = alpha; p.act = act; p.flags = flags;
    p.rowscale = rowscale; p.rows_per_group = rows_per_group; p.drop_seed = drop_seed; p.drop_p = drop_p;
    p.dbg = g_gemm_dbg;
    p.conv_taps = 0; p.conv_cpt = 1; p.conv_kw = 1; p.conv_bh = 1; p.conv_tiles_per_frame = 1; p.conv_bf = 1; p.conv_C = 0;
    p.conv_w8 = 0; p.conv_qw = 0; p.conv_qh = 0; p.conv_planes = 1; p.epi_tma = 0;

    CUtensorMap ma, mb;
    int rc;
    if (!a_mn) rc = make_map_2d(&ma, A, K, M, lda, BLOCK_K, BLOCK_M, CU_TENSOR_MAP_SWIZZLE_128B);
    else rc = make_map_2d(&ma, A, M, K, lda, 32, BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    if (!b_mn) rc = make_map_2d(&mb, B, K, N, ldb, BLOCK_K, two_cta ? BN / 2 : BN, CU_TENSOR_MAP_SWIZZLE_128B);
    else rc = make_map_2d(&mb, B, N, K, ldb, 32, BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;

    if (two_cta) {
        constexpr int ST2 = 7;
        constexpr int ST2_MN = 6;       // 192-column MN-major B stages are 28 KB: six of them beside the 32 KB of epilogue staging tiles
        // Outputs that are not split-K reductions leave through TMA bulk stores (epilogue_tile_tma).  With a residual operand and
        // K <= 1280 the residual is bulk-loaded as well: K <= 640 with four staging tiles per warp (loads two chunks ahead) and five
        // operand stages, longer K with three tiles and six stages; beyond that the main loop hides the per-lane epilogue and its
        // seventh operand stage is worth more (measured at M = 40960, N = 528: K = 528 91 -> 71 us, K = 1056 99 -> 91 us,
        // K = 2112 145 us per-lane against 150-160 us).
        static const bool no_tma_epi = [] { const char* e = getenv("VPTR_GEMM_EPI_STG"); return e && e[0] == '1'; }();
        const bool res_cfg = residual != nullptr && !wide && !a_mn && p.total_chunks <= 40;
        p.epi_tma = (no_tma_epi || (flags & 1)) ? 0 : (residual == nullptr ? 1 : (res_cfg ? 2 : 0));
        EpiMaps em{ma, ma, ma, ma};
        if (p.epi_tma) {
            rc = make_map_2d(&em.d, D, N, M, ldd, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc) return rc;
            rc = make_map_2d(&em.d16, D, N, M, ldd, 16, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc) return rc;
        }
        if (p.epi_tma == 2) {
            rc = make_map_2d(&em.r, residual, N, M, ldr, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc) return rc;
            rc = make_map_2d(&em.r16, residual, N, M, ldr, 16, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc) return rc;
        }
        if (wide) {
            if (!a_mn) return launch_gemm_2cta<256, 0, 0, 6>(ma, mb, em, p, stream);
            return launch_gemm_2cta<256, 1, 0, 6>(ma, mb, em, p, stream);
        }
        if (p.epi_tma == 2 && p.total_chunks <= 20) {      // short K: epilogue-bound, residual loads two chunks ahead
            if (!b_mn) return launch_gemm_2cta<176, 0, 0, 5, 4>(ma, mb, em, p, stream);
            return launch_gemm_2cta<192, 0, 1, 5, 4>(ma, mb, em, p, stream);
        }
        if (p.epi_tma == 2) {
            if (!b_mn) return launch_gemm_2cta<176, 0, 0, 6, 3>(ma, mb, em, p, stream);
            return launch_gemm_2cta<192, 0, 1, 6, 3>(ma, mb, em, p, stream);
        }
        if (!a_mn && !b_mn) return launch_gemm_2cta<176, 0, 0, ST2>(ma, mb, em, p, stream);
        if (!a_mn && b_mn) return launch_gemm_2cta<192, 0, 1, ST2_MN>(ma, mb, em, p, stream);
        if (a_mn && b_mn) return launch_gemm_2cta<192, 1, 1, ST2_MN>(ma, mb, em, p, stream);
        return launch_gemm_2cta<176, 1, 0, ST2>(ma, mb, em, p, stream);
    }
    constexpr int ST = 5;
    if (!a_mn && !b_mn) return launch_gemm<BN1, 0, 0, ST>(ma, mb, p, stream);
    if (!a_mn && b_mn) return launch_gemm<BN1, 0, 1, ST>(ma, mb, p, stream);
    if (a_mn && b_mn) return launch_gemm<BN1, 1, 1, ST>(ma, mb, p, stream);
    return launch_gemm<BN1, 1, 0, ST>(ma, mb, p, stream);
}

// The 3x3 stride-1 convolution of vptr_conv3x3_tf32 (below) for H, W multiples of 8 larger than 8 (the 16x16 grid of 128x128
// frames; reference model/ResNetAutoEncoder.py:138,151 at n_downsampling 3), on the raw-tile kernel: xq is the
// QUADRANT-tiled padded activation from vptr_pad_nhwc_quad -- every 8x8 quadrant of a frame with its own 1-pixel halo,
// [F * (H/8) * (W/8)][10][10][C] -- so each quadrant is an 8x8 "frame" of conv3x3_w8_kernel and only the epilogue's row mapping
// (epi_row) knows about the larger frame.  23 % more padded bytes than one (H+2)x(W+2) copy, but one TMA box per channel slice
// serves all 9 taps: the generic per-(tap, slice) path ran cfg4's encoder convs at 329 TFLOP/s executed against 848 on the 8x8 grid.
extern "C" int vptr_conv3x3_tf32_quad(const float* xq, const float* w, float* out, int F, int H, int W, int C, int Cout, const float* bias,
                                      const float* residual, int act, int flags, int w_planes, cudaStream_t stream) {
    VPTR_REQUIRE(w_planes == 1 || w_planes == 2, VPTR_ERR_SHAPE, "vptr_conv3x3_tf32_quad: w_planes must be 1 or 2");
    VPTR_REQUIRE(F > 0 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0 && C > 0 && Cout > 0, VPTR_ERR_SHAPE,
                 "vptr_conv3x3_tf32_quad: F=%d H=%d W=%d (H, W multiples of 8)", F, H, W);
    VPTR_REQUIRE(C % 4 == 0 && Cout % 4 == 0 && ((uintptr_t)xq % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                     ((uintptr_t)bias % 16 == 0) && ((uintptr_t)residual % 16 == 0),
                 VPTR_ERR_ALIGN, "vptr_conv3x3_tf32_quad: channels must be multiples of 4 and pointers 16-byte aligned");
    VPTR_REQUIRE(!(flags & 1), VPTR_ERR_UNSUPPORTED, "vptr_conv3x3_tf32_quad: accumulate mode not supported");
    VPTR_REQUIRE(act == 0 || act == 2, VPTR_ERR_UNSUPPORTED, "vptr_conv3x3_tf32_quad: act %d (the raw-tile epilogue has none / ReLU only)", act);
    const long long FQ = (long long)F * (H / 8) * (W / 8);
    VPTR_REQUIRE(FQ * 64 < 0x7fffffffLL, VPTR_ERR_SHAPE, "vptr_conv3x3_tf32_quad: too many rows");
    GemmParams p;
    p.M = (int)((long long)F * H * W); p.N = Cout; p.K = 9 * C;
    p.m_tiles = (int)((FQ + 3) / 4);              // pair tile = 4 quadrants = 256 output pixels
    p.n_tiles = vptr_cdiv(Cout, 176);
    p.conv_cpt = vptr_cdiv(C, BLOCK_K);
    p.total_chunks = w_planes * 9 * p.conv_cpt;
    p.k_splits = 1; p.chunks_per_split = p.total_chunks;
    p.D = out; p.ldd = Cout; p.bias = bias; p.residual = residual; p.ldr = Cout;
    p.alpha = 1.f; p.act = act; p.flags = flags;
    p.rowscale = nullptr; p.rows_per_group = 1; p.drop_seed = 0; p.drop_p = 0.f;
    p.dbg = g_gemm_dbg;
    p.conv_taps = 9; p.conv_kw = 3; p.conv_bh = 8; p.conv_tiles_per_frame = 1; p.conv_bf = 2; p.conv_C = C;
    p.conv_w8 = 1; p.conv_qw = W / 8; p.conv_qh = H / 8; p.conv_planes = w_planes; p.epi_tma = 0;
    CUtensorMap ma, mb;
    int rc = make_map_nhwc_w8(&ma, xq, FQ, C);
    if (rc) return rc;
    rc = make_map_2d(&mb, w, (long long)w_planes * 9 * C, Cout, (long long)w_planes * 9 * C, BLOCK_K, 176 / 2, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_w8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM_BYTES);
        VPTR_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(conv3x3_w8, smem=%d): %s", CW_SMEM_BYTES, cudaGetErrorString(e));
        attr_set = true;
    }
    const int total = p.m_tiles * p.n_tiles;
    int clusters = num_sms() / 2;
    if (total < clusters) clusters = total;
    conv3x3_w8_kernel<<<2 * clusters, NUM_THREADS, CW_SMEM_BYTES, stream>>>(ma, mb, p);
    return vptr_check_launch("conv3x3_w8_kernel(quad)");
}

// 3x3 stride-1 convolution, H and W multiples of 8, on the bf16x3 raw-tile kernel (conv3x3_w8_bf16x3_kernel): xq2 = the two bf16
// planes of the quadrant-tiled padded activation from vptr_pad_nhwc_quad_bf16x2, [2][F*(H/8)*(W/8)][10][10][C] bf16; w2 = the two
// bf16 planes of the packed weights from vptr_split_bf16x2, [Cout][2][9*C] bf16.  fp32 output, same epilogue options as
// vptr_conv3x3_tf32_quad (bias, ReLU, residual, tf32 rounding of the stored values).  C % 8 == 0 (16-byte TMA rows).
// Replaces nn.Conv2d(k3,s1) + folded eval BatchNorm of the ResnetBlocks (reference model/ResNetAutoEncoder.py:138,151).
extern "C" int vptr_conv3x3_bf16x3(const void* xq2, const void* w2, float* out, int F, int H, int W, int C, int Cout, const float* bias,
                                   const float* residual, int act, int flags, cudaStream_t stream) {
    VPTR_REQUIRE(F > 0 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0 && C > 0 && Cout > 0, VPTR_ERR_SHAPE,
                 "vptr_conv3x3_bf16x3: F=%d H=%d W=%d (H, W multiples of 8)", F, H, W);
    VPTR_REQUIRE(C % 8 == 0 && Cout % 4 == 0 && ((uintptr_t)xq2 % 16 == 0) && ((uintptr_t)w2 % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                     ((uintptr_t)bias % 16 == 0) && ((uintptr_t)residual % 16 == 0),
                 VPTR_ERR_ALIGN, "vptr_conv3x3_bf16x3: C %% 8, Cout %% 4 and 16-byte aligned pointers required");
    VPTR_REQUIRE(!(flags & 1), VPTR_ERR_UNSUPPORTED, "vptr_conv3x3_bf16x3: accumulate mode not supported");
    VPTR_REQUIRE(act == 0 || act == 2, VPTR_ERR_UNSUPPORTED, "vptr_conv3x3_bf16x3: act %d (the raw-tile epilogue has none / ReLU only)", act);
    const long long FQ = (long long)F * (H / 8) * (W / 8);
    VPTR_REQUIRE(FQ * 64 < 0x3fffffffLL, VPTR_ERR_SHAPE, "vptr_conv3x3_bf16x3: too many rows");
    GemmParams p;
    p.M = (int)((long long)F * H * W); p.N = Cout; p.K = 9 * C;
    p.m_tiles = (int)((FQ + 3) / 4);
    p.n_tiles = vptr_cdiv(Cout, 176);
    p.conv_cpt = vptr_cdiv(C, 64);
    p.total_chunks = 18 * p.conv_cpt;
    p.k_splits = 1; p.chunks_per_split = p.total_chunks;
    p.D = out; p.ldd = Cout; p.bias = bias; p.residual = residual; p.ldr = Cout;
    p.alpha = 1.f; p.act = act; p.flags = flags;
    p.rowscale = nullptr; p.rows_per_group = 1; p.drop_seed = 0; p.drop_p = 0.f;
    p.dbg = nullptr;
    p.conv_taps = 9; p.conv_kw = 3; p.conv_bh = 8; p.conv_tiles_per_frame = 1; p.conv_bf = 2; p.conv_C = C;
    p.conv_w8 = 1; p.conv_planes = 2; p.epi_tma = 0;
    p.conv_qw = (H == 8 && W == 8) ? 0 : W / 8; p.conv_qh = (H == 8 && W == 8) ? 0 : H / 8;
    CUtensorMap ma, mb;
    int rc = make_map_nhwc_w8_bf16(&ma, xq2, 2 * FQ, C);
    if (rc) return rc;
    rc = make_map_2d_bf16(&mb, w2, 2LL * 9 * C, Cout, 2LL * 9 * C, 64, 176 / 2);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_w8_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM_BYTES);
        VPTR_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(conv3x3_w8_bf16x3, smem=%d): %s", CB_SMEM_BYTES, cudaGetErrorString(e));
        attr_set = true;
    }
    const int total = p.m_tiles * p.n_tiles;
    int clusters = num_sms() / 2;
    if (total < clusters) clusters = total;
    conv3x3_w8_bf16x3_kernel<<<2 * clusters, NUM_THREADS, CB_SMEM_BYTES, stream>>>(ma, mb, p, (int)FQ);
    return vptr_check_launch("conv3x3_w8_bf16x3_kernel");
}

// Implicit-GEMM 3x3 stride-1 convolution on the tcgen05 kernel (ResnetBlock convs, reference model/ResNetAutoEncoder.py:138,151):
//   out[(f,oh,ow)][co] = act( sum_{kh,kw,ci} xpad[f][oh+kh][ow+kw][ci] * w[co][(kh,kw,ci)] + bias[co] ) (+ residual)
// xpad: NHWC activation already padded by 1 (zero / reflect / replicate: vptr_pad_nhwc) [F][H+2][W+2][C]; w: [Cout][9*C]
// (vptr_pack_conv_weight mode 0).  No im2col matrix is materialised: each k-chunk's A tile is a 4-D TMA box of xpad.
// Returns VPTR_ERR_UNSUPPORTED when the grid does not tile into 128-pixel boxes (caller falls back to im2col + vptr_gemm_tf32).
// w_planes = 2: w is [Cout][2][9*C] = tf32 hi part followed by the tf32 lo part of each weight (vptr_split_tf32): the contraction
// runs over both planes (2x the MMA work), which removes the weight-rounding half of the tf32 error (used by the frozen encoder,
// whose 21 chained convolutions otherwise land at 1.16e-3 relative, just outside the 1e-3 gate).
extern "C" int vptr_conv3x3_tf32(const float* xpad, const float* w, float* out, int F, int H, int W, int C, int Cout, const float* bias,
                                 const float* residual, int act, int flags, int w_planes, cudaStream_t stream) {
    VPTR_REQUIRE(w_planes == 1 || w_planes == 2, VPTR_ERR_SHAPE, "vptr_conv3x3_tf32: w_planes must be 1 or 2");
    VPTR_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0 && Cout > 0, VPTR_ERR_SHAPE, "vptr_conv3x3_tf32: empty problem");
    VPTR_REQUIRE(C % 4 == 0 && Cout % 4 == 0 && ((uintptr_t)xpad % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                     ((uintptr_t)bias % 16 == 0) && ((uintptr_t)residual % 16 == 0),
                 VPTR_ERR_ALIGN, "vptr_conv3x3_tf32: channels must be multiples of 4 and pointers 16-byte aligned");
    VPTR_REQUIRE(!(flags & 1), VPTR_ERR_UNSUPPORTED, "vptr_conv3x3_tf32: accumulate mode not supported");
    int bw = W, bh, bf, tiles_per_frame;
    const int px = H * W;
    if (px <= BLOCK_M) {
        if (BLOCK_M % px != 0 || W > 256 || H > 256) { vptr_set_error("vptr_conv3x3_tf32: %dx%d does not tile 128 pixels", H, W); return VPTR_ERR_UNSUPPORTED; }
        bh = H; bf = BLOCK_M / px; tiles_per_frame = 1;
    } else {
        if (BLOCK_M % W != 0 || H % (BLOCK_M / W) != 0) { vptr_set_error("vptr_conv3x3_tf32: %dx%d does not tile 128 pixels", H, W); return VPTR_ERR_UNSUPPORTED; }
        bh = BLOCK_M / W; bf = 1; tiles_per_frame = H / bh;
    }
    const long long M = (long long)F * px;
    const long long tiles = bf > 1 ? (F + bf - 1) / bf : (long long)F * tiles_per_frame;
    GemmParams p;
    p.M = (int)M; p.N = Cout; p.K = 9 * C;
    p.m_tiles = (int)((tiles + 1) / 2);
    p.n_tiles = vptr_cdiv(Cout, 176);
    p.conv_cpt = vptr_cdiv(C, BLOCK_K);
    p.total_chunks = w_planes * 9 * p.conv_cpt;
    p.k_splits = 1; p.chunks_per_split = p.total_chunks;
    p.D = out; p.ldd = Cout; p.bias = bias; p.residual = residual; p.ldr = Cout;
    p.alpha = 1.f; p.act = act; p.flags = flags;
    p.rowscale = nullptr; p.rows_per_group = 1; p.drop_seed = 0; p.drop_p = 0.f;
    p.dbg = g_gemm_dbg;
    p.conv_taps = 9; p.conv_kw = 3; p.conv_bh = bh; p.conv_tiles_per_frame = tiles_per_frame; p.conv_bf = bf; p.conv_C = C;
    p.conv_w8 = 0; p.conv_qw = 0; p.conv_qh = 0; p.conv_planes = w_planes; p.epi_tma = 0;
    // rows of a tile beyond F*H*W (frames past the end) are zero-filled by TMA and masked by the epilogue (m < M) only when tiles
    // map to whole frames in order, which holds for both tilings above.
    CUtensorMap ma, mb;
    static const bool generic_only = [] { const char* e = getenv("VPTR_CONV_GENERIC"); return e && e[0] == '1'; }();
    if (H == 8 && W == 8 && !generic_only && (act == 0 || act == 2)) {   // raw-tile kernel (its epilogue has no GELU): one TMA box per channel slice serves all 9 taps (and both planes)
        p.conv_w8 = 1; p.conv_planes = w_planes;
        p.m_tiles = vptr_cdiv(F, 4);           // pair tile = 4 frames = 256 output pixels
        int rc = make_map_nhwc_w8(&ma, xpad, F, C);
        if (rc) return rc;
        rc = make_map_2d(&mb, w, (long long)w_planes * 9 * C, Cout, (long long)w_planes * 9 * C, BLOCK_K, 176 / 2, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(conv3x3_w8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM_BYTES);
            VPTR_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(conv3x3_w8, smem=%d): %s", CW_SMEM_BYTES, cudaGetErrorString(e));
            attr_set = true;
        }
        const int total = p.m_tiles * p.n_tiles;
        int clusters = num_sms() / 2;
        if (total < clusters) clusters = total;
        conv3x3_w8_kernel<<<2 * clusters, NUM_THREADS, CW_SMEM_BYTES, stream>>>(ma, mb, p);
        return vptr_check_launch("conv3x3_w8_kernel");
    }
    int rc = make_map_nhwc(&ma, xpad, F, H + 2, W + 2, C, bw, bh, bf);
    if (rc) return rc;
    rc = make_map_2d(&mb, w, (long long)w_planes * 9 * C, Cout, (long long)w_planes * 9 * C, BLOCK_K, 176 / 2, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    return launch_gemm_2cta<176, 0, 0, 7>(ma, mb, EpiMaps{ma, ma, ma, ma}, p, stream);
}

VPTR_RNG_EPOCH_ACCESSOR(gemm_tcgen05)
