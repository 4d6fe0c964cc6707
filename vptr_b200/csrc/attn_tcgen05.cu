// Attention forward core on the 5th-generation tensor cores (tcgen05.mma kind::tf32, TMEM accumulators) fed by TMA.
//
// Replaces, for head_dim 66 and groups of <= 32 tokens, the QK^T + RPE -> softmax -> .V path of
//   SpatialLocalMultiheadAttention / MultiheadAttentionRPE   (model/VidHRFormer_modules.py:321-357, MultiHeadAttentionRPE.py:586-686)
//   the temporal / encoder-decoder nn.MultiheadAttention cores (model/VidHRFormer_modules.py:79-84,185-187,204-205).
//
// Work item = (tile of up to 128 token rows, head).  A tile is a set of WHOLE attention groups in their natural memory order:
//   window mode   : 128 consecutive tokens = whole frames (H*W <= 128) or whole bands of ws image rows -> 2-D TMA boxes;
//   temporal mode : P pixels x all T frames of one clip, rows ordered (pixel, t)                          -> 4-D TMA boxes.
// The window gather / scatter and the per-pixel temporal regrouping therefore cost nothing: TMA lands the rows where the
// tensor core wants them and attention structure becomes a MASK on the 128 x 128 score tile: S = Q K^T is computed for all
// row pairs (tcgen05.mma M=128, N=128, K=72), and a thread (= TMEM lane = query row) keeps only the columns of its own group
// (group id / in-group position per row and column are tile-independent and precomputed), adds the relative-position bias,
// applies the causal rule, and writes exp(s - max) as the K-major A operand of the second MMA, O = P V (N = 96, K = 128),
// whose result is normalised by 1/rowsum in the epilogue and leaves through a TMA store.
//
//   * head slices are 66 floats at 264-byte offsets, loaded as three 32-column boxes (96 columns, contraction over 72).  TMA
//     needs 16-byte aligned box starts, which an odd head (offset 264*h = 8 mod 16) does not have: its boxes start two floats
//     EARLY (66*h - 2).  Tile columns that belong to a neighbouring head (0..1 of an odd head, 66/68..71) are zeroed in the Q
//     tile by the issuing warp before the score MMA, so whatever K holds there contributes nothing; in V they only produce
//     output columns that are never stored.
//   * O leaves as two aligned 32-column TMA boxes (head columns 0..63 of an even head, 2..65 of an odd one) plus one 8-byte
//     generic store per row for the remaining pair: a clipped TMA STORE rewrites the whole 16-byte granule that the pair shares
//     with the neighbouring head (measured: it zeroed the neighbour's two columns), so no store box may end inside a granule.
//   * Q, K: K-major SWIZZLE_128B chunks of 32 floats; V: MN-major (SWIZZLE_128B_ATOM_32B) so the key index is the MMA's K.
//   * warp 4 issues TMA and MMA (one elected lane), warps 0-3 own TMEM lanes 0-127.  Q/K of item n+1 are prefetched as soon
//     as S(n) retires, V(n+1) as soon as the O(n) store has read the V buffer (which doubles as the O staging tile).
// Operands enter the tensor core as TF32: the producers round q / k / v to nearest tf32 (GEMM epilogue flag), P is rounded here.
// The result differs from the fp32 / 3xTF32 kernels by ~4e-4 relative (like any tf32 GEMM of the path); tests gate it at 2e-3
// against the fp32 oracle and the end-to-end parity gates stay at 1e-3.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int TC_D = 66;            // head dim
constexpr int TC_ROWS = 128;        // tile rows (UMMA M and N of the score MMA)
constexpr int TC_CHUNK = 16384;     // 128 rows x 128 B
constexpr int TC_THREADS = 160;
constexpr int TC_MAXHEADS = 8;

struct TcMaps {
    CUtensorMap q, k, v, o;   // ONE map per tensor: with a map per (tensor, head) the TMA unit re-fetched a descriptor for nearly
                              // every copy (32 maps cycling through its small descriptor cache) and an item took ~23 us
};
struct TcGeom {
    int mode;                 // 0 window, 1 temporal
    int nhead, Lq, Lk, causal, ws;
    int H, W, HW, nwh, nww;   // window mode
    int Tq, Tk, P, chunks_per_clip;   // temporal mode
    int tiles;                // number of tiles
    int wpt;                  // window mode: windows per tile
    int rows_q, rows_k;       // valid rows of a full tile
    float scale;
    int round_tf32;
    unsigned long long drop_seed;
    float drop_p;
    float* O;                 // output base / pitch / total rows for the two generic-store columns of odd heads
    long long ldo, total_rows_q;
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
// A protocol bug must not hang the GPU: after ~1e6 failed polls the wait records (barrier offset, parity, thread) in tc_err,
// raises a CTA-wide abort flag that makes every later wait fall through, and the kernel runs to completion with garbage output;
// the host reports tc_err (VPTR_ATTN_TC_DEBUG=1 checks it after every launch).
__device__ int tc_err[4];
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int code) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done || *abort_flag) break;
        if (++spins > (1u << 20)) {
            if (atomicCAS(&tc_err[0], 0, code) == 0) { tc_err[1] = (int)parity; tc_err[2] = (int)threadIdx.x; tc_err[3] = (int)blockIdx.x; }
            *abort_flag = 1;
            break;
        }
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor (sm_100): layout 2 = K-major SWIZZLE_128B, layout 1 = MN-major SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_128() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// (group id relative to the tile, position inside the group, validity) of tile row r -- identical for every tile
__device__ __forceinline__ void row_info(const TcGeom& g, int r, bool is_q, int& gid, int& pos, bool& valid) {
    if (g.mode == 0) {
        const int p = r % g.HW, f = r / g.HW;                 // tiles start on frame (HW <= 128) or band boundaries
        const int y = p / g.W, x = p - y * g.W;
        gid = (f * g.nwh + y / g.ws) * g.nww + x / g.ws;       // relative: band tiles have y < ws (p counts from the band start)
        pos = (y % g.ws) * g.ws + x % g.ws;
        valid = r < g.rows_q;
    } else {
        const int T = is_q ? g.Tq : g.Tk;
        gid = r / T;                 // rows ordered (pixel, t): a sequence's keys are T CONTIGUOUS columns of the score tile
        pos = r - gid * T;
        valid = r < T * g.P;
    }
}

// byte offset of element (row, col) of a 128-row K-major SWIZZLE_128B operand made of 32-float chunks
__device__ __forceinline__ uint32_t sw128_offset(int row, int col) {
    return (uint32_t)(col >> 5) * TC_CHUNK + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u +
           ((((uint32_t)(col & 31) >> 2) ^ (uint32_t)(row & 7)) << 4) + (uint32_t)(col & 3) * 4u;
}

// sequences a warp's 32 rows (tile rows 32w .. 32w+31, ordered (pixel, t)) can touch, maximum over the four warps
__host__ __device__ constexpr int seq_span(int TQ) {
    int m = 0;
    for (int w = 0; w < 4; ++w) {
        const int n = (32 * w + 31) / TQ - (32 * w) / TQ + 1;
        m = n > m ? n : m;
    }
    return m;
}
// FQ = FK = 0: window fast path (8x8 grid, 4x4 windows) or the generic mask path, chosen at run time.
// FQ, FK > 0 : temporal / enc-dec attention with compile-time query / key counts (T = 10 of cfg1, 29 of cfg2, 28 and 28 x 2 of
//              cfg3, 30 and 30 x 10 of cfg4): a row's keys are the FK contiguous columns of its own pixel sequence, picked out of a window of
//              seq_span(FQ) * FK columns the warp loads once from TMEM.
template <int FQ, int FK>
__global__ void __launch_bounds__(TC_THREADS, 1) attn_tc_fwd_kernel(const __grid_constant__ TcMaps maps, const float* __restrict__ rpe_table,
                                                                    const TcGeom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                       // 3 chunks, K-major SW128
    uint8_t* sK = sQ + 3 * TC_CHUNK;          // 3 chunks, K-major SW128
    uint8_t* sV = sK + 3 * TC_CHUNK;          // 3 groups, MN-major (also the O staging tile, K-major SW128, for the TMA store)
    uint8_t* sP = sV + 3 * TC_CHUNK;          // 4 chunks, K-major SW128
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * TC_CHUNK);
    uint64_t* qk_full = bars + 0;             // TMA: Q and K landed
    uint64_t* v_full = bars + 1;              // TMA: V landed
    uint64_t* s_ready = bars + 2;             // tcgen05.commit: S complete (Q, K buffers free)
    uint64_t* p_ready = bars + 3;             // 128 softmax threads: P written (and O TMEM drained)
    uint64_t* o_ready = bars + 4;             // tcgen05.commit: O complete
    uint64_t* v_free = bars + 5;              // O store has read the V / staging buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(bars + 9);
    unsigned char* colgid = reinterpret_cast<unsigned char*>(bars + 10);   // [128] group id of key column j
    unsigned char* colpos = colgid + 128;                                  // [128] in-group position (255 = invalid)
    float* sbias = reinterpret_cast<float*>(colpos + 128);                 // [nhead][Lq][Lk] (rpe only, Lq = Lk <= 16)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int items = g.tiles * g.nhead;

    // ---- one-time setup
    for (int e = threadIdx.x; e < (13 * TC_CHUNK) / 16; e += blockDim.x) reinterpret_cast<float4*>(smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x < 128) {
        int gid, pos; bool valid;
        row_info(g, threadIdx.x, false, gid, pos, valid);
        colgid[threadIdx.x] = (unsigned char)gid;
        colpos[threadIdx.x] = valid ? (unsigned char)pos : 255;
    }
    if (rpe_table) {
        for (int e = threadIdx.x; e < g.nhead * g.Lq * g.Lk; e += blockDim.x) {
            const int h = e / (g.Lq * g.Lk), r = e - h * g.Lq * g.Lk, i = r / g.Lk, j = r - i * g.Lk;
            const int ih = i / g.ws, iw = i - ih * g.ws, jh = j / g.ws, jw = j - jh * g.ws;
            sbias[e] = __ldg(rpe_table + ((ih - jh + g.ws - 1) * (2 * g.ws - 1) + (iw - jw + g.ws - 1)) * g.nhead + h);
        }
    }
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.v) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.o) : "memory");
        *abort_flag = 0;
        mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(s_ready, 1); mbar_init(p_ready, 128); mbar_init(o_ready, 1); mbar_init(v_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_async_smem();                      // the zero fill must be visible to the async proxy (MMA reads of never-loaded rows)
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

    if (warp == 4) {
        // ================= TMA + MMA issuer =================
        constexpr uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(128 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
        constexpr uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | (uint32_t(96 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
        auto load_qk = [&](int item) {
            const int tile = item / g.nhead, h = item - tile * g.nhead;
            if (elect_one()) {
                mbar_expect_tx(qk_full, (uint32_t)(g.rows_q + g.rows_k) * 3u * 128u);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (g.mode == 0) {
                        tma_load_2d(&maps.q, qk_full, sQ + c * TC_CHUNK, h * TC_D - 2 * (h & 1) + c * 32, tile * TC_ROWS);
                        tma_load_2d(&maps.k, qk_full, sK + c * TC_CHUNK, h * TC_D - 2 * (h & 1) + c * 32, tile * TC_ROWS);
                    } else {
                        const int n = tile / g.chunks_per_clip, p0 = (tile - n * g.chunks_per_clip) * g.P;
                        tma_load_4d(&maps.q, qk_full, sQ + c * TC_CHUNK, h * TC_D - 2 * (h & 1) + c * 32, 0, p0, n);
                        tma_load_4d(&maps.k, qk_full, sK + c * TC_CHUNK, h * TC_D - 2 * (h & 1) + c * 32, 0, p0, n);
                    }
                }
            }
            __syncwarp();
        };
        auto load_v = [&](int item) {
            const int tile = item / g.nhead, h = item - tile * g.nhead;
            if (elect_one()) {
                mbar_expect_tx(v_full, (uint32_t)g.rows_k * 3u * 128u);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (g.mode == 0) {
                        tma_load_2d(&maps.v, v_full, sV + c * TC_CHUNK, h * TC_D - 2 * (h & 1) + c * 32, tile * TC_ROWS);
                    } else {
                        const int n = tile / g.chunks_per_clip, p0 = (tile - n * g.chunks_per_clip) * g.P;
                        tma_load_4d(&maps.v, v_full, sV + c * TC_CHUNK, h * TC_D - 2 * (h & 1) + c * 32, 0, p0, n);
                    }
                }
            }
            __syncwarp();
        };
        uint32_t ph = 0;
        if ((int)blockIdx.x < items) { load_qk(blockIdx.x); load_v(blockIdx.x); }
        for (int item = blockIdx.x; item < items; item += gridDim.x, ph ^= 1) {
            const int next = item + gridDim.x;
            // ---- S = Q K^T : 9 k-steps of 8 over head_dim 72 (columns 66..71 are zero-filled by the clipped maps)
            mbar_wait(qk_full, ph, abort_flag, 1 + 10 * (warp == 4));
            {   // columns of the 72-wide contraction that do not belong to this head (the neighbours' data, or zero fill past the
                // tensor) are zeroed in the Q tile: even head -> 66..71; odd head (tile starts 2 columns early) -> 0..1 and 68..71
                const int odd_h = (item % g.nhead) & 1;
#pragma unroll
                for (int r = lane; r < TC_ROWS; r += 32) {
                    *reinterpret_cast<float2*>(sQ + sw128_offset(r, odd_h ? 0 : 66)) = make_float2(0.f, 0.f);
                    *reinterpret_cast<float4*>(sQ + sw128_offset(r, 68)) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                fence_async_smem();
                __syncwarp();
            }
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const uint32_t off = (uint32_t)(k >> 2) * TC_CHUNK + (uint32_t)(k & 3) * 32u;
                    umma_tf32(tmem_S, make_smem_desc(qa + off, 16, 1024, 2), make_smem_desc(ka + off, 16, 1024, 2), idesc_s, k > 0 ? 1u : 0u);
                }
                umma_commit(s_ready);
            }
            __syncwarp();
            // ---- Q, K buffers are free once S retired: prefetch the next item's
            mbar_wait(s_ready, ph, abort_flag, 2 + 10 * (warp == 4));
            if (next < items) load_qk(next);
            // ---- V of THIS item was requested one iteration ago (or in the prologue); the next one's after the O store below
            // ---- O = P V : 16 k-steps of 8 keys
            mbar_wait(p_ready, ph, abort_flag, 3 + 10 * (warp == 4));
            mbar_wait(v_full, ph, abort_flag, 4 + 10 * (warp == 4));
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint64_t da = make_smem_desc(pa + (uint32_t)(k >> 2) * TC_CHUNK + (uint32_t)(k & 3) * 32u, 16, 1024, 2);
                    const uint64_t db = make_smem_desc(va + (uint32_t)k * 1024u, TC_CHUNK, 512, 1);
                    umma_tf32(tmem_O, da, db, idesc_o, k > 0 ? 1u : 0u);
                }
                umma_commit(o_ready);
            }
            __syncwarp();
            // ---- the epilogue stages O over the V buffer and stores it; then V of the next item may land
            mbar_wait(v_free, ph, abort_flag, 6 + 10 * (warp == 4));
            if (next < items) load_v(next);
        }
    } else {
        // ================= softmax + epilogue: thread = query row = TMEM lane =================
        const int row = threadIdx.x;
        int gid_i, pos_i; bool valid_i;
        row_info(g, row, true, gid_i, pos_i, valid_i);
        // columns this row attends to, per 32-column chunk
        uint32_t vm[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t m = 0;
            for (int j = 0; j < 32; ++j) {
                const int col = c * 32 + j;
                const bool ok = valid_i && colpos[col] != 255 && colgid[col] == gid_i && (!g.causal || (int)colpos[col] <= pos_i);
                m |= (ok ? 1u : 0u) << j;
            }
            vm[c] = m;
        }
        uint32_t need = 0;   // chunks any row of this warp touches (warp-uniform): only those are read from TMEM and rewritten in P
#pragma unroll
        for (int c = 0; c < 4; ++c) need |= (__ballot_sync(0xffffffffu, vm[c] != 0) ? 1u : 0u) << c;
        const uint32_t lane_taddr = (uint32_t)(warp * 32) << 16;
        const float* brow = sbias + pos_i * g.Lk;             // + h * Lq * Lk per head
        // Fast path for THE window shape of the path (8x8 grid, 4x4 windows, tile = two frames): the 16 keys of a row's window are
        // four runs of four columns inside the warp's own 32-column chunk (image row r -> columns r*8 + xoff .. +3), so the whole
        // softmax is straight-line code on 16 registers -- no per-column bit tests, position look-ups or second TMEM pass.
        const bool fast_win = FQ == 0 && g.mode == 0 && g.HW == 64 && g.W == 8 && g.ws == 4;
        // sequence fast path (FQ > 0): first sequence of this warp's rows and my sequence's rank among them
        constexpr int SEQ_Q = FQ > 0 ? FQ : 1, SEQ_K = FK > 0 ? FK : 1;
        constexpr int NSEQ = seq_span(SEQ_Q), NLD = (NSEQ * SEQ_K + 31) / 32;
        const int seq_plo = (warp * 32) / SEQ_Q, seq_k = gid_i - seq_plo;
        const bool xhi = ((row & 7) >> 2) != 0;               // my window is the right-hand one of its image rows
        const uint32_t row_off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
        const uint32_t r7 = (uint32_t)(row & 7);
        uint32_t ph = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ph ^= 1) {
            const int tile = item / g.nhead, h = item - tile * g.nhead;
            // global batch index of this row (dropout stream shared with the other attention kernels)
            long long batch;
            if (g.mode == 0) batch = (long long)tile * g.wpt + gid_i;
            else { const int n = tile / g.chunks_per_clip; batch = (long long)n * g.HW + (tile - n * g.chunks_per_clip) * g.P + gid_i; }
            const unsigned long long drop_row = (((unsigned long long)batch * g.nhead + h) * g.Lq + pos_i) * g.Lk;
            mbar_wait(s_ready, ph, abort_flag, 2 + 10 * (warp == 4));
            tcgen05_fence_after();
            float sum = 0.f;
            if (fast_win) {
                float v[32], sc[16];
                tmem_ld32(tmem_S + lane_taddr + warp * 32, v);
                float4 bq[4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    bq[r] = rpe_table ? *reinterpret_cast<const float4*>(brow + h * 256 + r * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                tmem_ld_wait();
                float mx = -INFINITY;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    sc[r * 4 + 0] = fmaf(xhi ? v[r * 8 + 4] : v[r * 8 + 0], g.scale, bq[r].x);
                    sc[r * 4 + 1] = fmaf(xhi ? v[r * 8 + 5] : v[r * 8 + 1], g.scale, bq[r].y);
                    sc[r * 4 + 2] = fmaf(xhi ? v[r * 8 + 6] : v[r * 8 + 2], g.scale, bq[r].z);
                    sc[r * 4 + 3] = fmaf(xhi ? v[r * 8 + 7] : v[r * 8 + 3], g.scale, bq[r].w);
                    mx = fmaxf(fmaxf(mx, fmaxf(sc[r * 4], sc[r * 4 + 1])), fmaxf(sc[r * 4 + 2], sc[r * 4 + 3]));
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
                if (g.drop_p > 0.f) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {       // probabilities (row, 4r .. 4r+3): one hash per four (drop_row % 16 == 0)
                        const float4 k = vptr_drop_scale4(g.drop_seed, (drop_row >> 2) + r, g.drop_p);
                        sc[r * 4] *= k.x; sc[r * 4 + 1] *= k.y; sc[r * 4 + 2] *= k.z; sc[r * 4 + 3] *= k.w;
                    }
                }
                uint8_t* prow = sP + warp * TC_CHUNK + row_off;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 p4 = make_float4(vptr_round_tf32(sc[r * 4]), vptr_round_tf32(sc[r * 4 + 1]), vptr_round_tf32(sc[r * 4 + 2]),
                                                  vptr_round_tf32(sc[r * 4 + 3]));
                    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(prow + (((uint32_t)(2 * r) ^ r7) << 4)) = xhi ? z4 : p4;
                    *reinterpret_cast<float4*>(prow + (((uint32_t)(2 * r + 1) ^ r7) << 4)) = xhi ? p4 : z4;
                }
            } else if (FQ > 0) {
                float win[NLD * 32], sc[SEQ_K];
#pragma unroll
                for (int l = 0; l < NLD; ++l) tmem_ld32(tmem_S + lane_taddr + seq_plo * SEQ_K + l * 32, win + l * 32);
                tmem_ld_wait();
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < SEQ_K; ++j) {
                    float a = win[j];
#pragma unroll
                    for (int kk = 1; kk < NSEQ; ++kk)
                        if (seq_k == kk) a = win[kk * SEQ_K + j];
                    a *= g.scale;
                    if (g.causal && j > pos_i) a = -INFINITY;
                    sc[j] = a;
                    mx = fmaxf(mx, a);
                }
#pragma unroll
                for (int j = 0; j < SEQ_K; ++j) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
                if (g.drop_p > 0.f) {
                    const unsigned thr = vptr_drop_threshold(g.drop_p);
                    const float keep = 1.f / (1.f - g.drop_p);
                    if ((drop_row & 1) == 0) {                   // even start: elements (2jj, 2jj+1) share one hash group
#pragma unroll
                        for (int jj = 0; jj < (SEQ_K + 1) / 2; ++jj) {
                            const unsigned long long idx = drop_row + 2 * jj;
                            const unsigned z = (unsigned)(vptr_hash4(g.drop_seed, idx >> 2) >> (16 * (unsigned)(idx & 3)));
                            sc[2 * jj] *= (z & 0xFFFFu) >= thr ? keep : 0.f;
                            if (2 * jj + 1 < SEQ_K) sc[2 * jj + 1] *= ((z >> 16) & 0xFFFFu) >= thr ? keep : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < SEQ_K; ++j) sc[j] *= vptr_drop_scale(g.drop_seed, drop_row + j, g.drop_p);
                    }
                }
                if (valid_i) {
#pragma unroll
                    for (int j = 0; j < SEQ_K; ++j) *reinterpret_cast<float*>(sP + sw128_offset(row, gid_i * SEQ_K + j)) = vptr_round_tf32(sc[j]);
                }
            } else if (FQ == 0) {
                // pass 1: row maximum over the attended columns
                float mx = -INFINITY;
    #pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (!((need >> c) & 1)) continue;
                    float v[32];
                    tmem_ld32(tmem_S + lane_taddr + c * 32, v);
                    tmem_ld_wait();
    #pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if ((vm[c] >> j) & 1) {
                            float s = v[j] * g.scale;
                            if (rpe_table) s += brow[h * g.Lq * g.Lk + colpos[c * 32 + j]];
                            mx = fmaxf(mx, s);
                        }
                    }
                }
                // pass 2: p = exp(s - max) (x dropout keep-scale), row sum, P tile
    #pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (!((need >> c) & 1)) continue;
                    float v[32];
                    tmem_ld32(tmem_S + lane_taddr + c * 32, v);
                    tmem_ld_wait();
    #pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float p = 0.f;
                        if ((vm[c] >> j) & 1) {
                            float s = v[j] * g.scale;
                            if (rpe_table) s += brow[h * g.Lq * g.Lk + colpos[c * 32 + j]];
                            p = __expf(s - mx);
                            sum += p;
                            if (g.drop_p > 0.f) p *= vptr_drop_scale(g.drop_seed, drop_row + colpos[c * 32 + j], g.drop_p);
                        }
                        v[j] = vptr_round_tf32(p);
                    }
    #pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4)
                        *reinterpret_cast<float4*>(sP + sw128_offset(row, c * 32 + q4 * 4)) = make_float4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]);
                }
            }
            const float inv = valid_i ? 1.f / sum : 0.f;
            tcgen05_fence_before();
            fence_async_smem();                 // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
            mbar_arrive(p_ready);
            // ---- epilogue: O / rowsum -> staging tile (over V) -> TMA store
            mbar_wait(o_ready, ph, abort_flag, 5 + 10 * (warp == 4));
            tcgen05_fence_after();
            const int odd = h & 1;
            // global row of this thread (for the two generic-store columns of odd heads)
            long long grow = -1;
            if (g.mode == 0) { grow = (long long)tile * TC_ROWS + row; if (grow >= g.total_rows_q) grow = -1; }
            else {
                const int n = tile / g.chunks_per_clip, p0 = (tile - n * g.chunks_per_clip) * g.P;
                if (valid_i && p0 + gid_i < g.HW) grow = ((long long)n * g.Tq + pos_i) * g.HW + p0 + gid_i;
            }
            // 64 of the head's 66 columns leave through two aligned TMA boxes; the remaining pair (whose 16-byte granule is shared
            // with the neighbouring head -- a clipped TMA store rewrites the whole granule) is one generic store.  Accumulator
            // column of head column c: c (even head) or c + 2 (odd head, tiles start two columns early).
            uint8_t* orow = sV + row_off;
            auto stage = [&](int G, float4 o) {     // staged columns 4G .. 4G+3 of this row (K-major SWIZZLE_128B)
                *reinterpret_cast<float4*>(orow + (uint32_t)(G >> 3) * TC_CHUNK + ((((uint32_t)G & 7u) ^ r7) << 4)) = o;
            };
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v[32];
                if (c < 2) tmem_ld32(tmem_O + lane_taddr + c * 32, v);
                else {   // only accumulator columns 64..67 are needed from the last chunk
                    uint32_t* r = reinterpret_cast<uint32_t*>(v);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(tmem_O + lane_taddr + 64));
                }
                tmem_ld_wait();
#pragma unroll
                for (int q4 = 0; q4 < (c < 2 ? 8 : 1); ++q4) {
                    float4 o = make_float4(v[q4 * 4] * inv, v[q4 * 4 + 1] * inv, v[q4 * 4 + 2] * inv, v[q4 * 4 + 3] * inv);
                    if (g.round_tf32) { o.x = vptr_round_tf32(o.x); o.y = vptr_round_tf32(o.y); o.z = vptr_round_tf32(o.z); o.w = vptr_round_tf32(o.w); }
                    const int G = c * 8 + q4;                       // accumulator columns 4G..4G+3
                    if (!odd) {
                        if (G < 16) stage(G, o);                                                             // head columns 0..63
                        else if (grow >= 0) *reinterpret_cast<float2*>(g.O + grow * g.ldo + h * TC_D + 64) = make_float2(o.x, o.y);   // 64..65
                    } else if (G == 0) {                            // head columns 0..1 sit in accumulator columns 2..3
                        if (grow >= 0) *reinterpret_cast<float2*>(g.O + grow * g.ldo + h * TC_D) = make_float2(o.z, o.w);
                    } else {                                        // head columns 2..65 -> staged columns 0..63
                        stage(G - 1, o);
                    }
                }
            }
            tcgen05_fence_before();
            fence_async_smem();
            bar_sync_128();
            if (threadIdx.x == 0) {
                const int c0 = h * TC_D + 2 * odd;
                for (int c = 0; c < 2; ++c) {
                    if (g.mode == 0) {
                        tma_store_2d(&maps.o, sV + c * TC_CHUNK, c0 + c * 32, tile * TC_ROWS);
                    } else {
                        const int n = tile / g.chunks_per_clip, p0 = (tile - n * g.chunks_per_clip) * g.P;
                        tma_store_4d(&maps.o, sV + c * TC_CHUNK, c0 + c * 32, 0, p0, n);
                    }
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(v_free);
            }
        }
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 4) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// window: [rows][width]; temporal: [N][T][HW][width]  (width = nhead * 66 columns from `base`; boxes past it are zero-filled)
int make_tensor_map(CUtensorMap* m, const float* base, long long ld, int width_cols, const TcGeom& g, long long rows, int N, int T, CUtensorMapSwizzle sw) {
    EncodeTiledFn enc = encode_fn();
    VPTR_REQUIRE(enc != nullptr, VPTR_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t width = (cuuint64_t)width_cols;
    CUresult r;
    if (g.mode == 0) {
        cuuint64_t dims[2] = {width, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
        cuuint32_t box[2] = {32, TC_ROWS}, es[2] = {1, 1};
        r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        // dims ordered (column, t, pixel, clip): the box {32, T, P, 1} lands with t fastest, i.e. tile row = pixel * T + t
        cuuint64_t dims[4] = {width, (cuuint64_t)T, (cuuint64_t)g.HW, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)g.HW * ld * 4, (cuuint64_t)ld * 4, (cuuint64_t)T * g.HW * ld * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)T, (cuuint32_t)g.P, 1}, es[4] = {1, 1, 1, 1};
        r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    VPTR_REQUIRE(r == CUDA_SUCCESS, VPTR_ERR_DRIVER, "cuTensorMapEncodeTiled(attention map) failed (%d)", (int)r);
    return VPTR_OK;
}

}  // namespace

// Returns VPTR_ERR_UNSUPPORTED (without setting an error a caller must report) when the shape is outside the kernel's domain.
// mode 0: F_or_N = frames; mode 1: F_or_N = clips.  Same argument meaning as vptr_attn_fwd.
extern "C" int vptr_attn_fwd_tcgen05(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv, float* O, long long ldo,
                          const float* rpe_table, int mode, int F_or_N, int H, int W, int ws, int Tq, int Tk, int nhead, int d, int causal,
                          float scale, int round_tf32, unsigned long long drop_seed, float drop_p, cudaStream_t stream) {
    vptr_set_error("vptr_attn_fwd_tcgen05: shape outside the kernel's domain (head_dim 66, <= 8 heads, groups of <= 32 tokens, windows of <= 16)");
    if (d != TC_D || nhead > TC_MAXHEADS || nhead < 1 || mode < 0 || mode > 1) return VPTR_ERR_UNSUPPORTED;
    if ((ldq | ldk | ldv | ldo) % 4 != 0 || ((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O) % 16 != 0) return VPTR_ERR_UNSUPPORTED;
    TcGeom g{};
    g.mode = mode; g.nhead = nhead; g.causal = causal; g.scale = scale; g.round_tf32 = round_tf32; g.drop_seed = drop_seed; g.drop_p = drop_p;
    g.ws = ws; g.O = O; g.ldo = ldo;
    long long rows_q = 0, rows_k = 0;
    if (mode == 0) {
        if (ws <= 0 || H % ws || W % ws || ws * ws > 16) return VPTR_ERR_UNSUPPORTED;          // bias LUT sized for <= 16 positions
        const int HW = H * W;
        if (HW <= TC_ROWS) { if (TC_ROWS % HW) return VPTR_ERR_UNSUPPORTED; }
        else if ((ws * W) > TC_ROWS || TC_ROWS % (ws * W) || HW % TC_ROWS) return VPTR_ERR_UNSUPPORTED;
        g.H = H; g.W = W; g.HW = HW; g.nwh = H / ws; g.nww = W / ws; g.Lq = g.Lk = ws * ws;
        rows_q = rows_k = (long long)F_or_N * HW;
        g.tiles = (int)((rows_q + TC_ROWS - 1) / TC_ROWS);
        g.rows_q = g.rows_k = TC_ROWS;
        if (HW > TC_ROWS) g.nwh = TC_ROWS / (ws * W);   // band tiles: window rows per tile
        g.wpt = (HW <= TC_ROWS ? TC_ROWS / HW : 1) * g.nwh * g.nww;
    } else {
        if (Tq <= 0 || Tk <= 0 || Tq > 32 || Tk > 32 || (causal && Tq != Tk)) return VPTR_ERR_UNSUPPORTED;
        g.HW = H * W; g.Tq = Tq; g.Tk = Tk; g.Lq = Tq; g.Lk = Tk;
        const int Tm = Tq > Tk ? Tq : Tk;
        g.P = TC_ROWS / Tm;
        if (g.P > g.HW) g.P = g.HW;
        if (g.P > 255) return VPTR_ERR_UNSUPPORTED;
        g.chunks_per_clip = (g.HW + g.P - 1) / g.P;
        g.tiles = F_or_N * g.chunks_per_clip;
        g.rows_q = Tq * g.P; g.rows_k = Tk * g.P;
        rows_q = (long long)F_or_N * Tq * g.HW; rows_k = (long long)F_or_N * Tk * g.HW;
    }
    g.total_rows_q = rows_q;
    TcMaps maps;
    int rc = make_tensor_map(&maps.q, Q, ldq, nhead * TC_D, g, rows_q, F_or_N, Tq, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tensor_map(&maps.k, K, ldk, nhead * TC_D, g, rows_k, F_or_N, Tk, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tensor_map(&maps.v, V, ldv, nhead * TC_D, g, rows_k, F_or_N, Tk, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    rc = make_tensor_map(&maps.o, O, ldo, nhead * TC_D, g, rows_q, F_or_N, Tq, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const size_t smem = 13 * TC_CHUNK + 1024 + 512 + (rpe_table ? sizeof(float) * nhead * g.Lq * g.Lk : 0) + 64;
    VPTR_REQUIRE(smem <= 227 * 1024, VPTR_ERR_UNSUPPORTED, "vptr_attn_fwd_tcgen05: shared memory %zu", smem);
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const long long items = (long long)g.tiles * nhead;
    const int grid = (int)(items < sms ? items : sms);
    void (*kern)(const TcMaps, const float*, const TcGeom) = attn_tc_fwd_kernel<0, 0>;
    if (mode == 1 && g.P == TC_ROWS / (Tq > Tk ? Tq : Tk)) {   // full-width pixel chunks: the compile-time sequence fast paths
        if (Tq == 10 && Tk == 10) kern = attn_tc_fwd_kernel<10, 10>;
        else if (Tq == 29 && Tk == 29) kern = attn_tc_fwd_kernel<29, 29>;
        else if (Tq == 28 && Tk == 28) kern = attn_tc_fwd_kernel<28, 28>;
        else if (Tq == 28 && Tk == 2) kern = attn_tc_fwd_kernel<28, 2>;
        else if (Tq == 30 && Tk == 30) kern = attn_tc_fwd_kernel<30, 30>;     // cfg4 (10 -> 30 frames): decoder temporal self-attention
        else if (Tq == 30 && Tk == 10) kern = attn_tc_fwd_kernel<30, 10>;     // cfg4: encoder-decoder attention
    }
    cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    VPTR_REQUIRE(ea == cudaSuccess, (int)ea, "cudaFuncSetAttribute(attn_tc_fwd): %s", cudaGetErrorString(ea));
    kern<<<grid, TC_THREADS, smem, stream>>>(maps, rpe_table, g);
    static const bool debug = [] { const char* e = getenv("VPTR_ATTN_TC_DEBUG"); return e && e[0] == '1'; }();
    if (debug) {
        int err[4] = {0, 0, 0, 0};
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(err, tc_err, sizeof(err));
        VPTR_REQUIRE(e == cudaSuccess && err[0] == 0, VPTR_ERR_DRIVER, "attn_tc_fwd_kernel: %s; barrier timeout code %d parity %d thread %d block %d",
                     cudaGetErrorString(e), err[0], err[1], err[2], err[3]);
    }
    return vptr_check_launch("attn_tc_fwd_kernel");
}

VPTR_RNG_EPOCH_ACCESSOR(attn_tcgen05)
