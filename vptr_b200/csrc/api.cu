// Error reporting + version for the C-ABI (include/vptr_b200.h).
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void vptr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int vptr_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        vptr_set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return VPTR_OK;
}

extern "C" const char* vptr_last_error(void) { return g_err; }
extern "C" int vptr_version(void) { return 100; }

// ---- CUDA-graph support: device-side state that must change between replays of a captured step ---------------------------------
extern "C" unsigned long long* vptr_rng_epoch_addr_gemm_tcgen05(void);
extern "C" unsigned long long* vptr_rng_epoch_addr_gemm_simt(void);
extern "C" unsigned long long* vptr_rng_epoch_addr_norm(void);
extern "C" unsigned long long* vptr_rng_epoch_addr_attn(void);
extern "C" unsigned long long* vptr_rng_epoch_addr_attn_tcgen05(void);
extern "C" unsigned long long* vptr_rng_epoch_addr_elementwise(void);
namespace {
struct EpochPtrs { unsigned long long* p[6]; };
__global__ void rng_advance_kernel(EpochPtrs e, unsigned long long* counter, long long reset) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const unsigned long long v = reset >= 0 ? (unsigned long long)reset : *counter + 1;
        *counter = v;
        for (int i = 0; i < 6; ++i) *e.p[i] = v;
    }
}
__global__ void counter_add_kernel(long long* ctr, long long inc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *ctr += inc;
}
__device__ unsigned long long g_rng_counter = 0;
}  // namespace
// Dropout / DropPath epoch of every kernel in the library: reset >= 0 sets it (0 = the eager default), reset < 0 increments it.
// Capture it as the first node of a graphed training step (see common.cuh: g_vptr_rng_epoch).
extern "C" int vptr_rng_advance(long long reset, cudaStream_t stream) {
    static EpochPtrs e = [] {
        EpochPtrs t;
        t.p[0] = vptr_rng_epoch_addr_gemm_tcgen05(); t.p[1] = vptr_rng_epoch_addr_gemm_simt(); t.p[2] = vptr_rng_epoch_addr_norm();
        t.p[3] = vptr_rng_epoch_addr_attn(); t.p[4] = vptr_rng_epoch_addr_attn_tcgen05(); t.p[5] = vptr_rng_epoch_addr_elementwise();
        return t;
    }();
    for (int i = 0; i < 6; ++i) VPTR_REQUIRE(e.p[i] != nullptr, VPTR_ERR_DRIVER, "vptr_rng_advance: epoch symbol %d not found", i);
    unsigned long long* ctr = nullptr;
    cudaGetSymbolAddress(reinterpret_cast<void**>(&ctr), g_rng_counter);
    rng_advance_kernel<<<1, 32, 0, stream>>>(e, ctr, reset);
    return vptr_check_launch("rng_advance_kernel");
}
// *ctr += inc on the device (the optimizer's step count inside a captured graph)
extern "C" int vptr_counter_add(long long* ctr, long long inc, cudaStream_t stream) {
    counter_add_kernel<<<1, 32, 0, stream>>>(ctr, inc);
    return vptr_check_launch("counter_add_kernel");
}
