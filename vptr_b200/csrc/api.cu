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
