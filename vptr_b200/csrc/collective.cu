// Data-parallel gradient reduction behind the C-ABI (SURVEY.md 8b export list, 8e): ONE in-place NCCL all-reduce (mean) per finished
// slice of the flat gradient buffer, with the squared-norm accumulation that clip_grad_norm_ needs fused behind it on the same
// stream -- replaces DistributedDataParallel's bucketed reduction (reference train_NAR_mp.py:118,167-168; train_FAR_mp.py:132,178).
// NCCL is bound at run time (dlopen of the libnccl.so.2 the host process already carries), so the library has no link-time
// dependency on it and single-GPU users never touch it.
#include "common.cuh"
#include <dlfcn.h>
#include <string.h>

extern "C" int vptr_sqnorm_accumulate(const float* x, long long n, double* sqnorm_out, cudaStream_t stream);

namespace {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*CommDestroyFn)(NcclComm);
typedef const char* (*GetErrorStringFn)(int);
constexpr int kNcclFloat32 = 7, kNcclAvg = 4;

struct NcclApi {
    void* handle = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    AllReduceFn all_reduce = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    GetErrorStringFn error_string = nullptr;
};
NcclApi* nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            api.get_unique_id = (GetUniqueIdFn)dlsym(api.handle, "ncclGetUniqueId");
            api.comm_init_rank = (CommInitRankFn)dlsym(api.handle, "ncclCommInitRank");
            api.all_reduce = (AllReduceFn)dlsym(api.handle, "ncclAllReduce");
            api.comm_destroy = (CommDestroyFn)dlsym(api.handle, "ncclCommDestroy");
            api.error_string = (GetErrorStringFn)dlsym(api.handle, "ncclGetErrorString");
        }
    }
    return (api.handle && api.get_unique_id && api.comm_init_rank && api.all_reduce && api.comm_destroy) ? &api : nullptr;
}
const char* nccl_err(NcclApi* a, int rc) { return a->error_string ? a->error_string(rc) : "nccl error"; }

__global__ void __launch_bounds__(256) sqnorm_vec_kernel(const float4* __restrict__ x, long long n4, const float* __restrict__ tail, int ntail,
                                                         double* __restrict__ out) {
    __shared__ float red[32];
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = x[i];
        s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) s = fmaf(tail[threadIdx.x], tail[threadIdx.x], s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(out, (double)s);
}

}  // namespace

// 128-byte NCCL unique id for a new communicator (call on ONE rank, ship the bytes to the others by any means)
extern "C" int vptr_nccl_unique_id(unsigned char* out128) {
    NcclApi* a = nccl();
    VPTR_REQUIRE(a != nullptr, VPTR_ERR_UNSUPPORTED, "vptr_nccl_unique_id: libnccl.so.2 is not loadable in this process");
    NcclUniqueId id;
    const int rc = a->get_unique_id(&id);
    VPTR_REQUIRE(rc == 0, VPTR_ERR_DRIVER, "ncclGetUniqueId: %s", nccl_err(a, rc));
    memcpy(out128, id.internal, 128);
    return VPTR_OK;
}
// collective: every rank calls it with the same id; the calling thread's current CUDA device is the rank's GPU
extern "C" int vptr_nccl_comm_init(void** comm, int world, int rank, const unsigned char* id128) {
    NcclApi* a = nccl();
    VPTR_REQUIRE(a != nullptr, VPTR_ERR_UNSUPPORTED, "vptr_nccl_comm_init: libnccl.so.2 is not loadable in this process");
    VPTR_REQUIRE(comm != nullptr && world > 0 && rank >= 0 && rank < world, VPTR_ERR_SHAPE, "vptr_nccl_comm_init: world=%d rank=%d", world, rank);
    NcclUniqueId id;
    memcpy(id.internal, id128, 128);
    NcclComm c = nullptr;
    const int rc = a->comm_init_rank(&c, world, id, rank);
    VPTR_REQUIRE(rc == 0, VPTR_ERR_DRIVER, "ncclCommInitRank: %s", nccl_err(a, rc));
    *comm = c;
    return VPTR_OK;
}
extern "C" int vptr_nccl_comm_destroy(void* comm) {
    NcclApi* a = nccl();
    if (a && comm) a->comm_destroy(comm);
    return VPTR_OK;
}
// flat[0..n) <- mean over the ranks of `comm` (in place, ncclAvg over NVLink / NVSwitch), then -- when sqnorm_out != NULL -- the
// device double *sqnorm_out += sum(flat^2) of the REDUCED values, on the same stream (caller zeroes it once per step).
extern "C" int vptr_allreduce_grads(void* comm, float* flat, long long n, double* sqnorm_out, cudaStream_t stream) {
    NcclApi* a = nccl();
    VPTR_REQUIRE(a != nullptr, VPTR_ERR_UNSUPPORTED, "vptr_allreduce_grads: libnccl.so.2 is not loadable in this process");
    VPTR_REQUIRE(comm != nullptr && flat != nullptr && n >= 0, VPTR_ERR_SHAPE, "vptr_allreduce_grads: comm=%p flat=%p n=%lld", comm, (void*)flat, n);
    if (n == 0) return VPTR_OK;
    const int rc = a->all_reduce(flat, flat, (size_t)n, kNcclFloat32, kNcclAvg, comm, stream);
    VPTR_REQUIRE(rc == 0, VPTR_ERR_DRIVER, "ncclAllReduce: %s", nccl_err(a, rc));
    if (sqnorm_out) {
        const bool al = ((uintptr_t)flat & 15) == 0;
        const long long n4 = al ? n / 4 : 0;
        const int ntail = (int)(n - 4 * n4);
        if (ntail <= 256) {
            long long blocks = (n4 + 255) / 256;
            if (blocks < 1) blocks = 1;
            if (blocks > 148 * 8) blocks = 148 * 8;
            sqnorm_vec_kernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(flat), n4, flat + 4 * n4, ntail, sqnorm_out);
        } else {
            return vptr_sqnorm_accumulate(flat, n, sqnorm_out, stream);
        }
        return vptr_check_launch("sqnorm_vec_kernel");
    }
    return VPTR_OK;
}

// Scratch bytes the caller must provide to the entry points that take a workspace (the library never allocates):
//   op 0 vptr_norm_act_bwd (mode 0/2: 2*ch floats; mode 1: 2*frames floats)   op 1 vptr_bn_stats (2*ch doubles)
//   op 2 vptr_head_conv7x7_bwd (rows = F, ch = Ci, hw = H (= W), mode = Co)    anything else: 0
extern "C" long long vptr_workspace_bytes(int op, long long rows, int ch, int hw, int mode) {
    switch (op) {
        case 0: return (long long)sizeof(float) * 2 * (mode == 1 ? (hw > 0 ? rows / hw : 0) : ch);
        case 1: return (long long)sizeof(double) * 2 * ch;
        case 2: return (long long)sizeof(float) * (rows * (hw + 6) * (hw + 6) * ch + 49LL * mode * ch);
        default: return 0;
    }
}
