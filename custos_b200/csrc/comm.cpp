// comm.cpp — multi-GPU combine of reduction partials over NCCL (NVLink 5 / NVSwitch).
//
// The reference has no multi-device support (its one cross-device test is ignored:
// src/devices/cuda/cuda.rs:187-202).  Element-wise work shards by contiguous slice and
// needs no communication.  A sharded sum exchanges exactly one scalar per rank:
//   local deterministic two-pass sum -> ncclAllGather(1 element) -> fold in RANK ORDER
// on every rank, so all ranks hold the same bits and the order does not depend on
// arrival time (an all-reduce would leave the order to NCCL's algorithm choice).
// The message is 4-8 bytes: latency bound, NVLink bandwidth is irrelevant.
//
// NCCL is resolved with dlopen at first use: the library loads without it, and inside a
// torch process the already loaded libnccl.so.2 is reused.
#include <dlfcn.h>

#include <cstring>
#include <memory>

#include "device.h"

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7, ncclFloat64 = 8, ncclInt64 = 4 };

struct NcclApi {
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NcclApi &nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
    return api;
}

int32_t nccl_fail(int r, const char *what)
{
    return cb::fail(CB_ERR_NCCL + r, "%s: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "nccl error");
}

}  // namespace

struct cb_comm {
    cb_device *dev = nullptr;
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0;
    void *local = nullptr;     // this rank's partial (8 bytes)
    void *gathered = nullptr;  // n_ranks partials
};

using cb::fail;

extern "C" int32_t cb_comm_unique_id(uint8_t id[CB_COMM_ID_BYTES])
{
    CB_CHECK_ARG(id, "null id");
    static_assert(sizeof(ncclUniqueId) == CB_COMM_ID_BYTES, "ncclUniqueId size");
    if (!nccl().ok) return fail(CB_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
    ncclUniqueId u;
    int r = nccl().GetUniqueId(&u);
    if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
    std::memcpy(id, &u, sizeof u);
    return CB_OK;
}

extern "C" int32_t cb_comm_create(cb_device *dev, int32_t n_ranks, int32_t rank, const uint8_t id[CB_COMM_ID_BYTES],
                                  cb_comm **out)
{
    CB_CHECK_ARG(dev && out && id && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad argument");
    *out = nullptr;
    if (!nccl().ok) return fail(CB_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
    CB_TRY(dev->use());
    std::unique_ptr<cb_comm> c(new cb_comm());
    c->dev = dev;
    c->n_ranks = n_ranks;
    c->rank = rank;
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    int r = nccl().CommInitRank(&c->comm, n_ranks, u, rank);
    if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
    cudaError_t e = cudaMalloc(&c->local, 64);
    if (e == cudaSuccess) e = cudaMalloc(&c->gathered, (size_t)n_ranks * 8);
    if (e != cudaSuccess) {
        nccl().CommDestroy(c->comm);
        return dev->cuda_fail(e, "cudaMalloc (comm scratch)");
    }
    *out = c.release();
    return CB_OK;
}

extern "C" int32_t cb_comm_destroy(cb_comm *c)
{
    if (!c) return CB_OK;
    c->dev->use();
    cudaStreamSynchronize(c->dev->stream);
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->local) cudaFree(c->local);
    if (c->gathered) cudaFree(c->gathered);
    delete c;
    return CB_OK;
}

static int32_t comm_reduce(cb_comm *c, int32_t dtype, uint64_t in, size_t n_local, uint64_t out, size_t divisor)
{
    CB_CHECK_ARG(c && cb::valid_dtype(dtype) && out, "bad argument");
    cb_device *dev = c->dev;
    CB_TRY(dev->use());
    // a rank may own an empty slice (n smaller than the rank count): its partial is 0
    if (n_local == 0) {
        cudaError_t e = cudaMemsetAsync(c->local, 0, 8, dev->stream);
        if (e != cudaSuccess) return dev->cuda_fail(e, "cudaMemsetAsync");
    } else {
        CB_TRY(cb_sum(dev, dtype, in, n_local, reinterpret_cast<uint64_t>(c->local)));
    }
    const int nccl_type = (dtype == CB_F32 || dtype == CB_F16) ? ncclFloat32 : (dtype == CB_F64 ? ncclFloat64 : ncclInt64);
    int r = nccl().AllGather(c->local, c->gathered, 1, nccl_type, c->comm, dev->stream);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllGather");
    cudaError_t e = cb::launch_fold_ranks(dev->ctx(), dtype, c->gathered, c->n_ranks, reinterpret_cast<void *>(out), divisor);
    if (e != cudaSuccess) return dev->cuda_fail(e, "fold_ranks kernel");
    dev->launches += 1;
    return CB_OK;
}

extern "C" int32_t cb_comm_sum(cb_comm *c, int32_t dtype, uint64_t in, size_t n_local, uint64_t out)
{
    return comm_reduce(c, dtype, in, n_local, out, 0);
}

extern "C" int32_t cb_comm_mean(cb_comm *c, int32_t dtype, uint64_t in, size_t n_local, size_t n_global, uint64_t out)
{
    if (!n_global) return fail(CB_ERR_ZERO_LENGTH, "mean over a zero length buffer");
    return comm_reduce(c, dtype, in, n_local, out, n_global);
}
