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

#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

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
    // peer-memory exchange (the last block of the sum kernel folds and exchanges); NCCL is only used to set it up
    bool p2p = false;
    cb::XchgSlot *xchg = nullptr;                 // this rank's buffer: 2 parities x n_ranks slots
    std::vector<void *> opened;                   // peers' buffers mapped through CUDA IPC
    cb::XchgSlot **peer_tbl = nullptr;            // device array of n_ranks buffer addresses
    int *status = nullptr;                        // host-mapped flag: call number of an exchange in which a peer timed out
    int *status_dev = nullptr;                    // ... its device-side address
    long long timeout_cycles = 4000000000ll;      // ~2 s of SM clock (CB_COMM_TIMEOUT_MS)
    unsigned long long epoch = 0;
};

// Maps every rank's exchange buffer into this process.  Returns false (and leaves the communicator on the
// NCCL path) if any step is unavailable on ANY rank; the decision is taken collectively.
static bool setup_p2p(cb_comm *c)
{
    if (const char *v = std::getenv("CB_COMM_P2P"))
        if (v[0] == '0') return false;
    if (c->n_ranks > cb::kMaxRanks) return false;
    const size_t bytes = sizeof(cb::XchgSlot) * 2 * (size_t)c->n_ranks;
    // the timeout flag lives in mapped pinned host memory: the host can look at it without touching the stream
    bool ok = cudaMalloc(reinterpret_cast<void **>(&c->xchg), bytes) == cudaSuccess &&
              cudaMemset(c->xchg, 0, bytes) == cudaSuccess &&
              cudaHostAlloc(reinterpret_cast<void **>(&c->status), sizeof(int), cudaHostAllocMapped) == cudaSuccess &&
              cudaHostGetDevicePointer(reinterpret_cast<void **>(&c->status_dev), c->status, 0) == cudaSuccess;
    if (ok) *c->status = 0;
    c->timeout_cycles = (long long)cb::env_int("CB_COMM_TIMEOUT_MS", 2000, 1, 600000) * 2000000ll;  // ~2 GHz
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof mine);
    ok = ok && cudaIpcGetMemHandle(&mine, c->xchg) == cudaSuccess;
    // all-gather {ok flag, handle} through NCCL (bytes)
    struct Msg {
        unsigned char ok;
        unsigned char pad[7];
        cudaIpcMemHandle_t h;
    } msg;
    std::memset(&msg, 0, sizeof msg);
    msg.ok = ok ? 1 : 0;
    msg.h = mine;
    void *d_send = nullptr, *d_recv = nullptr;
    std::vector<Msg> all((size_t)c->n_ranks);
    bool comm_ok = cudaMalloc(&d_send, sizeof(Msg)) == cudaSuccess && cudaMalloc(&d_recv, sizeof(Msg) * (size_t)c->n_ranks) == cudaSuccess &&
                   cudaMemcpy(d_send, &msg, sizeof(Msg), cudaMemcpyHostToDevice) == cudaSuccess &&
                   nccl().AllGather(d_send, d_recv, sizeof(Msg), /*ncclUint8*/ 1, c->comm, c->dev->stream) == ncclSuccess &&
                   cudaStreamSynchronize(c->dev->stream) == cudaSuccess &&
                   cudaMemcpy(all.data(), d_recv, sizeof(Msg) * (size_t)c->n_ranks, cudaMemcpyDeviceToHost) == cudaSuccess;
    if (d_send) cudaFree(d_send);
    if (d_recv) cudaFree(d_recv);
    if (!comm_ok) {
        cudaGetLastError();
        return false;
    }
    for (const Msg &m : all) ok = ok && m.ok;
    std::vector<cb::XchgSlot *> tbl((size_t)c->n_ranks, nullptr);
    if (ok) {
        for (int r = 0; r < c->n_ranks && ok; r++) {
            if (r == c->rank) {
                tbl[(size_t)r] = c->xchg;
                continue;
            }
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                ok = false;
                break;
            }
            c->opened.push_back(p);
            tbl[(size_t)r] = static_cast<cb::XchgSlot *>(p);
        }
    }
    // second collective decision: did every rank manage to open every handle?
    unsigned char flag = ok ? 1 : 0;
    std::vector<unsigned char> flags((size_t)c->n_ranks, 0);
    void *d1 = nullptr, *dn = nullptr;
    comm_ok = cudaMalloc(&d1, 8) == cudaSuccess && cudaMalloc(&dn, (size_t)c->n_ranks + 8) == cudaSuccess &&
              cudaMemcpy(d1, &flag, 1, cudaMemcpyHostToDevice) == cudaSuccess &&
              nccl().AllGather(d1, dn, 1, 1, c->comm, c->dev->stream) == ncclSuccess &&
              cudaStreamSynchronize(c->dev->stream) == cudaSuccess &&
              cudaMemcpy(flags.data(), dn, (size_t)c->n_ranks, cudaMemcpyDeviceToHost) == cudaSuccess;
    if (d1) cudaFree(d1);
    if (dn) cudaFree(dn);
    for (unsigned char f : flags) ok = ok && f;
    if (!comm_ok || !ok) {
        cudaGetLastError();
        return false;
    }
    if (cudaMalloc(reinterpret_cast<void **>(&c->peer_tbl), sizeof(void *) * (size_t)c->n_ranks) != cudaSuccess ||
        cudaMemcpy(c->peer_tbl, tbl.data(), sizeof(void *) * (size_t)c->n_ranks, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        return false;  // cannot happen on one rank only in practice; the NCCL path still works everywhere
    }
    return true;
}

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
    c->p2p = setup_p2p(c.get());
    *out = c.release();
    return CB_OK;
}

extern "C" int32_t cb_comm_uses_peer_memory(cb_comm *c, int32_t *flag)
{
    CB_CHECK_ARG(c && flag, "null argument");
    *flag = c->p2p ? 1 : 0;
    return CB_OK;
}

extern "C" int32_t cb_comm_destroy(cb_comm *c)
{
    if (!c) return CB_OK;
    c->dev->use();
    cudaStreamSynchronize(c->dev->stream);
    for (void *p : c->opened) cudaIpcCloseMemHandle(p);
    if (c->peer_tbl) cudaFree(c->peer_tbl);
    if (c->status) cudaFreeHost(c->status);
    if (c->xchg) cudaFree(c->xchg);
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->local) cudaFree(c->local);
    if (c->gathered) cudaFree(c->gathered);
    delete c;
    return CB_OK;
}

// A peer that never arrived makes the exchange kernel give up (instead of hanging the GPU) and record the call number
// in the host-mapped flag; the value the kernel then wrote is NOT the global sum.  Reported here, at every later
// cb_comm_* call and by cb_comm_check, until the communicator is destroyed.
static int32_t comm_status(cb_comm *c)
{
    if (c->p2p && c->status && *reinterpret_cast<volatile int *>(c->status) != 0)
        return fail(CB_ERR_STATE, "cb_comm: a peer did not reach exchange %d within the timeout (CB_COMM_TIMEOUT_MS); "
                                  "the result of that call is not the global sum", *reinterpret_cast<volatile int *>(c->status));
    return CB_OK;
}

extern "C" int32_t cb_comm_check(cb_comm *c)
{
    CB_CHECK_ARG(c, "null communicator");
    CB_TRY(c->dev->use());
    cudaError_t e = cudaStreamSynchronize(c->dev->stream);
    if (e != cudaSuccess) return c->dev->cuda_fail(e, "cudaStreamSynchronize");
    return comm_status(c);
}

static int32_t comm_reduce(cb_comm *c, int32_t dtype, uint64_t in, size_t n_local, uint64_t out, size_t divisor)
{
    CB_CHECK_ARG(c && cb::valid_dtype(dtype) && out, "bad argument");
    CB_TRY(comm_status(c));
    if (dtype == CB_BOOL) return cb::fail(CB_ERR_UNSUPPORTED, "bool has no arithmetic in the reference (storage only)");
    cb_device *dev = c->dev;
    CB_TRY(dev->use());
    if (c->p2p) {
        // ONE kernel reduces the slice, folds the partials, exchanges the totals over NVLink peer memory and
        // folds them in rank order
        cb::XchgArgs x;
        x.peers = c->peer_tbl;
        x.n_ranks = c->n_ranks;
        x.rank = c->rank;
        x.epoch = ++c->epoch;
        x.timeout_cycles = c->timeout_cycles;
        x.status = c->status_dev;
        if (n_local && !in) return fail(CB_ERR_INVALID_ARG, "null buffer");
        cudaError_t e = cb::launch_sum_exchange(dev->ctx(), dtype, reinterpret_cast<const void *>(in), n_local, dev->sum_partials,
                                                dev->sum_ticket, reinterpret_cast<void *>(out), divisor, x, dev->next_sum_pdl(in, n_local * cb::dtype_size(dtype), out));
        if (e != cudaSuccess) return dev->cuda_fail(e, "sum + exchange kernel");
        dev->launches += 1;  // reduce, fold and exchange are one launch
        return CB_OK;
    }
    // a rank may own an empty slice (n smaller than the rank count): its partial is 0
    if (n_local == 0) {
        cudaError_t e = cudaMemsetAsync(c->local, 0, 8, dev->stream);
        if (e != cudaSuccess) return dev->cuda_fail(e, "cudaMemsetAsync");
    } else {
        CB_TRY(cb_sum(dev, dtype, in, n_local, reinterpret_cast<uint64_t>(c->local)));
    }
    const int nccl_type = (dtype == CB_F32 || cb::is_half_dtype(dtype)) ? ncclFloat32 : (dtype == CB_F64 ? ncclFloat64 : ncclInt64);
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

static size_t comm_acc_bytes(int32_t dtype) { return (dtype == CB_F32 || cb::is_half_dtype(dtype)) ? 4 : 8; }

static int32_t comm_reduce_host(cb_comm *c, int32_t dtype, uint64_t in, size_t n_local, void *host_out, size_t divisor)
{
    CB_CHECK_ARG(c && host_out, "null argument");
    cb_device *dev = c->dev;
    CB_TRY(comm_reduce(c, dtype, in, n_local, reinterpret_cast<uint64_t>(dev->sum_scalar), divisor));
    cudaError_t e = cudaMemcpyAsync(dev->sum_host, dev->sum_scalar, comm_acc_bytes(dtype), cudaMemcpyDeviceToHost, dev->stream);
    if (e != cudaSuccess) return dev->cuda_fail(e, "cudaMemcpyAsync");
    CB_TRY(cb_comm_check(c));  // synchronises; a timed-out exchange is an error, not a number
    std::memcpy(host_out, dev->sum_host, comm_acc_bytes(dtype));
    return CB_OK;
}

extern "C" int32_t cb_comm_sum_host(cb_comm *c, int32_t dtype, uint64_t in, size_t n_local, void *host_out)
{
    return comm_reduce_host(c, dtype, in, n_local, host_out, 0);
}

extern "C" int32_t cb_comm_mean_host(cb_comm *c, int32_t dtype, uint64_t in, size_t n_local, size_t n_global, void *host_out)
{
    if (!n_global) return fail(CB_ERR_ZERO_LENGTH, "mean over a zero length buffer");
    return comm_reduce_host(c, dtype, in, n_local, host_out, n_global);
}

extern "C" int32_t cb_comm_rank(cb_comm *c, int32_t *rank, int32_t *n_ranks)
{
    CB_CHECK_ARG(c, "null communicator");
    if (rank) *rank = c->rank;
    if (n_ranks) *n_ranks = c->n_ranks;
    return CB_OK;
}

extern "C" int32_t cb_comm_device(cb_comm *c, cb_device **dev)
{
    CB_CHECK_ARG(c && dev, "null argument");
    *dev = c->dev;
    return CB_OK;
}
