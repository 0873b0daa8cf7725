// device.h — the CUDA device behind the C ABI: stream, stream-ordered memory pool, the
// Cached-module allocation slots, pinned staging, the NVRTC kernel cache and CUDA-graph
// capture.  Replaces `CudaDevice` (src/devices/cuda/cuda_device.rs:14-127), `KernelCache`
// (kernel_cache.rs:10-73), `CudaSource` (source.rs:21-66) and `LazyCudaGraph` (lazy.rs:10-72).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.h"
#include "kernels.h"

// a compiled chain of expressions: one CUmodule holding the vector and the scalar variant
struct cb_expr {
    int32_t dtype = 0;
    int32_t kind = 0;
    int32_t n_progs = 0;
    uint64_t key = 0;
    CUmodule module = nullptr;
    CUfunction fn_vec = nullptr;
    CUfunction fn_scalar = nullptr;
    int threads = 256;
    int unroll = 4;       // 16-byte units per thread per tile of the vector kernel
    int resident_vec = 0;     // blocks of the vector / scalar kernel one SM holds (occupancy query):
    int resident_scalar = 0;  // the persistent grid is exactly one wave of them
    size_t cubin_bytes = 0;
    std::string ir;  // canonical bytes of (dtype, kind, programs): identity of the cached kernel
    // f16 / bf16 apply expressions: the chain's value for every one of the 65 536 inputs (128 KiB on the device),
    // filled by this expression's own arithmetic kernel; 0 = no table
    uint64_t lut = 0;
    bool lut_enabled = true;  // cb_expr_set_lookup
};

struct cb_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    size_t kernel_nodes = 0;
};

namespace cb {

// launch-shape tunables (environment, read once): CB_THREADS, CB_UNROLL, CB_MIN_BLOCKS,
// CB_BLOCKS_PER_SM, CB_LD_MOD, CB_ST_MOD.  Defaults were chosen from the ncu captures under profiles/.
struct Tunables {
    int threads = 256;
    int unroll = 4;
    int min_blocks = 5;  // __launch_bounds__ second argument: <= 48 registers, 5 resident blocks (profiles/r1_perf_ab4.log)
    int blocks_per_sm = 8;
    int waves = 16;  // CB_WAVES: grid cap = SMs x resident blocks x waves; >1 lets the hardware rebalance
                     // SMs that run slower (a single static wave left ~10% on the table, profiles/r1_tuning.md)
    int pair = 1;  // CB_PAIR: f32 chains evaluate two elements per thread on the packed f32x2 pipe
    int tile_redo = 0;  // CB_TILE_REDO=1: one fast-path / fallback decision per tile instead of per 16-byte unit
                        // (10 % fewer instructions but burstier stores: slower on B200, profiles/r1_tuning.md)
    int neg_xor = 0;    // CB_NEG_XOR=1: f32 pair neg flips the sign bits on the integer pipe (no gain measured)
    int fuse_scale_add = 1;  // CB_FUSE_SCALE_ADD: (u * 2^k) + C and (u + C) * 2^k become one exact fma in the f32 pair path
    int h_native = 1;  // CB_H_NATIVE: f16 add / sub / mul as single HFMA2s on packed halves
    int pdl = 1;       // CB_PDL: expression and binary kernels are launched with programmatic stream serialisation
    int sum_pdl = 1;   // CB_SUM_PDL: consecutive sums overlap through programmatic dependent launch
    int lut16 = 1;     // CB_LUT16: f16 / bf16 unary chains on large buffers run as a shared-memory table lookup
    int lut_shape = 0;  // CB_LUT_SHAPE (threads x units per tile x tiles per grab): 0 = 512x8x4 (default), 1 = 1024x4x8, 2 = 512x8x2, 3 = 1024x4x4, 4 = 256x16x2
    long long lut16_min_elems = 1ll << 22;  // CB_LUT16_MIN_ELEMS: shorter buffers keep the arithmetic kernel
    std::string ld_mod = ".cs";
    std::string st_mod = ".cs";
    std::string dump_dir;  // CB_DUMP_DIR: write generated sources and cubins here
};
const Tunables &tunables();
int env_int(const char *name, int dflt, int lo, int hi);

// Driver entry points, resolved through cudaGetDriverEntryPoint so the library carries no
// link-time dependency on libcuda (it must load on a machine without a driver).
struct DriverApi {
    CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             CUstream, void **, void **) = nullptr;
    CUresult (*LaunchKernelEx)(const CUlaunchConfig *, CUfunction, void **, void **) = nullptr;  // optional
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    CUresult (*OccupancyMaxActiveBlocks)(int *, CUfunction, int, size_t) = nullptr;
    bool loaded = false;
};

// NVRTC: full translation unit for (dtype, kind, programs) and its compilation to an sm_100a cubin
std::string build_kernel_source(int32_t dtype, int32_t kind, const cb_node *const *progs, const int32_t *n_nodes,
                                int32_t n_progs, const Tunables &t);
int32_t compile_cubin(const std::string &source, std::string *cubin, std::string *log);

}  // namespace cb

struct cb_device {
    int ordinal = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;       // compute stream (the reference's `stream`)
    cudaStream_t copy_stream = nullptr;  // transfer stream (the reference's `mem_transfer_stream`)
    cudaMemPool_t pool = nullptr;
    cb::DriverApi drv;
    uint64_t launches = 0;
    bool capturing = false;

    // kernel cache keyed by the IR hash; a bucket holds every expression with that hash and a hit is
    // confirmed against the stored IR bytes, so a 64-bit collision costs a compile, never a wrong kernel
    std::unordered_map<uint64_t, std::vector<std::unique_ptr<cb_expr>>> exprs;

    struct Slot {
        uint64_t ptr;
        size_t bytes;
    };
    std::unordered_map<uint64_t, Slot> cache_slots;  // Cached module: cursor -> allocation

    // pinned staging for pageable host memory
    void *stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;

    // host-operand pipeline (cb_apply_host): a ring of device chunks and a third stream for D2H
    static constexpr int kPipeSlots = 8;  // capacity; `pipe_slots` (CB_PIPE_SLOTS, default 3) are in use
    int pipe_slots = 3;
    cudaStream_t d2h_stream = nullptr;
    size_t pipe_chunk_bytes = 0;
    void *pipe_in[kPipeSlots] = {};
    void *pipe_out[kPipeSlots] = {};
    cudaEvent_t pipe_h2d[kPipeSlots] = {};
    cudaEvent_t pipe_k[kPipeSlots] = {};
    cudaEvent_t pipe_d2h[kPipeSlots] = {};

    // reduction scratch
    void *sum_partials = nullptr;  // kSumMaxBlocks x 8 bytes
    unsigned long long *lut_counters = nullptr;  // tile / finished-block counters of lut16_kernel, zero between launches
    unsigned int *sum_ticket = nullptr;  // two device counters of the one-launch sum (last block folds); zero between launches
    // programmatic dependent launch of back-to-back sums: every entry point that enqueues work bumps op_seq in use();
    // a sum may start under the tail of the previous kernel only if that kernel was the previous sum
    mutable uint64_t op_seq = 0;
    uint64_t sum_seq_mark = ~0ull;
    uint64_t sum_calls = 0;
    uint64_t last_sum_out = 0;  // where the previous sum writes its scalar: a sum that READS it must wait for it
    cb::SumPdl next_sum_pdl(uint64_t in, size_t in_bytes, uint64_t out);
    void *sum_scalar = nullptr;    // 8 bytes device
    void *sum_host = nullptr;      // 8 bytes pinned

    cb::LaunchCtx ctx() const
    {
        return cb::LaunchCtx{stream, sm_count * cb::tunables().blocks_per_sm * cb::tunables().waves, cb::tunables().pdl != 0};
    }
    int32_t cuda_fail(cudaError_t e, const char *what) const;
    int32_t drv_fail(CUresult r, const char *what) const;
    int32_t use() const;  // cudaSetDevice
};
