/*
 * custos_b200.h — the C ABI of the Blackwell-native custos backend.
 *
 * This is the drop-in boundary for the data-parallel hot path of
 * elftausend/custos (element-wise unary/binary ops, fused unary chains,
 * unary gradients, clear/copy, sum/mean).  A Rust `CUDA<Mods>` device would
 * bind exactly these entry points in place of the driver/NVRTC FFI it calls
 * today (reference: src/devices/cuda/api/ffi.rs:165-263 and
 * src/devices/cuda/api/nvrtc/ffi.rs:36-54).  INTEGRATION.md shows the
 * `extern "C"` block a maintainer would add.
 *
 * Conventions
 *  - every function returns an int32 status: 0 = CB_OK, otherwise a cb_status
 *    (values >= 1000 are `1000 + cudaError_t`, >= 2000 are `2000 + ncclResult_t`,
 *    >= 3000 are `3000 + nvrtcResult`); the text is available through
 *    cb_last_error() (thread local).  This mirrors `type Error = i32` of the
 *    reference device (src/devices/cuda/cuda.rs:70-73).
 *  - device memory is passed as raw 64-bit device addresses, exactly like
 *    `CUDAPtr.ptr` (src/devices/cuda/cuda_ptr.rs:7-15).
 *  - lengths are element counts (size_t, 64-bit indexing inside the kernels);
 *    the reference passes `usize` but its kernels declare `int`
 *    (src/devices/cuda/ops.rs:155,174).
 *  - a device handle is single-threaded, like the reference's RefCell-based
 *    device (src/devices/cuda/cuda_device.rs:14-29).
 *  - there is no CPU fallback anywhere behind this header: if no CUDA device
 *    is usable every compute entry point returns an error.
 */
#ifndef CUSTOS_B200_H
#define CUSTOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_ABI_VERSION 1

/* ------------------------------------------------------------------ status */
typedef enum cb_status {
    CB_OK = 0,
    CB_ERR_INVALID_ARG = 1,
    CB_ERR_ZERO_LENGTH = 2,     /* DeviceError::ZeroLengthBuffer / CudaErrorKind::InvalidAllocSize
                                   (src/devices/cuda/api/cuda.rs:69-71, src/devices/cpu/cpu_device.rs:135-137) */
    CB_ERR_NO_DEVICE = 3,       /* no usable CUDA device: never falls back to the CPU */
    CB_ERR_UNSUPPORTED = 4,
    CB_ERR_EXPR = 5,            /* malformed expression IR */
    CB_ERR_INVALID_LAZY_BUF = 6,/* DeviceError::InvalidLazyBuf (src/error.rs:29-56) */
    CB_ERR_MISSING_CACHE_TRACES = 7, /* DeviceError::MissingCacheTraces */
    CB_ERR_GRAPH_OPTIMIZATION = 8,   /* DeviceError::GraphOptimization */
    CB_ERR_SHAPE = 9,
    CB_ERR_STATE = 10,
    CB_ERR_TYPE_MISMATCH = 11,  /* Untyped: `matches_storage_type` failed (src/devices/untyped/storages.rs) */
    CB_ERR_PARSE = 12,          /* malformed serialised buffer */
    CB_ERR_CUDA = 1000,
    CB_ERR_NCCL = 2000,
    CB_ERR_NVRTC = 3000
} cb_status;

const char *cb_last_error(void);
int32_t cb_abi_version(void);

/* ------------------------------------------------------------------ dtypes */
/* Every CDatatype (src/devices/cdatatype.rs:3-62).  Values 0..6 are stable since ABI 1; 7..12
 * were appended later, which is why the list is not sorted by width. */
typedef enum cb_dtype {
    CB_F32 = 0,
    CB_F64 = 1,
    CB_F16 = 2,   /* IEEE binary16 storage; arithmetic in f32 with a round-to-nearest-even
                     after EVERY op, like `half` on the reference CPU device
                     (src/number.rs:543-608) */
    CB_I32 = 3,
    CB_I64 = 4,
    CB_U32 = 5,
    CB_U8  = 6,
    CB_BF16 = 7,  /* bfloat16 storage, same per-op f32 arithmetic + RNE (number.rs:611-676).  The
                     reference names it "half" in kernel source (cdatatype.rs:58-62, its own TODO
                     says that is wrong); this backend follows the CPU device, i.e. real bf16 */
    CB_I8  = 8,
    CB_I16 = 9,
    CB_U16 = 10,
    CB_U64 = 11,
    CB_BOOL = 12, /* one byte per element; storage only (alloc / clear / copy / read / write):
                     bool has no Number impl, so no expression can be built over it */
    CB_DTYPE_COUNT = 13
} cb_dtype;

size_t cb_dtype_size(int32_t dtype);

/* ------------------------------------------------- expression IR (two_way_ops) */
/* One node of a `Combiner` tree (src/two_way_ops/combiner.rs:8-116).  A program
 * is an array of nodes in topological order (operands before users); the last
 * node is the value of the expression.  `a`/`b` index earlier nodes. */
typedef enum cb_opcode {
    CB_OP_X = 0,        /* first marker / seed value: Resolve (resolve.rs:25-30)            */
    CB_OP_Y = 1,        /* second marker (two-argument closures, mod.rs:96-104)             */
    CB_OP_CONST = 2,    /* numeric literal operand                                          */
    CB_OP_ADD = 3,      /* ops.rs:53-103   "(a + b)"                                        */
    CB_OP_MUL = 4,      /* ops.rs:14-58    "(a * b)"                                        */
    CB_OP_SUB = 5,      /* ops.rs:105-148  "(a - b)"                                        */
    CB_OP_DIV = 6,      /* ops.rs:150-193  "(a / b)"                                        */
    CB_OP_POW = 7,      /* ops.rs:195-238  "pow(a, b)"                                      */
    CB_OP_MIN = 8,      /* ops.rs:240-276  "min(a, b)"  eval: if a < b {a} else {b}         */
    CB_OP_MAX = 9,      /* ops.rs:278-314  "max(a, b)"  eval: if a > b {a} else {b}         */
    CB_OP_SIN = 10,     /* ops/unary.rs                                                     */
    CB_OP_COS = 11,
    CB_OP_TAN = 12,     /* f16: evaluates cos() on the reference CPU (number.rs:575-577)    */
    CB_OP_TANH = 13,
    CB_OP_EXP = 14,
    CB_OP_LN = 15,      /* "log(a)"                                                         */
    CB_OP_ABS = 16,
    CB_OP_NEG = 17,     /* "-(a)"                                                           */
    CB_OP_IDENTITY = 18,
    CB_OP_GEQ = 19,     /* ops/cmps.rs:42-47   "(a >= b)" -> 0/1 in T                       */
    CB_OP_LEQ = 20,     /* ops/cmps.rs:87-92   "(a <= b)"                                   */
    CB_OP_EQ = 21,      /* ops/cmps.rs:132-137 source "(a == b)", but eval is a <= b (:135);
                           this backend follows the CPU eval                                */
    CB_OP_COUNT = 22
} cb_opcode;

typedef struct cb_node {
    int32_t op;   /* cb_opcode */
    int32_t a;    /* operand node index, -1 if unused */
    int32_t b;    /* operand node index, -1 if unused */
    int32_t _pad;
    double  fimm; /* CB_OP_CONST for float dtypes (exact for f16/f32/f64 literals) */
    int64_t iimm; /* CB_OP_CONST for integer dtypes */
} cb_node;

/* `to_cl_source()` of the tree, byte-for-byte in the reference's format
 * (src/two_way_ops/to_cl_source.rs:7-12 + the per-op format strings); markers
 * are given by the caller ("x", "x[idx]", "lhs[idx]" ...).  Writes a
 * NUL-terminated string, returns CB_ERR_INVALID_ARG if cap is too small. */
int32_t cb_expr_to_cl_source(int32_t dtype, const cb_node *nodes, int32_t n_nodes,
                             const char *marker_x, const char *marker_y,
                             char *out, size_t cap);

/* `operations_to_fused_src` (src/devices/fusing.rs:4-19): "x = <src>;\n" per op. */
int32_t cb_ops_to_fused_src(int32_t dtype, const cb_node *const *progs, const int32_t *n_nodes,
                            int32_t n_progs, char *out, size_t cap);

/* The sm_100a CUDA source this backend generates for a chain of programs
 * (typed literals, no double promotion, no FMA contraction).  kind: see cb_kernel_kind. */
typedef enum cb_kernel_kind {
    CB_KERNEL_APPLY = 0,       /* out[i] = fN(...f1(in[i]))              K1/K2, a1/a8 */
    CB_KERNEL_UNARY_GRAD = 1,  /* lhs_grad[i] += out_grad[i] * g(lhs[i]) K3, a2       */
    CB_KERNEL_BINARY = 2,      /* out[i] = f(lhs[i], rhs[i])             a5 (+ f2)    */
    CB_KERNEL_CHAIN_GRAD = 3   /* the backward of a fused chain of K unary_ew ops in ONE kernel:
                                  x_grad[i] += (((out_grad[i] * gK(x_{K-1})) ...) * g1(x_0)), x_k = f_k(x_{k-1})
                                  recomputed from x_0 = x[i] in registers.  Programs: f_1..f_K, then g_1..g_K
                                  (n_progs = 2K).  Same roundings, in the same order, as the K add_unary_grad
                                  calls the tape would replay (src/unary.rs:118-128,
                                  src/modules/autograd/tape.rs:39-47) with zeroed intermediate gradients */
} cb_kernel_kind;

int32_t cb_expr_cuda_source(int32_t dtype, int32_t kind, const cb_node *const *progs,
                            const int32_t *n_nodes, int32_t n_progs, char *out, size_t cap);

/* Compile (NVRTC, --gpu-architecture=sm_100a --fmad=false, no fast-math) without a
 * device: returns the cubin size.  Used by the build check and the CPU-side tests. */
int32_t cb_expr_compile_check(int32_t dtype, int32_t kind, const cb_node *const *progs,
                              const int32_t *n_nodes, int32_t n_progs, size_t *cubin_bytes);

/* ------------------------------------------------------------------ device */
typedef struct cb_device cb_device;
typedef struct cb_expr cb_expr;     /* a compiled (chain of) expression(s) */
typedef struct cb_graph cb_graph;   /* an instantiated CUDA graph */

/* CUDA::new(idx) (src/devices/cuda/cuda.rs:53-67).  ordinal < 0 = honour
 * CUSTOS_CU_DEVICE_IDX (src/devices/cuda/mod.rs:35-42), default 0. */
int32_t cb_device_create(int32_t ordinal, cb_device **out);
int32_t cb_device_destroy(cb_device *dev);
int32_t cb_device_ordinal(cb_device *dev, int32_t *ordinal);
int32_t cb_device_sm_count(cb_device *dev, int32_t *sms);
/* raw cudaStream_t of the compute stream (for event timing by the harness) */
int32_t cb_device_stream(cb_device *dev, void **stream);
int32_t cb_sync(cb_device *dev);

/* Alloc::alloc (src/devices/cuda/cuda.rs:114-123, cuda_ptr.rs:21-30).  zero != 0
 * gives the zero-initialised memory the CPU device hands out
 * (src/devices/cpu/cpu_ptr.rs:76-89); gradient buffers rely on it.
 * bytes == 0 -> CB_ERR_ZERO_LENGTH.  Memory comes from a stream-ordered pool
 * and is 256-byte aligned. */
int32_t cb_alloc(cb_device *dev, size_t bytes, int32_t zero, uint64_t *dptr);
int32_t cb_free(cb_device *dev, uint64_t dptr);          /* CUDAPtr::drop (cuda_ptr.rs:64-77) */
int32_t cb_mem_info(cb_device *dev, size_t *pool_reserved, size_t *pool_used);

/* Cached module slot (src/modules/cached.rs:173-229): the k-th retrieve of a loop
 * iteration returns the k-th allocation.  *hit is 1 when the slot existed. */
int32_t cb_cache_retrieve(cb_device *dev, uint64_t cursor, size_t bytes, uint64_t *dptr, int32_t *hit);
int32_t cb_cache_clear(cb_device *dev);

/* Read / WriteBuf::write / alloc_from_slice (src/devices/cuda/ops.rs:17-49,101-105,
 * cuda.rs:124-137).  Host pointers may be pageable; transfers are staged through
 * pinned buffers owned by the device.  cb_d2h synchronises.  cb_h2d returns once the source may
 * be reused: a pageable source has been copied into the staging buffers by then, a PINNED source
 * (cb_host_alloc) is read by the DMA engine directly, so cb_h2d waits for that copy to finish
 * (use cb_h2d_async to overlap it and synchronise yourself). */
int32_t cb_h2d(cb_device *dev, uint64_t dst, const void *src, size_t bytes);
int32_t cb_d2h(cb_device *dev, void *dst, uint64_t src, size_t bytes);
/* pinned host memory for zero-copy-staging callers */
int32_t cb_host_alloc(size_t bytes, void **out);
/* CB_HOST_WRITE_COMBINED: not cached by the CPU — the host writes it at full speed, reads it very slowly, and the
 * DMA engine reads it without snooping the CPU caches: for buffers the host only fills and the device only reads */
#define CB_HOST_WRITE_COMBINED 1u
int32_t cb_host_alloc_ex(size_t bytes, uint32_t flags, void **out);
int32_t cb_host_free(void *p);
/* asynchronous variants on the compute stream; host memory must be pinned */
int32_t cb_h2d_async(cb_device *dev, uint64_t dst, const void *src, size_t bytes);
int32_t cb_d2h_async(cb_device *dev, void *dst, uint64_t src, size_t bytes);

/* CopySlice::copy_slice_to / WriteBuf::write_buf / CloneBuf (src/devices/cuda/ops.rs:65-116,
 * cuda.rs:152-164): device to device copy of n elements with element offsets. */
int32_t cb_copy(cb_device *dev, int32_t dtype, uint64_t dst, size_t dst_off,
                uint64_t src, size_t src_off, size_t n);
/* ClearBuf::clear / ZeroGrad::zero_grad (src/devices/cuda/mod.rs:59-74) */
int32_t cb_clear(cb_device *dev, int32_t dtype, uint64_t dptr, size_t n);
/* device-side seed for backward(): replaces `vec![T::one(); len]` + write
 * (src/buffer/impl_autograd.rs:32, src/modules/autograd/tape.rs:53-64) */
int32_t cb_fill(cb_device *dev, int32_t dtype, uint64_t dptr, size_t n, double fvalue, int64_t ivalue);

/* ------------------------------------------------- compiled expression kernels */
/* Compile one expression (n_progs = 1, unfused apply_fn: K1) or a chain of unary
 * expressions applied in order (fused chain: K2, src/devices/cuda/fusing.rs:20-50)
 * into ONE sm_100a kernel.  Cached by (dtype, kind, IR hash) per device. */
int32_t cb_expr_compile(cb_device *dev, int32_t dtype, int32_t kind,
                        const cb_node *const *progs, const int32_t *n_nodes, int32_t n_progs,
                        cb_expr **out);
int32_t cb_expr_release(cb_expr *e);   /* expressions are owned by the device cache; no-op kept for symmetry */
/* f16 / bf16 apply expressions carry a 65 536-entry table of the chain's results (filled by the expression's own
 * arithmetic kernel at compile time); cb_apply of >= CB_LUT16_MIN_ELEMS (2^22) 16-byte-aligned elements runs as a
 * shared-memory table lookup — same bits, HBM-bound instead of FP32-pipe-bound.  f16 / bf16 UNARY_GRAD and
 * CHAIN_GRAD expressions carry the table of their backward term for out_grad = 1: cb_unary_grad_ex with
 * CB_GRAD_SEED_ONES on a large buffer is then one lookup and one 16-bit add per element (a single-op UNARY_GRAD
 * also with a general out_grad: lookup of g(lhs), 16-bit multiply and add).  Off per expression with
 * cb_expr_set_lookup(e, 0), off for the process with CB_LUT16=0. */
int32_t cb_expr_set_lookup(cb_expr *e, int32_t enabled);
int32_t cb_expr_has_lookup(cb_expr *e, int32_t *flag);

/* ApplyFunction::apply_fn body (src/devices/cuda/ops.rs:144-177) and the fused
 * chain body (src/devices/cuda/fusing.rs:28-49): out[i] = f(in[i]). */
int32_t cb_apply(cb_device *dev, cb_expr *f, uint64_t in, uint64_t out, size_t n);
/* UnaryGrad::add_unary_grad body (src/devices/cuda/ops.rs:196-234):
 * lhs_grad[i] += out_grad[i] * g(lhs[i]) — multiply then add, two roundings
 * (src/devices/cpu_stack_ops.rs:28). */
int32_t cb_unary_grad(cb_device *dev, cb_expr *g, uint64_t lhs, uint64_t lhs_grad,
                      uint64_t out_grad, size_t n);
/* The same for expressions of kind CB_KERNEL_UNARY_GRAD or CB_KERNEL_CHAIN_GRAD, with flags:
 * CB_GRAD_SEED_ONES — out_grad is not read but WRITTEN with T::one() and the kernel computes with 1: the seed of
 * `backward()` (`vec![T::one(); len]`, src/buffer/impl_autograd.rs:32) folded into the first grad kernel of the
 * replay instead of a separate fill pass over the buffer. */
#define CB_GRAD_SEED_ONES 1u
int32_t cb_unary_grad_ex(cb_device *dev, cb_expr *g, uint64_t lhs, uint64_t lhs_grad,
                         uint64_t out_grad, size_t n, uint32_t flags);
/* two-marker expression: out[i] = f(lhs[i], rhs[i]) */
int32_t cb_apply2(cb_device *dev, cb_expr *f, uint64_t lhs, uint64_t rhs, uint64_t out, size_t n);

/* Host-resident operands: out_host[i] = f(in_host[i]).  The reference's end-to-end path is
 * alloc_from_slice (pageable H2D, src/devices/cuda/cuda.rs:124-137) -> kernel -> read (D2H + two stream
 * syncs, src/devices/cuda/ops.rs:32-48), strictly one after the other.  Here the buffer is cut into
 * chunks that flow through three streams, so H2D of chunk i+1, the kernel of chunk i and D2H of chunk
 * i-1 overlap (full-duplex PCIe).  Pinned host memory (cb_host_alloc) gets the overlapped path,
 * pageable memory falls back to the staged serial one.  Synchronises before returning. */
int32_t cb_apply_host(cb_device *dev, cb_expr *f, const void *in_host, void *out_host, size_t n);

/* Binary element-wise add/mul/sub/div — the `AddEw`/`MulBuf` pattern
 * (README.md:96-122, tests/demo_impl/cuda/mod.rs:3-36, src/lib.rs:293-301). */
typedef enum cb_binop { CB_BIN_ADD = 0, CB_BIN_MUL = 1, CB_BIN_SUB = 2, CB_BIN_DIV = 3 } cb_binop;
int32_t cb_binary(cb_device *dev, int32_t dtype, int32_t op, uint64_t lhs, uint64_t rhs,
                  uint64_t out, size_t n);

/* ------------------------------------------------------------- reductions */
/* Deterministic sum in a fixed two-level order (no atomics on data, fixed grid): level 1 = per-block tree
 * over a contiguous chunk, level 2 = the partials folded in index order — both in ONE launch: the block that
 * finishes last folds (which block that is does not matter to the result).
 * The result is written to a device scalar of the accumulation type
 * (f32 -> f32, f64 -> f64, f16 -> f32, integers -> i64) at `out`.
 * The reference has no sum; the order is defined in DESIGN.md. */
int32_t cb_sum(cb_device *dev, int32_t dtype, uint64_t in, size_t n, uint64_t out);
int32_t cb_mean(cb_device *dev, int32_t dtype, uint64_t in, size_t n, uint64_t out);
/* convenience: reduce and copy the scalar to the host (synchronises) */
int32_t cb_sum_host(cb_device *dev, int32_t dtype, uint64_t in, size_t n, void *host_out);
int32_t cb_mean_host(cb_device *dev, int32_t dtype, uint64_t in, size_t n, void *host_out);
/* the fixed (hardware independent) reduction order for n elements: pass 1 runs
 * `blocks` blocks of `threads` threads over chunks of `chunk` elements reading
 * `vec` elements per access; pass 2 is one block of `threads2` threads. */
int32_t cb_sum_plan(int32_t dtype, size_t n, int32_t *blocks, size_t *chunk, int32_t *threads,
                    int32_t *vec, int32_t *threads2);

/* ------------------------------------------------ CUDA graph capture / replay */
/* LazyCudaGraph (src/devices/cuda/lazy.rs:10-72): capture the launches issued on
 * the compute stream between begin/end, instantiate once, replay many times. */
int32_t cb_graph_begin(cb_device *dev);
int32_t cb_graph_end(cb_device *dev, cb_graph **out);
int32_t cb_graph_launch(cb_device *dev, cb_graph *g);
int32_t cb_graph_destroy(cb_graph *g);
int32_t cb_graph_node_count(cb_graph *g, size_t *kernel_nodes);

/* CUDA events on the compute stream, for device-side timing by the harness
 * (torch.cuda.Event only sees torch's own streams) */
typedef struct cb_event cb_event;
int32_t cb_event_create(cb_device *dev, cb_event **out);
int32_t cb_event_record(cb_device *dev, cb_event *ev);
int32_t cb_event_sync(cb_event *ev);
int32_t cb_event_elapsed_ms(cb_event *start, cb_event *end, float *ms);
int32_t cb_event_destroy(cb_event *ev);

/* kernels launched by this library on this device since creation (the harness's
 * `gpu_launches` evidence) */
int32_t cb_launch_count(cb_device *dev, uint64_t *count);

/* ------------------------------------------------------ multi-GPU (NCCL) */
/* One process per GPU.  Element-wise work needs no communication; only the
 * reduction totals are exchanged — one scalar per rank, folded in rank order on every
 * rank (deterministic).  The exchange is fused into the last reduction kernel: it writes
 * the rank's total into every peer's buffer over NVLink (peer memory mapped through CUDA
 * IPC, system-scope release/acquire) and waits for the peers' totals; NCCL sets the
 * mapping up and remains the fallback (all-gather + fold). */
typedef struct cb_comm cb_comm;
#define CB_COMM_ID_BYTES 128
int32_t cb_comm_unique_id(uint8_t id[CB_COMM_ID_BYTES]);
int32_t cb_comm_create(cb_device *dev, int32_t n_ranks, int32_t rank,
                       const uint8_t id[CB_COMM_ID_BYTES], cb_comm **out);
int32_t cb_comm_destroy(cb_comm *c);
/* 1 when the totals are exchanged by the fused reduce+exchange kernel over NVLink peer memory (CUDA IPC),
 * 0 when the communicator fell back to ncclAllGather + fold (CB_COMM_P2P=0 forces the fallback) */
int32_t cb_comm_uses_peer_memory(cb_comm *c, int32_t *flag);
/* local deterministic sum of `in`, then exchange + rank-ordered fold, all in one kernel; out = device scalar */
int32_t cb_comm_sum(cb_comm *c, int32_t dtype, uint64_t in, size_t n_local, uint64_t out);
int32_t cb_comm_mean(cb_comm *c, int32_t dtype, uint64_t in, size_t n_local, size_t n_global,
                     uint64_t out);
/* the same, copied to a host scalar of the accumulation type (synchronises) */
int32_t cb_comm_sum_host(cb_comm *c, int32_t dtype, uint64_t in, size_t n_local, void *host_out);
int32_t cb_comm_mean_host(cb_comm *c, int32_t dtype, uint64_t in, size_t n_local, size_t n_global,
                          void *host_out);
/* Synchronises the device stream and reports a failed exchange: a peer that did not reach a cb_comm_sum / mean
 * within CB_COMM_TIMEOUT_MS (default 2000) makes the kernel give up instead of hanging the GPU, and the value it
 * wrote is then NOT the global sum -> CB_ERR_STATE here, from every later cb_comm_* call and from the *_host forms.
 * Call it before trusting a device scalar written by cb_comm_sum / cb_comm_mean. */
int32_t cb_comm_check(cb_comm *c);
int32_t cb_comm_rank(cb_comm *c, int32_t *rank, int32_t *n_ranks);
int32_t cb_comm_device(cb_comm *c, cb_device **dev);
/* contiguous slice [begin, end) of rank r out of n_ranks over n elements, 16-byte aligned starts */
int32_t cb_shard_range(size_t n, int32_t elem_bytes, int32_t n_ranks, int32_t rank,
                       size_t *begin, size_t *end);

/* ================================================================ module layer
 * A C++ restatement of custos' module stack around the device above
 * (Base, Cached, Lazy, Graph, Autograd: the .rs files under src/modules), exported so that the
 * parity tests can drive `CUDA<Graph<Lazy<Autograd<Base>>>>`-style devices
 * without a Rust toolchain.  A Rust build would keep its own module layer and
 * bind only the functions above. */
typedef struct cbm_device cbm_device;
typedef uint64_t cbm_buf;   /* buffer handle (not a device address) */

#define CBM_BASE     0u
#define CBM_CACHED   1u     /* Cached<..>   src/modules/cached.rs   */
#define CBM_LAZY     2u     /* Lazy<..>     src/modules/lazy.rs     */
#define CBM_GRAPH    4u     /* Graph<..>    src/modules/graph.rs    */
#define CBM_AUTOGRAD 8u     /* Autograd<..> src/modules/autograd.rs */

int32_t cbm_device_create(int32_t ordinal, uint32_t modules, int32_t dtype, cbm_device **out);
int32_t cbm_device_destroy(cbm_device *d);
int32_t cbm_device_raw(cbm_device *d, cb_device **raw);

/* Sharded device (one process per GPU; the reference is single-device, src/devices/cuda/cuda.rs:53-67,187-202).
 * After cbm_device_set_comm (the communicator must have been created on cbm_device_raw(d)) the *_sharded
 * constructors give every rank the contiguous slice cb_shard_range assigns it; every operator of this header then
 * works on the slice unchanged and without communication, results inherit the slice of their parents, and
 * cbm_sum / cbm_mean of a sharded buffer return the GLOBAL value — the rank's deterministic partial, one scalar per
 * rank exchanged over NVLink, folded in rank order: identical bits on every rank.  All ranks must call cbm_sum /
 * cbm_mean of sharded buffers in the same order (they are collective).  The communicator stays owned by the caller:
 * it must outlive the device (or be detached with cbm_device_set_comm(d, NULL) first); cbm_device_destroy does not
 * destroy it. */
int32_t cbm_device_set_comm(cbm_device *d, cb_comm *comm);
int32_t cbm_buffer_new_sharded(cbm_device *d, int32_t dtype, size_t global_len, cbm_buf *out);
/* `global_data` is the whole host array (global_len elements); only this rank's slice is read and uploaded */
int32_t cbm_buffer_from_host_sharded(cbm_device *d, int32_t dtype, const void *global_data, size_t global_len,
                                     cbm_buf *out);
/* [begin, end) of the global buffer this handle holds, and the global length (an unsharded buffer: [0, len), len) */
int32_t cbm_buffer_shard(cbm_device *d, cbm_buf b, size_t *begin, size_t *end, size_t *global_len);

/* Buffer::new / device.buffer([..]) (src/buffer.rs:80-93): allocated immediately, zeroed */
int32_t cbm_buffer_new(cbm_device *d, int32_t dtype, size_t len, cbm_buf *out);
int32_t cbm_buffer_from_host(cbm_device *d, int32_t dtype, const void *data, size_t len, cbm_buf *out);
int32_t cbm_buffer_drop(cbm_device *d, cbm_buf b);
int32_t cbm_buffer_len(cbm_device *d, cbm_buf b, size_t *len);
/* Buffer::replace().read(): resolves lazily retrieved buffers (src/modules/lazy.rs:430-455) */
int32_t cbm_buffer_read(cbm_device *d, cbm_buf b, void *host_out, size_t len);
int32_t cbm_buffer_write(cbm_device *d, cbm_buf b, const void *host_in, size_t len);
/* device address behind the handle after replace(); 0 when a lazy buffer is not allocated yet */
int32_t cbm_buffer_ptr(cbm_device *d, cbm_buf b, uint64_t *dptr);
/* HasId::id(): graph level id (cursor for retrieved buffers, address for eager ones) */
int32_t cbm_buffer_id(cbm_device *d, cbm_buf b, uint64_t *id);
int32_t cbm_buffer_require_grad(cbm_device *d, cbm_buf b);   /* Buffer::require_grad */
int32_t cbm_buffer_requires_grad(cbm_device *d, cbm_buf b, int32_t *flag);
int32_t cbm_buffer_checkpoint(cbm_device *d, cbm_buf b);     /* Buffer::checkpoint (src/buffer.rs:131-137) */

/* Retriever::retrieve (src/devices.rs:172-186) */
int32_t cbm_retrieve(cbm_device *d, int32_t dtype, size_t len, const cbm_buf *parents,
                     int32_t n_parents, cbm_buf *out);
/* ApplyFunction::apply_fn (src/unary.rs:7-28) */
int32_t cbm_apply_fn(cbm_device *d, cbm_buf in, const cb_node *nodes, int32_t n_nodes, cbm_buf *out);
/* UnaryGrad::add_unary_grad (src/unary.rs:31-58) */
int32_t cbm_add_unary_grad(cbm_device *d, cbm_buf lhs, cbm_buf lhs_grad, cbm_buf out_grad,
                           const cb_node *nodes, int32_t n_nodes);
/* UnaryElementWiseMayGrad::unary_ew (src/unary.rs:62-132) */
int32_t cbm_unary_ew(cbm_device *d, cbm_buf in, const cb_node *fwd, int32_t n_fwd,
                     const cb_node *grad, int32_t n_grad, cbm_buf *out);
/* binary element-wise op with the retrieve + add_op pattern of README.md:96-122 */
int32_t cbm_binary(cbm_device *d, int32_t op, cbm_buf lhs, cbm_buf rhs, cbm_buf *out);
/* the same op into a buffer the caller owns (the reference tests' own kernels: `launch_kernel1d(..,
 * &[&lhs, &rhs, &mut out, &len])`, src/devices/cuda/lazy.rs:96-141): recorded under Lazy, no retrieve */
int32_t cbm_binary_into(cbm_device *d, int32_t op, cbm_buf lhs, cbm_buf rhs, cbm_buf out);
int32_t cbm_clear(cbm_device *d, cbm_buf b);                  /* ClearBuf::clear, eager like the reference */
/* `add_op(&mut out, |out, _| out.clear())` (src/modules/lazy.rs:733-738): a clear that Lazy records */
int32_t cbm_clear_op(cbm_device *d, cbm_buf b);
int32_t cbm_copy_slice(cbm_device *d, cbm_buf src, size_t src_off, cbm_buf dst, size_t dst_off, size_t n);
int32_t cbm_clone_buf(cbm_device *d, cbm_buf src, cbm_buf *out);
/* sum / mean of a buffer to a host scalar of the accumulation type (executes now) */
int32_t cbm_sum(cbm_device *d, cbm_buf b, void *host_out);
int32_t cbm_mean(cbm_device *d, cbm_buf b, void *host_out);

/* Lazy: Run::run, ExecNow (src/modules/lazy.rs:143-198, src/features.rs:505-518) */
int32_t cbm_run(cbm_device *d);
int32_t cbm_exec_now(cbm_device *d, size_t begin, size_t end);   /* end = SIZE_MAX: to the last op */
int32_t cbm_exec_last_n(cbm_device *d, size_t n);
int32_t cbm_ops_count(cbm_device *d, size_t *n);
int32_t cbm_alloc_later(cbm_device *d);
int32_t cbm_set_lazy_enabled(cbm_device *d, int32_t enabled);
/* op hint of recorded op i as reference source ("sin(x)"), "" if none (src/op_hint.rs:44-83) */
int32_t cbm_op_hint_src(cbm_device *d, size_t i, char *out, size_t cap);
/* the compiled expression behind recorded op i after the fusing passes (NULL for no-ops and the AOT kernels) */
int32_t cbm_op_expr(cbm_device *d, size_t i, cb_expr **out);
/* replay through one captured CUDA graph instead of re-launching (src/devices/cuda/lazy.rs:31-49) */
int32_t cbm_set_graph_replay(cbm_device *d, int32_t enabled);
int32_t cbm_replay_kernel_nodes(cbm_device *d, size_t *n);

/* Graph: Optimize (src/modules/graph.rs:79-113) */
int32_t cbm_optimize_mem_graph(cbm_device *d);
int32_t cbm_unary_fusing(cbm_device *d);
/* beyond the reference: splice producers that have a single reader into their consumer
 * when the merged expression reads at most two buffers — binary ops fuse with neighbouring unary chains */
int32_t cbm_elementwise_fusing(cbm_device *d);
/* cache traces of the current graph: flattened as [cache_idx, k, use_0..use_{k-1}]* */
int32_t cbm_cache_traces(cbm_device *d, int64_t *out, size_t cap, size_t *written);

/* Cached: Cursor + range (src/features.rs:68-111, src/range.rs:10-60) */
int32_t cbm_cursor(cbm_device *d, uint64_t *cursor);
int32_t cbm_set_cursor(cbm_device *d, uint64_t cursor);

/* Autograd (src/buffer/impl_autograd.rs:20-197, src/modules/autograd/tape.rs:39-83) */
int32_t cbm_backward(cbm_device *d, cbm_buf out);
int32_t cbm_backward_with(cbm_device *d, cbm_buf out, const void *seed, size_t len);
int32_t cbm_grad(cbm_device *d, cbm_buf b, cbm_buf *grad);     /* Buffer::grad: allocates (zeroed) on first use */
int32_t cbm_zero_grad(cbm_device *d);
int32_t cbm_set_grad_enabled(cbm_device *d, int32_t enabled);  /* Autograd::{enable,disable}_grad */

/* -------------------------------------------- Untyped buffers and serde (f4) */
/* src/devices/untyped/: a buffer whose element type is a run-time tag (`UntypedData` = CpuStorage /
 * CudaStorage enums, storages.rs) instead of a type parameter.  Every cbm_buf already carries its
 * cb_dtype, so "untyped" here is a view: the tag can be queried, and typed access checks it the way
 * `to_typed` / `as_typed` / `read_typed` do (mod.rs:17-84: `None` on a mismatch). */
int32_t cbm_buffer_dtype(cbm_device *d, cbm_buf b, int32_t *dtype);           /* the storage tag */
/* AsType (matches_type.rs:28-71): the types an Untyped device accepts: u8, u32, i64, bf16, f16, f32, f64 */
int32_t cbm_untyped_supports(int32_t dtype);                                   /* 1 / 0 */
/* MatchesType::matches_storage_type: CB_OK or CB_ERR_TYPE_MISMATCH */
int32_t cbm_buffer_matches_type(cbm_device *d, cbm_buf b, int32_t dtype);
/* Buffer::read_typed::<OT>() (mod.rs:77-83) */
int32_t cbm_buffer_read_typed(cbm_device *d, cbm_buf b, int32_t dtype, void *host_out, size_t len);

/* serde of device buffers (src/devices/cuda/cuda_ptr.rs:122-157): a CUDAPtr<T> serialises as the
 * SEQUENCE of its elements (read back to the host first) and deserialises by allocating and writing.
 * serde is format-agnostic; two concrete encodings of that sequence are provided:
 *   CB_SER_JSON     what serde_json writes: "[1,2,3]", floats in ryu's shortest round-trip form
 *                   ("1.0", "0.1", "1e16", "1.5e-7"), non-finite floats as null (and, like serde_json,
 *                   a null does not deserialise into a float: CB_ERR_PARSE);
 *   CB_SER_BINCODE  what bincode 1.x (fixint, little endian) writes: u64 length, then the elements.
 * Only types that implement serde::Serialize in the reference build: the integers, f32, f64, bool
 * (half is built without its serde feature, Cargo.toml:36 -> f16 / bf16: CB_ERR_UNSUPPORTED).
 * Serialise: `*needed` receives the byte count; data is written only when cap >= *needed (call with
 * out = NULL to size the buffer).  JSON output is not NUL-terminated. */
typedef enum cb_ser_format { CB_SER_JSON = 0, CB_SER_BINCODE = 1 } cb_ser_format;
/* the codec alone, host memory to host memory (no device): `elems` are n elements of dtype */
int32_t cb_serde_encode(int32_t dtype, int32_t format, const void *elems, size_t n, void *out, size_t cap,
                        size_t *needed);
/* `*n` receives the element count; elements are written only when cap_elems >= *n */
int32_t cb_serde_decode(int32_t dtype, int32_t format, const void *in, size_t len, void *elems_out,
                        size_t cap_elems, size_t *n);
int32_t cbm_buffer_serialize(cbm_device *d, cbm_buf b, int32_t format, void *out, size_t cap, size_t *needed);
int32_t cbm_buffer_deserialize(cbm_device *d, int32_t dtype, int32_t format, const void *in, size_t len,
                               cbm_buf *out);

/* ------------------------------------------- OptGraph without a device (a10) */
/* src/modules/graph/opt_graph.rs:6-41 and opt_graph/optimize.rs:19-132 */
typedef struct cb_optgraph cb_optgraph;
int32_t cb_optgraph_create(cb_optgraph **out);
int32_t cb_optgraph_destroy(cb_optgraph *g);
int32_t cb_optgraph_add_leaf(cb_optgraph *g, size_t len, int64_t *idx);
int32_t cb_optgraph_add_node(cb_optgraph *g, size_t len, const int64_t *deps, int32_t n_deps, int64_t *idx);
int32_t cb_optgraph_set_skip(cb_optgraph *g, int64_t idx, int32_t skip);
int32_t cb_optgraph_is_path_optimizable(cb_optgraph *g, int64_t idx, int32_t *out);
int32_t cb_optgraph_trace_cache_path_raw(cb_optgraph *g, int64_t idx, int64_t *out, size_t cap, size_t *written);
int32_t cb_optgraph_cache_traces(cb_optgraph *g, int64_t *out, size_t cap, size_t *written);

#ifdef __cplusplus
}
#endif
#endif /* CUSTOS_B200_H */
