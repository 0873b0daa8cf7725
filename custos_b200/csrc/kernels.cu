// kernels.cu — the ahead-of-time compiled sm_100a kernels: clear / fill / copy, the binary
// element-wise ops, and the deterministic two-pass sum.  Compiled with
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (no fast-math)
// Everything here is HBM-bound streaming work: 128-bit coalesced accesses, persistent
// block-strided grids sized as a multiple of the SM count, 64-bit indexing, no shared
// memory except the block reduction, no tensor cores.
//
// Reference kernels these replace: `clear` (src/devices/cuda/mod.rs:59-74), the test-only
// `add`/`mul` element-wise kernels (tests/demo_impl/cuda/mod.rs:13-35, src/lib.rs:293-301),
// cuMemcpy D2D (src/devices/cuda/ops.rs:65-116).  Sum/mean have no reference kernel.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <type_traits>

#include "kernels.h"

namespace cb {
namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ 128-bit streaming IO
__device__ __forceinline__ uint4 ld16(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st16(uint4 *p, const uint4 &v)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

template <typename T>
union Pack {
    uint4 q;
    T v[16 / sizeof(T)];
};

// ------------------------------------------------------------------ element arithmetic
// f16 / bf16 are carried as their bit patterns; arithmetic is f32 with one RNE back (half crate).
struct half_bits {
    unsigned short b;
};
__device__ __forceinline__ float h2f(unsigned short h)
{
    float f;
    asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
    return f;
}
__device__ __forceinline__ unsigned short f2h(float f)
{
    unsigned short h;
    asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(f));
    return h;
}

// bf16: the upper half of an f32; half::bf16 arithmetic is f32 with one RNE back, like f16
struct bf16_bits {
    unsigned short b;
};
__device__ __forceinline__ float b2f(unsigned short h) { return __uint_as_float((unsigned int)h << 16); }
__device__ __forceinline__ unsigned short f2b(float f)
{
    unsigned short h;
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(f));
    return h;
}

template <typename T, int OP>
struct BinOp;
template <int OP>
struct BinOp<float, OP> {
    static __device__ __forceinline__ float apply(float a, float b)
    {
        if (OP == CB_BIN_ADD) return __fadd_rn(a, b);
        if (OP == CB_BIN_MUL) return __fmul_rn(a, b);
        if (OP == CB_BIN_SUB) return __fsub_rn(a, b);
        return __fdiv_rn(a, b);
    }
};
template <int OP>
struct BinOp<double, OP> {
    static __device__ __forceinline__ double apply(double a, double b)
    {
        if (OP == CB_BIN_ADD) return __dadd_rn(a, b);
        if (OP == CB_BIN_MUL) return __dmul_rn(a, b);
        if (OP == CB_BIN_SUB) return __dsub_rn(a, b);
        return __ddiv_rn(a, b);
    }
};
template <int OP>
struct BinOp<half_bits, OP> {
    static __device__ __forceinline__ half_bits apply(half_bits a, half_bits b)
    {
        return half_bits{f2h(BinOp<float, OP>::apply(h2f(a.b), h2f(b.b)))};
    }
};
template <int OP>
struct BinOp<bf16_bits, OP> {
    static __device__ __forceinline__ bf16_bits apply(bf16_bits a, bf16_bits b)
    {
        return bf16_bits{f2b(BinOp<float, OP>::apply(b2f(a.b), b2f(b.b)))};
    }
};
template <typename T, int OP>
struct BinOp {  // integers: wrapping, x / 0 = 0
    typedef typename std::make_unsigned<T>::type U;
    static __device__ __forceinline__ T apply(T a, T b)
    {
        if (OP == CB_BIN_ADD) return (T)((U)a + (U)b);
        if (OP == CB_BIN_MUL) return (T)((U)a * (U)b);
        if (OP == CB_BIN_SUB) return (T)((U)a - (U)b);
        if (b == (T)0) return (T)0;
        if ((T)-1 < (T)0 && b == (T)-1) return (T)((U)0 - (U)a);  // MIN / -1 wraps to MIN (the reference panics)
        return (T)(a / b);
    }
};

// ------------------------------------------------------------------ binary: out = lhs op rhs
// Algorithmic traffic: 3 * sizeof(T) per element.  UNROLL tiles of 16-byte units per thread.
template <typename T, int OP, int UNROLL>
__global__ void __launch_bounds__(kThreads) binary_vec_kernel(const T *lhs, const T *rhs, T *out, size_t n)
{
    constexpr int VEC = 16 / sizeof(T);
    const size_t nunits = n / VEC;
    const size_t tile_units = (size_t)kThreads * UNROLL;
    const size_t ntiles = nunits / tile_units;
    const uint4 *pl = reinterpret_cast<const uint4 *>(lhs);
    const uint4 *pr = reinterpret_cast<const uint4 *>(rhs);
    uint4 *po = reinterpret_cast<uint4 *>(out);
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t base = tile * tile_units + threadIdx.x;
        Pack<T> l[UNROLL], r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            l[u].q = ld16(pl + base + (size_t)u * kThreads);
            r[u].q = ld16(pr + base + (size_t)u * kThreads);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int j = 0; j < VEC; j++) l[u].v[j] = BinOp<T, OP>::apply(l[u].v[j], r[u].v[j]);
            st16(po + base + (size_t)u * kThreads, l[u].q);
        }
    }
    const size_t gid = (size_t)blockIdx.x * kThreads + threadIdx.x;
    const size_t gsz = (size_t)gridDim.x * kThreads;
    for (size_t u = ntiles * tile_units + gid; u < nunits; u += gsz) {
        Pack<T> l, r;
        l.q = ld16(pl + u);
        r.q = ld16(pr + u);
#pragma unroll
        for (int j = 0; j < VEC; j++) l.v[j] = BinOp<T, OP>::apply(l.v[j], r.v[j]);
        st16(po + u, l.q);
    }
    for (size_t i = nunits * VEC + gid; i < n; i += gsz) out[i] = BinOp<T, OP>::apply(lhs[i], rhs[i]);
}

template <typename T, int OP>
__global__ void __launch_bounds__(kThreads) binary_scalar_kernel(const T *lhs, const T *rhs, T *out, size_t n)
{
    const size_t gsz = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += gsz)
        out[i] = BinOp<T, OP>::apply(lhs[i], rhs[i]);
}

// ------------------------------------------------------------------ fill / clear: 1 write per element
__global__ void __launch_bounds__(kThreads) fill16_kernel(uint4 *out, size_t nunits, uint4 pattern)
{
    constexpr int UNROLL = 4;
    const size_t tile_units = (size_t)kThreads * UNROLL;
    const size_t ntiles = nunits / tile_units;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t base = tile * tile_units + threadIdx.x;
#pragma unroll
        for (int u = 0; u < UNROLL; u++) st16(out + base + (size_t)u * kThreads, pattern);
    }
    const size_t gid = (size_t)blockIdx.x * kThreads + threadIdx.x;
    const size_t gsz = (size_t)gridDim.x * kThreads;
    for (size_t u = ntiles * tile_units + gid; u < nunits; u += gsz) st16(out + u, pattern);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) fill_scalar_kernel(T *out, size_t n, T value)
{
    const size_t gsz = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += gsz) out[i] = value;
}

// ------------------------------------------------------------------ copy: 1 read + 1 write
__global__ void __launch_bounds__(kThreads) copy16_kernel(const uint4 *in, uint4 *out, size_t nunits)
{
    constexpr int UNROLL = 4;
    const size_t tile_units = (size_t)kThreads * UNROLL;
    const size_t ntiles = nunits / tile_units;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t base = tile * tile_units + threadIdx.x;
        uint4 r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) r[u] = ld16(in + base + (size_t)u * kThreads);
#pragma unroll
        for (int u = 0; u < UNROLL; u++) st16(out + base + (size_t)u * kThreads, r[u]);
    }
    const size_t gid = (size_t)blockIdx.x * kThreads + threadIdx.x;
    const size_t gsz = (size_t)gridDim.x * kThreads;
    for (size_t u = ntiles * tile_units + gid; u < nunits; u += gsz) st16(out + u, ld16(in + u));
}

__global__ void __launch_bounds__(kThreads) copy_bytes_kernel(const unsigned char *in, unsigned char *out, size_t n)
{
    const size_t gsz = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += gsz) out[i] = in[i];
}

// ------------------------------------------------------------------ deterministic sum
// Order (restated on the CPU by oracle/oracle.c: orc_sum_two_pass, and in DESIGN.md):
//  pass 1: block b owns elements [b*chunk, min(n, (b+1)*chunk)); thread t owns the vector
//          units t, t+256, t+512, ... of that chunk and keeps one accumulator per vector
//          lane; lanes are folded left to right; thread totals go through the xor-shuffle
//          butterfly 16,8,4,2,1 (all lanes agree because fp add is commutative); warp
//          totals go through the same butterfly in warp 0 (missing warps contribute +0).
//  pass 2: one block, thread t folds partials t, t+256, ... then the same tree.
// No atomics, fixed grid -> bitwise run-to-run reproducible, independent of the SM count.
template <typename ACC>
__device__ __forceinline__ ACC acc_add(ACC a, ACC b);
template <>
__device__ __forceinline__ float acc_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <>
__device__ __forceinline__ double acc_add<double>(double a, double b) { return __dadd_rn(a, b); }
template <>
__device__ __forceinline__ long long acc_add<long long>(long long a, long long b)
{
    return (long long)((unsigned long long)a + (unsigned long long)b);
}

template <typename T, typename ACC>
__device__ __forceinline__ ACC to_acc(T v) { return (ACC)v; }
template <>
__device__ __forceinline__ float to_acc<half_bits, float>(half_bits v) { return h2f(v.b); }
template <>
__device__ __forceinline__ float to_acc<bf16_bits, float>(bf16_bits v) { return b2f(v.b); }

template <typename ACC>
__device__ __forceinline__ ACC shfl_xor(ACC v, int off) { return __shfl_xor_sync(0xffffffffu, v, off); }

template <typename ACC, int THREADS>
__device__ __forceinline__ ACC block_tree(ACC s)
{
    __shared__ ACC warp_tot[32];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s = acc_add<ACC>(s, shfl_xor<ACC>(s, off));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_tot[warp] = s;
    __syncthreads();
    ACC v = (lane < THREADS / 32) ? warp_tot[lane] : (ACC)0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = acc_add<ACC>(v, shfl_xor<ACC>(v, off));
    return v;  // valid in warp 0
}

template <typename T, typename ACC, bool ALIGNED, int UNROLL>
__global__ void __launch_bounds__(kThreads) sum_pass1_kernel(const T *in, size_t n, size_t chunk, ACC *partials)
{
    constexpr int VEC = 16 / sizeof(T);
    const size_t begin = (size_t)blockIdx.x * chunk;
    ACC acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; j++) acc[j] = (ACC)0;
    if (begin < n) {
        const size_t count = (n - begin < chunk) ? n - begin : chunk;
        const T *base = in + begin;
        const size_t nunits = count / VEC;
        size_t u = threadIdx.x;
        if (ALIGNED) {
            const uint4 *p = reinterpret_cast<const uint4 *>(base);
            for (; u + (size_t)(UNROLL - 1) * kThreads < nunits; u += (size_t)UNROLL * kThreads) {
                Pack<T> r[UNROLL];
#pragma unroll
                for (int k = 0; k < UNROLL; k++) r[k].q = ld16(p + u + (size_t)k * kThreads);
#pragma unroll
                for (int k = 0; k < UNROLL; k++)
#pragma unroll
                    for (int j = 0; j < VEC; j++) acc[j] = acc_add<ACC>(acc[j], to_acc<T, ACC>(r[k].v[j]));
            }
            for (; u < nunits; u += kThreads) {
                Pack<T> r;
                r.q = ld16(p + u);
#pragma unroll
                for (int j = 0; j < VEC; j++) acc[j] = acc_add<ACC>(acc[j], to_acc<T, ACC>(r.v[j]));
            }
        } else {
            for (; u < nunits; u += kThreads)
#pragma unroll
                for (int j = 0; j < VEC; j++) acc[j] = acc_add<ACC>(acc[j], to_acc<T, ACC>(base[u * VEC + j]));
        }
        const size_t rem = count - nunits * VEC;
        if (threadIdx.x < rem) acc[0] = acc_add<ACC>(acc[0], to_acc<T, ACC>(base[nunits * VEC + threadIdx.x]));
    }
    ACC s = acc[0];
#pragma unroll
    for (int j = 1; j < VEC; j++) s = acc_add<ACC>(s, acc[j]);
    s = block_tree<ACC, kThreads>(s);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// divisor > 0: mean = sum / (ACC)divisor, one IEEE division (integers: truncating)
template <typename ACC>
__global__ void __launch_bounds__(kThreads) sum_pass2_kernel(const ACC *partials, int nblocks, ACC *out, size_t divisor)
{
    ACC s = (ACC)0;
    for (int i = threadIdx.x; i < nblocks; i += kThreads) s = acc_add<ACC>(s, partials[i]);
    s = block_tree<ACC, kThreads>(s);
    if (threadIdx.x == 0) *out = divisor ? s / (ACC)divisor : s;
}

// rank-ordered fold of the gathered per-rank partials (multi-GPU combine): sequential
template <typename ACC>
__global__ void fold_ranks_kernel(const ACC *gathered, int n_ranks, ACC *out, size_t divisor)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ACC s = gathered[0];
        for (int r = 1; r < n_ranks; r++) s = acc_add<ACC>(s, gathered[r]);
        *out = divisor ? s / (ACC)divisor : s;
    }
}

// ---- fused pass 2 + cross-GPU exchange over NVLink peer memory -----------------------------------
// Every rank owns a small exchange buffer that all peers map (CUDA IPC).  The single block that folds
// the pass-1 partials also publishes the rank's total into every peer's buffer with system-scope
// release stores (threads 0..R-1 each serve one peer: R NVLink stores in flight), waits for the R
// totals addressed to this rank with acquire loads, and folds them in rank order.  One kernel instead
// of pass 2 + ncclAllGather + fold: the exchange costs one NVLink round trip (~2-4 us), not a
// collective launch.  Slots are double buffered by the parity of the call number: a rank can only be one
// call ahead of a peer (it needs the peer's value to finish a call), so parity p of call k+2 is never
// written before every rank has finished reading parity p of call k.
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <typename ACC>
__global__ void __launch_bounds__(kThreads)
sum_pass2_exchange_kernel(const ACC *partials, int nblocks, XchgSlot *const *peers, int n_ranks, int rank,
                          unsigned long long epoch, ACC *out, size_t divisor, long long timeout_cycles, int *status)
{
    __shared__ ACC vals[kMaxRanks];
    __shared__ ACC local;
    ACC s = (ACC)0;
    for (int i = threadIdx.x; i < nblocks; i += kThreads) s = acc_add<ACC>(s, partials[i]);
    s = block_tree<ACC, kThreads>(s);
    if (threadIdx.x == 0) local = s;
    __syncthreads();
    const int parity = (int)(epoch & 1ull);
    if ((int)threadIdx.x < n_ranks) {
        const int r = threadIdx.x;
        unsigned long long bits = 0;
        const ACC mine = local;
        memcpy(&bits, &mine, sizeof(ACC));
        XchgSlot *dst = peers[r] + parity * n_ranks + rank;  // my slot in rank r's buffer (peer memory)
        *reinterpret_cast<volatile unsigned long long *>(&dst->value) = bits;
        st_release_sys(&dst->epoch, epoch);
        const XchgSlot *src = peers[rank] + parity * n_ranks + r;  // rank r's slot in my buffer
        const long long t0 = clock64();
        bool ok = true;
        while (ld_acquire_sys(&src->epoch) != epoch) {
            if (clock64() - t0 > timeout_cycles) {  // a peer never arrived: fail instead of hanging the GPU
                ok = false;
                break;
            }
        }
        unsigned long long got = *reinterpret_cast<const volatile unsigned long long *>(&src->value);
        if (!ok) {
            got = 0;
            atomicExch(status, 1);
        }
        ACC v;
        memcpy(&v, &got, sizeof(ACC));
        vals[r] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ACC t = vals[0];
        for (int r = 1; r < n_ranks; r++) t = acc_add<ACC>(t, vals[r]);  // rank order, like fold_ranks_kernel
        *out = divisor ? t / (ACC)divisor : t;
    }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int grid_for(size_t work_items, size_t per_block, int max_blocks)
{
    size_t g = (work_items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > (size_t)max_blocks) g = (size_t)max_blocks;
    return (int)g;
}

template <typename T, int OP>
cudaError_t launch_binary_t(const LaunchCtx &ctx, const void *lhs, const void *rhs, void *out, size_t n)
{
    constexpr int UNROLL = 2;
    constexpr int VEC = 16 / sizeof(T);
    if (aligned16(lhs) && aligned16(rhs) && aligned16(out)) {
        const int grid = grid_for(n / VEC + 1, (size_t)kThreads * UNROLL, ctx.max_blocks);
        binary_vec_kernel<T, OP, UNROLL><<<grid, kThreads, 0, ctx.stream>>>((const T *)lhs, (const T *)rhs, (T *)out, n);
    } else {
        const int grid = grid_for(n, (size_t)kThreads * 4, ctx.max_blocks);
        binary_scalar_kernel<T, OP><<<grid, kThreads, 0, ctx.stream>>>((const T *)lhs, (const T *)rhs, (T *)out, n);
    }
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_binary_op(const LaunchCtx &ctx, int op, const void *lhs, const void *rhs, void *out, size_t n)
{
    switch (op) {
    case CB_BIN_ADD: return launch_binary_t<T, CB_BIN_ADD>(ctx, lhs, rhs, out, n);
    case CB_BIN_MUL: return launch_binary_t<T, CB_BIN_MUL>(ctx, lhs, rhs, out, n);
    case CB_BIN_SUB: return launch_binary_t<T, CB_BIN_SUB>(ctx, lhs, rhs, out, n);
    default: return launch_binary_t<T, CB_BIN_DIV>(ctx, lhs, rhs, out, n);
    }
}

template <typename T, typename ACC>
cudaError_t launch_sum_xchg_t(const LaunchCtx &ctx, const void *in, size_t n, int blocks, size_t chunk, void *partials,
                              void *out, size_t divisor, const XchgArgs &x)
{
    if (n == 0) {  // an empty slice contributes 0 but still takes part in the exchange
        cudaError_t e = cudaMemsetAsync(partials, 0, sizeof(ACC), ctx.stream);
        if (e != cudaSuccess) return e;
        blocks = 1;
    } else if (aligned16(in)) {
        sum_pass1_kernel<T, ACC, true, 4><<<blocks, kThreads, 0, ctx.stream>>>((const T *)in, n, chunk, (ACC *)partials);
    } else {
        sum_pass1_kernel<T, ACC, false, 1><<<blocks, kThreads, 0, ctx.stream>>>((const T *)in, n, chunk, (ACC *)partials);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sum_pass2_exchange_kernel<ACC><<<1, kThreads, 0, ctx.stream>>>((const ACC *)partials, blocks, x.peers, x.n_ranks, x.rank,
                                                                   x.epoch, (ACC *)out, divisor, x.timeout_cycles, x.status);
    return cudaGetLastError();
}

template <typename T, typename ACC>
cudaError_t launch_sum_t(const LaunchCtx &ctx, const void *in, size_t n, int blocks, size_t chunk, void *partials,
                         void *out, size_t divisor)
{
    if (aligned16(in))
        sum_pass1_kernel<T, ACC, true, 4><<<blocks, kThreads, 0, ctx.stream>>>((const T *)in, n, chunk, (ACC *)partials);
    else
        sum_pass1_kernel<T, ACC, false, 1><<<blocks, kThreads, 0, ctx.stream>>>((const T *)in, n, chunk, (ACC *)partials);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sum_pass2_kernel<ACC><<<1, kThreads, 0, ctx.stream>>>((const ACC *)partials, blocks, (ACC *)out, divisor);
    return cudaGetLastError();
}

}  // namespace

// ------------------------------------------------------------------ host launchers
cudaError_t launch_binary(const LaunchCtx &ctx, int dtype, int op, const void *lhs, const void *rhs, void *out, size_t n)
{
    (void)cudaGetLastError();  // a stale error of an earlier, unchecked call must not be blamed on this launch
    switch (dtype) {
    case CB_F32: return launch_binary_op<float>(ctx, op, lhs, rhs, out, n);
    case CB_F64: return launch_binary_op<double>(ctx, op, lhs, rhs, out, n);
    case CB_F16: return launch_binary_op<half_bits>(ctx, op, lhs, rhs, out, n);
    case CB_I32: return launch_binary_op<int>(ctx, op, lhs, rhs, out, n);
    case CB_I64: return launch_binary_op<long long>(ctx, op, lhs, rhs, out, n);
    case CB_U32: return launch_binary_op<unsigned int>(ctx, op, lhs, rhs, out, n);
    case CB_U8: return launch_binary_op<unsigned char>(ctx, op, lhs, rhs, out, n);
    case CB_BF16: return launch_binary_op<bf16_bits>(ctx, op, lhs, rhs, out, n);
    case CB_I8: return launch_binary_op<signed char>(ctx, op, lhs, rhs, out, n);
    case CB_I16: return launch_binary_op<short>(ctx, op, lhs, rhs, out, n);
    case CB_U16: return launch_binary_op<unsigned short>(ctx, op, lhs, rhs, out, n);
    case CB_U64: return launch_binary_op<unsigned long long>(ctx, op, lhs, rhs, out, n);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_fill(const LaunchCtx &ctx, void *out, size_t n, int elem_bytes, uint64_t pattern)
{
    (void)cudaGetLastError();  // a stale error of an earlier, unchecked call must not be blamed on this launch
    // replicate the element pattern over 16 bytes
    uint64_t p64 = pattern;
    if (elem_bytes == 1) p64 = (pattern & 0xffu) * 0x0101010101010101ull;
    else if (elem_bytes == 2) p64 = (pattern & 0xffffu) * 0x0001000100010001ull;
    else if (elem_bytes == 4) p64 = (pattern & 0xffffffffu) * 0x0000000100000001ull;
    const uint4 pat = make_uint4((unsigned)p64, (unsigned)(p64 >> 32), (unsigned)p64, (unsigned)(p64 >> 32));
    unsigned char *base = (unsigned char *)out;
    const size_t bytes = n * (size_t)elem_bytes;
    auto scalar = [&](unsigned char *p, size_t nbytes) -> cudaError_t {
        if (!nbytes) return cudaSuccess;
        const size_t cnt = nbytes / (size_t)elem_bytes;
        const int grid = grid_for(cnt, (size_t)kThreads * 4, ctx.max_blocks);
        switch (elem_bytes) {
        case 1: fill_scalar_kernel<unsigned char><<<grid, kThreads, 0, ctx.stream>>>(p, cnt, (unsigned char)pattern); break;
        case 2: fill_scalar_kernel<unsigned short><<<grid, kThreads, 0, ctx.stream>>>((unsigned short *)p, cnt, (unsigned short)pattern); break;
        case 4: fill_scalar_kernel<unsigned int><<<grid, kThreads, 0, ctx.stream>>>((unsigned int *)p, cnt, (unsigned int)pattern); break;
        default: fill_scalar_kernel<unsigned long long><<<grid, kThreads, 0, ctx.stream>>>((unsigned long long *)p, cnt, (unsigned long long)pattern); break;
        }
        return cudaGetLastError();
    };
    if (!aligned16(base)) return scalar(base, bytes);  // sub-slices that do not start on a 16-byte boundary
    const size_t units = bytes / 16;
    if (units) {
        const int grid = grid_for(units, (size_t)kThreads * 4, ctx.max_blocks);
        fill16_kernel<<<grid, kThreads, 0, ctx.stream>>>((uint4 *)base, units, pat);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return scalar(base + units * 16, bytes - units * 16);
}

int launch_fill_count(const void *out, size_t n, int elem_bytes)
{
    const size_t bytes = n * (size_t)elem_bytes;
    if (!bytes) return 0;
    if (!aligned16(out)) return 1;
    return (bytes / 16 ? 1 : 0) + (bytes % 16 ? 1 : 0);
}

cudaError_t launch_copy(const LaunchCtx &ctx, void *dst, const void *src, size_t bytes)
{
    (void)cudaGetLastError();  // a stale error of an earlier, unchecked call must not be blamed on this launch
    if (!bytes) return cudaSuccess;
    if (aligned16(dst) && aligned16(src)) {
        const size_t units = bytes / 16;
        if (units) {
            const int grid = grid_for(units, (size_t)kThreads * 4, ctx.max_blocks);
            copy16_kernel<<<grid, kThreads, 0, ctx.stream>>>((const uint4 *)src, (uint4 *)dst, units);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
        const size_t rest = bytes - units * 16;
        if (rest) {
            copy_bytes_kernel<<<1, kThreads, 0, ctx.stream>>>((const unsigned char *)src + units * 16,
                                                              (unsigned char *)dst + units * 16, rest);
            return cudaGetLastError();
        }
        return cudaSuccess;
    }
    const int grid = grid_for(bytes, (size_t)kThreads * 4, ctx.max_blocks);
    copy_bytes_kernel<<<grid, kThreads, 0, ctx.stream>>>((const unsigned char *)src, (unsigned char *)dst, bytes);
    return cudaGetLastError();
}

int launch_copy_count(const void *dst, const void *src, size_t bytes)
{
    if (!bytes) return 0;
    if (aligned16(dst) && aligned16(src)) return (bytes / 16 ? 1 : 0) + (bytes % 16 ? 1 : 0);
    return 1;
}

void sum_plan(int dtype, size_t n, int *blocks, size_t *chunk, int *threads, int *vec, int *threads2)
{
    static const int sizes[CB_DTYPE_COUNT] = {4, 8, 2, 4, 8, 4, 1, 2, 1, 2, 2, 8, 1};
    const int v = 16 / sizes[dtype];
    const size_t unit = (size_t)kThreads * (size_t)v;
    size_t b = (n + unit - 1) / unit;
    if (b < 1) b = 1;
    if (b > (size_t)kSumMaxBlocks) b = (size_t)kSumMaxBlocks;
    size_t c = (n + b - 1) / b;
    c = (c + unit - 1) / unit * unit;
    if (c == 0) c = unit;
    b = (n + c - 1) / c;
    if (b < 1) b = 1;
    *blocks = (int)b;
    *chunk = c;
    *threads = kThreads;
    *vec = v;
    *threads2 = kThreads;
}

cudaError_t launch_sum(const LaunchCtx &ctx, int dtype, const void *in, size_t n, void *partials, void *out, size_t divisor)
{
    (void)cudaGetLastError();  // a stale error of an earlier, unchecked call must not be blamed on this launch
    int blocks, threads, vec, threads2;
    size_t chunk;
    sum_plan(dtype, n, &blocks, &chunk, &threads, &vec, &threads2);
    switch (dtype) {
    case CB_F32: return launch_sum_t<float, float>(ctx, in, n, blocks, chunk, partials, out, divisor);
    case CB_F64: return launch_sum_t<double, double>(ctx, in, n, blocks, chunk, partials, out, divisor);
    case CB_F16: return launch_sum_t<half_bits, float>(ctx, in, n, blocks, chunk, partials, out, divisor);
    case CB_I32: return launch_sum_t<int, long long>(ctx, in, n, blocks, chunk, partials, out, divisor);
    case CB_I64: return launch_sum_t<long long, long long>(ctx, in, n, blocks, chunk, partials, out, divisor);
    case CB_U32: return launch_sum_t<unsigned int, long long>(ctx, in, n, blocks, chunk, partials, out, divisor);
    case CB_U8: return launch_sum_t<unsigned char, long long>(ctx, in, n, blocks, chunk, partials, out, divisor);
    case CB_BF16: return launch_sum_t<bf16_bits, float>(ctx, in, n, blocks, chunk, partials, out, divisor);
    case CB_I8: return launch_sum_t<signed char, long long>(ctx, in, n, blocks, chunk, partials, out, divisor);
    case CB_I16: return launch_sum_t<short, long long>(ctx, in, n, blocks, chunk, partials, out, divisor);
    case CB_U16: return launch_sum_t<unsigned short, long long>(ctx, in, n, blocks, chunk, partials, out, divisor);
    case CB_U64: return launch_sum_t<unsigned long long, long long>(ctx, in, n, blocks, chunk, partials, out, divisor);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_sum_exchange(const LaunchCtx &ctx, int dtype, const void *in, size_t n, void *partials, void *out,
                                size_t divisor, const XchgArgs &x)
{
    (void)cudaGetLastError();
    int blocks = 1, threads, vec, threads2;
    size_t chunk = 0;
    if (n) sum_plan(dtype, n, &blocks, &chunk, &threads, &vec, &threads2);
    switch (dtype) {
    case CB_F32: return launch_sum_xchg_t<float, float>(ctx, in, n, blocks, chunk, partials, out, divisor, x);
    case CB_F64: return launch_sum_xchg_t<double, double>(ctx, in, n, blocks, chunk, partials, out, divisor, x);
    case CB_F16: return launch_sum_xchg_t<half_bits, float>(ctx, in, n, blocks, chunk, partials, out, divisor, x);
    case CB_I32: return launch_sum_xchg_t<int, long long>(ctx, in, n, blocks, chunk, partials, out, divisor, x);
    case CB_I64: return launch_sum_xchg_t<long long, long long>(ctx, in, n, blocks, chunk, partials, out, divisor, x);
    case CB_U32: return launch_sum_xchg_t<unsigned int, long long>(ctx, in, n, blocks, chunk, partials, out, divisor, x);
    case CB_U8: return launch_sum_xchg_t<unsigned char, long long>(ctx, in, n, blocks, chunk, partials, out, divisor, x);
    case CB_BF16: return launch_sum_xchg_t<bf16_bits, float>(ctx, in, n, blocks, chunk, partials, out, divisor, x);
    case CB_I8: return launch_sum_xchg_t<signed char, long long>(ctx, in, n, blocks, chunk, partials, out, divisor, x);
    case CB_I16: return launch_sum_xchg_t<short, long long>(ctx, in, n, blocks, chunk, partials, out, divisor, x);
    case CB_U16: return launch_sum_xchg_t<unsigned short, long long>(ctx, in, n, blocks, chunk, partials, out, divisor, x);
    case CB_U64: return launch_sum_xchg_t<unsigned long long, long long>(ctx, in, n, blocks, chunk, partials, out, divisor, x);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_fold_ranks(const LaunchCtx &ctx, int dtype, const void *gathered, int n_ranks, void *out, size_t divisor)
{
    (void)cudaGetLastError();  // a stale error of an earlier, unchecked call must not be blamed on this launch
    switch (dtype) {
    case CB_F32: case CB_F16: case CB_BF16:
        fold_ranks_kernel<float><<<1, 32, 0, ctx.stream>>>((const float *)gathered, n_ranks, (float *)out, divisor);
        break;
    case CB_F64:
        fold_ranks_kernel<double><<<1, 32, 0, ctx.stream>>>((const double *)gathered, n_ranks, (double *)out, divisor);
        break;
    default:
        fold_ranks_kernel<long long><<<1, 32, 0, ctx.stream>>>((const long long *)gathered, n_ranks, (long long *)out, divisor);
        break;
    }
    return cudaGetLastError();
}

}  // namespace cb
