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
        // (8- and 16-bit operands are widened to unsigned int explicitly: `U * U` would promote to SIGNED int and
        // 65535 * 65535 overflows it)
        typedef typename std::conditional<(sizeof(T) < 4), unsigned int, U>::type W;
        if (OP == CB_BIN_ADD) return (T)(U)((W)(U)a + (W)(U)b);
        if (OP == CB_BIN_MUL) return (T)(U)((W)(U)a * (W)(U)b);
        if (OP == CB_BIN_SUB) return (T)(U)((W)(U)a - (W)(U)b);
        if (b == (T)0) return (T)0;
        if ((T)-1 < (T)0 && b == (T)-1) return (T)((U)0 - (U)a);  // MIN / -1 wraps to MIN (the reference panics)
        return (T)(a / b);
    }
};

// ------------------------------------------------------------------ binary: out = lhs op rhs
// Algorithmic traffic: 3 * sizeof(T) per element.  UNROLL tiles of 16-byte units per thread.
// programmatic dependent launch, see cb_pdl_enter() in kernels/skeleton.cuh
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
}

// launches a kernel that begins with pdl_enter() (or handles the dependency itself) with programmatic stream serialisation
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(const LaunchCtx &ctx, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, Args... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx.pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename T, int OP, int UNROLL>
__global__ void __launch_bounds__(kThreads) binary_vec_kernel(const T *lhs, const T *rhs, T *out, size_t n)
{
    pdl_enter();
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
    pdl_enter();
    const size_t gsz = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += gsz)
        out[i] = BinOp<T, OP>::apply(lhs[i], rhs[i]);
}

// ------------------------------------------------------------------ fill / clear: 1 write per element
__global__ void __launch_bounds__(kThreads) fill16_kernel(uint4 *out, size_t nunits, uint4 pattern)
{
    pdl_enter();
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
    pdl_enter();
    const size_t gsz = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += gsz) out[i] = value;
}

// ------------------------------------------------------------------ copy: 1 read + 1 write
__global__ void __launch_bounds__(kThreads) copy16_kernel(const uint4 *in, uint4 *out, size_t nunits)
{
    pdl_enter();
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
    pdl_enter();
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
template <>
__device__ __forceinline__ unsigned long long acc_add<unsigned long long>(unsigned long long a, unsigned long long b)
{
    return a + b;  // u64 sums wrap as unsigned and the mean divides unsigned
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

// cross-GPU exchange of the rank totals through peer-mapped memory (see sum_kernel)
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
__device__ __forceinline__ ACC ld_partial(const ACC *p)
{
    return *reinterpret_cast<const volatile ACC *>(p);  // written by other blocks of this launch: never from L1
}

// ONE launch per sum.  Every block reduces its chunk (pass 1) and publishes the partial; the block that takes the
// last ticket of a device counter folds all partials (pass 2) — in INDEX order, whichever block that happens to be,
// so the result does not depend on block scheduling — and, in the multi-GPU form (XCHG), also exchanges the rank's
// total with the peers:
//   every rank owns a small exchange buffer that all peers map (CUDA IPC).  Threads 0..R-1 of the last block each
//   publish the total into one peer's buffer with a system-scope release store (R NVLink stores in flight), wait
//   for the R totals addressed to this rank with acquire loads, and thread 0 folds them in RANK order: all ranks
//   end with identical bits.  Slots are double buffered by the parity of the call number: a rank can only be one
//   call ahead of a peer (it needs the peer's value to finish a call), so parity p of call k+2 is never written
//   before every rank has finished reading parity p of call k.  A peer that does not arrive within
//   `timeout_cycles` makes the kernel store the call number to `status` (host-mapped memory) instead of hanging
//   the GPU; the host reports it as an error at the next synchronisation point (cb_comm_check).
// divisor > 0: mean = sum / (ACC)divisor, one IEEE division (integers: truncating)
template <typename T, typename ACC, bool ALIGNED, int UNROLL, bool XCHG>
__global__ void __launch_bounds__(kThreads)
sum_kernel(const T *in, size_t n, size_t chunk, ACC *partials, unsigned int *ticket, ACC *out, size_t divisor, XchgArgs x,
           unsigned int pdl)
{
    constexpr int VEC = 16 / sizeof(T);
    // Programmatic dependent launch (the kernel is launched with programmatic stream serialisation): when the kernel
    // before this one on the stream is ANOTHER sum (pdl bit 0; the host alternates two sets of partials / tickets, and the
    // exchange slots alternate by call parity anyway) the streaming pass below may start while that sum's last block is
    // still folding and waiting for its peers — the SMs would otherwise idle through every exchange round trip.  After
    // anything else the blocks wait for the preceding kernel first, exactly like a plain launch.
    if (!(pdl & 1u)) asm volatile("griddepcontrol.wait;" ::: "memory");
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

    // publish the partial, then take a ticket: the block holding the last one sees every partial
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s;
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) {
        asm volatile("griddepcontrol.launch_dependents;");  // this block is done with everything a following sum shares
        return;
    }
    // the last block first makes sure the preceding sum is complete (its `out`, its ticket reset), then lets the next
    // one start streaming under this call's fold + exchange
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    __threadfence();
    if (threadIdx.x == 0) *ticket = 0u;  // ready for the launch after next (which uses this set again)

    // pass 2: thread t folds partials t, t + 256, ... then the same tree
    ACC t = (ACC)0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) t = acc_add<ACC>(t, ld_partial<ACC>(partials + i));
    t = block_tree<ACC, kThreads>(t);
    if (!XCHG) {
        if (threadIdx.x == 0) *out = divisor ? t / (ACC)divisor : t;
        return;
    }
    __shared__ ACC vals[kMaxRanks];
    __shared__ ACC local;
    if (threadIdx.x == 0) local = t;
    __syncthreads();
    const int parity = (int)(x.epoch & 1ull);
    if ((int)threadIdx.x < x.n_ranks) {
        const int r = threadIdx.x;
        unsigned long long bits = 0;
        const ACC mine = local;
        memcpy(&bits, &mine, sizeof(ACC));
        XchgSlot *dst = x.peers[r] + parity * x.n_ranks + x.rank;  // my slot in rank r's buffer (peer memory)
        *reinterpret_cast<volatile unsigned long long *>(&dst->value) = bits;
        st_release_sys(&dst->epoch, x.epoch);
        const XchgSlot *src = x.peers[x.rank] + parity * x.n_ranks + r;  // rank r's slot in my buffer
        const long long t0 = clock64();
        bool ok = true;
        while (ld_acquire_sys(&src->epoch) != x.epoch) {
            if (clock64() - t0 > x.timeout_cycles) {  // a peer never arrived: fail instead of hanging the GPU
                ok = false;
                break;
            }
        }
        unsigned long long got = *reinterpret_cast<const volatile unsigned long long *>(&src->value);
        if (!ok) {
            got = 0;
            *reinterpret_cast<volatile int *>(x.status) = (int)x.epoch;  // host-mapped: read by cb_comm_check
            __threadfence_system();
        }
        ACC v;
        memcpy(&v, &got, sizeof(ACC));
        vals[r] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ACC tot = vals[0];
        for (int r = 1; r < x.n_ranks; r++) tot = acc_add<ACC>(tot, vals[r]);  // rank order, like fold_ranks_kernel
        *out = divisor ? tot / (ACC)divisor : tot;
    }
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

// ------------------------------------------------------------------ 16-bit unary chains as a table lookup
// A fused unary chain over f16 / bf16 is a function of one 16-bit value: 65 536 possible inputs.  Evaluated with
// arithmetic it follows the reference — f32 math and a round to 16 bits after EVERY op (src/number.rs:543-676) —
// and is bound by the FP32 pipe at ~0.46 of the 16-bit HBM roofline (profiles/r1_half_kernels_ncu.md).  The table
// is filled ONCE per compiled chain by running that very arithmetic kernel over all 65 536 bit patterns, so a
// lookup returns bit for bit what the arithmetic kernel would have computed, NaN patterns included.
//  * the 128 KiB table lives in shared memory (one persistent 1024-thread block per SM; 228 KB per SM on B200);
//  * 128-bit streaming loads / stores, 8 lookups (LDS.U16) per 16-byte unit;
//  * tiles are handed out by a device counter, so an SM that runs slower (or starts later) simply takes fewer tiles;
//    the block that finishes last resets the counters for the next launch.
constexpr size_t kLutBytes = 65536 * 2;

__global__ void iota16_kernel(unsigned short *out)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 65536u) out[i] = (unsigned short)i;
}

__device__ __forceinline__ unsigned int lut2(const unsigned short *lut, unsigned int w)
{
    const unsigned int lo = lut[w & 0xffffu], hi = lut[w >> 16];
    unsigned int r;
    asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ uint4 lut8(const unsigned short *lut, uint4 q)
{
    return make_uint4(lut2(lut, q.x), lut2(lut, q.y), lut2(lut, q.z), lut2(lut, q.w));
}

// g + t for two packed 16-bit lanes, as the reference adds them: through f32, one round-to-nearest-even back
// (exact for f16: an f32 holds the sum of two binary16 values exactly enough that the single rounding is the correct
// one; bf16 is defined that way by `half`)
template <bool BF16>
__device__ __forceinline__ unsigned int add16x2(unsigned int g, unsigned int t)
{
    float g0, g1, t0, t1;
    if (BF16) {
        g0 = __uint_as_float(g << 16), g1 = __uint_as_float(g & 0xffff0000u);
        t0 = __uint_as_float(t << 16), t1 = __uint_as_float(t & 0xffff0000u);
    } else {
        g0 = h2f((unsigned short)(g & 0xffffu)), g1 = h2f((unsigned short)(g >> 16));
        t0 = h2f((unsigned short)(t & 0xffffu)), t1 = h2f((unsigned short)(t >> 16));
    }
    const float s0 = __fadd_rn(g0, t0), s1 = __fadd_rn(g1, t1);
    unsigned int r;
    if (BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(s1), "f"(s0));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(s1), "f"(s0));
    return r;
}
// o * t for two packed 16-bit lanes: the f32 product (exact for binary16 operands) rounded once to 16 bits, as `half` does
template <bool BF16>
__device__ __forceinline__ unsigned int mul16x2(unsigned int o, unsigned int t)
{
    float o0, o1, t0, t1;
    if (BF16) {
        o0 = __uint_as_float(o << 16), o1 = __uint_as_float(o & 0xffff0000u);
        t0 = __uint_as_float(t << 16), t1 = __uint_as_float(t & 0xffff0000u);
    } else {
        o0 = h2f((unsigned short)(o & 0xffffu)), o1 = h2f((unsigned short)(o >> 16));
        t0 = h2f((unsigned short)(t & 0xffffu)), t1 = h2f((unsigned short)(t >> 16));
    }
    const float p0 = __fmul_rn(o0, t0), p1 = __fmul_rn(o1, t1);
    unsigned int r;
    if (BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(p1), "f"(p0));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(p1), "f"(p0));
    return r;
}
template <bool BF16>
__device__ __forceinline__ uint4 mul16x8(uint4 o, uint4 t)
{
    return make_uint4(mul16x2<BF16>(o.x, t.x), mul16x2<BF16>(o.y, t.y), mul16x2<BF16>(o.z, t.z), mul16x2<BF16>(o.w, t.w));
}
template <bool BF16>
__device__ __forceinline__ uint4 add16x8(uint4 g, uint4 t)
{
    return make_uint4(add16x2<BF16>(g.x, t.x), add16x2<BF16>(g.y, t.y), add16x2<BF16>(g.z, t.z), add16x2<BF16>(g.w, t.w));
}

// kLutThreads threads per block (one block per SM), kLutUnroll 16-byte units per thread per tile (and as many again
// prefetched), kLutGrab tiles per ticket.
// MODE 0: out[i] = table[in[i]]                                  (a fused unary chain, K1/K2)
// MODE 1 / 2 (f16 / bf16): the SEEDED backward of a fused chain — out_grad = ones makes the whole backward term a
//   function of x alone, table[x] = (((1 * gK(x_{K-1})) ...) * g1(x)) with every 16-bit rounding of the reference:
//   grad[i] = grad[i] + table[in[i]] (one 16-bit add, through f32 like `half`) and seed[i] = 1.
//   4 x 2 bytes per element instead of the FP32-pipe-bound chain-grad arithmetic.
// MODE 3 / 4 (f16 / bf16): add_unary_grad of ONE op with a general out_grad — the same table holds g(lhs) (1 * g is g):
//   grad[i] = grad[i] + seed[i] * table[in[i]], multiply then add, each rounded to 16 bits (`seed` is read here).
template <int kLutThreads, int kLutUnroll, int kLutGrab, int MODE>
__global__ void __launch_bounds__(kLutThreads, 1)
lut16_kernel(const unsigned short *in, unsigned short *out, unsigned short *seed, size_t n, const uint4 *table,
             unsigned long long *counters)
{
    extern __shared__ uint4 lut_q[];
    // the table was written when the expression was compiled: loading it does not depend on the kernel before this one,
    // so it happens BEFORE the dependent-launch wait and overlaps that kernel's tail
    for (int i = threadIdx.x; i < (int)(kLutBytes / 16); i += kLutThreads) lut_q[i] = table[i];
    pdl_enter();
    __syncthreads();
    const unsigned short *lut = reinterpret_cast<const unsigned short *>(lut_q);
    const size_t nunits = n / 8;
    // Work is handed out per WARP (no block-wide barrier in the streaming loop): a warp tile is kLutUnroll rows of 32
    // consecutive 16-byte units (512 contiguous bytes per row) and a warp takes kLutGrab consecutive tiles at a time —
    // the first grab is static, the following ones come from a device counter (one atomic per grab: handing out
    // single 2 KB tiles made the kernel atomic-bound at ~0.5 G atomics/s on the one address, profiles/r2_lut16_shapes.log).
    // The loads of the NEXT tile are issued before the lookups of the current one, so global-memory latency overlaps
    // the shared-memory work (the block-tile version was latency bound: LSU data pipe at 67 %, 16 long-scoreboard
    // stall cycles per issue, profiles/r2_lut16_ncu.txt).
    constexpr bool kGrad = MODE != 0;
    constexpr bool kMul = MODE >= 3;  // out_grad is an input, not the seed to write
    constexpr bool kBf16 = MODE == 2 || MODE == 4;
    constexpr unsigned int kOne = kBf16 ? 0x3f803f80u : 0x3c003c00u;
    const uint4 ones = make_uint4(kOne, kOne, kOne, kOne);
    const size_t tile_units = (size_t)32 * kLutUnroll;
    const size_t ntiles = nunits / tile_units;
    const uint4 *pin = reinterpret_cast<const uint4 *>(in);
    uint4 *pout = reinterpret_cast<uint4 *>(out);
    uint4 *pseed = reinterpret_cast<uint4 *>(seed);
    const unsigned int lane = threadIdx.x & 31u;
    const size_t warps_total = (size_t)gridDim.x * (kLutThreads / 32);
    size_t tile = ((size_t)blockIdx.x * (kLutThreads / 32) + (threadIdx.x >> 5)) * kLutGrab;
    size_t grab_end = tile + kLutGrab;
    uint4 cur[kLutUnroll], nxt[kLutUnroll], gcur[kGrad ? kLutUnroll : 1], gnxt[kGrad ? kLutUnroll : 1];
    uint4 ocur[kMul ? kLutUnroll : 1], onxt[kMul ? kLutUnroll : 1];
    if (tile < ntiles) {
#pragma unroll
        for (int u = 0; u < kLutUnroll; u++) {
            cur[u] = ld16(pin + tile * tile_units + (size_t)u * 32 + lane);
            if (kGrad) gcur[u] = ld16(pout + tile * tile_units + (size_t)u * 32 + lane);
            if (kMul) ocur[u] = ld16(pseed + tile * tile_units + (size_t)u * 32 + lane);
        }
    }
    while (tile < ntiles) {
        size_t next = tile + 1;
        if (next == grab_end) {  // warp-uniform
            unsigned int got = 0;  // (a 32-bit ticket: 2^32 grabs are far beyond any buffer)
            if (lane == 0) got = atomicAdd(reinterpret_cast<unsigned int *>(&counters[0]), 1u);
            next = ((size_t)__shfl_sync(0xffffffffu, got, 0) + warps_total) * kLutGrab;
            grab_end = next + kLutGrab;
        }
        if (next < ntiles) {
#pragma unroll
            for (int u = 0; u < kLutUnroll; u++) {
                nxt[u] = ld16(pin + next * tile_units + (size_t)u * 32 + lane);
                if (kGrad) gnxt[u] = ld16(pout + next * tile_units + (size_t)u * 32 + lane);
                if (kMul) onxt[u] = ld16(pseed + next * tile_units + (size_t)u * 32 + lane);
            }
        }
#pragma unroll
        for (int u = 0; u < kLutUnroll; u++) {
            const size_t at = tile * tile_units + (size_t)u * 32 + lane;
            if (kMul) {
                st16(pout + at, add16x8<kBf16>(gcur[u], mul16x8<kBf16>(ocur[u], lut8(lut, cur[u]))));
            } else if (kGrad) {
                st16(pout + at, add16x8<kBf16>(gcur[u], lut8(lut, cur[u])));
                st16(pseed + at, ones);
            } else {
                st16(pout + at, lut8(lut, cur[u]));
            }
        }
#pragma unroll
        for (int u = 0; u < kLutUnroll; u++) {
            cur[u] = nxt[u];
            if (kGrad) gcur[u] = gnxt[u];
            if (kMul) ocur[u] = onxt[u];
        }
        tile = next;
    }
    // ragged end: units that do not fill a tile, then the < 8 element tail
    const size_t gid = (size_t)blockIdx.x * kLutThreads + threadIdx.x;
    const size_t gsz = (size_t)gridDim.x * kLutThreads;
    for (size_t u = ntiles * tile_units + gid; u < nunits; u += gsz) {
        if (kMul) {
            st16(pout + u, add16x8<kBf16>(ld16(pout + u), mul16x8<kBf16>(ld16(pseed + u), lut8(lut, ld16(pin + u)))));
        } else if (kGrad) {
            st16(pout + u, add16x8<kBf16>(ld16(pout + u), lut8(lut, ld16(pin + u))));
            st16(pseed + u, ones);
        } else {
            st16(pout + u, lut8(lut, ld16(pin + u)));
        }
    }
    for (size_t i = nunits * 8 + gid; i < n; i += gsz) {
        if (kMul) {
            out[i] = (unsigned short)(add16x2<kBf16>(out[i], mul16x2<kBf16>(seed[i], lut[in[i]])) & 0xffffu);
        } else if (kGrad) {
            out[i] = (unsigned short)(add16x2<kBf16>(out[i], lut[in[i]]) & 0xffffu);
            seed[i] = (unsigned short)(kOne & 0xffffu);
        } else {
            out[i] = lut[in[i]];
        }
    }
    // the last block to get here leaves the counters at zero for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&counters[1], 1ull) == gridDim.x - 1) {
            counters[0] = 0ull;
            counters[1] = 0ull;
            __threadfence();
        }
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
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = ctx.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx.pdl ? 1 : 0;
    if (aligned16(lhs) && aligned16(rhs) && aligned16(out)) {
        cfg.gridDim = dim3((unsigned)grid_for(n / VEC + 1, (size_t)kThreads * UNROLL, ctx.max_blocks));
        return cudaLaunchKernelEx(&cfg, binary_vec_kernel<T, OP, UNROLL>, (const T *)lhs, (const T *)rhs, (T *)out, n);
    }
    cfg.gridDim = dim3((unsigned)grid_for(n, (size_t)kThreads * 4, ctx.max_blocks));
    return cudaLaunchKernelEx(&cfg, binary_scalar_kernel<T, OP>, (const T *)lhs, (const T *)rhs, (T *)out, n);
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

template <typename T, typename ACC, bool XCHG>
cudaError_t launch_sum_t(const LaunchCtx &ctx, const void *in, size_t n, int blocks, size_t chunk, void *partials,
                         unsigned int *ticket, void *out, size_t divisor, const XchgArgs &x, const SumPdl &pdl)
{
    // n == 0 (an empty slice of a sharded buffer): one block, nothing to read, the partial is 0 — the rank still takes
    // part in the exchange
    if (n == 0) blocks = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = ctx.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl.enabled ? 1 : 0;
    // the two parities own separate partials and tickets
    ACC *parts = (ACC *)partials + (pdl.parity ? kSumMaxBlocks : 0);
    unsigned int *tick = ticket + (pdl.parity ? 1 : 0);
    const unsigned int flags = (pdl.enabled && pdl.after_sum) ? 1u : 0u;
    const size_t ch = chunk ? chunk : 1;
    if (aligned16(in)) return cudaLaunchKernelEx(&cfg, sum_kernel<T, ACC, true, 4, XCHG>, (const T *)in, n, ch, parts, tick, (ACC *)out, divisor, x, flags);
    return cudaLaunchKernelEx(&cfg, sum_kernel<T, ACC, false, 1, XCHG>, (const T *)in, n, ch, parts, tick, (ACC *)out, divisor, x, flags);
}

template <bool XCHG>
cudaError_t launch_sum_dtype(const LaunchCtx &ctx, int dtype, const void *in, size_t n, int blocks, size_t chunk, void *partials,
                             unsigned int *ticket, void *out, size_t divisor, const XchgArgs &x, const SumPdl &pdl)
{
    switch (dtype) {
    case CB_F32: return launch_sum_t<float, float, XCHG>(ctx, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
    case CB_F64: return launch_sum_t<double, double, XCHG>(ctx, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
    case CB_F16: return launch_sum_t<half_bits, float, XCHG>(ctx, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
    case CB_I32: return launch_sum_t<int, long long, XCHG>(ctx, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
    case CB_I64: return launch_sum_t<long long, long long, XCHG>(ctx, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
    case CB_U32: return launch_sum_t<unsigned int, long long, XCHG>(ctx, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
    case CB_U8: return launch_sum_t<unsigned char, long long, XCHG>(ctx, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
    case CB_BF16: return launch_sum_t<bf16_bits, float, XCHG>(ctx, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
    case CB_I8: return launch_sum_t<signed char, long long, XCHG>(ctx, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
    case CB_I16: return launch_sum_t<short, long long, XCHG>(ctx, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
    case CB_U16: return launch_sum_t<unsigned short, long long, XCHG>(ctx, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
    case CB_U64: return launch_sum_t<unsigned long long, unsigned long long, XCHG>(ctx, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
    default: return cudaErrorInvalidValue;
    }
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
        case 1: return launch_pdl(ctx, fill_scalar_kernel<unsigned char>, grid, kThreads, 0, p, cnt, (unsigned char)pattern);
        case 2: return launch_pdl(ctx, fill_scalar_kernel<unsigned short>, grid, kThreads, 0, (unsigned short *)p, cnt, (unsigned short)pattern);
        case 4: return launch_pdl(ctx, fill_scalar_kernel<unsigned int>, grid, kThreads, 0, (unsigned int *)p, cnt, (unsigned int)pattern);
        default: return launch_pdl(ctx, fill_scalar_kernel<unsigned long long>, grid, kThreads, 0, (unsigned long long *)p, cnt, (unsigned long long)pattern);
        }
    };
    if (!aligned16(base)) return scalar(base, bytes);  // sub-slices that do not start on a 16-byte boundary
    const size_t units = bytes / 16;
    if (units) {
        const int grid = grid_for(units, (size_t)kThreads * 4, ctx.max_blocks);
        cudaError_t e = launch_pdl(ctx, fill16_kernel, grid, kThreads, 0, (uint4 *)base, units, pat);
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
            cudaError_t e = launch_pdl(ctx, copy16_kernel, grid, kThreads, 0, (const uint4 *)src, (uint4 *)dst, units);
            if (e != cudaSuccess) return e;
        }
        const size_t rest = bytes - units * 16;
        if (rest) {
            return launch_pdl(ctx, copy_bytes_kernel, 1, kThreads, 0, (const unsigned char *)src + units * 16,
                              (unsigned char *)dst + units * 16, rest);
        }
        return cudaSuccess;
    }
    const int grid = grid_for(bytes, (size_t)kThreads * 4, ctx.max_blocks);
    return launch_pdl(ctx, copy_bytes_kernel, grid, kThreads, 0, (const unsigned char *)src, (unsigned char *)dst, bytes);
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

cudaError_t launch_sum(const LaunchCtx &ctx, int dtype, const void *in, size_t n, void *partials, unsigned int *ticket,
                       void *out, size_t divisor, const SumPdl &pdl)
{
    (void)cudaGetLastError();  // a stale error of an earlier, unchecked call must not be blamed on this launch
    int blocks, threads, vec, threads2;
    size_t chunk;
    sum_plan(dtype, n, &blocks, &chunk, &threads, &vec, &threads2);
    XchgArgs none;
    memset(&none, 0, sizeof none);
    return launch_sum_dtype<false>(ctx, dtype, in, n, blocks, chunk, partials, ticket, out, divisor, none, pdl);
}

cudaError_t launch_sum_exchange(const LaunchCtx &ctx, int dtype, const void *in, size_t n, void *partials, unsigned int *ticket,
                                void *out, size_t divisor, const XchgArgs &x, const SumPdl &pdl)
{
    (void)cudaGetLastError();
    int blocks = 1, threads, vec, threads2;
    size_t chunk = 0;
    if (n) sum_plan(dtype, n, &blocks, &chunk, &threads, &vec, &threads2);
    return launch_sum_dtype<true>(ctx, dtype, in, n, blocks, chunk, partials, ticket, out, divisor, x, pdl);
}

cudaError_t launch_iota16(const LaunchCtx &ctx, void *out)
{
    (void)cudaGetLastError();
    iota16_kernel<<<65536 / 256, 256, 0, ctx.stream>>>((unsigned short *)out);
    return cudaGetLastError();
}

template <int THREADS, int UNROLL, int GRAB, int MODE>
static cudaError_t launch_lut16_t(const LaunchCtx &ctx, int sm_count, const void *in, void *out, void *seed, size_t n,
                                  const void *table, unsigned long long *counters)
{
    cudaError_t e = cudaFuncSetAttribute(lut16_kernel<THREADS, UNROLL, GRAB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kLutBytes);
    if (e != cudaSuccess) return e;
    const size_t block_tiles = n / 8 / ((size_t)THREADS * UNROLL);
    int grid = sm_count;
    if ((size_t)grid > block_tiles + 1) grid = (int)(block_tiles + 1);
    return launch_pdl(ctx, lut16_kernel<THREADS, UNROLL, GRAB, MODE>, grid, THREADS, kLutBytes, (const unsigned short *)in,
                      (unsigned short *)out, (unsigned short *)seed, n, (const uint4 *)table, counters);
}

cudaError_t launch_lut16(const LaunchCtx &ctx, int sm_count, const void *in, void *out, size_t n, const void *table,
                         unsigned long long *counters, int shape)
{
    (void)cudaGetLastError();
    switch (shape) {  // CB_LUT_SHAPE: launch shapes kept for A/B measurements (profiles/r2_lut16_shapes*.log)
    case 1: return launch_lut16_t<1024, 4, 8, 0>(ctx, sm_count, in, out, nullptr, n, table, counters);
    case 2: return launch_lut16_t<512, 8, 2, 0>(ctx, sm_count, in, out, nullptr, n, table, counters);
    case 3: return launch_lut16_t<1024, 4, 4, 0>(ctx, sm_count, in, out, nullptr, n, table, counters);
    case 4: return launch_lut16_t<256, 16, 2, 0>(ctx, sm_count, in, out, nullptr, n, table, counters);
    default: return launch_lut16_t<512, 8, 4, 0>(ctx, sm_count, in, out, nullptr, n, table, counters);
    }
}

cudaError_t launch_lut16_grad_seed(const LaunchCtx &ctx, int sm_count, int dtype, const void *x, void *x_grad, void *out_grad,
                                   size_t n, const void *table, unsigned long long *counters)
{
    (void)cudaGetLastError();
    if (dtype == CB_BF16) return launch_lut16_t<512, 4, 8, 2>(ctx, sm_count, x, x_grad, out_grad, n, table, counters);
    return launch_lut16_t<512, 4, 8, 1>(ctx, sm_count, x, x_grad, out_grad, n, table, counters);
}

cudaError_t launch_lut16_grad_mul(const LaunchCtx &ctx, int sm_count, int dtype, const void *x, void *x_grad, const void *out_grad,
                                  size_t n, const void *table, unsigned long long *counters)
{
    (void)cudaGetLastError();
    void *og = const_cast<void *>(out_grad);  // read only in these modes
    if (dtype == CB_BF16) return launch_lut16_t<512, 4, 8, 4>(ctx, sm_count, x, x_grad, og, n, table, counters);
    return launch_lut16_t<512, 4, 8, 3>(ctx, sm_count, x, x_grad, og, n, table, counters);
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
    case CB_U64:
        fold_ranks_kernel<unsigned long long><<<1, 32, 0, ctx.stream>>>((const unsigned long long *)gathered, n_ranks,
                                                                       (unsigned long long *)out, divisor);
        break;
    default:
        fold_ranks_kernel<long long><<<1, 32, 0, ctx.stream>>>((const long long *)gathered, n_ranks, (long long *)out, divisor);
        break;
    }
    return cudaGetLastError();
}

}  // namespace cb
