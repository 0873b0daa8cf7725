// kernels.h — host launchers of the ahead-of-time compiled kernels in kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/custos_b200.h"

namespace cb {

struct LaunchCtx {
    cudaStream_t stream;
    int max_blocks;  // persistent-grid cap: SM count x resident blocks per SM
    bool pdl;        // launch the kernels that wait in pdl_enter() with programmatic stream serialisation
};

// fixed, hardware-independent upper bound on pass-1 blocks of the sum (8 per SM on a 148-SM B200)
constexpr int kSumMaxBlocks = 1184;

cudaError_t launch_binary(const LaunchCtx &ctx, int dtype, int op, const void *lhs, const void *rhs, void *out, size_t n);
cudaError_t launch_fill(const LaunchCtx &ctx, void *out, size_t n, int elem_bytes, uint64_t pattern);
int launch_fill_count(const void *out, size_t n, int elem_bytes);
cudaError_t launch_copy(const LaunchCtx &ctx, void *dst, const void *src, size_t bytes);
int launch_copy_count(const void *dst, const void *src, size_t bytes);
void sum_plan(int dtype, size_t n, int *blocks, size_t *chunk, int *threads, int *vec, int *threads2);
// partials: device scratch of kSumMaxBlocks accumulators; out: device scalar of the accumulation type
// Programmatic dependent launch of consecutive sums (see sum_kernel): `parity` selects one of the two sets of partials
// / tickets, `after_sum` says that the kernel before this one on the stream is a sum of the other parity.
struct SumPdl {
    bool enabled;
    bool after_sum;
    int parity;
};
// (`partials`: 2 x kSumMaxBlocks accumulators; `ticket`: two zero-initialised device counters owned by the device, left at zero)
cudaError_t launch_sum(const LaunchCtx &ctx, int dtype, const void *in, size_t n, void *partials, unsigned int *ticket, void *out,
                       size_t divisor, const SumPdl &pdl);
// cross-GPU exchange of reduction totals through peer-mapped memory (see kernels.cu)
constexpr int kMaxRanks = 64;
struct XchgSlot {
    unsigned long long value;  // bit pattern of the rank's total (accumulation type)
    unsigned long long epoch;  // call number that wrote it
};
struct XchgArgs {
    XchgSlot *const *peers;  // device array: peers[r] = rank r's exchange buffer (2 x n_ranks slots), peer mapped
    int n_ranks, rank;
    unsigned long long epoch;
    long long timeout_cycles;
    int *status;  // host-mapped flag: the kernel stores the call number here when a peer never arrived
};
cudaError_t launch_sum_exchange(const LaunchCtx &ctx, int dtype, const void *in, size_t n, void *partials, unsigned int *ticket,
                                void *out, size_t divisor, const XchgArgs &x, const SumPdl &pdl);
// 16-bit unary chains as a table lookup (see kernels.cu): table = 65 536 results, counters = 2 zeroed u64
cudaError_t launch_iota16(const LaunchCtx &ctx, void *out);
cudaError_t launch_lut16(const LaunchCtx &ctx, int sm_count, const void *in, void *out, size_t n, const void *table,
                         unsigned long long *counters, int shape);
// seeded backward of a 16-bit chain: x_grad[i] += table[x[i]], out_grad[i] = 1 (dtype CB_F16 or CB_BF16)
cudaError_t launch_lut16_grad_seed(const LaunchCtx &ctx, int sm_count, int dtype, const void *x, void *x_grad, void *out_grad,
                                   size_t n, const void *table, unsigned long long *counters);
// add_unary_grad of one 16-bit op with a general out_grad: x_grad[i] += out_grad[i] * table[x[i]] (table = g(x))
cudaError_t launch_lut16_grad_mul(const LaunchCtx &ctx, int sm_count, int dtype, const void *x, void *x_grad, const void *out_grad,
                                  size_t n, const void *table, unsigned long long *counters);
cudaError_t launch_fold_ranks(const LaunchCtx &ctx, int dtype, const void *gathered, int n_ranks, void *out, size_t divisor);

}  // namespace cb
