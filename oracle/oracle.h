/*
 * oracle.h — CPU restatement of the custos reference's CPU device for the hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under custos_b200/ may include, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, and only as the checker / reported baseline.
 *
 * Why a restatement: the reference is a Rust crate (edition 2024) and this image has
 * no rustc/cargo, so it cannot be compiled into oracle/_ref.  Parity is pinned by the
 * reference's own known-answer tests instead (tests/test_oracle_kat.py lists each
 * with its reference file:line).  Parity unpinned for: f16 results (the reference
 * has no f16 result test; arithmetic follows the published `half` 2.x algorithm) and
 * sum/mean (the reference has no such op; the order is defined in DESIGN.md).
 */
#ifndef CUSTOS_ORACLE_H
#define CUSTOS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same numbering as include/custos_b200.h (cb_dtype / cb_opcode / cb_node) */
enum { ORC_F32 = 0, ORC_F64 = 1, ORC_F16 = 2, ORC_I32 = 3, ORC_I64 = 4, ORC_U32 = 5, ORC_U8 = 6,
       ORC_BF16 = 7, ORC_I8 = 8, ORC_I16 = 9, ORC_U16 = 10, ORC_U64 = 11, ORC_BOOL = 12, ORC_DTYPE_COUNT = 13 };
enum {
    ORC_OP_X = 0, ORC_OP_Y, ORC_OP_CONST, ORC_OP_ADD, ORC_OP_MUL, ORC_OP_SUB, ORC_OP_DIV, ORC_OP_POW,
    ORC_OP_MIN, ORC_OP_MAX, ORC_OP_SIN, ORC_OP_COS, ORC_OP_TAN, ORC_OP_TANH, ORC_OP_EXP, ORC_OP_LN,
    ORC_OP_ABS, ORC_OP_NEG, ORC_OP_IDENTITY, ORC_OP_GEQ, ORC_OP_LEQ, ORC_OP_EQ, ORC_OP_COUNT
};
typedef struct orc_node {
    int32_t op, a, b, _pad;
    double fimm;
    int64_t iimm;
} orc_node;

#define ORC_MAX_NODES 256
enum { ORC_OK = 0, ORC_ERR_ARG = 1, ORC_ERR_UNSUPPORTED = 4, ORC_ERR_EXPR = 5 };

size_t orc_dtype_size(int dtype);

/* half crate 2.x software conversions (round to nearest even) */
uint16_t orc_f32_to_f16(float v);
float orc_f16_to_f32(uint16_t h);
/* ulp distance statistics between two f32 arrays (test helper; see oracle.c) */
uint32_t orc_ulp_stats_f32(const float *got, const float *want, size_t n, uint64_t hist[6], size_t *argmax);
/* (u * P) + C == fmaf(u, P, C) [mode 0] / (u + C) * P == fmaf(u, P, C * P) [mode 1] over a range of f32 bit patterns */
void orc_fmaf_array(const float *a, float b, float c, float *out, size_t n);
uint64_t orc_check_scale_add(uint64_t first, uint64_t count, float P, float C, int mode, int threads, uint32_t *first_bad);
uint16_t orc_f32_to_bf16(float v); /* half::bf16::from_f32 */
float orc_bf16_to_f32(uint16_t h);

/* Eval::eval of one expression on scalars (src/two_way_ops/eval.rs, ops.rs, ops/unary.rs, ops/cmps.rs) */
int orc_eval(int dtype, const orc_node *nodes, int n, const void *x, const void *y, void *out);

/* apply_fn_slice (src/devices/cpu_stack_ops.rs:7-15) */
int orc_apply_fn(int dtype, const orc_node *nodes, int n, const void *x, void *out, size_t len);
/* CPU UnaryFusing::unary_fuse_op (src/devices/cpu/cpu_device.rs:204-232): ops applied in order per element */
int orc_apply_chain(int dtype, const orc_node *const *progs, const int *n_nodes, int n_progs,
                    const void *x, void *out, size_t len);
/* same, split over `threads` OS threads (NOT what the reference does; labelled in bench output) */
int orc_apply_chain_mt(int dtype, const orc_node *const *progs, const int *n_nodes, int n_progs,
                       const void *x, void *out, size_t len, int threads);
/* f32 only — the two single-threaded CPU baselines of SURVEY.md §8(d): (i) the fused path as the reference runs it, a
 * heap-boxed dyn op built, evaluated and dropped per element and per op (src/devices/cpu/cpu_device.rs:217-229,
 * src/op_hint.rs:30-33); (ii) one apply_fn_slice loop per recorded op (src/devices/cpu_stack_ops.rs:7-15).  Same bits
 * as orc_apply_chain. */
int orc_apply_chain_boxed_f32(const orc_node *const *progs, const int *n_nodes, int n_progs, const float *x, float *out,
                              size_t len);
int orc_apply_chain_unfused_f32(const orc_node *const *progs, const int *n_nodes, int n_progs, const float *x, float *out,
                                size_t len);
/* add_unary_grad (src/devices/cpu_stack_ops.rs:18-30): lhs_grad += out * g(lhs) */
int orc_add_unary_grad(int dtype, const orc_node *nodes, int n, const void *lhs, const void *out_grad,
                       void *lhs_grad, size_t len);
/* two-marker closure evaluated element-wise */
int orc_apply2(int dtype, const orc_node *nodes, int n, const void *lhs, const void *rhs, void *out,
               size_t len);
/* binary add/mul/sub/div (tests/demo_impl/cpu.rs:12-43, src/lib.rs:293-301); op: 0 add 1 mul 2 sub 3 div */
int orc_binary(int dtype, int op, const void *lhs, const void *rhs, void *out, size_t len);
/* clear_slice (src/devices/cpu_stack_ops.rs:33-37) */
int orc_clear(int dtype, void *buf, size_t len);

/* sums: sequential left-to-right in the accumulation type (f32->f32, f64->f64, f16->f32, ints->i64) */
int orc_sum_seq(int dtype, const void *in, size_t len, void *out);
/* ground truth: sequential fp64 accumulation of the converted values */
double orc_sum_f64(int dtype, const void *in, size_t len);
/* the device's documented two-pass order restated on the CPU (bit-exact target):
 * `blocks` chunks of `chunk` elements; inside a block, thread t owns the vector
 * units t, t+threads, ... (a unit = `vec` consecutive elements) and keeps one
 * accumulator per vector lane, lanes are folded left to right, thread totals go
 * through the xor-shuffle tree (16,8,4,2,1), warp totals through the same tree in
 * warp 0, and pass 2 folds the block partials with the same scheme (vec = 1). */
int orc_sum_two_pass(int dtype, const void *in, size_t len, int blocks, size_t chunk, int threads,
                     int vec, int threads2, void *out);

/* OptGraph (src/modules/graph/opt_graph.rs:6-41, opt_graph/optimize.rs:19-132) */
typedef struct orc_graph orc_graph;
orc_graph *orc_graph_new(void);
void orc_graph_free(orc_graph *g);
int64_t orc_graph_add_leaf(orc_graph *g, size_t len);
int64_t orc_graph_add_node(orc_graph *g, size_t len, const int64_t *deps, int n_deps);
void orc_graph_set_skip(orc_graph *g, int64_t idx, int skip);
int orc_graph_is_leaf(const orc_graph *g, int64_t idx);
int orc_graph_is_path_optimizable(const orc_graph *g, int64_t idx);
size_t orc_graph_trace_cache_path_raw(const orc_graph *g, int64_t idx, int64_t *out, size_t cap);
/* flattened [cache_idx, k, use_0..use_{k-1}]* ; returns number of int64 written */
size_t orc_graph_cache_traces(const orc_graph *g, int64_t *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
