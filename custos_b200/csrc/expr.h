// expr.h — the two_way_ops expression IR on the host: validation, the reference's
// `to_cl_source()` rendering, and the typed CUDA source this backend compiles.
#pragma once
#include <string>
#include <vector>

#include "common.h"

namespace cb {

constexpr int kMaxNodes = 256;

// true for ops taking two operands
bool op_is_binary(int32_t op);
bool op_is_unary(int32_t op);

// Checks the node array (topological order, operand indices, ops legal for the dtype,
// markers legal for the kernel kind).  kind < 0 = no kernel-kind restriction.
int32_t expr_validate(int32_t dtype, int32_t kind, const cb_node *nodes, int32_t n);
int32_t chain_validate(int32_t dtype, int32_t kind, const cb_node *const *progs, const int32_t *n_nodes,
                       int32_t n_progs);

// Rust `{:?}` of a literal of the dtype (src/two_way_ops/to_cl_source.rs:7-12)
std::string rust_debug_literal(int32_t dtype, const cb_node &c);

// `to_cl_source()` of the tree, the reference's format strings
std::string expr_to_cl_source(int32_t dtype, const cb_node *nodes, int32_t n, const char *marker_x,
                              const char *marker_y);

// The generated part of the translation unit: `cb_fn(x, y)` applying the programs in order.
std::string expr_cuda_function(int32_t dtype, int32_t kind, const cb_node *const *progs, const int32_t *n_nodes,
                               int32_t n_progs,
                               bool fuse_scale_add = false);

// CB_KERNEL_CHAIN_GRAD: K forward programs + K grad programs joined into the one two-marker expression
// (X = chain input, Y = out_grad) whose value the kernel adds to the input's gradient; see expr.cpp
std::vector<cb_node> chain_grad_tree(const cb_node *const *progs, const int32_t *n_nodes, int32_t n_progs, bool canon_all);

// 64-bit key of (dtype, kind, programs) for the kernel cache
// canonical byte string of (dtype, kind, programs): what chain_hash hashes, kept to confirm cache hits
std::string chain_bytes(int32_t dtype, int32_t kind, const cb_node *const *progs, const int32_t *n_nodes,
                        int32_t n_progs);
uint64_t chain_hash(int32_t dtype, int32_t kind, const cb_node *const *progs, const int32_t *n_nodes,
                    int32_t n_progs);

// IEEE binary16 <-> binary32, round to nearest even (host side, for f16 literals)
uint16_t host_f32_to_f16(float v);
float host_f16_to_f32(uint16_t h);
uint16_t host_f32_to_bf16(float v);
float host_bf16_to_f32(uint16_t h);
uint16_t host_f32_to_half(int32_t dtype, float v);  // dtype CB_F16 or CB_BF16
float host_half_to_f32(int32_t dtype, uint16_t h);

}  // namespace cb
