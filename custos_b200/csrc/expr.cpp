// expr.cpp — host side of the two_way_ops expression IR.
//
// Reference behaviour mirrored here:
//  * the Combiner tree and its `to_cl_source()` strings (src/two_way_ops/ops.rs,
//    ops/unary.rs, ops/cmps.rs) — expr_to_cl_source reproduces them byte for byte,
//    including Rust's `{:?}` rendering of literals (to_cl_source.rs:7-12);
//  * deliberately NOT mirrored: the CUDA we compile uses typed
//    literal bit patterns (no double promotion), ternary min/max, `<=` for Eq — i.e. the
//    semantics of the CPU `Eval` impls, which are the parity target.
#include "expr.h"

#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>

namespace cb {

bool op_is_binary(int32_t op)
{
    return (op >= CB_OP_ADD && op <= CB_OP_MAX) || (op >= CB_OP_GEQ && op <= CB_OP_EQ);
}
bool op_is_unary(int32_t op) { return op >= CB_OP_SIN && op <= CB_OP_IDENTITY; }

static const char *op_name(int32_t op)
{
    static const char *names[CB_OP_COUNT] = {"x",   "y",   "const", "add", "mul",  "sub", "div", "pow",
                                             "min", "max", "sin",   "cos", "tan",  "tanh", "exp", "ln",
                                             "abs", "neg", "identity", "geq", "leq", "eq"};
    return (op >= 0 && op < CB_OP_COUNT) ? names[op] : "?";
}

// Which ops exist for which T in the reference: Add/Mul/Sub/Div need core::ops (all
// numbers); Neg needs core::ops::Neg (floats, signed ints); GEq/LEq/Eq need Number; every
// other op is bounded by `T: Float` (ops.rs:233,271,309; ops/unary.rs).
static bool op_supported(int32_t dtype, int32_t op)
{
    if (is_float_dtype(dtype)) return true;
    if (dtype == CB_BOOL) return false;  // bool is a CDatatype but not a Number: storage only
    switch (op) {
    case CB_OP_X: case CB_OP_Y: case CB_OP_CONST: case CB_OP_ADD: case CB_OP_MUL: case CB_OP_SUB:
    case CB_OP_DIV: case CB_OP_GEQ: case CB_OP_LEQ: case CB_OP_EQ:
        return true;
    case CB_OP_NEG: return is_signed_int_dtype(dtype);
    default: return false;
    }
}

int32_t expr_validate(int32_t dtype, int32_t kind, const cb_node *nodes, int32_t n)
{
    if (!valid_dtype(dtype)) return fail(CB_ERR_INVALID_ARG, "invalid dtype %d", dtype);
    if (!nodes || n <= 0) return fail(CB_ERR_EXPR, "empty expression");
    if (n > kMaxNodes) return fail(CB_ERR_EXPR, "expression has %d nodes (max %d)", n, kMaxNodes);
    for (int32_t i = 0; i < n; i++) {
        const cb_node &c = nodes[i];
        if (c.op < 0 || c.op >= CB_OP_COUNT) return fail(CB_ERR_EXPR, "node %d: unknown opcode %d", i, c.op);
        if (!op_supported(dtype, c.op))
            return fail(CB_ERR_UNSUPPORTED, "node %d: op '%s' is not implemented for dtype %s in the reference",
                        i, op_name(c.op), dtype_name(dtype));
        if (c.op == CB_OP_Y && kind >= 0 && kind != CB_KERNEL_BINARY)
            return fail(CB_ERR_EXPR, "node %d: second marker in a unary expression", i);
        if (op_is_binary(c.op) || op_is_unary(c.op)) {
            if (c.a < 0 || c.a >= i) return fail(CB_ERR_EXPR, "node %d: operand a=%d out of order", i, c.a);
            if (op_is_binary(c.op) && (c.b < 0 || c.b >= i))
                return fail(CB_ERR_EXPR, "node %d: operand b=%d out of order", i, c.b);
        }
        // unused operand slots must be "none" (negative): every consumer may then index by a / b whenever it is >= 0
        if (!op_is_binary(c.op) && c.b >= 0) return fail(CB_ERR_EXPR, "node %d: op '%s' takes no operand b (got %d)", i, op_name(c.op), c.b);
        if (!op_is_binary(c.op) && !op_is_unary(c.op) && c.a >= 0)
            return fail(CB_ERR_EXPR, "node %d: '%s' is a leaf and takes no operand a (got %d)", i, op_name(c.op), c.a);
        if (c.op == CB_OP_CONST && is_half_dtype(dtype)) {
            if (std::isfinite(c.fimm) && (double)host_half_to_f32(dtype, host_f32_to_half(dtype, (float)c.fimm)) != c.fimm)
                return fail(CB_ERR_EXPR, "node %d: literal %.17g is not representable in %s (round it first)", i,
                            c.fimm, dtype_name(dtype));
        }
        if (c.op == CB_OP_CONST && dtype == CB_F32 && std::isfinite(c.fimm) && (double)(float)c.fimm != c.fimm)
            return fail(CB_ERR_EXPR, "node %d: literal %.17g is not representable in f32 (round it first)", i, c.fimm);
    }
    return CB_OK;
}

int32_t chain_validate(int32_t dtype, int32_t kind, const cb_node *const *progs, const int32_t *n_nodes,
                       int32_t n_progs)
{
    if (kind < CB_KERNEL_APPLY || kind > CB_KERNEL_CHAIN_GRAD) return fail(CB_ERR_INVALID_ARG, "invalid kernel kind %d", kind);
    if (!progs || !n_nodes || n_progs <= 0) return fail(CB_ERR_EXPR, "no programs");
    if (kind != CB_KERNEL_APPLY && kind != CB_KERNEL_CHAIN_GRAD && n_progs != 1)
        return fail(CB_ERR_EXPR, "only apply and chain-grad kernels take a chain of programs");
    if (kind == CB_KERNEL_CHAIN_GRAD && (n_progs % 2) != 0)
        return fail(CB_ERR_EXPR, "a chain-grad kernel takes K forward programs followed by their K grad programs (got %d)", n_progs);
    if (n_progs > 128 || (kind == CB_KERNEL_APPLY && n_progs > 64)) return fail(CB_ERR_EXPR, "chain of %d programs (max 64)", n_progs);
    // the programs of a chain-grad kernel are unary closures (one marker), like the ops of a fused chain
    const int32_t prog_kind = kind == CB_KERNEL_CHAIN_GRAD ? CB_KERNEL_APPLY : kind;
    for (int32_t k = 0; k < n_progs; k++) CB_TRY(expr_validate(dtype, prog_kind, progs[k], n_nodes[k]));
    if (kind == CB_KERNEL_CHAIN_GRAD && chain_grad_tree(progs, n_nodes, n_progs, true).size() > (size_t)kMaxNodes * 4)
        return fail(CB_ERR_EXPR, "chain-grad expression too large");
    return CB_OK;
}

// ------------------------------------------------------------------ binary16 (host)
uint16_t host_f32_to_f16(float v)
{
    uint32_t u;
    std::memcpy(&u, &v, 4);
    const uint32_t sign = (u >> 16) & 0x8000u;
    u &= 0x7fffffffu;
    if (u >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (u > 0x7f800000u ? 0x0200u | ((u >> 13) & 0x3ffu) : 0u));
    if (u >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);  // rounds to >= 65520 -> inf
    if (u < 0x33000001u) return (uint16_t)sign;                // <= 2^-25 -> zero (tie goes to even = 0)
    // scale so that the 11-bit result sits in an integer, then round to nearest even
    const int e = (int)(u >> 23);
    uint32_t sig = (u & 0x7fffffu) | 0x800000u;
    int shift = (e >= 113) ? 13 : (126 - e);  // normal halfs drop 13 bits, subnormals more
    const uint32_t q = sig >> shift;
    const uint32_t rem = sig & ((1u << shift) - 1u);
    const uint32_t half = 1u << (shift - 1);
    uint32_t r = q + ((rem > half || (rem == half && (q & 1u))) ? 1u : 0u);
    if (e >= 113) r += (uint32_t)(e - 113) << 10;  // r holds the implicit bit at 0x400: exponent field e-112
    return (uint16_t)(sign | r);
}

float host_f16_to_f32(uint16_t h)
{
    const int s = (h >> 15) & 1, e = (h >> 10) & 0x1f, m = h & 0x3ff;
    float r;
    if (e == 0) r = std::ldexp((float)m, -24);
    else if (e == 31) r = m ? NAN : INFINITY;
    else r = std::ldexp((float)(m | 0x400), e - 25);
    return s ? -r : r;
}

// ------------------------------------------------------------------ bfloat16 (host)
// half::bf16::from_f32: NaN keeps its top payload bits with the quiet bit forced, everything else is
// round-to-nearest-even on the upper 16 bits (overflow carries into infinity by itself).
uint16_t host_f32_to_bf16(float v)
{
    uint32_t u;
    std::memcpy(&u, &v, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);
    const uint32_t round_bit = 0x8000u;
    if ((u & round_bit) && (u & (3u * round_bit - 1u))) return (uint16_t)((u >> 16) + 1u);
    return (uint16_t)(u >> 16);
}

float host_bf16_to_f32(uint16_t h)
{
    const uint32_t u = (uint32_t)h << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

uint16_t host_f32_to_half(int32_t dtype, float v) { return dtype == CB_BF16 ? host_f32_to_bf16(v) : host_f32_to_f16(v); }
float host_half_to_f32(int32_t dtype, uint16_t h) { return dtype == CB_BF16 ? host_bf16_to_f32(h) : host_f16_to_f32(h); }

// ------------------------------------------------------------ Rust `{:?}` for numbers
// core::fmt::float: shortest digits that round-trip; decimal with at least one
// fractional digit when 1e-4 <= |v| < 1e16 (or v == 0), exponential ("1e16", "1.5e-7")
// otherwise; "NaN", "inf", "-inf".
static std::string rust_debug_float(double v, bool single)
{
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    std::string out;
    if (std::signbit(v)) out += "-";
    const double a = std::fabs(v);
    if (a == 0.0) return out + "0.0";

    char buf[64];
    int prec = 1;
    for (; prec <= 17; prec++) {
        std::snprintf(buf, sizeof buf, "%.*e", prec - 1, a);
        if (single ? (std::strtof(buf, nullptr) == (float)a) : (std::strtod(buf, nullptr) == a)) break;
    }
    // buf = d.ddddde[+-]XX
    std::string digits;
    const char *p = buf;
    for (; *p && *p != 'e'; p++)
        if (*p != '.') digits += *p;
    const int exp10 = std::atoi(p + 1);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int nd = (int)digits.size();

    if (a >= 1e-4 && a < 1e16) {
        if (exp10 >= nd - 1) {
            out += digits + std::string((size_t)(exp10 - (nd - 1)), '0') + ".0";
        } else if (exp10 >= 0) {
            out += digits.substr(0, (size_t)exp10 + 1) + "." + digits.substr((size_t)exp10 + 1);
        } else {
            out += "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
        }
    } else {
        out += digits.substr(0, 1);
        if (nd > 1) out += "." + digits.substr(1);
        out += "e" + std::to_string(exp10);
    }
    return out;
}

std::string rust_debug_literal(int32_t dtype, const cb_node &c)
{
    switch (dtype) {
    case CB_F32: return rust_debug_float((double)(float)c.fimm, true);
    case CB_F64: return rust_debug_float(c.fimm, false);
    case CB_F16: case CB_BF16:  // half: Debug via f32
        return rust_debug_float((double)host_half_to_f32(dtype, host_f32_to_half(dtype, (float)c.fimm)), true);
    case CB_U32: return std::to_string((uint32_t)c.iimm);
    case CB_U8: return std::to_string((unsigned)(uint8_t)c.iimm);
    case CB_U16: return std::to_string((unsigned)(uint16_t)c.iimm);
    case CB_U64: return std::to_string((unsigned long long)c.iimm);
    case CB_I8: return std::to_string((int)(int8_t)c.iimm);
    case CB_I16: return std::to_string((int)(int16_t)c.iimm);
    case CB_I32: return std::to_string((int32_t)c.iimm);
    default: return std::to_string((long long)c.iimm);
    }
}

std::string expr_to_cl_source(int32_t dtype, const cb_node *nodes, int32_t n, const char *marker_x,
                              const char *marker_y)
{
    std::vector<std::string> s((size_t)n);
    for (int32_t i = 0; i < n; i++) {
        const cb_node &c = nodes[i];
        const std::string a = c.a >= 0 ? s[(size_t)c.a] : std::string();
        const std::string b = c.b >= 0 ? s[(size_t)c.b] : std::string();
        switch (c.op) {
        case CB_OP_X: s[i] = marker_x ? marker_x : "x"; break;
        case CB_OP_Y: s[i] = marker_y ? marker_y : "y"; break;
        case CB_OP_CONST: s[i] = rust_debug_literal(dtype, c); break;
        case CB_OP_ADD: s[i] = "(" + a + " + " + b + ")"; break;
        case CB_OP_MUL: s[i] = "(" + a + " * " + b + ")"; break;
        case CB_OP_SUB: s[i] = "(" + a + " - " + b + ")"; break;
        case CB_OP_DIV: s[i] = "(" + a + " / " + b + ")"; break;
        case CB_OP_POW: s[i] = "pow(" + a + ", " + b + ")"; break;
        case CB_OP_MIN: s[i] = "min(" + a + ", " + b + ")"; break;
        case CB_OP_MAX: s[i] = "max(" + a + ", " + b + ")"; break;
        case CB_OP_SIN: s[i] = "sin(" + a + ")"; break;
        case CB_OP_COS: s[i] = "cos(" + a + ")"; break;
        case CB_OP_TAN: s[i] = "tan(" + a + ")"; break;
        case CB_OP_TANH: s[i] = "tanh(" + a + ")"; break;
        case CB_OP_EXP: s[i] = "exp(" + a + ")"; break;
        case CB_OP_LN: s[i] = "log(" + a + ")"; break;
        case CB_OP_ABS: s[i] = "abs(" + a + ")"; break;
        case CB_OP_NEG: s[i] = "-(" + a + ")"; break;
        case CB_OP_IDENTITY: s[i] = a; break;
        case CB_OP_GEQ: s[i] = "(" + a + " >= " + b + ")"; break;
        case CB_OP_LEQ: s[i] = "(" + a + " <= " + b + ")"; break;
        case CB_OP_EQ: s[i] = "(" + a + " == " + b + ")"; break;
        default: break;
        }
    }
    return s[(size_t)n - 1];
}

// ------------------------------------------------------------------ CUDA codegen
// A literal is emitted as its exact bit pattern in the compute dtype, so `x * 2.0`
// is an f32 multiply by 2.0f (the reference CUDA path would promote to double).
static std::string cuda_literal(int32_t dtype, const cb_node &c)
{
    char buf[96];
    switch (dtype) {
    case CB_F32: {
        const float f = (float)c.fimm;
        uint32_t u;
        std::memcpy(&u, &f, 4);
        std::snprintf(buf, sizeof buf, "__uint_as_float(0x%08xu)", u);
    } break;
    case CB_F64: {
        uint64_t u;
        std::memcpy(&u, &c.fimm, 8);
        std::snprintf(buf, sizeof buf, "__longlong_as_double((long long)0x%016llxULL)", (unsigned long long)u);
    } break;
    case CB_F16: case CB_BF16:
        std::snprintf(buf, sizeof buf, "((T)0x%04xu)", (unsigned)host_f32_to_half(dtype, (float)c.fimm));
        break;
    case CB_I8: std::snprintf(buf, sizeof buf, "((T)0x%02xu)", (unsigned)(uint8_t)c.iimm); break;
    case CB_I16: case CB_U16: std::snprintf(buf, sizeof buf, "((T)0x%04xu)", (unsigned)(uint16_t)c.iimm); break;
    case CB_I32: std::snprintf(buf, sizeof buf, "((T)0x%08xu)", (unsigned)(uint32_t)(int32_t)c.iimm); break;
    case CB_U32: std::snprintf(buf, sizeof buf, "((T)0x%08xu)", (unsigned)(uint32_t)c.iimm); break;
    case CB_U8: std::snprintf(buf, sizeof buf, "((T)0x%02xu)", (unsigned)(uint8_t)c.iimm); break;
    default: std::snprintf(buf, sizeof buf, "((T)0x%016llxULL)", (unsigned long long)c.iimm); break;
    }
    return buf;
}

static const char *cuda_fn(int32_t op)
{
    switch (op) {
    case CB_OP_ADD: return "cb_add";
    case CB_OP_MUL: return "cb_mul";
    case CB_OP_SUB: return "cb_sub";
    case CB_OP_DIV: return "cb_div";
    case CB_OP_POW: return "cb_pow";
    case CB_OP_MIN: return "cb_min";
    case CB_OP_MAX: return "cb_max";
    case CB_OP_SIN: return "cb_sin";
    case CB_OP_COS: return "cb_cos";
    case CB_OP_TAN: return "cb_tan";
    case CB_OP_TANH: return "cb_tanh";
    case CB_OP_EXP: return "cb_exp";
    case CB_OP_LN: return "cb_ln";
    case CB_OP_ABS: return "cb_abs";
    case CB_OP_NEG: return "cb_neg";
    case CB_OP_IDENTITY: return "cb_identity";
    case CB_OP_GEQ: return "cb_geq";
    case CB_OP_LEQ: return "cb_leq";
    case CB_OP_EQ: return "cb_eq";
    default: return "cb_bad";
    }
}


// ------------------------------------------------------------------ joined trees (hash-consed)
// Several programs joined into ONE tree: the marker of a program is the value of an earlier root.  Identical
// nodes (same op, same operands, same literal) are shared, so `exp(x)` computed by the forward op and again by its
// grad closure, or `tanh(x)` twice inside `1 - tanh(x) * tanh(x)`, is evaluated once.  Every op is a pure function
// of its operands, so sharing never changes a result.
namespace {
struct Dag {
    std::vector<cb_node> nodes;
    std::map<std::array<uint64_t, 5>, int32_t> index;
    int32_t add(cb_node c)
    {
        uint64_t fbits;
        std::memcpy(&fbits, &c.fimm, 8);
        if (c.op != CB_OP_CONST) {
            fbits = 0;
            c.fimm = 0.0;
            c.iimm = 0;
        }
        c._pad = 0;
        const std::array<uint64_t, 5> key = {(uint64_t)(uint32_t)c.op, (uint64_t)(uint32_t)c.a, (uint64_t)(uint32_t)c.b, fbits,
                                             (uint64_t)c.iimm};
        auto it = index.find(key);
        if (it != index.end()) return it->second;
        nodes.push_back(c);
        index.emplace(key, (int32_t)nodes.size() - 1);
        return (int32_t)nodes.size() - 1;
    }
    int32_t leaf(int32_t op, double f = 0.0, int64_t i = 0)
    {
        cb_node c;
        std::memset(&c, 0, sizeof c);
        c.op = op;
        c.a = c.b = -1;
        c.fimm = f;
        c.iimm = i;
        return add(c);
    }
    int32_t binary(int32_t op, int32_t a, int32_t b)
    {
        cb_node c;
        std::memset(&c, 0, sizeof c);
        c.op = op;
        c.a = a;
        c.b = b;
        return add(c);
    }
    // appends a unary program whose marker X is the node `x_root`; returns the index of its value
    int32_t append(const cb_node *prog, int32_t n, int32_t x_root)
    {
        std::vector<int32_t> map((size_t)n, -1);
        for (int32_t i = 0; i < n; i++) {
            cb_node c = prog[i];
            if (c.op == CB_OP_X) {
                map[(size_t)i] = x_root;
                continue;
            }
            if (c.a >= 0) c.a = map[(size_t)c.a];
            if (c.b >= 0) c.b = map[(size_t)c.b];
            map[(size_t)i] = add(c);
        }
        return map[(size_t)n - 1];
    }
    // the nodes `root` depends on, in topological order, re-indexed
    std::vector<cb_node> reachable(int32_t root) const
    {
        std::vector<char> keep(nodes.size(), 0);
        keep[(size_t)root] = 1;
        for (int32_t i = root; i >= 0; i--) {
            if (!keep[(size_t)i]) continue;
            if (nodes[(size_t)i].a >= 0) keep[(size_t)nodes[(size_t)i].a] = 1;
            if (nodes[(size_t)i].b >= 0) keep[(size_t)nodes[(size_t)i].b] = 1;
        }
        std::vector<int32_t> idx(nodes.size(), -1);
        std::vector<cb_node> out;
        for (int32_t i = 0; i <= root; i++) {
            if (!keep[(size_t)i]) continue;
            cb_node c = nodes[(size_t)i];
            if (c.a >= 0) c.a = idx[(size_t)c.a];
            if (c.b >= 0) c.b = idx[(size_t)c.b];
            idx[(size_t)i] = (int32_t)out.size();
            out.push_back(c);
        }
        return out;
    }
};
}  // namespace

// The backward of a fused unary chain as ONE two-marker expression (X = the chain's input x0, Y = out_grad[i]):
// what the K grad functions of `unary_ew` (src/unary.rs:118-128) accumulate into x0's gradient when the tape
// replays them in reverse (src/modules/autograd/tape.rs:39-47), with every intermediate recomputed from x0:
//     x_k = f_k(x_{k-1})                                   k = 1 .. K-1      (forward, in registers)
//     t_K = Y;   t_{k-1} = 0 + t_k * g_k(x_{k-1})          k = K .. 2        (`lhs_grad += out_grad * g(lhs)` into
//                                                                             a zero-initialised gradient buffer,
//                                                                             src/devices/cpu_stack_ops.rs:18-30)
//     value = t_1 * g_1(x_0)                                                  (the kernel adds it to x0.grad)
// Multiply and add stay separately rounded operations in exactly this order, so the result is bit-identical to the
// K unfused `add_unary_grad` kernels.  `0 + v` only turns a -0 product into +0; `canon_all = false` keeps that add
// for t_1 only: a zero stays a zero (of either sign) through every later multiply by a finite factor and becomes a
// NaN in both forms for an infinite one, so canonicalising once, before the last multiply, gives the same bits
// (NaN payloads are not preserved by either form).
std::vector<cb_node> chain_grad_tree(const cb_node *const *progs, const int32_t *n_nodes, int32_t n_progs, bool canon_all)
{
    const int32_t K = n_progs / 2;
    Dag dag;
    std::vector<int32_t> xs((size_t)K, -1);
    xs[0] = dag.leaf(CB_OP_X);
    for (int32_t k = 1; k < K; k++) xs[(size_t)k] = dag.append(progs[k - 1], n_nodes[k - 1], xs[(size_t)k - 1]);
    int32_t t = dag.leaf(CB_OP_Y);
    for (int32_t k = K; k >= 1; k--) {
        const int32_t g = dag.append(progs[K + k - 1], n_nodes[K + k - 1], xs[(size_t)k - 1]);
        t = dag.binary(CB_OP_MUL, t, g);
        if (k > 1 && (canon_all || k == 2)) t = dag.binary(CB_OP_ADD, dag.leaf(CB_OP_CONST, 0.0, 0), t);
    }
    return dag.reachable(t);
}

// right-hand side of one node of the f32 pair function
static std::string pair_rhs(const cb_node &c, const std::string &ta, const std::string &tb)
{
    if (c.op == CB_OP_X) return "x";
    if (c.op == CB_OP_Y) return "y";
    if (c.op == CB_OP_CONST) return "cb2_splat(" + cuda_literal(CB_F32, c) + ")";
    const std::string name = "cb2_" + std::string(cuda_fn(c.op) + 3);
    if (op_is_binary(c.op)) return name + "(" + ta + ", " + tb + ")";
    if (c.op == CB_OP_SIN || c.op == CB_OP_COS) return name + "(" + ta + ", redo)";  // fast path only; `redo` asks for the scalar forms
    return name + "(" + ta + ")";
}

// |v| = 2^k with k >= min_k (and finite): multiplying by it is exact apart from overflow / underflow
static bool is_pow2(double v, int min_k, int *k_out)
{
    if (!std::isfinite(v) || v == 0.0) return false;
    int e = 0;
    const double m = std::frexp(std::fabs(v), &e);  // |v| = m * 2^e, m in [0.5, 1)
    if (m != 0.5) return false;
    if (e - 1 < min_k) return false;
    if (k_out) *k_out = e - 1;
    return true;
}

// The f32 pair function of a whole chain with scale-and-shift steps strength-reduced to ONE fma, bit for bit:
//   A.  (u * P) + C  ->  fma(u, P, C)       P = +-2^k, k >= 0: u * P is exact (or overflows, and then both forms give
//                                            the same infinity because |C| <= 2^100 cannot bring the sum back)
//   B.  (u + C) * P  ->  fma(u, P, C * P)   P = +2^k, any k: scaling by a power of two commutes with rounding as long
//                                            as the result is normal, zero or overflows; a non-zero u + C is at least
//                                            2^(e_C - 24) in magnitude, so e_C - 24 + k >= -126 keeps it normal;
//                                            |C| <= 2^100 keeps u + C itself from overflowing, and P must be positive
//                                            because u = -C gives (+0) * P = -0 for a negative P but +0 when fused
// (the last two conditions were found by tests/test_scale_add_fusion.py, which interprets the generated text on the CPU)
// The chain is first joined into one tree (the marker of op k+1 is the value of op k), so a `mul(2.0)` op followed by
// an `add(1.0)` op fuses as well.  Only the pair function is rewritten: the scalar `cb_fn` (tails, unaligned
// buffers, the slow-path redo) keeps the two separately rounded operations, which makes every test that compares the
// vector path with the scalar path a check of the equivalence claimed here.
static std::string fused_pair_function(const cb_node *const *progs, const int32_t *n_nodes, int32_t n_progs)
{
    std::vector<cb_node> all;
    int32_t prev_root = -1;
    for (int32_t k = 0; k < n_progs; k++) {
        std::vector<int32_t> map((size_t)n_nodes[k], -1);
        for (int32_t i = 0; i < n_nodes[k]; i++) {
            cb_node c = progs[k][i];
            if (c.op == CB_OP_X && prev_root >= 0) {
                map[(size_t)i] = prev_root;
                continue;
            }
            if (c.a >= 0) c.a = map[(size_t)c.a];
            if (c.b >= 0) c.b = map[(size_t)c.b];
            map[(size_t)i] = (int32_t)all.size();
            all.push_back(c);
        }
        prev_root = map[(size_t)n_nodes[k] - 1];
    }
    const int32_t n = (int32_t)all.size();
    std::vector<int> uses((size_t)n, 0);
    for (const cb_node &c : all) {
        if (c.a >= 0) uses[(size_t)c.a]++;
        if (c.b >= 0) uses[(size_t)c.b]++;
    }
    uses[(size_t)prev_root]++;
    auto lit = [&all](int32_t i, double *v) {
        if (i < 0 || all[(size_t)i].op != CB_OP_CONST) return false;
        *v = (double)(float)all[(size_t)i].fimm;
        return true;
    };
    auto literal_text = [](double v) {
        cb_node c;
        std::memset(&c, 0, sizeof c);
        c.op = CB_OP_CONST;
        c.fimm = v;
        return "cb2_splat(" + cuda_literal(CB_F32, c) + ")";
    };
    std::vector<char> absorbed((size_t)n, 0), rewritten((size_t)n, 0);  // rewritten: already the outer node of a fused pair
    std::vector<std::string> rhs((size_t)n);
    for (int32_t i = 0; i < n; i++) {
        const cb_node &c = all[(size_t)i];
        rhs[(size_t)i] = pair_rhs(c, "t" + std::to_string(c.a), "t" + std::to_string(c.b));
        if (c.op != CB_OP_ADD && c.op != CB_OP_MUL) continue;
        for (int side = 0; side < 2; side++) {
            const int32_t inner = side ? c.b : c.a, other = side ? c.a : c.b;
            double outer_lit;
            if (!lit(other, &outer_lit) || uses[(size_t)inner] != 1 || rewritten[(size_t)inner]) continue;
            const cb_node &in = all[(size_t)inner];
            if (in.op != (c.op == CB_OP_ADD ? CB_OP_MUL : CB_OP_ADD)) continue;
            for (int iside = 0; iside < 2; iside++) {
                const int32_t u = iside ? in.b : in.a, ilit = iside ? in.a : in.b;
                double inner_lit;
                if (!lit(ilit, &inner_lit) || all[(size_t)u].op == CB_OP_CONST) continue;
                double P, C, addend;
                int k = 0;
                if (c.op == CB_OP_ADD) {  // A: (u * P) + C
                    P = inner_lit, C = outer_lit;
                    if (!is_pow2(P, 0, &k) || !std::isfinite(C) || std::fabs(C) > 0x1p100) continue;
                    addend = C;
                } else {  // B: (u + C) * P
                    P = outer_lit, C = inner_lit;
                    int ec = 0;
                    // P > 0: for u = -C the sum is +0 and (+0) * P keeps the sign of P, while the fused form gives +0;
                    // |C| <= 2^100: u + C must not overflow where the scaled-down result would be finite
                    if (!is_pow2(P, -200, &k) || P < 0.0 || !std::isfinite(C) || C == 0.0 || std::fabs(C) > 0x1p100) continue;
                    std::frexp(C, &ec);  // |C| in [2^(ec-1), 2^ec)
                    addend = C * P;
                    if ((ec - 1) - 24 + k < -126 || !std::isfinite(addend) || (double)(float)addend != addend ||
                        std::fabs(addend) < 0x1p-126)
                        continue;
                }
                rhs[(size_t)i] = "cb2_fmap(t" + std::to_string(u) + ", " + literal_text(P) + ", " + literal_text(addend) + ")";
                absorbed[(size_t)inner] = 1;
                rewritten[(size_t)i] = 1;
                side = 2;
                break;
            }
        }
    }
    // u * 1.0 is u for every u (a NaN stays a NaN; payloads are not preserved anywhere): the constant grad closures
    // of `add` ops (`|_| 1.0`) cost nothing in a chain-grad kernel.  Done last, so that `(u * 1.0) + C` above still
    // becomes one fma.
    for (int32_t i = 0; i < n; i++) {
        const cb_node &c = all[(size_t)i];
        if (c.op != CB_OP_MUL || absorbed[(size_t)i] || rewritten[(size_t)i]) continue;
        double one;
        const bool a_one = lit(c.a, &one) && one == 1.0, b_one = lit(c.b, &one) && one == 1.0;
        if (!a_one && !b_one) continue;
        const int32_t other = b_one ? c.a : c.b;
        if (all[(size_t)other].op == CB_OP_CONST || absorbed[(size_t)other]) continue;
        rhs[(size_t)i] = "cb2_identity(t" + std::to_string(other) + ")";
    }
    // literals that only fed an absorbed node are dead; the compiler drops them
    std::string s = "#if CB_PAIR\n__device__ __forceinline__ cb_f2 cb_fn2(cb_f2 x, cb_f2 y, bool &redo)\n{\n";
    s += "    const cb_f2 x_in = x;\n    (void)x_in;\n";
    for (int32_t i = 0; i < n; i++) {
        if (absorbed[(size_t)i]) continue;
        s += "    const cb_f2 t" + std::to_string(i) + " = " + rhs[(size_t)i] + ";\n";
    }
    s += "    return t" + std::to_string(prev_root) + ";\n}\n#endif\n";
    return s;
}

std::string expr_cuda_function(int32_t dtype, int32_t kind, const cb_node *const *progs, const int32_t *n_nodes,
                               int32_t n_progs, bool fuse_scale_add)
{
    if (kind == CB_KERNEL_CHAIN_GRAD) {
        // the scalar function (tails, unaligned buffers, the slow-path redo, every dtype but f32) adds the zero of the
        // intermediate gradient buffer after EVERY op; the f32 pair function only before the last multiply (see
        // chain_grad_tree) — tests that compare the vector path with the scalar one check that the two agree
        const std::vector<cb_node> full = chain_grad_tree(progs, n_nodes, n_progs, true);
        const cb_node *p1[1] = {full.data()};
        const int32_t c1[1] = {(int32_t)full.size()};
        std::string out = expr_cuda_function(dtype, CB_KERNEL_BINARY, p1, c1, 1, false);
        if (dtype == CB_F32) {
            const std::vector<cb_node> lean = chain_grad_tree(progs, n_nodes, n_progs, !fuse_scale_add);
            const cb_node *p2[1] = {lean.data()};
            const int32_t c2[1] = {(int32_t)lean.size()};
            const size_t cut = out.find("#if CB_PAIR");
            const size_t end = out.rfind("}  // namespace CB_NS");
            if (cut != std::string::npos && end != std::string::npos) {
                std::string pair = fuse_scale_add ? fused_pair_function(p2, c2, 1)
                                                  : expr_cuda_function(dtype, CB_KERNEL_BINARY, p2, c2, 1, false);
                if (!fuse_scale_add) {
                    const size_t a = pair.find("#if CB_PAIR"), b = pair.rfind("}  // namespace CB_NS");
                    pair = pair.substr(a, b - a);
                }
                out = out.substr(0, cut) + pair + out.substr(end);
            }
        }
        return out;
    }
    std::string s;
    s += "// generated from the recorded Combiner trees; one block per recorded op, applied in order\n";
    s += "namespace CB_NS {\n__device__ __forceinline__ T cb_fn(T x, T y)\n{\n";
    for (int32_t k = 0; k < n_progs; k++) {
        const cb_node *nd = progs[k];
        const int32_t n = n_nodes[k];
        // (the rendered source of a joined tree with shared nodes can be huge: only small programs get the comment)
        s += "    { // op " + std::to_string(k) + ": x = " + (n <= 32 ? expr_to_cl_source(dtype, nd, n, "x", "y") : std::to_string(n) + " nodes") + "\n";
        for (int32_t i = 0; i < n; i++) {
            const cb_node &c = nd[i];
            const std::string name = "t" + std::to_string(i);
            std::string rhs;
            if (c.op == CB_OP_X) rhs = "x";
            else if (c.op == CB_OP_Y) rhs = "y";
            else if (c.op == CB_OP_CONST) rhs = cuda_literal(dtype, c) + " /* " + rust_debug_literal(dtype, c) + " */";
            else if (op_is_binary(c.op))
                rhs = std::string(cuda_fn(c.op)) + "(t" + std::to_string(c.a) + ", t" + std::to_string(c.b) + ")";
            else
                rhs = std::string(cuda_fn(c.op)) + "(t" + std::to_string(c.a) + ")";
            s += "        const T " + name + " = " + rhs + ";\n";
        }
        s += "        x = t" + std::to_string(n - 1) + ";\n    }\n";
    }
    s += "    return x;\n}\n";
    if (dtype == CB_F32 && fuse_scale_add) {
        s += fused_pair_function(progs, n_nodes, n_progs);
    } else if (dtype == CB_F32) {
        // the same programs on two elements at a time (one f32x2 register pair, see skeleton.cuh)
        s += "#if CB_PAIR\n__device__ __forceinline__ cb_f2 cb_fn2(cb_f2 x, cb_f2 y, bool &redo)\n{\n";
        for (int32_t k = 0; k < n_progs; k++) {
            const cb_node *nd = progs[k];
            const int32_t n = n_nodes[k];
            s += "    { // op " + std::to_string(k) + "\n";
            for (int32_t i = 0; i < n; i++) {
                const cb_node &c = nd[i];
                s += "        const cb_f2 t" + std::to_string(i) + " = " + pair_rhs(c, "t" + std::to_string(c.a), "t" + std::to_string(c.b)) + ";\n";
            }
            s += "        x = t" + std::to_string(n - 1) + ";\n    }\n";
        }
        s += "    return x;\n}\n#endif\n";
    }
    if (is_half_dtype(dtype)) {
        // two binary16 / bfloat16 elements per 32-bit word (skeleton.cuh: cbw_*); ops with a literal operand take the
        // mixed-precision forms that need no unpack
        auto f32_bits_of = [dtype](const cb_node &c, bool negate) {
            float f = host_half_to_f32(dtype, host_f32_to_half(dtype, (float)c.fimm));
            if (negate) f = -f;
            uint32_t u;
            std::memcpy(&u, &f, 4);
            char buf[48];
            std::snprintf(buf, sizeof buf, "__uint_as_float(0x%08xu)", u);
            return std::string(buf);
        };
        auto half_bits_of = [dtype](const cb_node &c) {
            char buf[32];
            std::snprintf(buf, sizeof buf, "(T)0x%04xu", (unsigned)host_f32_to_half(dtype, (float)c.fimm));
            return std::string(buf);
        };
        s += "#if CB_PAIR\n__device__ __forceinline__ cb_w cb_fnw(cb_w x, cb_w y, bool &redo)\n{\n";
        for (int32_t k = 0; k < n_progs; k++) {
            const cb_node *nd = progs[k];
            const int32_t n = n_nodes[k];
            s += "    { // op " + std::to_string(k) + "\n";
            for (int32_t i = 0; i < n; i++) {
                const cb_node &c = nd[i];
                const std::string ta = "t" + std::to_string(c.a), tb = "t" + std::to_string(c.b);
                const bool a_const = c.a >= 0 && nd[c.a].op == CB_OP_CONST, b_const = c.b >= 0 && nd[c.b].op == CB_OP_CONST;
                std::string rhs;
                if (c.op == CB_OP_X) rhs = "x";
                else if (c.op == CB_OP_Y) rhs = "y";
                else if (c.op == CB_OP_CONST) {
                    char buf[32];
                    std::snprintf(buf, sizeof buf, "cbw_lit(0x%04xu)", (unsigned)host_f32_to_half(dtype, (float)c.fimm));
                    rhs = buf;
                } else if (c.op == CB_OP_ADD && b_const) rhs = "cbw_add_c(" + ta + ", " + f32_bits_of(nd[c.b], false) + ")";
                else if (c.op == CB_OP_ADD && a_const) rhs = "cbw_add_c(" + tb + ", " + f32_bits_of(nd[c.a], false) + ")";
                else if (c.op == CB_OP_SUB && b_const) rhs = "cbw_add_c(" + ta + ", " + f32_bits_of(nd[c.b], true) + ")";
                else if (c.op == CB_OP_MUL && b_const) rhs = "cbw_mul_c(" + ta + ", " + half_bits_of(nd[c.b]) + ")";
                else if (c.op == CB_OP_MUL && a_const) rhs = "cbw_mul_c(" + tb + ", " + half_bits_of(nd[c.a]) + ")";
                else if (c.op == CB_OP_SIN || c.op == CB_OP_COS || c.op == CB_OP_TAN)
                    rhs = "cbw_" + std::string(cuda_fn(c.op) + 3) + "(" + ta + ", redo)";
                else if (op_is_binary(c.op)) rhs = "cbw_" + std::string(cuda_fn(c.op) + 3) + "(" + ta + ", " + tb + ")";
                else rhs = "cbw_" + std::string(cuda_fn(c.op) + 3) + "(" + ta + ")";
                s += "        const cb_w t" + std::to_string(i) + " = " + rhs + ";\n";
            }
            s += "        x = t" + std::to_string(n - 1) + ";\n    }\n";
        }
        s += "    return x;\n}\n#endif\n";
    }
    s += "}  // namespace CB_NS\n";
    return s;
}

std::string chain_bytes(int32_t dtype, int32_t kind, const cb_node *const *progs, const int32_t *n_nodes,
                        int32_t n_progs)
{
    std::string out;
    auto put = [&out](uint64_t v) {
        for (int i = 0; i < 8; i++) out.push_back((char)((v >> (8 * i)) & 0xffu));
    };
    put((uint64_t)dtype);
    put((uint64_t)kind);
    put((uint64_t)n_progs);
    for (int32_t k = 0; k < n_progs; k++) {
        put((uint64_t)n_nodes[k]);
        for (int32_t i = 0; i < n_nodes[k]; i++) {
            const cb_node &c = progs[k][i];
            put((uint64_t)(uint32_t)c.op);
            put((uint64_t)(uint32_t)c.a);
            put((uint64_t)(uint32_t)c.b);
            if (c.op == CB_OP_CONST) {
                uint64_t bits;
                if (is_float_dtype(dtype)) std::memcpy(&bits, &c.fimm, 8);
                else bits = (uint64_t)c.iimm;
                put(bits);
            }
        }
    }
    return out;
}

// FNV-1a over chain_bytes
uint64_t chain_hash(int32_t dtype, int32_t kind, const cb_node *const *progs, const int32_t *n_nodes,
                    int32_t n_progs)
{
    uint64_t h = 1469598103934665603ull;
    for (unsigned char ch : chain_bytes(dtype, kind, progs, n_nodes, n_progs)) {
        h ^= ch;
        h *= 1099511628211ull;
    }
    return h;
}

}  // namespace cb
