/*
 * oracle.c — CPU restatement of the custos reference CPU device (hot path only).
 * TEST INFRASTRUCTURE ONLY — see oracle.h.  Build: `make -C oracle` (gcc -O2
 * -ffp-contract=off, no fast-math, so every + and * rounds once like rustc's output).
 *
 * The transcendental functions call glibc's libm — the same symbols Rust's
 * f32::exp / sin / ... lower to on x86_64-unknown-linux-gnu (src/number.rs:273-346).
 */
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

size_t orc_dtype_size(int dtype)
{
    switch (dtype) {
    case ORC_F32: return 4;
    case ORC_F64: return 8;
    case ORC_F16: return 2;
    case ORC_I32: return 4;
    case ORC_I64: return 8;
    case ORC_U32: return 4;
    case ORC_U8: return 1;
    case ORC_BF16: return 2;
    case ORC_I8: return 1;
    case ORC_I16: return 2;
    case ORC_U16: return 2;
    case ORC_U64: return 8;
    case ORC_BOOL: return 1;
    default: return 0;
    }
}

/* ------------------------------------------------------------------ binary16
 * `half` crate 2.x (Cargo.toml:36, version unpinned: no Cargo.lock in the tree):
 * f16 <-> f32 in software, round to nearest, ties to even; NaNs stay NaNs (quiet
 * bit forced), overflow goes to infinity, results below half the smallest
 * subnormal flush to signed zero. */
uint16_t orc_f32_to_f16(float v)
{
    uint32_t bits;
    memcpy(&bits, &v, 4);
    const uint16_t sign = (uint16_t)((bits >> 16) & 0x8000u);
    const uint32_t expo = (bits >> 23) & 0xffu;
    const uint32_t frac = bits & 0x7fffffu;

    if (expo == 0xffu) { /* inf / nan */
        if (frac == 0) return (uint16_t)(sign | 0x7c00u);
        return (uint16_t)(sign | 0x7c00u | 0x0200u | (frac >> 13));
    }
    const int e = (int)expo - 127 + 15; /* re-biased exponent */
    if (e >= 31) return (uint16_t)(sign | 0x7c00u);
    if (e <= 0) {
        /* subnormal half (or zero): shift the 24-bit significand right */
        const int shift = 14 - e; /* >= 14 */
        if (shift > 24) return sign;
        const uint32_t sig = frac | 0x800000u;
        uint32_t q = sig >> shift;
        const uint32_t rem = sig & ((1u << shift) - 1u);
        const uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (q & 1u))) q += 1;
        return (uint16_t)(sign | q); /* a carry into bit 10 is the smallest normal: correct */
    }
    uint32_t q = ((uint32_t)e << 10) | (frac >> 13);
    const uint32_t rem = frac & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) q += 1; /* may carry into the exponent / inf */
    return (uint16_t)(sign | q);
}

float orc_f16_to_f32(uint16_t h)
{
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    const uint32_t expo = (h >> 10) & 0x1fu;
    uint32_t frac = h & 0x3ffu;
    uint32_t bits;
    if (expo == 0x1fu) {
        bits = sign | 0x7f800000u | (frac << 13);
        if (frac) bits |= 0x00400000u;
    } else if (expo == 0) {
        if (frac == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do {
                frac <<= 1;
                e++;
            } while (!(frac & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((frac & 0x3ffu) << 13);
        }
    } else {
        bits = sign | ((expo + 127 - 15) << 23) | (frac << 13);
    }
    float out;
    memcpy(&out, &bits, 4);
    return out;
}

/* ------------------------------------------------------------------ bfloat16
 * half::bf16::from_f32 / to_f32: the upper 16 bits of the f32, round to nearest even on the
 * dropped half; a NaN keeps its top payload bits and gets the quiet bit (0x0040). */
uint16_t orc_f32_to_bf16(float v)
{
    uint32_t bits;
    memcpy(&bits, &v, 4);
    if ((bits & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((bits >> 16) | 0x0040u);
    const uint32_t lower = bits & 0xffffu;
    uint32_t q = bits >> 16;
    if (lower > 0x8000u || (lower == 0x8000u && (q & 1u))) q += 1; /* may carry into the exponent / inf */
    return (uint16_t)q;
}

float orc_bf16_to_f32(uint16_t h)
{
    uint32_t bits = (uint32_t)h << 16;
    if ((h & 0x7fffu) > 0x7f80u) bits |= 0x00400000u;
    float out;
    memcpy(&out, &bits, 4);
    return out;
}

/* ------------------------------------------------------------- program check */
static int check_prog(const orc_node *nd, int n)
{
    if (!nd || n <= 0 || n > ORC_MAX_NODES) return ORC_ERR_EXPR;
    for (int i = 0; i < n; i++) {
        const int op = nd[i].op;
        if (op < 0 || op >= ORC_OP_COUNT) return ORC_ERR_EXPR;
        const int binary = (op >= ORC_OP_ADD && op <= ORC_OP_MAX) || (op >= ORC_OP_GEQ);
        if (op >= ORC_OP_ADD) {
            if (nd[i].a < 0 || nd[i].a >= i) return ORC_ERR_EXPR;
            if (binary && (nd[i].b < 0 || nd[i].b >= i)) return ORC_ERR_EXPR;
        }
        /* unused operand slots must be negative: the evaluators index by a / b whenever it is >= 0 */
        if (!binary && nd[i].b >= 0) return ORC_ERR_EXPR;
        if (op < ORC_OP_ADD && nd[i].a >= 0) return ORC_ERR_EXPR;
    }
    return ORC_OK;
}

/* --------------------------------------------------- Eval for f32 / f64
 * src/two_way_ops/ops.rs:53-58,98-103,143-148,188-193,233-238,271-276,309-314
 * src/two_way_ops/ops/unary.rs (Identity, Exp, Sin, Cos, Tan, Tanh, Neg, Ln, Abs)
 * src/two_way_ops/ops/cmps.rs:42-47,87-92,132-137 (Eq evaluates `le`, line 135)
 * min/max: Number::min/max by comparison, src/number.rs:202-209. */
#define DEFINE_FLOAT_EVAL(NAME, T, EXP, LOG, SIN, COS, TAN, TANH, POW, FABS)                       \
    static T NAME(const orc_node *nd, int n, T x, T y)                                             \
    {                                                                                              \
        T v[ORC_MAX_NODES];                                                                        \
        for (int i = 0; i < n; i++) {                                                              \
            const orc_node *c = &nd[i];                                                            \
            const T a = c->a >= 0 ? v[c->a] : (T)0;                                                \
            const T b = c->b >= 0 ? v[c->b] : (T)0;                                                \
            T r;                                                                                   \
            switch (c->op) {                                                                       \
            case ORC_OP_X: r = x; break;                                                           \
            case ORC_OP_Y: r = y; break;                                                           \
            case ORC_OP_CONST: r = (T)c->fimm; break;                                              \
            case ORC_OP_ADD: r = a + b; break;                                                     \
            case ORC_OP_MUL: r = a * b; break;                                                     \
            case ORC_OP_SUB: r = a - b; break;                                                     \
            case ORC_OP_DIV: r = a / b; break;                                                     \
            case ORC_OP_POW: r = POW(a, b); break;                                                 \
            case ORC_OP_MIN: r = (a < b) ? a : b; break;                                           \
            case ORC_OP_MAX: r = (a > b) ? a : b; break;                                           \
            case ORC_OP_SIN: r = SIN(a); break;                                                    \
            case ORC_OP_COS: r = COS(a); break;                                                    \
            case ORC_OP_TAN: r = TAN(a); break;                                                    \
            case ORC_OP_TANH: r = TANH(a); break;                                                  \
            case ORC_OP_EXP: r = EXP(a); break;                                                    \
            case ORC_OP_LN: r = LOG(a); break;                                                     \
            case ORC_OP_ABS: r = FABS(a); break;                                                   \
            case ORC_OP_NEG: r = -a; break;                                                        \
            case ORC_OP_IDENTITY: r = a; break;                                                    \
            case ORC_OP_GEQ: r = (T)(a >= b ? 1 : 0); break;                                       \
            case ORC_OP_LEQ: r = (T)(a <= b ? 1 : 0); break;                                       \
            default: /* ORC_OP_EQ */ r = (T)(a <= b ? 1 : 0); break;                               \
            }                                                                                      \
            v[i] = r;                                                                              \
        }                                                                                          \
        return v[n - 1];                                                                           \
    }

DEFINE_FLOAT_EVAL(eval_f32, float, expf, logf, sinf, cosf, tanf, tanhf, powf, fabsf)
DEFINE_FLOAT_EVAL(eval_f64, double, exp, log, sin, cos, tan, tanh, pow, fabs)

/* --------------------------------------------------- Eval for f16 / bf16
 * Float for half::f16 (src/number.rs:543-608) and half::bf16 (:611-676): every function goes
 * to f32 and back (`Self::from_f32(self.to_f32().exp())`), `tan` calls cos (lines 575-577 and
 * 643-645); + - * / are `half`'s operators (f32 arithmetic, rounded back after each op).
 * Number::max for f16 (number.rs:507-510) AND for bf16 (number.rs:536-539) forwards to half's
 * inherent `max`, which keeps `self` unless `other > self` (half 2.x: `if other > self &&
 * !other.is_nan() { other } else { self }` — so a NaN `self` stays, a NaN `other` is ignored,
 * and +0 / -0 ties keep `self`); both mins use the trait default
 * `if self < rhs { self } else { rhs }` (:207-209).  (Round 1 gave bf16 the trait default
 * `if self > rhs { self } else { rhs }` for max: a misreading of number.rs:515-540.) */
static uint16_t eval_half(int bf, const orc_node *nd, int n, uint16_t x, uint16_t y)
{
#define TO_F32(h) (bf ? orc_bf16_to_f32(h) : orc_f16_to_f32(h))
#define FROM_F32(f) (bf ? orc_f32_to_bf16(f) : orc_f32_to_f16(f))
    uint16_t v[ORC_MAX_NODES];
    for (int i = 0; i < n; i++) {
        const orc_node *c = &nd[i];
        const uint16_t ha = c->a >= 0 ? v[c->a] : 0;
        const uint16_t hb = c->b >= 0 ? v[c->b] : 0;
        const float a = TO_F32(ha), b = TO_F32(hb);
        uint16_t r;
        switch (c->op) {
        case ORC_OP_X: r = x; break;
        case ORC_OP_Y: r = y; break;
        case ORC_OP_CONST: r = FROM_F32((float)c->fimm); break;
        case ORC_OP_ADD: r = FROM_F32(a + b); break;
        case ORC_OP_MUL: r = FROM_F32(a * b); break;
        case ORC_OP_SUB: r = FROM_F32(a - b); break;
        case ORC_OP_DIV: r = FROM_F32(a / b); break;
        case ORC_OP_POW: r = FROM_F32(powf(a, b)); break;
        case ORC_OP_MIN: r = (a < b) ? ha : hb; break;
        case ORC_OP_MAX: r = (b > a) ? hb : ha; break;
        case ORC_OP_SIN: r = FROM_F32(sinf(a)); break;
        case ORC_OP_COS: r = FROM_F32(cosf(a)); break;
        case ORC_OP_TAN: r = FROM_F32(cosf(a)); break; /* sic */
        case ORC_OP_TANH: r = FROM_F32(tanhf(a)); break;
        case ORC_OP_EXP: r = FROM_F32(expf(a)); break;
        case ORC_OP_LN: r = FROM_F32(logf(a)); break;
        case ORC_OP_ABS: r = FROM_F32(fabsf(a)); break;
        case ORC_OP_NEG: r = (uint16_t)(ha ^ 0x8000u); break; /* half: Neg flips the sign bit */
        case ORC_OP_IDENTITY: r = ha; break;
        case ORC_OP_GEQ: r = FROM_F32(a >= b ? 1.0f : 0.0f); break;
        case ORC_OP_LEQ: r = FROM_F32(a <= b ? 1.0f : 0.0f); break;
        default: r = FROM_F32(a <= b ? 1.0f : 0.0f); break;
        }
        v[i] = r;
    }
    return v[n - 1];
}
static uint16_t eval_f16(const orc_node *nd, int n, uint16_t x, uint16_t y) { return eval_half(0, nd, n, x, y); }
static uint16_t eval_bf16(const orc_node *nd, int n, uint16_t x, uint16_t y) { return eval_half(1, nd, n, x, y); }

/* --------------------------------------------------- Eval for integers
 * Only Add/Mul/Sub/Div/Neg(signed)/GEq/LEq/Eq have integer impls; the Float-bounded
 * ops do not compile for integers in the reference.  Arithmetic wraps (release build). */
static int int_op_supported(int op, int is_signed)
{
    switch (op) {
    case ORC_OP_X: case ORC_OP_Y: case ORC_OP_CONST: case ORC_OP_ADD: case ORC_OP_MUL:
    case ORC_OP_SUB: case ORC_OP_DIV: case ORC_OP_GEQ: case ORC_OP_LEQ: case ORC_OP_EQ:
        return 1;
    case ORC_OP_NEG: return is_signed;
    default: return 0;
    }
}

#define DEFINE_INT_EVAL(NAME, T, UT, SIGNED)                                                               \
    static T NAME(const orc_node *nd, int n, T x, T y)                                             \
    {                                                                                              \
        T v[ORC_MAX_NODES];                                                                        \
        for (int i = 0; i < n; i++) {                                                              \
            const orc_node *c = &nd[i];                                                            \
            const T a = c->a >= 0 ? v[c->a] : (T)0;                                                \
            const T b = c->b >= 0 ? v[c->b] : (T)0;                                                \
            T r;                                                                                   \
            switch (c->op) {                                                                       \
            case ORC_OP_X: r = x; break;                                                           \
            case ORC_OP_Y: r = y; break;                                                           \
            case ORC_OP_CONST: r = (T)c->iimm; break;                                              \
            case ORC_OP_ADD: r = (T)((UT)a + (UT)b); break;                                        \
            case ORC_OP_MUL: r = (T)((UT)a * (UT)b); break;                                        \
            case ORC_OP_SUB: r = (T)((UT)a - (UT)b); break;                                        \
            case ORC_OP_DIV: /* x / 0 and MIN / -1 panic in Rust; defined here as 0 and MIN */     \
                r = (b == 0) ? (T)0 : ((SIGNED && b == (T)-1) ? (T)((UT)0 - (UT)a) : (T)(a / b)); \
                break;                                                                             \
            case ORC_OP_NEG: r = (T)((UT)0 - (UT)a); break;                                        \
            case ORC_OP_GEQ: r = (T)(a >= b); break;                                               \
            case ORC_OP_LEQ: r = (T)(a <= b); break;                                               \
            default: r = (T)(a <= b); break;                                                       \
            }                                                                                      \
            v[i] = r;                                                                              \
        }                                                                                          \
        return v[n - 1];                                                                           \
    }

DEFINE_INT_EVAL(eval_i32, int32_t, uint32_t, 1)
DEFINE_INT_EVAL(eval_i64, int64_t, uint64_t, 1)
DEFINE_INT_EVAL(eval_u32, uint32_t, uint32_t, 0)
DEFINE_INT_EVAL(eval_u8, uint8_t, uint8_t, 0)
DEFINE_INT_EVAL(eval_i8, int8_t, uint8_t, 1)
DEFINE_INT_EVAL(eval_i16, int16_t, uint16_t, 1)
DEFINE_INT_EVAL(eval_u16, uint16_t, uint16_t, 0)
DEFINE_INT_EVAL(eval_u64, uint64_t, uint64_t, 0)

static int check_dtype_prog(int dtype, const orc_node *nd, int n)
{
    int rc = check_prog(nd, n);
    if (rc) return rc;
    if (dtype == ORC_F32 || dtype == ORC_F64 || dtype == ORC_F16 || dtype == ORC_BF16) return ORC_OK;
    if (dtype < 0 || dtype >= ORC_DTYPE_COUNT) return ORC_ERR_ARG;
    if (dtype == ORC_BOOL) return ORC_ERR_UNSUPPORTED; /* bool: CDatatype but not Number */
    const int is_signed = (dtype == ORC_I32 || dtype == ORC_I64 || dtype == ORC_I8 || dtype == ORC_I16);
    for (int i = 0; i < n; i++)
        if (!int_op_supported(nd[i].op, is_signed)) return ORC_ERR_UNSUPPORTED;
    return ORC_OK;
}

/* one element: out = f(x, y) */
static inline void eval_one(int dtype, const orc_node *nd, int n, const void *x, const void *y, void *out)
{
    switch (dtype) {
    case ORC_F32: *(float *)out = eval_f32(nd, n, *(const float *)x, y ? *(const float *)y : 0.f); break;
    case ORC_F64: *(double *)out = eval_f64(nd, n, *(const double *)x, y ? *(const double *)y : 0.0); break;
    case ORC_F16: *(uint16_t *)out = eval_f16(nd, n, *(const uint16_t *)x, y ? *(const uint16_t *)y : 0); break;
    case ORC_I32: *(int32_t *)out = eval_i32(nd, n, *(const int32_t *)x, y ? *(const int32_t *)y : 0); break;
    case ORC_I64: *(int64_t *)out = eval_i64(nd, n, *(const int64_t *)x, y ? *(const int64_t *)y : 0); break;
    case ORC_U32: *(uint32_t *)out = eval_u32(nd, n, *(const uint32_t *)x, y ? *(const uint32_t *)y : 0); break;
    case ORC_BF16: *(uint16_t *)out = eval_bf16(nd, n, *(const uint16_t *)x, y ? *(const uint16_t *)y : 0); break;
    case ORC_I8: *(int8_t *)out = eval_i8(nd, n, *(const int8_t *)x, y ? *(const int8_t *)y : 0); break;
    case ORC_I16: *(int16_t *)out = eval_i16(nd, n, *(const int16_t *)x, y ? *(const int16_t *)y : 0); break;
    case ORC_U16: *(uint16_t *)out = eval_u16(nd, n, *(const uint16_t *)x, y ? *(const uint16_t *)y : 0); break;
    case ORC_U64: *(uint64_t *)out = eval_u64(nd, n, *(const uint64_t *)x, y ? *(const uint64_t *)y : 0); break;
    default: *(uint8_t *)out = eval_u8(nd, n, *(const uint8_t *)x, y ? *(const uint8_t *)y : 0); break;
    }
}

int orc_eval(int dtype, const orc_node *nodes, int n, const void *x, const void *y, void *out)
{
    int rc = check_dtype_prog(dtype, nodes, n);
    if (rc) return rc;
    eval_one(dtype, nodes, n, x, y, out);
    return ORC_OK;
}

/* apply_fn_slice: `for (x, out) in x.iter().zip(out.iter_mut()) { *out = f((*x).to_val()).eval(); }` */
int orc_apply_fn(int dtype, const orc_node *nodes, int n, const void *x, void *out, size_t len)
{
    int rc = check_dtype_prog(dtype, nodes, n);
    if (rc) return rc;
    const size_t sz = orc_dtype_size(dtype);
    const char *px = (const char *)x;
    char *po = (char *)out;
    for (size_t i = 0; i < len; i++) eval_one(dtype, nodes, n, px + i * sz, NULL, po + i * sz);
    return ORC_OK;
}

/* cpu_device.rs:217-229: per element, current_val threads through every op in order. */
static void chain_range(int dtype, const orc_node *const *progs, const int *n_nodes, int n_progs,
                        const char *px, char *po, size_t begin, size_t end)
{
    const size_t sz = orc_dtype_size(dtype);
    for (size_t i = begin; i < end; i++) {
        uint64_t cur = 0, nxt = 0;
        memcpy(&cur, px + i * sz, sz);
        for (int k = 0; k < n_progs; k++) {
            eval_one(dtype, progs[k], n_nodes[k], &cur, NULL, &nxt);
            cur = nxt;
        }
        memcpy(po + i * sz, &cur, sz);
    }
}

int orc_apply_chain(int dtype, const orc_node *const *progs, const int *n_nodes, int n_progs,
                    const void *x, void *out, size_t len)
{
    for (int k = 0; k < n_progs; k++) {
        int rc = check_dtype_prog(dtype, progs[k], n_nodes[k]);
        if (rc) return rc;
    }
    chain_range(dtype, progs, n_nodes, n_progs, (const char *)x, (char *)out, 0, len);
    return ORC_OK;
}

struct chain_job {
    int dtype, n_progs;
    const orc_node *const *progs;
    const int *n_nodes;
    const char *px;
    char *po;
    size_t begin, end;
};

static void *chain_worker(void *arg)
{
    struct chain_job *j = (struct chain_job *)arg;
    chain_range(j->dtype, j->progs, j->n_nodes, j->n_progs, j->px, j->po, j->begin, j->end);
    return NULL;
}

int orc_apply_chain_mt(int dtype, const orc_node *const *progs, const int *n_nodes, int n_progs,
                       const void *x, void *out, size_t len, int threads)
{
    for (int k = 0; k < n_progs; k++) {
        int rc = check_dtype_prog(dtype, progs[k], n_nodes[k]);
        if (rc) return rc;
    }
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    struct chain_job jobs[256];
    const size_t per = (len + (size_t)threads - 1) / (size_t)threads;
    int started = 0;
    for (int t = 0; t < threads; t++) {
        size_t b = per * (size_t)t, e = b + per;
        if (b >= len) break;
        if (e > len) e = len;
        jobs[t] = (struct chain_job){dtype, n_progs, progs, n_nodes, (const char *)x, (char *)out, b, e};
        if (pthread_create(&tid[t], NULL, chain_worker, &jobs[t]) != 0) {
            chain_worker(&jobs[t]);
            tid[t] = 0;
        }
        started = t + 1;
    }
    for (int t = 0; t < started; t++)
        if (tid[t]) pthread_join(tid[t], NULL);
    return ORC_OK;
}

/* ------------------------------------------------- the two CPU baselines SURVEY.md §8(d) names, for f32
 * (i)  the FAITHFUL fused path: CPU::unary_fuse_op (src/devices/cpu/cpu_device.rs:217-229) walks the elements and,
 *      PER ELEMENT AND PER OP, calls the op hint, which builds the op's expression and boxes it —
 *      `let op: Box<dyn TwoWay<T>> = Box::new(op(x));` (src/op_hint.rs:30-33) — then `.eval()`s it through the vtable and
 *      drops the box.  Restated: one heap allocation, one indirect call and one free per op per element; the
 *      arithmetic inside `eval` is the monomorphised expression (here: the evaluator specialised by a switch on the
 *      program's root — programs are walked, not re-validated).
 * (ii) the MONOMORPHISED unfused path: apply_fn_slice (src/devices/cpu_stack_ops.rs:7-15) once per recorded op, each a
 *      tight loop over the whole buffer (n_progs passes over memory).
 * Both produce the bits of orc_apply_chain (tests/test_oracle_kat.py). */
struct boxed_op {
    float (*eval)(const struct boxed_op *self); /* the vtable slot */
    const orc_node *nd;
    int n;
    float val; /* Resolve { val, marker } */
};
static float boxed_eval(const struct boxed_op *self) { return eval_f32(self->nd, self->n, self->val, 0.f); }

int orc_apply_chain_boxed_f32(const orc_node *const *progs, const int *n_nodes, int n_progs, const float *x, float *out,
                              size_t len)
{
    for (int k = 0; k < n_progs; k++) {
        int rc = check_dtype_prog(ORC_F32, progs[k], n_nodes[k]);
        if (rc) return rc;
    }
    for (size_t i = 0; i < len; i++) {
        float cur = x[i];
        for (int k = 0; k < n_progs; k++) {
            struct boxed_op *volatile op = (struct boxed_op *)malloc(sizeof(struct boxed_op)); /* Box::new(op(resolve)) */
            if (!op) return ORC_ERR_ARG;
            op->eval = boxed_eval;
            op->nd = progs[k];
            op->n = n_nodes[k];
            op->val = cur;
            cur = op->eval(op); /* dyn TwoWay::eval */
            free((void *)op);   /* the box is dropped */
        }
        out[i] = cur;
    }
    return ORC_OK;
}

int orc_apply_chain_unfused_f32(const orc_node *const *progs, const int *n_nodes, int n_progs, const float *x, float *out,
                                size_t len)
{
    for (int k = 0; k < n_progs; k++) {
        int rc = check_dtype_prog(ORC_F32, progs[k], n_nodes[k]);
        if (rc) return rc;
    }
    float *tmp = (float *)malloc(len * sizeof(float) + 4);
    if (!tmp) return ORC_ERR_ARG;
    const float *src = x;
    for (int k = 0; k < n_progs; k++) {
        float *dst = (k == n_progs - 1) ? out : ((k & 1) ? out : tmp); /* a new buffer per op in the reference */
        if (dst == src) dst = (dst == out) ? tmp : out;
        for (size_t i = 0; i < len; i++) dst[i] = eval_f32(progs[k], n_nodes[k], src[i], 0.f);
        src = dst;
    }
    if (src != out) memcpy(out, src, len * sizeof(float));
    free(tmp);
    return ORC_OK;
}

/* cpu_stack_ops.rs:18-30: `*lhs_grad += *out * lhs_grad_fn((*lhs).to_val()).eval();`
 * — a multiply, then an add, each rounded (no FMA). */
int orc_add_unary_grad(int dtype, const orc_node *nodes, int n, const void *lhs, const void *out_grad,
                       void *lhs_grad, size_t len)
{
    int rc = check_dtype_prog(dtype, nodes, n);
    if (rc) return rc;
    for (size_t i = 0; i < len; i++) {
        switch (dtype) {
        case ORC_F32: {
            float g = eval_f32(nodes, n, ((const float *)lhs)[i], 0.f);
            float m = ((const float *)out_grad)[i] * g;
            ((float *)lhs_grad)[i] = ((float *)lhs_grad)[i] + m;
        } break;
        case ORC_F64: {
            double g = eval_f64(nodes, n, ((const double *)lhs)[i], 0.0);
            double m = ((const double *)out_grad)[i] * g;
            ((double *)lhs_grad)[i] = ((double *)lhs_grad)[i] + m;
        } break;
        case ORC_F16: {
            uint16_t g = eval_f16(nodes, n, ((const uint16_t *)lhs)[i], 0);
            uint16_t m = orc_f32_to_f16(orc_f16_to_f32(((const uint16_t *)out_grad)[i]) * orc_f16_to_f32(g));
            ((uint16_t *)lhs_grad)[i] =
                orc_f32_to_f16(orc_f16_to_f32(((uint16_t *)lhs_grad)[i]) + orc_f16_to_f32(m));
        } break;
        case ORC_I32: {
            uint32_t g = (uint32_t)eval_i32(nodes, n, ((const int32_t *)lhs)[i], 0);
            ((int32_t *)lhs_grad)[i] =
                (int32_t)((uint32_t)((int32_t *)lhs_grad)[i] + (uint32_t)((const int32_t *)out_grad)[i] * g);
        } break;
        case ORC_I64: {
            uint64_t g = (uint64_t)eval_i64(nodes, n, ((const int64_t *)lhs)[i], 0);
            ((int64_t *)lhs_grad)[i] =
                (int64_t)((uint64_t)((int64_t *)lhs_grad)[i] + (uint64_t)((const int64_t *)out_grad)[i] * g);
        } break;
        case ORC_U32: {
            uint32_t g = eval_u32(nodes, n, ((const uint32_t *)lhs)[i], 0);
            ((uint32_t *)lhs_grad)[i] = ((uint32_t *)lhs_grad)[i] + ((const uint32_t *)out_grad)[i] * g;
        } break;
        case ORC_BF16: {
            uint16_t g = eval_bf16(nodes, n, ((const uint16_t *)lhs)[i], 0);
            uint16_t m = orc_f32_to_bf16(orc_bf16_to_f32(((const uint16_t *)out_grad)[i]) * orc_bf16_to_f32(g));
            ((uint16_t *)lhs_grad)[i] =
                orc_f32_to_bf16(orc_bf16_to_f32(((uint16_t *)lhs_grad)[i]) + orc_bf16_to_f32(m));
        } break;
        case ORC_I8: {
            uint8_t g = (uint8_t)eval_i8(nodes, n, ((const int8_t *)lhs)[i], 0);
            ((uint8_t *)lhs_grad)[i] = (uint8_t)(((uint8_t *)lhs_grad)[i] + (uint8_t)(((const uint8_t *)out_grad)[i] * g));
        } break;
        case ORC_I16: {
            uint16_t g = (uint16_t)eval_i16(nodes, n, ((const int16_t *)lhs)[i], 0);
            ((uint16_t *)lhs_grad)[i] = (uint16_t)(((uint16_t *)lhs_grad)[i] + (uint16_t)((uint32_t)((const uint16_t *)out_grad)[i] * g));
        } break;
        case ORC_U16: {
            uint16_t g = eval_u16(nodes, n, ((const uint16_t *)lhs)[i], 0);
            ((uint16_t *)lhs_grad)[i] = (uint16_t)(((uint16_t *)lhs_grad)[i] + (uint16_t)((uint32_t)((const uint16_t *)out_grad)[i] * g));
        } break;
        case ORC_U64: {
            uint64_t g = eval_u64(nodes, n, ((const uint64_t *)lhs)[i], 0);
            ((uint64_t *)lhs_grad)[i] = ((uint64_t *)lhs_grad)[i] + ((const uint64_t *)out_grad)[i] * g;
        } break;
        default: {
            uint8_t g = eval_u8(nodes, n, ((const uint8_t *)lhs)[i], 0);
            ((uint8_t *)lhs_grad)[i] = (uint8_t)(((uint8_t *)lhs_grad)[i] + (uint8_t)(((const uint8_t *)out_grad)[i] * g));
        } break;
        }
    }
    return ORC_OK;
}

int orc_apply2(int dtype, const orc_node *nodes, int n, const void *lhs, const void *rhs, void *out,
               size_t len)
{
    int rc = check_dtype_prog(dtype, nodes, n);
    if (rc) return rc;
    const size_t sz = orc_dtype_size(dtype);
    for (size_t i = 0; i < len; i++)
        eval_one(dtype, nodes, n, (const char *)lhs + i * sz, (const char *)rhs + i * sz, (char *)out + i * sz);
    return ORC_OK;
}

/* tests/demo_impl/cpu.rs:12-43 / src/lib.rs:293-301: out[i] = lhs[i] op rhs[i] */
int orc_binary(int dtype, int op, const void *lhs, const void *rhs, void *out, size_t len)
{
    if (op < 0 || op > 3) return ORC_ERR_ARG;
    orc_node prog[3] = {{ORC_OP_X, -1, -1, 0, 0.0, 0}, {ORC_OP_Y, -1, -1, 0, 0.0, 0}, {ORC_OP_ADD + op, 0, 1, 0, 0.0, 0}};
    return orc_apply2(dtype, prog, 3, lhs, rhs, out, len);
}

/* clear_slice: `*value = T::default()` */
int orc_clear(int dtype, void *buf, size_t len)
{
    const size_t sz = orc_dtype_size(dtype);
    if (!sz) return ORC_ERR_ARG;
    memset(buf, 0, sz * len); /* T::default() is all-zero bits for every supported type */
    return ORC_OK;
}

/* --------------------------------------------------------------- reductions */
int orc_sum_seq(int dtype, const void *in, size_t len, void *out)
{
    switch (dtype) {
    case ORC_F32: { float s = 0.f; for (size_t i = 0; i < len; i++) s = s + ((const float *)in)[i]; *(float *)out = s; } break;
    case ORC_F64: { double s = 0.0; for (size_t i = 0; i < len; i++) s = s + ((const double *)in)[i]; *(double *)out = s; } break;
    case ORC_F16: { float s = 0.f; for (size_t i = 0; i < len; i++) s = s + orc_f16_to_f32(((const uint16_t *)in)[i]); *(float *)out = s; } break;
    case ORC_I32: { int64_t s = 0; for (size_t i = 0; i < len; i++) s += ((const int32_t *)in)[i]; *(int64_t *)out = s; } break;
    case ORC_I64: { uint64_t s = 0; for (size_t i = 0; i < len; i++) s += (uint64_t)((const int64_t *)in)[i]; *(int64_t *)out = (int64_t)s; } break;
    case ORC_U32: { int64_t s = 0; for (size_t i = 0; i < len; i++) s += ((const uint32_t *)in)[i]; *(int64_t *)out = s; } break;
    case ORC_U8: { int64_t s = 0; for (size_t i = 0; i < len; i++) s += ((const uint8_t *)in)[i]; *(int64_t *)out = s; } break;
    case ORC_BF16: { float s = 0.f; for (size_t i = 0; i < len; i++) s = s + orc_bf16_to_f32(((const uint16_t *)in)[i]); *(float *)out = s; } break;
    case ORC_I8: { int64_t s = 0; for (size_t i = 0; i < len; i++) s += ((const int8_t *)in)[i]; *(int64_t *)out = s; } break;
    case ORC_I16: { int64_t s = 0; for (size_t i = 0; i < len; i++) s += ((const int16_t *)in)[i]; *(int64_t *)out = s; } break;
    case ORC_U16: { int64_t s = 0; for (size_t i = 0; i < len; i++) s += ((const uint16_t *)in)[i]; *(int64_t *)out = s; } break;
    case ORC_U64: { uint64_t s = 0; for (size_t i = 0; i < len; i++) s += ((const uint64_t *)in)[i]; *(int64_t *)out = (int64_t)s; } break;
    default: return ORC_ERR_ARG;
    }
    return ORC_OK;
}

double orc_sum_f64(int dtype, const void *in, size_t len)
{
    double s = 0.0;
    switch (dtype) {
    case ORC_F32: for (size_t i = 0; i < len; i++) s += (double)((const float *)in)[i]; break;
    case ORC_F64: for (size_t i = 0; i < len; i++) s += ((const double *)in)[i]; break;
    case ORC_F16: for (size_t i = 0; i < len; i++) s += (double)orc_f16_to_f32(((const uint16_t *)in)[i]); break;
    case ORC_I32: for (size_t i = 0; i < len; i++) s += (double)((const int32_t *)in)[i]; break;
    case ORC_I64: for (size_t i = 0; i < len; i++) s += (double)((const int64_t *)in)[i]; break;
    case ORC_U32: for (size_t i = 0; i < len; i++) s += (double)((const uint32_t *)in)[i]; break;
    case ORC_U8: for (size_t i = 0; i < len; i++) s += (double)((const uint8_t *)in)[i]; break;
    case ORC_BF16: for (size_t i = 0; i < len; i++) s += (double)orc_bf16_to_f32(((const uint16_t *)in)[i]); break;
    case ORC_I8: for (size_t i = 0; i < len; i++) s += (double)((const int8_t *)in)[i]; break;
    case ORC_I16: for (size_t i = 0; i < len; i++) s += (double)((const int16_t *)in)[i]; break;
    case ORC_U16: for (size_t i = 0; i < len; i++) s += (double)((const uint16_t *)in)[i]; break;
    case ORC_U64: for (size_t i = 0; i < len; i++) s += (double)((const uint64_t *)in)[i]; break;
    default: break;
    }
    return s;
}

/* Block-level order of the device reduction, for float (f32 accumulate) and double.
 * `get(i)` converts element i of the block's chunk to the accumulation type. */
#define DEFINE_BLOCK_SUM(NAME, ACC)                                                                \
    static ACC NAME(ACC (*get)(const void *, size_t), const void *base, size_t count, int threads, \
                    int vec)                                                                       \
    {                                                                                              \
        ACC *tot = (ACC *)malloc(sizeof(ACC) * (size_t)threads);                                   \
        const size_t nunits = count / (size_t)vec;                                                 \
        const size_t rem = count % (size_t)vec;                                                    \
        for (int t = 0; t < threads; t++) {                                                        \
            ACC acc[16] = {0};                                                                         \
            for (int j = 0; j < vec; j++) acc[j] = (ACC)0;                                         \
            for (size_t u = (size_t)t; u < nunits; u += (size_t)threads)                           \
                for (int j = 0; j < vec; j++) acc[j] = acc[j] + get(base, u * (size_t)vec + (size_t)j); \
            if ((size_t)t < rem) acc[0] = acc[0] + get(base, nunits * (size_t)vec + (size_t)t);    \
            ACC s = acc[0];                                                                        \
            for (int j = 1; j < vec; j++) s = s + acc[j];                                          \
            tot[t] = s;                                                                            \
        }                                                                                          \
        /* warp xor-shuffle trees: every lane ends with the same value */                          \
        const int nwarps = threads / 32;                                                           \
        ACC warp_tot[32];                                                                          \
        for (int w = 0; w < 32; w++) warp_tot[w] = (ACC)0;                                         \
        for (int w = 0; w < nwarps; w++) {                                                         \
            ACC lane[32], nxt[32];                                                                 \
            for (int l = 0; l < 32; l++) lane[l] = tot[w * 32 + l];                                \
            for (int off = 16; off >= 1; off >>= 1) {                                              \
                for (int l = 0; l < 32; l++) nxt[l] = lane[l] + lane[l ^ off];                     \
                for (int l = 0; l < 32; l++) lane[l] = nxt[l];                                     \
            }                                                                                      \
            warp_tot[w] = lane[0];                                                                 \
        }                                                                                          \
        ACC lane[32], nxt[32];                                                                     \
        for (int l = 0; l < 32; l++) lane[l] = warp_tot[l];                                        \
        for (int off = 16; off >= 1; off >>= 1) {                                                  \
            for (int l = 0; l < 32; l++) nxt[l] = lane[l] + lane[l ^ off];                         \
            for (int l = 0; l < 32; l++) lane[l] = nxt[l];                                         \
        }                                                                                          \
        free(tot);                                                                                 \
        return lane[0];                                                                            \
    }

DEFINE_BLOCK_SUM(block_sum_f32, float)
DEFINE_BLOCK_SUM(block_sum_f64, double)

static float get_f32(const void *p, size_t i) { return ((const float *)p)[i]; }
static float get_f16(const void *p, size_t i) { return orc_f16_to_f32(((const uint16_t *)p)[i]); }
static float get_bf16(const void *p, size_t i) { return orc_bf16_to_f32(((const uint16_t *)p)[i]); }
static double get_f64(const void *p, size_t i) { return ((const double *)p)[i]; }

int orc_sum_two_pass(int dtype, const void *in, size_t len, int blocks, size_t chunk, int threads,
                     int vec, int threads2, void *out)
{
    if (blocks <= 0 || threads % 32 || threads2 % 32 || vec < 1 || vec > 16) return ORC_ERR_ARG;
    if (dtype == ORC_F32 || dtype == ORC_F16 || dtype == ORC_BF16) {
        float *partials = (float *)calloc((size_t)blocks, sizeof(float));
        const size_t sz = orc_dtype_size(dtype);
        for (int b = 0; b < blocks; b++) {
            size_t begin = (size_t)b * chunk;
            if (begin >= len) { partials[b] = 0.f; continue; }
            size_t cnt = len - begin < chunk ? len - begin : chunk;
            partials[b] = block_sum_f32(dtype == ORC_F32 ? get_f32 : (dtype == ORC_F16 ? get_f16 : get_bf16), (const char *)in + begin * sz, cnt, threads, vec);
        }
        *(float *)out = block_sum_f32(get_f32, partials, (size_t)blocks, threads2, 1);
        free(partials);
        return ORC_OK;
    }
    if (dtype == ORC_F64) {
        double *partials = (double *)calloc((size_t)blocks, sizeof(double));
        for (int b = 0; b < blocks; b++) {
            size_t begin = (size_t)b * chunk;
            if (begin >= len) { partials[b] = 0.0; continue; }
            size_t cnt = len - begin < chunk ? len - begin : chunk;
            partials[b] = block_sum_f64(get_f64, (const char *)in + begin * 8, cnt, threads, vec);
        }
        *(double *)out = block_sum_f64(get_f64, partials, (size_t)blocks, threads2, 1);
        free(partials);
        return ORC_OK;
    }
    /* integer sums are exact in i64: order is irrelevant */
    return orc_sum_seq(dtype, in, len, out);
}

/* ------------------------------------------------------------------ ulp statistics (f32)
 * Distance in units in the last place between two f32 arrays on the monotonic integer line; NaN vs NaN
 * counts as 0, NaN vs number as 2^31.  hist[k] counts distances k = 0..4, hist[5] everything above.
 * Returns the maximum and writes the index of its first occurrence. */
uint32_t orc_ulp_stats_f32(const float *got, const float *want, size_t n, uint64_t hist[6], size_t *argmax)
{
    uint32_t worst = 0;
    size_t where = 0;
    for (size_t i = 0; i < n; i++) {
        uint32_t a, b;
        memcpy(&a, &got[i], 4);
        memcpy(&b, &want[i], 4);
        const int na = (a & 0x7fffffffu) > 0x7f800000u, nb = (b & 0x7fffffffu) > 0x7f800000u;
        uint32_t d;
        if (na || nb) {
            d = (na && nb) ? 0u : 0x80000000u;
        } else {
            const int64_t oa = (a & 0x80000000u) ? -(int64_t)(a & 0x7fffffffu) : (int64_t)a;
            const int64_t ob = (b & 0x80000000u) ? -(int64_t)(b & 0x7fffffffu) : (int64_t)b;
            const int64_t diff = oa > ob ? oa - ob : ob - oa;
            d = diff > 0x7fffffff ? 0x7fffffffu : (uint32_t)diff;
        }
        hist[d < 5 ? d : 5]++;
        if (d > worst) {
            worst = d;
            where = i;
        }
    }
    if (argmax) *argmax = where;
    return worst;
}

/* ------------------------------------------------------------------ scale-and-shift equivalence (f32)
 * Checks, over the f32 bit patterns [first, first + count), that two separately rounded operations equal ONE fused
 * multiply-add (glibc fmaf, correctly rounded) bit for bit (NaN == NaN):
 *   mode 0:  (u * P) + C  ==  fmaf(u, P, C)
 *   mode 1:  (u + C) * P  ==  fmaf(u, P, C * P)
 * Returns the number of mismatches and writes the first offending pattern.  This is the machine check of the
 * strength reduction in custos_b200/csrc/expr.cpp (fused_pair_function); built with -ffp-contract=off, so the
 * left-hand sides really round twice. */
struct sa_job {
    uint64_t first, count;
    float P, C;
    int mode;
    uint64_t bad;
    uint32_t first_bad;
};

static void *sa_worker(void *arg)
{
    struct sa_job *j = (struct sa_job *)arg;
    const float P = j->P, C = j->C, CP = C * P;
    for (uint64_t k = 0; k < j->count; k++) {
        const uint32_t bits = (uint32_t)(j->first + k);
        float u;
        memcpy(&u, &bits, 4);
        volatile float two_step;
        float fused;
        if (j->mode == 0) {
            volatile float t = u * P;
            two_step = t + C;
            fused = fmaf(u, P, C);
        } else {
            volatile float t = u + C;
            two_step = t * P;
            fused = fmaf(u, P, CP);
        }
        const float a = two_step;
        uint32_t x, y;
        memcpy(&x, &a, 4);
        memcpy(&y, &fused, 4);
        const int nx = (x & 0x7fffffffu) > 0x7f800000u, ny = (y & 0x7fffffffu) > 0x7f800000u;
        if ((nx || ny) ? (nx != ny) : (x != y)) {
            if (!j->bad) j->first_bad = bits;
            j->bad++;
        }
    }
    return NULL;
}

/* out[i] = fmaf(a[i], b, c): glibc's correctly rounded fused multiply-add, element by element */
void orc_fmaf_array(const float *a, float b, float c, float *out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = fmaf(a[i], b, c);
}

uint64_t orc_check_scale_add(uint64_t first, uint64_t count, float P, float C, int mode, int threads, uint32_t *first_bad)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    struct sa_job jobs[256];
    const uint64_t per = (count + (uint64_t)threads - 1) / (uint64_t)threads;
    int started = 0;
    for (int t = 0; t < threads; t++) {
        const uint64_t b = per * (uint64_t)t;
        if (b >= count) break;
        jobs[t] = (struct sa_job){first + b, (b + per > count) ? count - b : per, P, C, mode, 0, 0};
        if (pthread_create(&tid[t], NULL, sa_worker, &jobs[t]) != 0) {
            sa_worker(&jobs[t]);
            tid[t] = 0;
        }
        started = t + 1;
    }
    uint64_t bad = 0;
    for (int t = 0; t < started; t++) {
        if (tid[t]) pthread_join(tid[t], NULL);
        if (jobs[t].bad && !bad && first_bad) *first_bad = jobs[t].first_bad;
        bad += jobs[t].bad;
    }
    return bad;
}

/* ------------------------------------------------------------------ OptGraph
 * src/modules/graph/node.rs:2-42, opt_graph.rs:6-41, opt_graph/optimize.rs:19-132 */
struct orc_gnode {
    int64_t *deps;
    int n_deps;
    size_t len;
    int skip;
};
struct orc_graph {
    struct orc_gnode *nodes;
    size_t n, cap;
};

orc_graph *orc_graph_new(void) { return (orc_graph *)calloc(1, sizeof(orc_graph)); }

void orc_graph_free(orc_graph *g)
{
    if (!g) return;
    for (size_t i = 0; i < g->n; i++) free(g->nodes[i].deps);
    free(g->nodes);
    free(g);
}

int64_t orc_graph_add_node(orc_graph *g, size_t len, const int64_t *deps, int n_deps)
{
    if (g->n == g->cap) {
        g->cap = g->cap ? g->cap * 2 : 16;
        g->nodes = (struct orc_gnode *)realloc(g->nodes, g->cap * sizeof(struct orc_gnode));
    }
    struct orc_gnode *nd = &g->nodes[g->n];
    nd->n_deps = n_deps;
    nd->deps = n_deps ? (int64_t *)malloc(sizeof(int64_t) * (size_t)n_deps) : NULL;
    for (int i = 0; i < n_deps; i++) nd->deps[i] = deps[i];
    nd->len = len;
    nd->skip = 0;
    return (int64_t)g->n++;
}

int64_t orc_graph_add_leaf(orc_graph *g, size_t len) { return orc_graph_add_node(g, len, NULL, 0); }

void orc_graph_set_skip(orc_graph *g, int64_t idx, int skip) { g->nodes[idx].skip = skip; }

/* Node::is_leaf: no deps, or every dep is the node itself (node.rs:36-42) */
int orc_graph_is_leaf(const orc_graph *g, int64_t idx)
{
    const struct orc_gnode *nd = &g->nodes[idx];
    for (int i = 0; i < nd->n_deps; i++)
        if (nd->deps[i] != idx) return 0;
    return 1;
}

static int has_dep(const struct orc_gnode *nd, int64_t idx)
{
    for (int i = 0; i < nd->n_deps; i++)
        if (nd->deps[i] == idx) return 1;
    return 0;
}

/* optimize.rs:110-131: at most one later node of the same length may consume it */
int orc_graph_is_path_optimizable(const orc_graph *g, int64_t idx)
{
    if (orc_graph_is_leaf(g, idx)) return 0;
    int occurrences = 0;
    for (size_t k = (size_t)idx + 1; k < g->n; k++) {
        const struct orc_gnode *c = &g->nodes[k];
        if (g->nodes[idx].len != c->len || !has_dep(c, idx)) continue;
        if (occurrences >= 1) return 0;
        occurrences++;
    }
    return 1;
}

/* optimize.rs:57-90 */
size_t orc_graph_trace_cache_path_raw(const orc_graph *g, int64_t start, int64_t *out, size_t cap)
{
    if (!orc_graph_is_path_optimizable(g, start)) return 0;
    size_t w = 0;
    int64_t idx = start;
    for (size_t k = (size_t)start + 1; k < g->n; k++) {
        const struct orc_gnode *c = &g->nodes[k];
        if (c->skip) continue;
        if (!has_dep(c, idx)) continue;
        if (g->nodes[start].len != c->len) continue;
        idx = (int64_t)k;
        if (w < cap) out[w] = idx;
        w++;
        if (!orc_graph_is_path_optimizable(g, idx)) break;
    }
    return w;
}

/* optimize.rs:19-54 */
size_t orc_graph_cache_traces(const orc_graph *g, int64_t *out, size_t cap)
{
    size_t w = 0;
    char *visited = (char *)calloc(g->n + 1, 1);
    int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * (g->n + 1));
    for (size_t i = 0; i < g->n; i++) {
        if (orc_graph_is_leaf(g, (int64_t)i)) continue;
        if (g->nodes[i].skip) continue;
        if (visited[i]) continue;
        size_t k = orc_graph_trace_cache_path_raw(g, (int64_t)i, tmp, g->n);
        if (k == 0) continue;
        size_t hdr = w;
        if (w + 2 <= cap) { out[w] = (int64_t)i; out[w + 1] = 0; }
        w += 2;
        int64_t kept = 0;
        for (size_t j = 0; j < k; j++) {
            if (visited[tmp[j]]) continue;
            visited[tmp[j]] = 1;
            if (w < cap) out[w] = tmp[j];
            w++;
            kept++;
        }
        if (hdr + 1 < cap) out[hdr + 1] = kept;
    }
    free(visited);
    free(tmp);
    return w;
}
