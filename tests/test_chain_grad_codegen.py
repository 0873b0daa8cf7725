"""CB_KERNEL_CHAIN_GRAD on the CPU: the generated backward expression of a fused chain of K `unary_ew` ops
(custos_b200/csrc/expr.cpp: chain_grad_tree) against the reference's semantics restated by the oracle —
K separate `add_unary_grad` calls replayed in reverse over zero-initialised intermediate gradients
(src/unary.rs:118-128, src/modules/autograd/tape.rs:39-47, src/devices/cpu_stack_ops.rs:18-30).

The f32 pair function adds the zero of the intermediate gradient buffer only once (before the last multiply) and
drops multiplications by 1.0; the scalar function keeps every add.  Both are interpreted here from the emitted text
and must give the oracle's bits, zeros of either sign, infinities and subnormals included."""
import random
import re

import numpy as np
import pytest

from custos_b200 import _native as N
from custos_b200 import expr as E
from custos_b200.expr import Combiner, Resolve
from custos_b200.workloads import CHAIN8, CHAIN8_GRADS
from oracle import oracle as orc
from tests.helpers import assert_bit_exact, edge_values
from tests.test_scale_add_fusion import interpret_pair_function


def chain_grad_source(fwd, grads, dtype=N.F32):
    return E.cuda_source(list(fwd) + list(grads), dtype, N.KERNEL_CHAIN_GRAD)


def pair_body(src):
    start = src.index("cb_fn2(cb_f2 x, cb_f2 y, bool &redo)")
    return src[start:src.index("#endif", start)]


def scalar_body_as_pair_text(src):
    """The scalar `cb_fn` uses the same straight-line shape with `cb_` instead of `cb2_` names and bare literals:
    rewrite it into the pair dialect so that one interpreter serves both."""
    start = src.index("T cb_fn(T x, T y)")
    body = src[start:src.index("    return x;", start)]
    out = []
    for line in body.splitlines():
        m = re.match(r"\s*const T (t\d+) = (.*);$", line)
        if not m:
            r = re.match(r"\s*x = (t\d+);", line)
            if r:
                out.append(f"    return {r.group(1)};")
            continue
        name, rhs = m.group(1), m.group(2)
        lit = re.match(r"(__uint_as_float\(0x[0-9a-f]{8}u\)) /\*.*\*/$", rhs)
        if lit:
            rhs = f"cb2_splat({lit.group(1)})"
        elif rhs not in ("x", "y"):
            rhs = rhs.replace("cb_", "cb2_", 1)
        out.append(f"    const cb_f2 {name} = {rhs};")
    return "\n".join(out)


def oracle_backward(fwd, grads, x, out_grad, x_grad):
    """What the tape does: activations op by op, then the grad functions in reverse, every intermediate gradient
    starting from zero."""
    acts = [x]
    for f in fwd[:-1]:
        acts.append(orc.apply_fn(f, orc.F32, acts[-1]))
    g = out_grad
    for k in reversed(range(len(fwd))):
        into = x_grad if k == 0 else np.zeros_like(x)
        g = orc.add_unary_grad(grads[k], orc.F32, acts[k], into, g)
    return g


def test_chain8_backward_expression_shape():
    src = chain_grad_source(CHAIN8, CHAIN8_GRADS)
    body = pair_body(src)
    # forward recomputation: the two scale-and-shift steps as fmas, exp / sin / tanh once each; cos and the shared exp
    for name, count in (("cb2_exp(", 1), ("cb2_sin(", 1), ("cb2_cos(", 1), ("cb2_tanh(", 1), ("cb2_fmap(", 2)):
        assert body.count(name) == count, (name, body)
    assert body.count("cb2_add(") == 1      # the zero of the intermediate gradient buffers, once
    assert body.count("cb2_identity(") == 2  # the two `* 1.0` grad closures of the add ops
    scalar = src[src.index("T cb_fn(T x, T y)"):src.index("cb_fn2")]
    assert scalar.count("cb_add(") == 2 + 7  # forward adds + one per intermediate gradient
    assert scalar.count("cb_exp(") == 1 and scalar.count("cb_tanh(") == 1  # shared between forward and grad closures
    assert E.compile_check(CHAIN8 + CHAIN8_GRADS, N.F32, N.KERNEL_CHAIN_GRAD) > 1000


def test_chain_grad_compiles_for_every_float_dtype_and_ints():
    for dt in (N.F32, N.F64, N.F16, N.BF16):
        assert E.compile_check(CHAIN8 + CHAIN8_GRADS, dt, N.KERNEL_CHAIN_GRAD) > 1000
    ints = [lambda x: x.add(1), lambda x: x.mul(3)]
    int_grads = [lambda x: 1, lambda x: 3]
    for dt in (N.I32, N.I64, N.U8, N.I16):
        assert E.compile_check(ints + int_grads, dt, N.KERNEL_CHAIN_GRAD) > 1000


def test_chain_grad_rejects_an_odd_program_count_and_second_markers():
    with pytest.raises(N.CustosError):
        E.compile_check(CHAIN8 + CHAIN8_GRADS[:-1], N.F32, N.KERNEL_CHAIN_GRAD)
    with pytest.raises(N.CustosError):
        E.cuda_source([lambda x, y: x.add(y), lambda x, y: 1.0], N.F32, N.KERNEL_CHAIN_GRAD, 2)


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_generated_backward_equals_the_replayed_tape_on_random_chains(seed, cases=250):
    rng = random.Random(seed)
    lits = [0.5, 2.0, -1.5, 3.0, 0.25, 1.0, -0.0, 8.0, -0.75, 4.0, -1.0, 0.0, 2.0 ** 100, 1e-30, -2.0, 2.0 ** -126]
    data = np.random.default_rng(seed)
    x = np.concatenate([data.uniform(-4, 4, 400).astype(np.float32), edge_values(np.float32),
                        np.array([1e-45, -1e-45, 3e-39, -3e-39, 1.7e38, -1.7e38, 2e-38, 0.0, -0.0, 0.0, -0.0], np.float32)])
    out_grad = np.concatenate([data.uniform(-2, 2, x.size - 8).astype(np.float32),
                               np.array([0.0, -0.0, 1.0, -1.0, np.inf, 1e-45, -0.0, 0.0], np.float32)])
    x_grad = data.permutation(np.concatenate([data.uniform(-1, 1, x.size - 6).astype(np.float32),
                                              np.array([0.0, -0.0, -0.0, 0.0, 3e-39, -0.0], np.float32)]))

    def tree(depth):
        roll = rng.random()
        if depth == 0 or roll < 0.2:
            return Resolve("x") if rng.random() < 0.6 else Combiner._wrap(rng.choice(lits))
        if roll < 0.35:
            return getattr(tree(depth - 1), rng.choice(["neg", "abs", "identity"]))()
        a = tree(depth - 1)
        return getattr(a, rng.choice(["add", "mul", "sub", "add", "mul", "min", "max", "geq"]))(tree(depth - 1))

    def closure(depth):
        t = tree(depth)
        return t if isinstance(t, Resolve) or rng.random() < 0.8 else Combiner._wrap(rng.choice(lits))
    for case in range(cases):
        K = rng.randint(1, 6)
        fwd = [closure(rng.randint(0, 2)) for _ in range(K)]
        grads = [closure(rng.randint(0, 2)) for _ in range(K)]
        want = oracle_backward(fwd, grads, x, out_grad, x_grad)
        src = chain_grad_source(fwd, grads)
        for which, body in (("pair", pair_body(src)), ("scalar", scalar_body_as_pair_text(src))):
            term = interpret_pair_function(body, x, out_grad)
            with np.errstate(all="ignore"):
                got = (x_grad + term).astype(np.float32)
            assert_bit_exact(got, want, f"case {case} ({which}, K={K}): {body}")


def test_seeded_backward_of_chain8_matches_the_golden_gradient():
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / "oracle_vectors.npz")
    x = g["chain8_x_f32"]
    src = chain_grad_source(CHAIN8[:2] + CHAIN8[4:6] + [CHAIN8[7]], CHAIN8_GRADS[:2] + CHAIN8_GRADS[4:6] + [CHAIN8_GRADS[7]])
    # (the exact-op sub-chain only: the interpreter has no transcendentals) ((x+1)*0.5*2+1 negated: grad = -1 exactly)
    term = interpret_pair_function(pair_body(src), x, np.ones_like(x))
    assert np.all(term == np.float32(-1.0))
    assert g["chain8_grad_f32"].shape == x.shape


@pytest.mark.parametrize("dt", [N.F16, N.BF16])
def test_generated_16bit_backward_word_function_equals_the_replayed_tape(dt, cases=150):
    """The f16 / bf16 chain-grad kernel runs the joined backward expression two lanes per 32-bit word (`cb_fnw`): the
    emitted text interpreted with NumPy (every op: widen to f32, one IEEE operation, round to 16 bits) against the
    oracle's op-by-op replay, bit for bit — the same check the f32 pair function gets above."""
    from custos_b200.expr import bf16_from_f32  # noqa: F401
    from tests.test_half_codegen import Half, interpret_word_function, same_bits
    h = Half(dt)
    rng = random.Random(40 + dt)
    lits = [0.5, 2.0, -1.5, 3.0, 0.25, 1.0, -0.0, 8.0, -0.75, 0.0, -1.0, 100.0, 6.1035e-05]
    data = np.random.default_rng(40 + dt)
    etype = np.float16 if dt == N.F16 else np.float32
    x = h.narrow(np.concatenate([data.uniform(-4, 4, 300).astype(np.float32), edge_values(etype).astype(np.float32)]))
    og = h.narrow(np.concatenate([data.uniform(-2, 2, x.size - 6).astype(np.float32), np.array([0.0, -0.0, 1.0, -1.0, np.inf, 6e-8], np.float32)]))
    x_grad = data.permutation(h.narrow(np.concatenate([data.uniform(-1, 1, x.size - 4).astype(np.float32), np.array([0.0, -0.0, -0.0, 0.0], np.float32)])))
    as_dt = (lambda b: b.view(np.float16)) if dt == N.F16 else (lambda b: b)

    def tree(depth):
        roll = rng.random()
        if depth == 0 or roll < 0.2:
            return Resolve("x") if rng.random() < 0.6 else Combiner._wrap(rng.choice(lits))
        if roll < 0.35:
            return getattr(tree(depth - 1), rng.choice(["neg", "abs", "identity"]))()
        return getattr(tree(depth - 1), rng.choice(["add", "mul", "sub", "add", "mul", "min", "max", "geq"]))(tree(depth - 1))
    for case in range(cases):
        K = rng.randint(1, 5)
        fwd = [tree(rng.randint(0, 2)) for _ in range(K)]
        grads = [tree(rng.randint(0, 2)) for _ in range(K)]
        acts = [as_dt(x)]
        for f in fwd[:-1]:
            acts.append(orc.apply_fn(f, dt, acts[-1]))
        want = as_dt(og)
        for k in reversed(range(K)):
            into = as_dt(x_grad) if k == 0 else np.zeros_like(as_dt(x))
            want = orc.add_unary_grad(grads[k], dt, acts[k], into, want)
        src = chain_grad_source(fwd, grads, dt)
        start = src.index("cb_fnw(cb_w x, cb_w y, bool &redo)")
        body = src[start:src.index("#endif", start)]
        term = interpret_word_function(body, h, x, og)
        with np.errstate(all="ignore"):
            got = h.narrow(h.widen(x_grad) + h.widen(term))
        assert same_bits(h, got, np.asarray(want).view(np.uint16)), f"dtype {dt} case {case} (K={K}):\n{body}"
