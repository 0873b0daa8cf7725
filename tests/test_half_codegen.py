"""The generated 16-bit "word" function (`cb_fnw`: two f16 / bf16 lanes per 32-bit register), interpreted on the CPU.

The code generator special-cases literal operands (`cbw_add_c` with an f32 literal, `x - c` as `x + (-c)`, `cbw_mul_c`
with the 16-bit pattern) and lifts the rest lane by lane.  This test executes the emitted text with NumPy — every op is
"convert to f32, one IEEE operation, round to the 16-bit type", the reference's semantics — and compares with the
oracle's evaluation of the IR, bit for bit, on random exact-op trees and edge values.  It is the CPU-side twin of
tests/test_gpu_fuzz_expr.py for the half types."""
import random
import re
import struct

import numpy as np
import pytest

from custos_b200 import _native as N
from custos_b200 import expr as E
from custos_b200.expr import Combiner, Resolve, bf16_from_f32, bf16_to_f32
from oracle import oracle as orc
from tests.helpers import edge_values


class Half:
    """16-bit storage <-> f32, for f16 (NumPy's binary16) and bf16 (the package's RNE helper = half::bf16::from_f32)."""

    def __init__(self, dt):
        self.dt = dt
        self.one = 0x3C00 if dt == N.F16 else 0x3F80

    def widen(self, bits):
        bits = np.asarray(bits, np.uint16)
        return bits.view(np.float16).astype(np.float32) if self.dt == N.F16 else bf16_to_f32(bits)

    def narrow(self, f):
        f = np.asarray(f, np.float32)
        with np.errstate(over="ignore"):
            return f.astype(np.float16).view(np.uint16) if self.dt == N.F16 else bf16_from_f32(f)


def interpret_word_function(body: str, h: Half, x: np.ndarray, y=None) -> np.ndarray:
    env = {"x": x, "y": y}

    def f32_lit(hex_bits):
        return np.float32(struct.unpack("<f", struct.pack("<I", int(hex_bits, 16)))[0])

    def arith(fn, *args):
        with np.errstate(all="ignore"):
            return h.narrow(fn(*[h.widen(a) for a in args]))

    def select(cond_fn, a, b, if_true, if_false):
        with np.errstate(all="ignore"):
            return np.where(cond_fn(h.widen(a), h.widen(b)), if_true, if_false).astype(np.uint16)
    one, zero = np.uint16(h.one), np.uint16(0)
    for line in body.splitlines():
        m = re.match(r"\s*const cb_w (\w+) = (.*);$", line)
        if not m:
            r = re.match(r"\s*x = (\w+);", line)
            if r:
                env["x"] = env[r.group(1)]
            continue
        name, rhs = m.group(1), m.group(2)
        lit = re.fullmatch(r"cbw_lit\(0x([0-9a-f]{4})u\)", rhs)
        add_c = re.fullmatch(r"cbw_add_c\((\w+), __uint_as_float\(0x([0-9a-f]{8})u\)\)", rhs)
        mul_c = re.fullmatch(r"cbw_mul_c\((\w+), \(T\)0x([0-9a-f]{4})u\)", rhs)
        call = re.fullmatch(r"cbw_(\w+)\((\w+)(?:, (\w+))?\)", rhs)
        if lit:
            env[name] = np.full(x.shape, int(lit.group(1), 16), np.uint16)
        elif rhs in ("x", "y"):
            env[name] = env[rhs]
        elif add_c:
            c = f32_lit(add_c.group(2))
            env[name] = arith(lambda a: a + c, env[add_c.group(1)])
        elif mul_c:
            c = h.widen(np.array([int(mul_c.group(2), 16)], np.uint16))[0]
            env[name] = arith(lambda a: a * c, env[mul_c.group(1)])
        elif call:
            op, a, b = call.group(1), env[call.group(2)], env.get(call.group(3)) if call.group(3) else None
            if op in ("add", "sub", "mul", "div"):
                env[name] = arith({"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide}[op], a, b)
            elif op == "neg":
                env[name] = (a ^ np.uint16(0x8000)).astype(np.uint16)
            elif op == "identity":
                env[name] = a
            elif op == "abs":
                env[name] = arith(np.abs, a)
            elif op == "min":
                env[name] = select(lambda p, q: p < q, a, b, a, b)
            elif op == "max":  # f16 and bf16 (number.rs:507-510, 536-539): half's inherent max keeps `self` unless other > self
                env[name] = select(lambda p, q: q > p, a, b, b, a)
            elif op in ("geq", "leq", "eq"):
                cmp = (lambda p, q: p >= q) if op == "geq" else (lambda p, q: p <= q)
                env[name] = select(cmp, a, b, one, zero)
            else:
                raise AssertionError(f"cannot interpret: {line}")
        else:
            raise AssertionError(f"cannot interpret: {line}")
    return env["x"]


def word_function(fs, dt, n_args=1, kind=N.KERNEL_APPLY):
    src = E.cuda_source(fs if isinstance(fs, list) else [fs], dt, kind, n_args)
    start = src.index("cb_fnw(cb_w x, cb_w y, bool &redo)")
    return src[start:src.index("#endif", start)]


def same_bits(h, got, want):
    nan = lambda b: (b & 0x7FFF) > (0x7C00 if h.dt == N.F16 else 0x7F80)  # noqa: E731
    return np.all((got == want) | (nan(got) & nan(want)))


@pytest.mark.parametrize("dt", [N.F16, N.BF16])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_generated_word_function_equals_the_oracle_on_random_trees(dt, seed):
    h = Half(dt)
    rng = random.Random(seed)
    lits = [0.5, 2.0, -1.5, 3.0, 0.25, 1.0, -0.0, 8.0, -0.75, 0.0, 100.0, 6.1035e-05, 3e4]
    bins = ["add", "mul", "sub", "div", "add", "mul", "sub", "min", "max", "geq", "leq", "eq"]
    data = np.random.default_rng(seed)
    xf = np.concatenate([data.uniform(-4, 4, 500).astype(np.float32), edge_values(np.float16 if dt == N.F16 else np.float32).astype(np.float32)])
    x = h.narrow(xf)
    y = data.permutation(x)

    def tree(depth, leaves):
        roll = rng.random()
        if depth == 0 or roll < 0.15:
            return rng.choice(leaves) if rng.random() < 0.6 else Combiner._wrap(rng.choice(lits))
        if roll < 0.25:
            return getattr(tree(depth - 1, leaves), rng.choice(["neg", "abs", "identity"]))()
        a = tree(depth - 1, leaves)
        return getattr(a, rng.choice(bins))(a if rng.random() < 0.1 else tree(depth - 1, leaves))
    special = 0
    for case in range(400):
        if case % 4 == 3:
            fs = [tree(rng.randint(1, 3), [Resolve("x")]) for _ in range(rng.randint(2, 4))]
            body, want, got_y = word_function(fs, dt), orc.apply_chain(fs, dt, x.view(np.float16) if dt == N.F16 else x), None
        elif case % 4 == 2:
            f = tree(rng.randint(1, 4), [Resolve("x"), Resolve("y")])
            body = word_function(f, dt, 2, N.KERNEL_BINARY)
            want, got_y = orc.apply2(f, dt, x.view(np.float16) if dt == N.F16 else x, y.view(np.float16) if dt == N.F16 else y), y
        else:
            f = tree(rng.randint(1, 4), [Resolve("x")])
            body, want, got_y = word_function(f, dt), orc.apply_fn(f, dt, x.view(np.float16) if dt == N.F16 else x), None
        special += body.count("cbw_add_c(") + body.count("cbw_mul_c(")
        got = interpret_word_function(body, h, x, got_y)
        want = np.asarray(want).view(np.uint16)
        assert same_bits(h, got, want), f"dtype {dt} case {case}:\n{body}"
    assert special > 100  # the literal-operand forms are exercised
