"""The exact scale-and-shift fusion of the f32 pair code generator (custos_b200/csrc/expr.cpp: fused_pair_function).

Two halves, both on the CPU: (1) which expressions the generator rewrites into one fma (read off the generated CUDA);
(2) a machine check, over ALL 2^32 f32 values of `u`, that for those constants two separately rounded operations and
one fused multiply-add give the same bits — and that the check does find mismatches where the generator refuses."""
import re

import numpy as np

import pytest

from custos_b200 import _native as N
from custos_b200 import expr as E
from custos_b200.workloads import CHAIN8
from oracle import oracle as orc


def pair_function(fs, n_args=1, kind=N.KERNEL_APPLY):
    src = E.cuda_source(fs if isinstance(fs, list) else [fs], N.F32, kind, n_args)
    start = src.index("cb_fn2(cb_f2 x, cb_f2 y, bool &redo)")
    return src[start:src.index("#endif", start)]


def fmas(body):
    """[(P bits, addend bits)] of the fused fmas in a pair function"""
    return [(int(p, 16), int(c, 16)) for p, c in
            re.findall(r"cb2_fmap\(t\d+, cb2_splat\(__uint_as_float\(0x([0-9a-f]{8})u\)\), cb2_splat\(__uint_as_float\(0x([0-9a-f]{8})u\)\)\)", body)]


def test_chain8_fuses_its_two_scale_and_shift_steps():
    body = pair_function(CHAIN8)
    assert fmas(body) == [(0x3f000000, 0x3f000000), (0x40000000, 0x3f800000)]  # fma(x, 0.5, 0.5), fma(y, 2, 1)
    assert body.count("cb2_add(") == 0 and body.count("cb2_mul(") == 0
    for name in ("cb2_exp(", "cb2_sin(", "cb2_tanh(", "cb2_neg("):
        assert body.count(name) == 1
    scalar = E.cuda_source(CHAIN8, N.F32)
    scalar = scalar[scalar.index("T cb_fn(T x, T y)"):scalar.index("cb_fn2")]
    assert scalar.count("cb_add(") == 2 and scalar.count("cb_mul(") == 2  # the scalar form keeps both roundings


FUSED = {  # expression -> the one fma it must become
    "x*2+1": (lambda x: x.mul(2.0).add(1.0), (2.0, 1.0)),
    "1+x*2": (lambda x: E.Combiner._wrap(1.0).add(x.mul(2.0)), (2.0, 1.0)),
    "x*-4+3.5": (lambda x: x.mul(-4.0).add(3.5), (-4.0, 3.5)),
    "x*1+0.1": (lambda x: x.mul(1.0).add(0.1), (1.0, float.fromhex("0x1.99999ap-4"))),
    "(x+1)*0.5": (lambda x: x.add(1.0).mul(0.5), (0.5, 0.5)),
    "(x+3)*8": (lambda x: x.add(3.0).mul(8.0), (8.0, 24.0)),
    "0.25*(x+-1.5)": (lambda x: E.Combiner._wrap(0.25).mul(x.add(-1.5)), (0.25, -0.375)),
    "(x+1e-30)*0.25": (lambda x: x.add(1e-30).mul(0.25), (0.25, None)),  # 2^-100 - 24 - 2 = -126: just allowed
}
NOT_FUSED = {
    "x*3+1 (3 is not a power of two)": lambda x: x.mul(3.0).add(1.0),
    "x*0.5+2^-140 (the product is inexact for odd subnormals)": lambda x: x.mul(0.5).add(float.fromhex("0x1p-140")),
    "x*0.5+1 (same rule, although this addend would hide the lost bit)": lambda x: x.mul(0.5).add(1.0),
    "(x+1)*3": lambda x: x.add(1.0).mul(3.0),
    "(x+1e-30)*2^-80 (the result could be subnormal)": lambda x: x.add(1e-30).mul(2.0 ** -80),
    "(x+1e-30)*0.125 (one binade too far: 2^-127 is possible)": lambda x: x.add(1e-30).mul(0.125),
    "x*2+1e35 (the addend could undo an overflow)": lambda x: x.mul(2.0).add(1e35),
    "(x+0)*2": lambda x: x.add(0.0).mul(2.0),
    "(x+1)*-2 (x = -1 gives -0 in two steps, +0 fused)": lambda x: x.add(1.0).mul(-2.0),
    "(x+1e35)*2^-80 (x + 1e35 can overflow although the scaled result is finite)": lambda x: x.add(1e35).mul(2.0 ** -80),
    "x*x+1": lambda x: x.mul(x).add(1.0),
}


@pytest.mark.parametrize("name", sorted(FUSED))
def test_fused_patterns(name):
    f, (P, addend) = FUSED[name]
    got = fmas(pair_function(f))
    assert len(got) == 1, name
    import struct
    assert got[0][0] == struct.unpack("<I", struct.pack("<f", P))[0]
    if addend is not None:
        assert got[0][1] == struct.unpack("<I", struct.pack("<f", addend))[0]


@pytest.mark.parametrize("name", sorted(NOT_FUSED))
def test_patterns_that_must_stay_two_operations(name):
    assert fmas(pair_function(NOT_FUSED[name])) == [], name


def test_a_shared_inner_node_is_not_absorbed():
    def f(x):
        t = x.mul(2.0)
        return t.add(1.0).mul(t)  # t has two readers: it must stay a node of its own
    body = pair_function(f)
    assert fmas(body) == [] and body.count("cb2_mul(") == 2 and body.count("cb2_add(") == 1


def test_nested_patterns_fuse_once_and_still_compile():
    # ((x + 1) * 2) + 3: the inner pair becomes fma(x, 2, 2); the outer add must then stay an add of that value
    for f, n_fma, n_add, n_mul in ((lambda x: x.add(1.0).mul(2.0).add(3.0), 1, 1, 0),
                                   (lambda x: x.mul(2.0).add(1.0).mul(4.0), 1, 0, 1),
                                   (lambda x: x.mul(2.0).add(1.0).mul(4.0).add(0.5).mul(x), 2, 0, 1),
                                   (lambda x: x.add(1.0).mul(2.0).add(x.mul(4.0).add(2.0)), 2, 1, 0)):
        body = pair_function(f)
        assert (len(fmas(body)), body.count("cb2_add("), body.count("cb2_mul(")) == (n_fma, n_add, n_mul), body
        assert E.compile_check([f], N.F32) > 1000
        assert E.compile_check([f], N.F32, N.KERNEL_UNARY_GRAD) > 1000


def test_random_trees_still_compile_with_the_fusion_on():
    import random
    from custos_b200.expr import Combiner, Resolve
    rng = random.Random(5)
    lits = [0.5, 2.0, -1.5, 3.0, 0.25, 1.0, -0.0, 8.0, -0.75, 4.0]

    def tree(depth):
        if depth == 0 or rng.random() < 0.2:
            return Resolve("x") if rng.random() < 0.6 else Combiner._wrap(rng.choice(lits))
        a, b = tree(depth - 1), tree(depth - 1)
        return getattr(a, rng.choice(["add", "mul", "add", "mul", "sub"]))(b)
    for _ in range(25):
        assert E.compile_check([tree(rng.randint(2, 5))], N.F32) > 1000


def test_fusion_across_recorded_ops_and_in_the_other_kernel_kinds():
    assert len(fmas(pair_function([lambda x: x.sin(), lambda x: x.mul(4.0), lambda x: x.add(0.5), lambda x: x.cos()]))) == 1
    assert len(fmas(pair_function(lambda x: x.mul(2.0).add(1.0), kind=N.KERNEL_UNARY_GRAD))) == 1
    assert len(fmas(pair_function(lambda x, y: x.mul(y).add(1.0), n_args=2, kind=N.KERNEL_BINARY))) == 0
    assert len(fmas(pair_function(lambda x, y: x.add(y).mul(2.0).add(1.0), n_args=2, kind=N.KERNEL_BINARY))) == 1
    # f64 and the 16-bit types are generated as before
    def generated(dt):
        src = E.cuda_source([lambda x: x.mul(2.0).add(1.0)], dt)
        start = src.index("// generated from the recorded Combiner trees")
        return src[start:src.index("}  // namespace CB_NS", start)]
    assert "cb_fn2" not in generated(N.F64) and "cb2_fmap" not in generated(N.F64)
    assert "cbw_mul_c" in generated(N.F16) and "cb2_fmap" not in generated(N.F16)


EXHAUSTIVE = [  # (P, C, mode): mode 0 = (u * P) + C, mode 1 = (u + C) * P — every u in 2^32
    (2.0, 1.0, 0), (0.5, 1.0, 1),                      # CHAIN8
    (-4.0, 3.5, 0), (1.0, float.fromhex("0x1.99999ap-4"), 0), (2.0 ** 100, 2.0 ** 100, 0),
    (8.0, 3.0, 1), (0.25, -1.5, 1), (0.25, float.fromhex("0x1.4484cp-100"), 1),  # the last: C = f32(1e-30), the limit case
]


@pytest.mark.parametrize("P,C,mode", EXHAUSTIVE)
def test_two_roundings_equal_one_fma_for_every_f32(P, C, mode):
    bad, where = orc.check_scale_add(P, C, mode)
    assert bad == 0, f"{bad} mismatches, first at u = {where:#010x}"


def test_the_check_finds_the_cases_the_generator_refuses():
    # (ranges around the first mismatch an exhaustive run reported, to keep the CPU suite short)
    assert orc.check_scale_add(0.5, float.fromhex("0x1p-140"), 0, 0x80000000, 1 << 20)[0] > 0  # x * 0.5 + tiny: a subnormal x loses a bit first
    assert orc.check_scale_add(3.0, 1.0, 0, 0x32000000, 1 << 24)[0] > 0                        # not a power of two
    assert orc.check_scale_add(2.0 ** -80, 1e-30, 1, 0x80000000, 1 << 20)[0] > 0               # (x + C) * P lands in the subnormals
    assert orc.check_scale_add(2.0, float.fromhex("0x1p127"), 0, 0xFEF00000, 1 << 21)[0] > 0   # the addend brings an overflowed product back
    assert orc.check_scale_add(-2.0, 1.0, 1, 0xBF800000, 1) == (1, 0xBF800000)                 # (x + 1) * -2 at x = -1: -0 vs +0
    assert orc.check_scale_add(2.0 ** -80, 1e35, 1, 0x7f000000, 1 << 23)[0] > 0                # x + 1e35 overflows first


# ------------------------------------------------------------------ the generated pair function, interpreted on the CPU
def interpret_pair_function(body: str, x: np.ndarray, y: np.ndarray | None = None) -> np.ndarray:
    """Executes the straight-line `cb_fn2` text the generator emitted, in NumPy f32 (every IEEE op rounds once, like the
    device intrinsics; the fused multiply-adds go through glibc's fmaf).  Exact ops only."""
    import struct
    env = {"x": x, "y": y}

    def lit(bits):
        return np.float32(struct.unpack("<f", struct.pack("<I", int(bits, 16)))[0])

    def value(tok):
        tok = tok.strip()
        m = re.fullmatch(r"cb2_splat\(__uint_as_float\(0x([0-9a-f]{8})u\)\)", tok)
        return lit(m.group(1)) if m else env[tok]

    one = np.float32(1.0)
    ops2 = {"add": lambda a, b: a + b, "mul": lambda a, b: a * b, "sub": lambda a, b: a - b, "div": lambda a, b: a / b,
            "min": lambda a, b: np.where(a < b, a, b), "max": lambda a, b: np.where(a > b, a, b),
            "geq": lambda a, b: np.where(a >= b, one, np.float32(0)), "leq": lambda a, b: np.where(a <= b, one, np.float32(0)),
            "eq": lambda a, b: np.where(a <= b, one, np.float32(0))}
    ops1 = {"neg": lambda a: -a, "abs": np.abs, "identity": lambda a: a}
    result = None
    with np.errstate(all="ignore"):
        for line in body.splitlines():
            m = re.match(r"\s*const cb_f2 (\w+) = (.*);$", line)
            if not m:
                r = re.match(r"\s*return (\w+);", line)
                if r:
                    result = env[r.group(1)]
                continue
            name, rhs = m.group(1), m.group(2)
            if name == "x_in":
                continue
            f = re.fullmatch(r"cb2_fmap\((\w+), (cb2_splat\(.*?\)\)), (cb2_splat\(.*?\)\))\)", rhs)
            c = re.fullmatch(r"cb2_(\w+)\((.*)\)", rhs)
            if f:
                env[name] = orc.fmaf_array(np.broadcast_to(env[f.group(1)], x.shape), float(value(f.group(2))), float(value(f.group(3))))
            elif rhs.startswith("cb2_splat("):
                env[name] = value(rhs)
            elif rhs in ("x", "y"):
                env[name] = env[rhs]
            elif c and c.group(1) in ops2:
                a, b = c.group(2).split(", ")
                env[name] = np.asarray(ops2[c.group(1)](np.asarray(env[a], np.float32), np.asarray(env[b], np.float32)), np.float32)
            elif c and c.group(1) in ops1:
                env[name] = np.asarray(ops1[c.group(1)](np.asarray(env[c.group(2)], np.float32)), np.float32)
            else:
                raise AssertionError(f"cannot interpret: {line}")
    return np.broadcast_to(np.asarray(result, np.float32), x.shape).copy()


@pytest.mark.parametrize("seed", [9, 10, 11, 12])
def test_generated_pair_function_equals_the_oracle_on_random_trees(seed, cases=1200):
    """What the GPU fuzzer checks on the device, checked here on the generated text: thousands of random exact-op trees
    (nested scale-and-shift patterns included), evaluated from the emitted `cb_fn2` and by the oracle from the IR."""
    import random
    from custos_b200.expr import Combiner, Resolve
    from tests.helpers import assert_bit_exact, edge_values
    rng = random.Random(seed)
    lits = [0.5, 2.0, -1.5, 3.0, 0.25, 1.0, -0.0, 8.0, -0.75, 4.0, 1e-30, 0.125, 2.0 ** -80, 1e35, 2.0 ** 100, -2.0, -0.5, 0.0,
            2.0 ** 127, 1e-45, 2.0 ** -126, -4.0, 16.0]
    bins = ["add", "mul", "add", "mul", "sub", "div", "min", "max", "geq", "leq", "eq"]
    data = np.random.default_rng(3)
    x = np.concatenate([data.uniform(-4, 4, 600).astype(np.float32), edge_values(np.float32),
                        np.array([1e-45, -1e-45, 3e-39, -3e-39, 1.7e38, -1.7e38, 3.4e38, -3.4e38, 2e-38, 1e-30 * 0.99], np.float32)])
    y = data.permutation(x)

    def tree(depth, leaves):
        roll = rng.random()
        if depth == 0 or roll < 0.15:
            return rng.choice(leaves) if rng.random() < 0.65 else Combiner._wrap(rng.choice(lits))
        if roll < 0.25:
            return getattr(tree(depth - 1, leaves), rng.choice(["neg", "abs", "identity"]))()
        if roll < 0.5:  # a scale-and-shift step around a sub-tree: the shapes the rewrite looks for (and must refuse)
            u, p, c = tree(depth - 1, leaves), Combiner._wrap(rng.choice(lits)), Combiner._wrap(rng.choice(lits))
            form = rng.randrange(4)
            return (u.mul(p).add(c), c.add(p.mul(u)), u.add(c).mul(p), p.mul(c.add(u)))[form]
        a = tree(depth - 1, leaves)
        return getattr(a, rng.choice(bins))(a if rng.random() < 0.1 else tree(depth - 1, leaves))
    fused_seen = 0
    for case in range(cases):
        two = case % 3 == 0
        leaves = [Resolve("x"), Resolve("y")] if two else [Resolve("x")]
        if case % 5 == 4:  # a chain of recorded ops
            fs = [tree(rng.randint(1, 3), [Resolve("x")]) for _ in range(rng.randint(2, 4))]
            body, want = pair_function(fs), orc.apply_chain(fs, orc.F32, x)
            got = interpret_pair_function(body, x)
        elif two:
            f = tree(rng.randint(1, 5), leaves)
            body, want = pair_function(f, 2, N.KERNEL_BINARY), orc.apply2(f, orc.F32, x, y)
            got = interpret_pair_function(body, x, y)
        else:
            f = tree(rng.randint(1, 5), leaves)
            body, want = pair_function(f), orc.apply_fn(f, orc.F32, x)
            got = interpret_pair_function(body, x)
        fused_seen += len(fmas(body))
        assert_bit_exact(got, want, f"case {case}: {body}")
    assert fused_seen > cases // 12  # the rewrite really is exercised
