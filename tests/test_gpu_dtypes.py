"""GPU parity tests for the dtypes beyond f32/f64/f16/i32/i64/u32/u8: every remaining `CDatatype` of the
reference (src/devices/cdatatype.rs:3-62) — bf16, i8, i16, u16, u64 and storage-only bool.

Same bars as tests/test_gpu_kernels.py: bit-exact for arithmetic, comparisons, copy, clear and integer sums;
bf16 transcendentals are the f32 functions (<= 4 ulp of f32) rounded once to bf16, so they may differ from the
oracle by at most 1 ulp(bf16) and only at a rounding boundary.
"""
import numpy as np
import pytest

from custos_b200 import _native as N
from custos_b200 import CustosError
from custos_b200.expr import bf16_from_f32, bf16_to_f32
from custos_b200.raw import sum_plan
from oracle import oracle as orc
from tests.helpers import (CHAIN8, CHEAP8, NP, assert_bf16_bit_exact, assert_bit_exact, assert_same, bf16_is_nan,
                           bf16_ulp_distance, edge_values, random_inputs)
from tests.test_gpu_kernels import EXACT_OPS, SIZES, TRANSCENDENTAL, run_apply

pytestmark = pytest.mark.gpu

NEW_INTS = [N.I8, N.I16, N.U16, N.U64]
NEW_NUMBERS = [N.BF16] + NEW_INTS


def bf16_inputs(n, seed, lo=-4.0, hi=4.0):
    return np.concatenate([random_inputs(N.BF16, n, seed, lo, hi), bf16_from_f32(edge_values(np.float32))])


# ------------------------------------------------------------------ binary ops (a5)
@pytest.mark.parametrize("dt", NEW_NUMBERS)
@pytest.mark.parametrize("op", [N.BIN_ADD, N.BIN_MUL, N.BIN_SUB, N.BIN_DIV])
def test_binary_bit_exact(raw_device, dt, op):
    dev = raw_device
    for n in SIZES:
        lhs, rhs = random_inputs(dt, n, 2, -1, 1), random_inputs(dt, n, 3, -1, 1)
        if dt != N.BF16 and op == N.BIN_DIV:
            rhs[rhs == 0] = 1
        a, b = dev.upload(lhs), dev.upload(rhs)
        o = dev.alloc(lhs.nbytes)
        dev.binary(dt, op, a, b, o, n)
        assert_same(dt, dev.d2h(o, n, dt), orc.binary(op, dt, lhs, rhs), f"binary op {op} dtype {dt} n {n}")
        for p in (a, b, o):
            dev.free(p)


def test_integer_extremes_wrap_like_release_rust(raw_device):
    dev = raw_device
    for dt in (N.I8, N.I16):
        info = np.iinfo(NP[dt])
        x = np.array([info.min, info.max, -1, 0, 1, info.min + 1], NP[dt])
        for f in (lambda v: v.add(1), lambda v: v.sub(1), lambda v: v.mul(v), lambda v: v.neg(), lambda v: v.div(-1),
                  lambda v: v.div(0), lambda v: v.mul(2).add(v)):
            assert_bit_exact(run_apply(dev, f, dt, x), orc.apply_fn(f, dt, x), f"extremes dtype {dt}")
    x = np.array([0, 1, 2 ** 63, 2 ** 64 - 1, 2 ** 64 - 2], np.uint64)
    for f in (lambda v: v.add(2), lambda v: v.sub(3), lambda v: v.mul(v), lambda v: v.div(7), lambda v: v.geq(2 ** 63)):
        assert_bit_exact(run_apply(dev, f, N.U64, x), orc.apply_fn(f, N.U64, x), "u64 extremes")


# ------------------------------------------------------------------ apply_fn (a1)
@pytest.mark.parametrize("name", sorted(EXACT_OPS))
def test_apply_exact_ops_bf16(raw_device, name):
    f = EXACT_OPS[name]
    x = bf16_inputs(70_001, 11)
    assert_bf16_bit_exact(run_apply(raw_device, f, N.BF16, x), orc.apply_fn(f, N.BF16, x), f"bf16 {name}")


@pytest.mark.parametrize("dt", NEW_INTS)
def test_apply_integer_ops(raw_device, dt):
    x = random_inputs(dt, 50_003, 12)
    for f in (lambda x: x.add(3), lambda x: x.mul(2).add(1), lambda x: x.geq(4), lambda x: x.eq(3),
              lambda x: x.sub(1).mul(x), lambda x: x.div(3), lambda x: x.add(2).add(x.mul(8)), lambda x: x.leq(x.mul(2))):
        assert_bit_exact(run_apply(raw_device, f, dt, x), orc.apply_fn(f, dt, x), f"int dtype {dt}")
    if dt in (N.I8, N.I16):
        assert_bit_exact(run_apply(raw_device, lambda x: x.neg(), dt, x), orc.apply_fn(lambda x: x.neg(), dt, x))
    else:
        with pytest.raises(CustosError) as ei:
            raw_device.compile(lambda x: x.neg(), dt)
        assert ei.value.code == N.CB_ERR_UNSUPPORTED


def test_bool_is_storage_only(raw_device):
    dev, n = raw_device, 70_001
    x = random_inputs(N.BOOL, n, 3)
    p, q = dev.upload(x), dev.alloc(n, zero=False)
    dev.copy(N.BOOL, q, 0, p, 0, n)
    assert np.array_equal(dev.d2h(q, n, N.BOOL), x)
    dev.fill(N.BOOL, q, n, 1)
    assert np.all(dev.d2h(q, n, N.BOOL))
    dev.clear(N.BOOL, q, n)
    assert not np.any(dev.d2h(q, n, N.BOOL))
    for call in (lambda: dev.compile(lambda v: v.add(1), N.BOOL), lambda: dev.binary(N.BOOL, N.BIN_ADD, p, p, q, n),
                 lambda: dev.sum(N.BOOL, p, n)):
        with pytest.raises(CustosError):
            call()
    dev.free(p)
    dev.free(q)


@pytest.mark.parametrize("name", ["exp", "ln", "sin", "cos", "tanh"])
def test_transcendental_bf16_every_value(raw_device, name):
    f = TRANSCENDENTAL[name][0]
    x = np.arange(65536, dtype=np.uint16)
    if name in ("sin", "cos"):
        x = x[np.abs(bf16_to_f32(x)) < 1e9]  # same domain as the f32 test (|x| >= 1e9 is outside the 4-ulp claim)
    got, want = run_apply(raw_device, f, N.BF16, x), orc.apply_fn(f, N.BF16, x)
    d = bf16_ulp_distance(got, want)
    worst = int(np.argmax(d))
    assert d[worst] <= 1, f"bf16 {name}: {d[worst]:.0f} ulp at input {int(x[worst]):#06x}: got {int(got[worst]):#06x} want {int(want[worst]):#06x}"
    frac_exact = float(np.mean((got == want) | (bf16_is_nan(got) & bf16_is_nan(want))))
    assert frac_exact > 0.999, f"bf16 {name}: only {frac_exact:.5f} of results identical"


def test_f16_max_keeps_self_on_ties(raw_device):
    # src/number.rs:507-510 and :536-539: Number::max for f16 AND bf16 is half's inherent max (`other > self ? other :
    # self`); f32 / f64 take the trait default (`self > rhs ? self : rhs`).  They differ on +0 / -0.
    dev = raw_device
    a, b = np.array([0.0, -0.0, 1.0, 2.0], np.float16), np.array([-0.0, 0.0, 2.0, 1.0], np.float16)
    for dt, (x, y) in ((N.F16, (a, b)), (N.F32, (a.astype(np.float32), b.astype(np.float32))),
                       (N.BF16, (bf16_from_f32(a.astype(np.float32)), bf16_from_f32(b.astype(np.float32))))):
        for f in (lambda p, q: p.max(q), lambda p, q: p.min(q)):
            e = dev.compile(f, dt, N.KERNEL_BINARY)
            px, py = dev.upload(x), dev.upload(y)
            po = dev.alloc(x.nbytes)
            dev.apply2(e, px, py, po, x.size)
            got, want = dev.d2h(po, x.size, dt), orc.apply2(f, dt, x, y)
            assert got.tobytes() == want.tobytes(), (dt, got, want)
            for p in (px, py, po):
                dev.free(p)


# ------------------------------------------------------------------ fused chains (a8)
def test_chain8_bf16_stepwise(raw_device):
    dev, dt = raw_device, N.BF16
    x = random_inputs(dt, (1 << 20) + 17, 4, -4, 4)
    cur = x
    for k, (f, limit) in enumerate(zip(CHAIN8, [0, 0, 1, 1, 0, 0, 1, 0])):
        nxt, want = run_apply(dev, f, dt, cur), orc.apply_fn(f, dt, cur)
        if limit == 0:
            assert_bf16_bit_exact(nxt, want, f"op {k}")
        else:
            assert float(bf16_ulp_distance(nxt, want).max()) <= limit, f"op {k}"
        cur = nxt
    fused = run_apply(dev, CHAIN8, dt, x)
    assert_bf16_bit_exact(fused, cur, "fused kernel vs op-by-op on the device")
    want = orc.apply_chain(CHAIN8, dt, x)
    err = np.abs(bf16_to_f32(fused).astype(np.float64) - bf16_to_f32(want).astype(np.float64))
    assert float(err.max()) < 3e-2, float(err.max())  # outputs in (-1, 1); one bf16 ulp near 1 is 7.8e-3
    print(f"chain8 bf16: max abs err {err.max():.3e}, bit-identical {np.mean(fused == want):.4f}")


def test_cheap_chain_bf16_bit_exact_and_unaligned(raw_device):
    dev = raw_device
    x = bf16_inputs(300_007, 5)
    assert_bf16_bit_exact(run_apply(dev, CHEAP8, N.BF16, x), orc.apply_chain(CHEAP8, N.BF16, x), "cheap8 bf16")
    e = dev.compile(CHEAP8, N.BF16)
    p, q = dev.upload(x), dev.alloc(x.nbytes)
    dev.apply(e, p + 2, q + 6, x.size - 5)  # not 16-byte aligned: scalar kernel
    assert_bf16_bit_exact(dev.d2h(q + 6, x.size - 5, N.BF16), orc.apply_chain(CHEAP8, N.BF16, x[1:-4]), "unaligned")
    dev.free(p)
    dev.free(q)


# ------------------------------------------------------------------ unary_grad (a2)
@pytest.mark.parametrize("dt", NEW_NUMBERS)
def test_unary_grad_is_mul_then_add(raw_device, dt):
    dev, n = raw_device, 300_011
    lhs, og, lg = random_inputs(dt, n, 31), random_inputs(dt, n, 32), random_inputs(dt, n, 33)
    pl, pg, po = dev.upload(lhs), dev.upload(lg), dev.upload(og)
    two = 2.0 if dt == N.BF16 else 2
    for g in (lambda x: x.mul(two).add(two), lambda x: two, lambda x: x.mul(x).mul(two)):
        dev.h2d(pg, lg)
        dev.unary_grad(dev.compile(g, dt, N.KERNEL_UNARY_GRAD), pl, pg, po, n)
        assert_same(dt, dev.d2h(pg, n, dt), orc.add_unary_grad(g, dt, lhs, lg, og), f"unary_grad dtype {dt}")
    for p in (pl, pg, po):
        dev.free(p)


# ------------------------------------------------------------------ clear / fill / copy (a6, a7)
@pytest.mark.parametrize("dt", NEW_NUMBERS)
def test_clear_fill_copy(raw_device, dt):
    dev = raw_device
    for n in (1, 6, 1000, 70_001):
        x = random_inputs(dt, n, 41)
        p = dev.upload(x)
        q = dev.alloc(x.nbytes, zero=False)
        dev.copy(dt, q, 0, p, 0, n)
        assert np.array_equal(dev.d2h(q, n, dt), x), "copy"
        dev.copy(dt, q, 1, p, 0, n - 1)  # element offsets, not bytes
        assert np.array_equal(dev.d2h(q, n, dt)[1:], x[:-1]), "copy_slice"
        dev.clear(dt, p, n)
        assert not np.any(dev.d2h(p, n, dt).view(np.uint8)), "clear"
        dev.fill(dt, p, n, 1)
        one = 0x3F80 if dt == N.BF16 else 1
        assert np.all(dev.d2h(p, n, dt) == NP[dt](one)), "fill"
        dev.free(p)
        dev.free(q)


# ------------------------------------------------------------------ sum / mean (a13)
def test_sums(raw_device):
    dev, n = raw_device, (1 << 20) + 3
    y = random_inputs(N.BF16, n, 7, 0, 1)
    q = dev.upload(y)
    plan = sum_plan(N.BF16, n)
    want = orc.sum_two_pass(N.BF16, y, plan["blocks"], plan["chunk"], plan["threads"], plan["vec"], plan["threads2"])
    got = dev.sum(N.BF16, q, n)
    assert got.tobytes() == want.tobytes(), (got, want)  # f32 accumulation in the stated order
    assert abs(float(got) - orc.sum_f64(N.BF16, y)) <= 1e-6 * orc.sum_f64(N.BF16, y)
    assert dev.mean(N.BF16, q, n) == np.float32(got / np.float32(n))
    dev.free(q)
    for dt in NEW_INTS:
        info = np.iinfo(NP[dt])
        x = np.random.default_rng(8).integers(info.min, min(info.max, 2 ** 40), n, dtype=np.int64).astype(NP[dt])
        q = dev.upload(x)
        assert dev.sum(dt, q, n) == orc.sum_seq(dt, x) == np.int64(x.astype(np.int64).sum()), dt  # exact in i64
        dev.free(q)


# ------------------------------------------------------------------ module stack with a non-default T
def test_lazy_graph_fused_chain_and_backward_bf16():
    # Lazy<'a, Mods, T> / Graph<Mods, T> are typed by the module's T (lazy.rs:37, graph.rs:21): a bf16 stack
    from custos_b200.device import CUDA
    n = 50_003
    x = random_inputs(N.BF16, n, 4, -4, 4)
    with CUDA("Lazy", "Graph", "Base", dtype=N.BF16) as dev:
        cur = dev.buffer(x, dtype=N.BF16)
        for f in CHEAP8:
            cur = dev.apply_fn(cur, f)
        dev.optimize_mem_graph()
        dev.unary_fusing()
        before = dev.raw.launches
        dev.run()
        assert dev.raw.launches - before == 1, "eight recorded bf16 ops must run as ONE kernel"
        assert_bf16_bit_exact(cur.replace().read(), orc.apply_chain(CHEAP8, orc.BF16, x), "fused cheap8 bf16")
    with CUDA("Autograd", "Base", dtype=N.BF16) as dev:
        buf = dev.buffer(x, dtype=N.BF16).require_grad()
        out = dev.unary_ew(buf, lambda v: v.mul(3.0), lambda v: 3.0)
        out = dev.unary_ew(out, lambda v: v.mul(v), lambda v: v.mul(2.0))
        out.backward()  # seed = ones (bf16 0x3f80), grads: 2*(3x) then *3, each rounded to bf16
        a1 = orc.apply_fn(lambda v: v.mul(3.0), orc.BF16, x)
        g1 = orc.add_unary_grad(lambda v: v.mul(2.0), orc.BF16, a1, np.zeros(n, np.uint16), np.full(n, 0x3F80, np.uint16))
        g0 = orc.add_unary_grad(lambda v: 3.0, orc.BF16, x, np.zeros(n, np.uint16), g1)
        assert_bf16_bit_exact(buf.grad().read(), g0, "bf16 backward")


def test_module_buffers_of_every_new_dtype():
    from custos_b200.device import CUDA
    with CUDA("Base") as dev:
        for dt in NEW_INTS:
            x = random_inputs(dt, 1000, 9)
            b = dev.buffer(x, dtype=dt)
            out = dev.apply_fn(b, lambda v: v.mul(3).add(1))
            assert_bit_exact(out.read(), orc.apply_fn(lambda v: v.mul(3).add(1), dt, x), f"apply_fn dtype {dt}")
            c = b.clone()
            c.clear()
            assert not np.any(c.read()) and np.array_equal(b.read(), x)
        flags = dev.buffer(np.array([True, False, True]), dtype=N.BOOL)
        assert flags.read().tolist() == [True, False, True]
        flags.clear()
        assert flags.read().tolist() == [False] * 3


# ------------------------------------------------------------------ 16-bit chains as a table lookup
@pytest.mark.parametrize("dt", [N.F16, N.BF16])
@pytest.mark.parametrize("chain_name", ["chain8", "cheap8", "single_tan"])
def test_large_16bit_buffers_take_the_lookup_kernel_and_match_the_arithmetic_kernel(raw_device, dt, chain_name):
    """cb_apply switches to lut16_kernel at CB_LUT16_MIN_ELEMS (2^22) elements.  The table is filled by the arithmetic
    kernel itself, so both paths must agree bit for bit on every one of the 65 536 inputs — NaN patterns included —
    whatever the buffer length, alignment of the tail, or aliasing of input and output."""
    dev = raw_device
    chain = {"chain8": CHAIN8, "cheap8": CHEAP8, "single_tan": [lambda v: v.tan()]}[chain_name]
    e = dev.compile(chain, dt)
    n = (1 << 22) + 8 * 1024 * 3 + 5  # whole tiles + ragged units + a scalar tail
    rng = np.random.default_rng(31)
    x = np.concatenate([np.arange(65536, dtype=np.uint16), rng.integers(0, 65536, n - 65536).astype(np.uint16)])
    px, po = dev.upload(x), dev.alloc(n * 2)
    before = dev.launches
    dev.apply(e, px, po, n)  # lookup kernel
    assert dev.launches - before == 1
    got = dev.d2h(po, n, N.U16)
    pa = dev.alloc(n * 2)
    assert dev.has_lut(e)
    dev.set_lut(e, False)  # the arithmetic kernel on the same buffer
    dev.apply(e, px, pa, n)
    dev.set_lut(e, True)
    want = dev.d2h(pa, n, N.U16)
    assert got.tobytes() == want.tobytes(), f"{int(np.sum(got != want))} of {n} elements differ"
    # and against the oracle on the 65 536 distinct inputs (exact-op chains only; transcendentals are <= 1 ulp)
    if chain_name == "cheap8":
        ref = orc.apply_chain(chain, dt, x[:65536].view(np.float16) if dt == N.F16 else x[:65536])
        ref = np.asarray(ref).view(np.uint16)
        nan = (lambda b: ((b & 0x7c00) == 0x7c00) & ((b & 0x3ff) != 0)) if dt == N.F16 else bf16_is_nan
        assert np.all((got[:65536] == ref) | (nan(got[:65536]) & nan(ref)))
    dev.apply(e, px, px, n)  # in place
    assert dev.d2h(px, n, N.U16).tobytes() == want.tobytes()
    dev.h2d(px, x)
    dev.apply(e, px + 2, po + 2, n - 1)  # not 16-byte aligned: stays on the arithmetic (scalar) kernel, same bits
    assert dev.d2h(po, n - 1, N.U16, offset_bytes=2).tobytes() == want[1:].tobytes()
    for p in (px, po, pa):
        dev.free(p)


def test_narrow_unsigned_multiply_wraps_without_signed_overflow(raw_device):
    # u16 operands are promoted to (signed) int by C++: 65535 * 65535 overflows it; the kernels multiply as unsigned int
    dev = raw_device
    for dt, np_t, hi in ((N.U16, np.uint16, 65535), (N.U8, np.uint8, 255)):
        a = np.array([hi, hi, hi - 1, 2, 0, hi], np_t)
        b = np.array([hi, 2, hi, hi, hi, 1], np_t)
        pa, pb, po = dev.upload(a), dev.upload(b), dev.alloc(a.nbytes)
        dev.binary(dt, N.BIN_MUL, pa, pb, po, a.size)
        want = (a.astype(np.uint64) * b.astype(np.uint64)).astype(np_t)
        assert dev.d2h(po, a.size, dt).tolist() == want.tolist() == orc.binary(1, dt, a, b).tolist()
        e = dev.compile(lambda x, y: x.mul(y), dt, N.KERNEL_BINARY)
        dev.apply2(e, pa, pb, po, a.size)
        assert dev.d2h(po, a.size, dt).tolist() == want.tolist()
        for p in (pa, pb, po):
            dev.free(p)


def test_u64_sum_wraps_unsigned_and_mean_divides_unsigned(raw_device):
    dev = raw_device
    x = np.full(1000, (1 << 63) // 750, np.uint64)  # the sum passes 2^63: a signed accumulator would go negative
    p = dev.upload(x)
    total = int(x.astype(object).sum())
    assert (1 << 63) <= total < (1 << 64)
    assert int(dev.sum(N.U64, p, x.size)) == total
    assert int(dev.mean(N.U64, p, x.size)) == total // x.size
    dev.free(p)


@pytest.mark.parametrize("dt", [N.F16, N.BF16])
@pytest.mark.parametrize("kind", ["chain_grad", "unary_grad"])
def test_seeded_16bit_backward_by_lookup_matches_the_arithmetic_kernel(raw_device, dt, kind):
    """`backward()` seeds out_grad with ones; the backward term of a 16-bit expression is then a function of lhs alone
    and large buffers take lut16_kernel<MODE 1 / 2>: lhs_grad += table[lhs], out_grad = 1.  Same bits as the arithmetic
    kernel for every input pattern and every start value of the gradient, signed zeros included."""
    from custos_b200.workloads import CHAIN8_GRADS
    dev = raw_device
    if kind == "chain_grad":
        e = dev.compile(CHAIN8 + CHAIN8_GRADS, dt, N.KERNEL_CHAIN_GRAD)
    else:
        e = dev.compile(lambda v: v.cos().mul(0.5), dt, N.KERNEL_UNARY_GRAD)
    assert dev.has_lut(e)
    n = (1 << 22) + 8 * 512 * 5 + 3
    rng = np.random.default_rng(33)
    x = np.concatenate([np.arange(65536, dtype=np.uint16), rng.integers(0, 65536, n - 65536).astype(np.uint16)])
    one = 0x3c00 if dt == N.F16 else 0x3f80
    g0 = rng.integers(0, 65536, n).astype(np.uint16)
    g0[:65536:3] = 0x8000  # -0 start values: a -0 term must survive
    g0[1:65536:3] = 0x0000
    px, pg, po = dev.upload(x), dev.upload(g0), dev.alloc(n * 2)
    before = dev.launches
    dev.unary_grad_ex(e, px, pg, po, n, N.GRAD_SEED_ONES)  # lookup kernel
    assert dev.launches - before == 1
    got_g, got_o = dev.d2h(pg, n, N.U16), dev.d2h(po, n, N.U16)
    dev.h2d(pg, g0)
    dev.clear(N.U16, po, n)
    dev.set_lut(e, False)
    dev.unary_grad_ex(e, px, pg, po, n, N.GRAD_SEED_ONES)  # arithmetic kernel
    dev.set_lut(e, True)
    want_g, want_o = dev.d2h(pg, n, N.U16), dev.d2h(po, n, N.U16)
    assert np.all(want_o == one) and np.all(got_o == one)
    assert got_g.tobytes() == want_g.tobytes(), f"{int(np.sum(got_g != want_g))} of {n} gradients differ"
    # unseeded with out_grad = 1 gives the same gradient (a chain keeps the arithmetic kernel, one op looks g(lhs) up)
    dev.h2d(pg, g0)
    dev.fill(dt, po, n, 1.0)
    dev.unary_grad_ex(e, px, pg, po, n, 0)
    assert dev.d2h(pg, n, N.U16).tobytes() == want_g.tobytes()
    # a general out_grad: for ONE op the table holds g(lhs) and the kernel multiplies and adds in 16 bits
    og = rng.integers(0, 65536, n).astype(np.uint16)
    og[:8] = [0x0000, 0x8000, one, one | 0x8000, 0x7c00 if dt == N.F16 else 0x7f80, 0x0001, 0x7e00 if dt == N.F16 else 0x7fc0, 0x3800]
    results = []
    for use_lut in (True, False):
        dev.h2d(pg, g0)
        dev.h2d(po, og)
        dev.set_lut(e, use_lut)
        before = dev.launches
        dev.unary_grad_ex(e, px, pg, po, n, 0)
        assert dev.launches - before == 1
        results.append(dev.d2h(pg, n, N.U16))
    dev.set_lut(e, True)
    nan = (lambda b: ((b & 0x7c00) == 0x7c00) & ((b & 0x3ff) != 0)) if dt == N.F16 else bf16_is_nan
    same = (results[0] == results[1]) | (nan(results[0]) & nan(results[1]))
    assert np.all(same), f"{int(np.sum(~same))} of {n} gradients differ with a general out_grad ({kind})"
    for p in (px, pg, po):
        dev.free(p)
