"""GPU parity tests of the device-level C ABI (cb_*) against the CPU oracle.

Bars (BASELINE.json north_star): bit-exact for add/mul/sub/div/neg/abs/min/max/cmp/copy/clear,
<= 4 ulp per transcendental (exp/ln/sin/cos/tanh; tan and pow reported with their own
limits), f32 sums bit-exact against the restated two-pass order and within 1e-6 relative
of the fp64 sum.
"""
import numpy as np
import pytest

from custos_b200 import _native as N
from custos_b200 import CustosError
from custos_b200.raw import sum_plan
from oracle import oracle as orc
from tests.helpers import (CHAIN8, CHAIN8_GRADS, CHEAP8, CONFIG1, NP, assert_bit_exact, assert_ulp, edge_values,
                           random_inputs)

pytestmark = pytest.mark.gpu

SIZES = [1, 3, 255, 4099, (1 << 20) + 5]
FLOATS = [N.F32, N.F64, N.F16]
ALL = [N.F32, N.F64, N.F16, N.I32, N.I64, N.U32, N.U8]


def run_apply(dev, fs, dt, x):
    e = dev.compile(fs, dt)
    src = dev.upload(x)
    dst = dev.alloc(x.nbytes)
    dev.apply(e, src, dst, x.size)
    out = dev.d2h(dst, x.size, dt)
    dev.free(src)
    dev.free(dst)
    return out


# ------------------------------------------------------------------ binary ops (a5)
@pytest.mark.parametrize("dt", ALL)
@pytest.mark.parametrize("op", [N.BIN_ADD, N.BIN_MUL, N.BIN_SUB, N.BIN_DIV])
def test_binary_bit_exact(raw_device, dt, op):
    dev = raw_device
    for n in SIZES:
        lhs, rhs = random_inputs(dt, n, 2, -1, 1), random_inputs(dt, n, 3, -1, 1)
        if np.dtype(NP[dt]).kind != "f" and op == N.BIN_DIV:
            rhs[rhs == 0] = 1
        a, b = dev.upload(lhs), dev.upload(rhs)
        o = dev.alloc(lhs.nbytes)
        dev.binary(dt, op, a, b, o, n)
        assert_bit_exact(dev.d2h(o, n, dt), orc.binary(op, dt, lhs, rhs), f"binary op {op} dtype {dt} n {n}")
        for p in (a, b, o):
            dev.free(p)


def test_binary_demo_655360(raw_device):
    # tests/demo_impl/cuda/mod.rs:40-66: 1 + 4 == 5 over 655 360 f32, 100 times
    dev, n = raw_device, 655_360
    a, b = dev.upload(np.full(n, 1, np.float32)), dev.upload(np.full(n, 4, np.float32))
    o = dev.alloc(n * 4)
    for _ in range(100):
        dev.binary(N.F32, N.BIN_ADD, a, b, o, n)
    assert np.all(dev.d2h(o, n, N.F32) == 5.0)
    for p in (a, b, o):
        dev.free(p)


def test_binary_unaligned_slices(raw_device):
    dev, n = raw_device, 10_000
    lhs, rhs = random_inputs(N.F32, n + 8, 7), random_inputs(N.F32, n + 8, 8)
    a, b = dev.upload(lhs), dev.upload(rhs)
    o = dev.alloc((n + 8) * 4)
    for off in (1, 2, 3):
        dev.binary(N.F32, N.BIN_MUL, a + 4 * off, b + 4 * off, o + 4 * off, n)
        assert_bit_exact(dev.d2h(o + 4 * off, n, N.F32), lhs[off:off + n] * rhs[off:off + n], f"offset {off}")
    for p in (a, b, o):
        dev.free(p)


# ------------------------------------------------------------------ apply_fn (a1): exact class
EXACT_OPS = {
    "add": lambda x: x.add(1.5), "mul": lambda x: x.mul(0.75), "sub": lambda x: x.sub(2.25),
    "div": lambda x: x.div(3.0), "neg": lambda x: x.neg(), "abs": lambda x: x.abs(),
    "min": lambda x: x.min(0.5), "max": lambda x: x.max(-0.5), "identity": lambda x: x.identity(),
    "geq": lambda x: x.geq(0.25), "leq": lambda x: x.leq(0.25), "eq": lambda x: x.eq(0.25),
    "relu": lambda x: x.geq(0.0).mul(x), "poly": lambda x: x.add(2.0).mul(x).add(x.mul(8.0)).mul(5.0),
    "x*2+1": lambda x: x.mul(2.0).add(1.0),
}


@pytest.mark.parametrize("dt", FLOATS)
@pytest.mark.parametrize("name", sorted(EXACT_OPS))
def test_apply_exact_ops(raw_device, dt, name):
    f = EXACT_OPS[name]
    x = np.concatenate([random_inputs(dt, 70_001, 11), edge_values(NP[dt])])
    assert_bit_exact(run_apply(raw_device, f, dt, x), orc.apply_fn(f, dt, x), f"{name} dtype {dt}")


@pytest.mark.parametrize("dt", [N.I32, N.I64, N.U32, N.U8])
def test_apply_integer_ops(raw_device, dt):
    x = random_inputs(dt, 50_003, 12)
    for f in (lambda x: x.add(3), lambda x: x.mul(2).add(1), lambda x: x.geq(4), lambda x: x.eq(3),
              lambda x: x.sub(1).mul(x), lambda x: x.div(3), lambda x: x.add(2).add(x.mul(8))):
        assert_bit_exact(run_apply(raw_device, f, dt, x), orc.apply_fn(f, dt, x), f"int dtype {dt}")
    if dt in (N.I32, N.I64):
        assert_bit_exact(run_apply(raw_device, lambda x: x.neg(), dt, x), orc.apply_fn(lambda x: x.neg(), dt, x))


def test_reference_kats_on_device(raw_device):
    dev = raw_device
    # src/devices/cuda/ops.rs:252-258
    assert run_apply(dev, lambda x: x.add(1.0), N.F32, np.array([1, 2, 3, 4, 5, 6], np.float32)).tolist() == [2, 3, 4, 5, 6, 7]
    # src/two_way_ops/mod.rs:237-246
    assert run_apply(dev, lambda x: x.add(3), N.I32, np.array([3, 3, 4, 5, 3, 2], np.int32)).tolist() == [6, 6, 7, 8, 6, 5]
    # src/two_way_ops/mod.rs:206-219 (clip)
    out = run_apply(dev, lambda x: x.max(3.).min(5.), N.F64, np.array([1., 3., 4., 6., 3., 2.]))
    assert out.tolist() == [3., 3., 4., 5., 3., 3.]
    # src/two_way_ops/mod.rs:265-283, reference tolerance is roughly_eq (0.1); ours is tight
    out = run_apply(dev, lambda x: x.mul(2.).add(4.).sin().mul(x).add(1.), N.F64, np.array([3., 3., 4., 5., 3., 2.]))
    np.testing.assert_allclose(out, [-0.6320633326681093, -0.6320633326681093, -1.1462916720017398,
                                     5.953036778474352, -0.6320633326681093, 2.978716493246764], rtol=1e-14)
    # src/two_way_ops/mod.rs:64-68: exp(1) == E
    assert run_apply(dev, lambda x: x.exp(), N.F32, np.array([1.0], np.float32)).view(np.uint32)[0] == 0x402DF854


def test_unsupported_ops_rejected(raw_device):
    with pytest.raises(CustosError) as ei:
        raw_device.compile(lambda x: x.sin(), N.I32)
    assert ei.value.code == N.CB_ERR_UNSUPPORTED


# ------------------------------------------------------------------ transcendentals: <= 4 ulp each
TRANSCENDENTAL = {
    "exp": (lambda x: x.exp(), 4, (-90.0, 90.0)),
    "ln": (lambda x: x.ln(), 4, (1e-30, 1e30)),
    "sin": (lambda x: x.sin(), 4, (-100.0, 100.0)),
    "cos": (lambda x: x.cos(), 4, (-100.0, 100.0)),
    "tanh": (lambda x: x.tanh(), 4, (-12.0, 12.0)),
    "tan": (lambda x: x.tan(), 6, (-1.5, 1.5)),
}


@pytest.mark.parametrize("name", sorted(TRANSCENDENTAL))
def test_transcendental_ulp_f32(raw_device, name):
    f, limit, (lo, hi) = TRANSCENDENTAL[name]
    rng = np.random.default_rng(21)
    x = np.concatenate([rng.uniform(lo, hi, 400_000), rng.uniform(-4, 4, 200_000)]).astype(np.float32)
    if name == "ln":
        x = np.abs(x) + np.float32(1e-30)
    x = np.concatenate([x, edge_values(np.float32)])
    if name in ("sin", "cos"):
        x = np.concatenate([x[np.abs(x) < 1e9], rng.uniform(-1e6, 1e6, 100_000).astype(np.float32)])
    if name == "tan":
        x = x[np.abs(x) < 1e4]
    mx, mean = assert_ulp(run_apply(raw_device, f, N.F32, x), orc.apply_fn(f, N.F32, x), limit, name)
    print(f"{name}: max {mx:.0f} ulp, mean {mean:.3f} ulp over {x.size} inputs")


@pytest.mark.parametrize("name", ["exp", "ln", "sin", "tanh"])
def test_transcendental_ulp_f64(raw_device, name):
    f, limit, (lo, hi) = TRANSCENDENTAL[name]
    rng = np.random.default_rng(22)
    x = rng.uniform(max(lo, -50), min(hi, 50), 100_000)
    if name == "ln":
        x = np.abs(x) + 1e-300
    assert_ulp(run_apply(raw_device, f, N.F64, x), orc.apply_fn(f, N.F64, x), limit, name)


@pytest.mark.parametrize("name", ["exp", "ln", "sin", "cos", "tanh"])
def test_transcendental_f16_every_value(raw_device, name):
    # all 65536 binary16 inputs; the f32 function is within 4 ulp(f32), so after rounding to
    # f16 the result can differ from the oracle by at most 1 ulp(f16) at a rounding boundary
    f = TRANSCENDENTAL[name][0]
    x = np.arange(65536, dtype=np.uint16).view(np.float16)
    got, want = run_apply(raw_device, f, N.F16, x), orc.apply_fn(f, N.F16, x)
    mx, _ = assert_ulp(got, want, 1, f"f16 {name}")
    frac_exact = float(np.mean((got.view(np.uint16) == want.view(np.uint16)) | (np.isnan(got) & np.isnan(want))))
    assert frac_exact > 0.995, f"f16 {name}: only {frac_exact:.4f} of results identical"


def test_signed_zero_goes_through_odd_functions(raw_device):
    # sin(-0) = tan(-0) = tanh(-0) = -0 and f(+0) = +0, like glibc (and exp(-0) = cos(-0) = 1): the packed forms keep
    # the sign of a zero through the range reduction and the polynomial
    z = np.array([0.0, -0.0], np.float32)
    for dt, x in ((N.F32, z), (N.F64, z.astype(np.float64)), (N.F16, z.astype(np.float16))):
        for name in ("sin", "tanh", "tan") if dt != N.F16 else ("sin", "tanh"):
            f = TRANSCENDENTAL[name][0]
            got = run_apply(raw_device, f, dt, x)
            assert got.tobytes() == x.tobytes(), (dt, name, got)
            assert got.tobytes() == orc.apply_fn(f, dt, x).tobytes()
        for name in ("exp", "cos"):
            assert run_apply(raw_device, TRANSCENDENTAL[name][0], dt, x).tolist() == [1.0, 1.0]


def test_pow_ulp(raw_device):
    rng = np.random.default_rng(23)
    a, b = rng.uniform(0.01, 20, 200_000).astype(np.float32), rng.uniform(-5, 5, 200_000).astype(np.float32)
    dev = raw_device
    e = dev.compile(lambda x, y: x.pow(y), N.F32, N.KERNEL_BINARY)
    pa, pb = dev.upload(a), dev.upload(b)
    po = dev.alloc(a.nbytes)
    dev.apply2(e, pa, pb, po, a.size)
    assert_ulp(dev.d2h(po, a.size, N.F32), orc.apply2(lambda x, y: x.pow(y), N.F32, a, b), 6, "pow")
    # src/two_way_ops/mod.rs:96-101
    e2 = dev.compile(lambda x, y: x.mul(3.).pow(y.add(1.)), N.F32, N.KERNEL_BINARY)
    dev.h2d(pa, np.array([3.0], np.float32))
    dev.h2d(pb, np.array([2.0], np.float32))
    dev.apply2(e2, pa, pb, po, 1)
    assert abs(float(dev.d2h(po, 1, N.F32)[0]) - 729.0) <= 729.0 * 4 * 2 ** -23
    for p in (pa, pb, po):
        dev.free(p)


# ------------------------------------------------------------------ fused chains (a8)
def check_chain_stepwise(dev, chain, limits, dt, x):
    """The rigorous form of "a fused chain matches the reference": composition can amplify a
    1-ulp difference without bound (cancellation in `2*sin(y)+1`), so every op is checked on
    the DEVICE's own input to that op — bit-exact for arithmetic ops, <= limit ulp for
    transcendentals — and the fused kernel must then equal the op-by-op device result bit for bit."""
    cur = x
    for k, (f, limit) in enumerate(zip(chain, limits)):
        nxt = run_apply(dev, f, dt, cur)
        want = orc.apply_fn(f, dt, cur)
        if limit == 0:
            assert_bit_exact(nxt, want, f"op {k}")
        else:
            assert_ulp(nxt, want, limit, f"op {k}")
        cur = nxt
    fused = run_apply(dev, chain, dt, x)
    assert_bit_exact(fused, cur, "fused kernel vs op-by-op on the device")
    return fused


@pytest.mark.parametrize("dt", FLOATS)
def test_chain8_against_oracle(raw_device, dt):
    x = random_inputs(dt, (1 << 20) + 17, 4, -4, 4)
    t = 1 if dt == N.F16 else 4
    fused = check_chain_stepwise(raw_device, CHAIN8, [0, 0, t, t, 0, 0, t, 0], dt, x)
    # end to end against the oracle's own chain: outputs live in (-1, 1), so an absolute bound is meaningful
    want = orc.apply_chain(CHAIN8, dt, x)
    err = np.abs(fused.astype(np.float64) - want.astype(np.float64))
    bound = {N.F32: 1e-5, N.F64: 1e-13, N.F16: 4e-3}[dt]
    assert float(err.max()) < bound, float(err.max())
    d = np.abs(fused.astype(np.float64) - want.astype(np.float64)) / np.maximum(np.abs(want.astype(np.float64)), 1e-30)
    print(f"chain8 dtype {dt}: max abs err {err.max():.3e}, median rel err {np.median(d):.3e}, "
          f"bit-identical {np.mean(fused == want):.4f}")


@pytest.mark.parametrize("dt", FLOATS)
def test_cheap_chain_bit_exact(raw_device, dt):
    x = np.concatenate([random_inputs(dt, 300_007, 5), edge_values(NP[dt])])
    assert_bit_exact(run_apply(raw_device, CHEAP8, dt, x), orc.apply_chain(CHEAP8, dt, x), "cheap8")


def test_chain_equals_unfused_ops_on_device(raw_device):
    # one fused kernel == eight separate launches, bit for bit (same device functions)
    dev = raw_device
    x = random_inputs(N.F32, 500_003, 9)
    fused = run_apply(dev, CHAIN8, N.F32, x)
    cur = x
    for f in CHAIN8:
        cur = run_apply(dev, f, N.F32, cur)
    assert_bit_exact(fused, cur, "fused vs unfused")


def test_config1_chain(raw_device):
    # BASELINE configs[0]: exp().sin()*2+1 on 1M f32, U[-2,2), seed 1
    x = random_inputs(N.F32, 1 << 20, 1, -2, 2)
    fused = check_chain_stepwise(raw_device, CONFIG1, [4, 4, 0, 0], N.F32, x)
    want = orc.apply_chain(CONFIG1, N.F32, x)
    assert float(np.max(np.abs(fused.astype(np.float64) - want.astype(np.float64)))) < 2e-6  # values in [-1, 3]


def test_apply_in_place_and_unaligned(raw_device):
    dev = raw_device
    x = random_inputs(N.F32, 100_003, 13)
    e = dev.compile(CHEAP8, N.F32)
    p = dev.upload(x)
    dev.apply(e, p, p, x.size)  # aliasing produced by optimize_mem_graph
    assert_bit_exact(dev.d2h(p, x.size, N.F32), orc.apply_chain(CHEAP8, N.F32, x), "in place")
    dev.h2d(p, x)
    q = dev.alloc(x.nbytes)
    dev.apply(e, p + 4, q + 12, x.size - 5)  # scalar kernel
    assert_bit_exact(dev.d2h(q + 12, x.size - 5, N.F32), orc.apply_chain(CHEAP8, N.F32, x[1:-4]), "unaligned")
    dev.free(p)
    dev.free(q)


def test_in_place_with_lanes_on_the_slow_path(raw_device):
    # lanes with |x| > 1e5 leave the packed fast path of sin / cos; their tile is reloaded and redone with the
    # scalar forms — that must also hold when the kernel runs in place (nothing of the tile is stored before)
    dev = raw_device
    rng = np.random.default_rng(14)
    x = rng.uniform(-4, 4, 300_007).astype(np.float32)
    x[rng.integers(0, x.size, 2000)] = rng.uniform(-1e6, 1e6, 2000).astype(np.float32)   # scattered slow lanes
    chain = [lambda v: v.mul(1.5), lambda v: v.sin(), lambda v: v.add(0.25), lambda v: v.cos()]
    separate = run_apply(dev, chain, N.F32, x)
    p = dev.upload(x)
    dev.apply(dev.compile(chain, N.F32), p, p, x.size)
    in_place = dev.d2h(p, x.size, N.F32)
    dev.free(p)
    assert_bit_exact(in_place, separate, "in place == out of place")
    cur = x
    for f in chain:  # and both equal the op-by-op device result (every op <= 4 ulp of the oracle on its own input)
        nxt = run_apply(dev, f, N.F32, cur)
        lim = 0 if f in (chain[0], chain[2]) else 4
        (assert_bit_exact if lim == 0 else lambda a, b, w: assert_ulp(a, b, lim, w))(nxt, orc.apply_fn(f, N.F32, cur), "op")
        cur = nxt
    assert_bit_exact(separate, cur, "fused == op by op")
    h = x.astype(np.float16)
    h[::97] = np.float16(60000.0)  # f16 slow lanes (|x| <= 65504 < 1e5 never leaves the fast path: still exact)
    assert_bit_exact(run_apply(dev, [lambda v: v.sin()], N.F16, h), run_apply(dev, lambda v: v.sin(), N.F16, h))


def test_kernel_cache_reuses_compiled_chain(raw_device):
    e1 = raw_device.compile(CHAIN8, N.F32)
    e2 = raw_device.compile(CHAIN8, N.F32)
    assert e1.handle.value == e2.handle.value


# ------------------------------------------------------------------ unary_grad (a2)
def test_unary_grad_int_kats(raw_device):
    dev = raw_device
    lhs = np.array([1, 2, 3, 4, 5, 6], np.int32)
    pl, pg, po = dev.upload(lhs), dev.upload(lhs), dev.upload(np.ones(6, np.int32))
    dev.unary_grad(dev.compile(lambda x: x.mul(2).add(1), N.I32, N.KERNEL_UNARY_GRAD), pl, pg, po, 6)
    assert dev.d2h(pg, 6, N.I32).tolist() == [4, 7, 10, 13, 16, 19]  # src/devices/cuda/ops.rs:261-275
    dev.h2d(pg, lhs)
    dev.unary_grad(dev.compile(lambda x: x.add(2), N.I32, N.KERNEL_UNARY_GRAD), pl, pg, po, 6)
    assert dev.d2h(pg, 6, N.I32).tolist() == [4, 6, 8, 10, 12, 14]  # src/devices/cuda/ops.rs:279-294
    for p in (pl, pg, po):
        dev.free(p)


@pytest.mark.parametrize("dt", FLOATS)
def test_unary_grad_is_mul_then_add(raw_device, dt):
    # lhs_grad += out_grad * g(lhs): two roundings, never an FMA (cpu_stack_ops.rs:28).
    dev, n = raw_device, 300_011
    lhs, og, lg = random_inputs(dt, n, 31), random_inputs(dt, n, 32), random_inputs(dt, n, 33)
    pl, pg, po = dev.upload(lhs), dev.upload(lg), dev.upload(og)
    for g in (lambda x: x.mul(2.0).add(1.0), lambda x: 2.0, lambda x: x.mul(x).mul(3.0), lambda x: x.neg()):
        dev.h2d(pg, lg)
        dev.unary_grad(dev.compile(g, dt, N.KERNEL_UNARY_GRAD), pl, pg, po, n)
        assert_bit_exact(dev.d2h(pg, n, dt), orc.add_unary_grad(g, dt, lhs, lg, og), f"unary_grad dtype {dt}")
    # transcendental g: isolate the mul/add from the function error by feeding the oracle the
    # device's own g(lhs)
    g = lambda x: x.cos()
    gdev = run_apply(dev, g, dt, lhs)
    assert_ulp(gdev, orc.apply_fn(g, dt, lhs), 4 if dt != N.F16 else 1, "cos in grad")
    dev.h2d(pg, lg)
    dev.unary_grad(dev.compile(g, dt, N.KERNEL_UNARY_GRAD), pl, pg, po, n)
    want = orc.add_unary_grad(lambda x: x.mul(1.0), dt, gdev, lg, og)  # lg + og * gdev
    assert_bit_exact(dev.d2h(pg, n, dt), want, "unary_grad(cos) composition")
    for p in (pl, pg, po):
        dev.free(p)


def test_chain8_backward_against_oracle(raw_device):
    # unary_ew per op + backward with seed ones: x.grad = prod_k g_k(x_k)  
    dev, n, dt = raw_device, 200_003, N.F32
    x = random_inputs(dt, n, 4, -4, 4)
    acts_dev, acts_orc = [x], [x]
    for f in CHAIN8:
        acts_dev.append(run_apply(dev, f, dt, acts_dev[-1]))
        acts_orc.append(orc.apply_fn(f, dt, acts_orc[-1]))
    grad_dev = dev.alloc(n * 4)
    dev.fill(dt, grad_dev, n, 1.0)
    grad_orc = np.ones(n, np.float32)
    for k in reversed(range(len(CHAIN8))):
        nxt = dev.alloc(n * 4)  # zeroed, like a fresh gradient buffer
        pl = dev.upload(acts_dev[k])
        dev.unary_grad(dev.compile(CHAIN8_GRADS[k], dt, N.KERNEL_UNARY_GRAD), pl, nxt, grad_dev, n)
        dev.free(pl)
        dev.free(grad_dev)
        grad_dev = nxt
        grad_orc = orc.add_unary_grad(CHAIN8_GRADS[k], dt, acts_orc[k], np.zeros(n, np.float32), grad_orc)
    got = dev.d2h(grad_dev, n, dt)
    dev.free(grad_dev)
    # the composed gradient has cancellations (2*sin(y)+1 near 0, 1 - tanh^2 near 0): the bar is a mixed
    # absolute/relative one; the per-op bars are checked in test_unary_grad_is_mul_then_add
    err = np.abs(got.astype(np.float64) - grad_orc.astype(np.float64))
    assert np.all(err <= 2e-5 + 1e-4 * np.abs(grad_orc.astype(np.float64))), float(err.max())
    assert float(np.median(err / np.maximum(np.abs(grad_orc.astype(np.float64)), 1e-12))) < 1e-6


# ------------------------------------------------------------------ clear / fill / copy (a6, a7)
@pytest.mark.parametrize("dt", ALL)
def test_clear_fill_copy(raw_device, dt):
    dev = raw_device
    for n in (1, 6, 1000, 70_001):
        x = random_inputs(dt, n, 41)
        p = dev.upload(x)
        q = dev.alloc(x.nbytes, zero=False)
        dev.copy(dt, q, 0, p, 0, n)
        assert_bit_exact(dev.d2h(q, n, dt), x, "copy")
        dev.clear(dt, p, n)
        assert not np.any(dev.d2h(p, n, dt).view(np.uint8)), "clear"
        dev.fill(dt, p, n, 1)
        assert np.all(dev.d2h(p, n, dt) == NP[dt](1)), "fill"
        dev.free(p)
        dev.free(q)


def test_copy_slice_kats(raw_device):
    # src/op_traits.rs:34-93 doc examples of copy_slice_to / copy_slice_all
    dev = raw_device
    src = dev.upload(np.array([1., 2., 3., 4., 5.], np.float32))
    dst = dev.alloc(6 * 4)
    dev.copy(N.F32, dst, 3, src, 1, 2)  # source 1..3 -> dest 3..5
    assert dev.d2h(dst, 6, N.F32).tolist() == [0., 0., 0., 2., 3., 0.]
    dev.clear(N.F32, dst, 6)
    dev.copy(N.F32, dst, 4, src, 2, 2)
    dev.copy(N.F32, dst, 0, src, 0, 2)
    assert dev.d2h(dst, 6, N.F32).tolist() == [1., 2., 0., 0., 3., 4.]
    dev.free(src)
    dev.free(dst)


def test_alloc_semantics(raw_device):
    dev = raw_device
    with pytest.raises(CustosError) as ei:
        dev.alloc(0)
    assert ei.value.code == N.CB_ERR_ZERO_LENGTH  # src/devices/cuda/api/cuda.rs:69-71
    p = dev.alloc(12345)
    assert p % 256 == 0
    assert not np.any(dev.d2h(p, 12345, N.U8))  # zeroed like the CPU device (cpu_ptr.rs:76-89)
    dev.free(p)
    a, hit = dev.cache_retrieve(1000, 4096)
    b, hit2 = dev.cache_retrieve(1000, 4096)
    assert (not hit) and hit2 and a == b  # src/modules/cached.rs:184-196


def test_h2d_d2h_roundtrip_large_pageable(raw_device):
    dev = raw_device
    x = np.random.default_rng(50).integers(0, 255, (40 << 20) + 123, dtype=np.uint8)  # > 2 staging buffers
    p = dev.upload(x)
    assert np.array_equal(dev.d2h(p, x.size, N.U8), x)
    dev.free(p)


# ------------------------------------------------------------------ sum / mean (a13)
@pytest.mark.parametrize("n", [1, 5, 1023, 1024, 1025, 20_000, (1 << 22) + 77])
def test_sum_f32_exact_order_and_accuracy(raw_device, n):
    dev = raw_device
    x = np.random.default_rng(5).random(n, dtype=np.float32)
    p = dev.upload(x)
    got = dev.sum(N.F32, p, n)
    plan = sum_plan(N.F32, n)
    want = orc.sum_two_pass(N.F32, x, plan["blocks"], plan["chunk"], plan["threads"], plan["vec"], plan["threads2"])
    assert got.view(np.uint32) == want.view(np.uint32), (got, want, plan)
    truth = orc.sum_f64(N.F32, x)
    assert abs(float(got) - truth) <= 1e-6 * abs(truth)
    mean = dev.mean(N.F32, p, n)
    assert mean == np.float32(got / np.float32(n))
    dev.free(p)


def test_sum_signed_and_other_dtypes(raw_device):
    dev = raw_device
    n = (1 << 20) + 3
    x = np.random.default_rng(6).uniform(-1, 1, n).astype(np.float32)
    p = dev.upload(x)
    got = dev.sum(N.F32, p, n)
    plan = sum_plan(N.F32, n)
    want = orc.sum_two_pass(N.F32, x, plan["blocks"], plan["chunk"], plan["threads"], plan["vec"], plan["threads2"])
    assert got.view(np.uint32) == want.view(np.uint32)
    assert abs(float(got) - orc.sum_f64(N.F32, x)) <= 1e-6 * float(np.sum(np.abs(x.astype(np.float64))))
    dev.free(p)
    for dt in (N.F64, N.F16):
        y = random_inputs(dt, n, 7, 0, 1)
        q = dev.upload(y)
        plan = sum_plan(dt, n)
        want = orc.sum_two_pass(dt, y, plan["blocks"], plan["chunk"], plan["threads"], plan["vec"], plan["threads2"])
        got = dev.sum(dt, q, n)
        assert got.tobytes() == want.tobytes(), (dt, got, want)
        dev.free(q)
    # tests/cuda/gpu_or_cpu.rs:34-64: i32 sum of 0..20000
    q = dev.upload(np.arange(20000, dtype=np.int32))
    assert dev.sum(N.I32, q, 20000) == 199_990_000
    dev.free(q)


def test_sum_unaligned_same_bits(raw_device):
    dev, n = raw_device, 300_001
    x = np.random.default_rng(8).random(n + 4, dtype=np.float32)
    p = dev.upload(x)
    q = dev.upload(x[1:])
    assert dev.sum(N.F32, p + 4, n).tobytes() == dev.sum(N.F32, q, n).tobytes()
    dev.free(p)
    dev.free(q)


def test_sum_is_run_to_run_deterministic(raw_device):
    dev, n = raw_device, (1 << 24) + 1
    x = np.random.default_rng(9).uniform(-1, 1, n).astype(np.float32)
    p = dev.upload(x)
    first = dev.sum(N.F32, p, n).tobytes()
    for _ in range(10):
        assert dev.sum(N.F32, p, n).tobytes() == first
    dev.free(p)


# ------------------------------------------------------------------ CUDA graph replay (K7)
def test_graph_capture_and_replay(raw_device):
    dev, n = raw_device, 4096
    x = random_inputs(N.F32, n, 60)
    exprs = [dev.compile(f, N.F32) for f in CHAIN8]  # compiled before capture
    a, b = dev.upload(x), dev.alloc(n * 4)
    dev.sync()
    before = dev.launches
    dev.graph_begin()
    src, dst = a, b
    for e in exprs:
        dev.apply(e, src, dst, n)
        src, dst = dst, src
    g = dev.graph_end()
    assert dev.graph_kernel_nodes(g) == 8
    assert dev.launches - before == 8  # captured, not executed
    dev.graph_launch(g)
    final = src
    got = dev.d2h(final, n, N.F32)
    # eager, same kernels
    cur = x
    for f in CHAIN8:
        cur = run_apply(dev, f, N.F32, cur)
    assert_bit_exact(got, cur, "graph replay vs eager")
    dev.graph_destroy(g)
    dev.free(a)
    dev.free(b)


# ------------------------------------------------------------------ host operands (row f3: H2D/D2H staging)
def test_apply_host_pipelined_and_pageable(raw_device):
    import ctypes
    dev = raw_device
    n = (40 << 20) // 4 + 12345  # several 16 MiB chunks plus a ragged tail
    x = random_inputs(N.F32, n, 77)
    want = orc.apply_chain(CHEAP8, orc.F32, x)
    e = dev.compile(CHEAP8, N.F32)
    # pinned: overlapped three-stream path
    h_in, h_out = dev.host_alloc(n * 4), dev.host_alloc(n * 4)
    a_in = np.ctypeslib.as_array((ctypes.c_float * n).from_address(h_in))
    a_out = np.ctypeslib.as_array((ctypes.c_float * n).from_address(h_out))
    a_in[:] = x
    before = dev.launches
    for _ in range(2):  # second call reuses ring slots that are still draining from the first
        a_out[:] = 0
        dev.apply_host(e, h_in, h_out, n)
        assert_bit_exact(a_out.copy(), want, "cb_apply_host (pinned)")
    assert dev.launches - before == 2 * 3  # one kernel per 16 MiB chunk
    dev.host_free(h_in)
    dev.host_free(h_out)
    # pageable: staged serial path, same results
    out = np.zeros(n, np.float32)
    dev.apply_host(e, x.ctypes.data, out.ctypes.data, n)
    assert_bit_exact(out, want, "cb_apply_host (pageable)")


def test_back_to_back_sums_overlap_without_losing_a_dependency(raw_device):
    """Consecutive sums are launched with programmatic stream serialisation: the streaming pass of a sum may start while
    the previous sum's last block is still folding.  Results must not change — also when a sum READS the scalar the
    previous one writes (row sums into a buffer, then the sum of that buffer), and when other kernels sit in between."""
    dev = raw_device
    rows, n = 64, 70_001
    rng = np.random.default_rng(77)
    data = rng.uniform(-1, 1, (rows, n)).astype(np.float32)
    p = dev.upload(data.reshape(-1))
    sums = dev.alloc(rows * 4)
    total = dev.alloc(64)
    plan = sum_plan(N.F32, n)
    want_rows = np.array([orc.sum_two_pass(orc.F32, data[r], plan["blocks"], plan["chunk"], plan["threads"], plan["vec"],
                                           plan["threads2"]) for r in range(rows)], np.float32)
    plan2 = sum_plan(N.F32, rows)
    want_total = orc.sum_two_pass(orc.F32, want_rows, plan2["blocks"], plan2["chunk"], plan2["threads"], plan2["vec"], plan2["threads2"])
    for rep in range(20):
        dev.clear(N.F32, sums, rows)
        for r in range(rows):  # 64 sums back to back, each into its own slot
            dev.sum_into(N.F32, p + 4 * r * n, n, sums + 4 * r)
        dev.sum_into(N.F32, sums, rows, total)  # reads what the sum just before it wrote
        got_rows, got_total = dev.d2h(sums, rows, N.F32), dev.d2h(total, 1, N.F32)[0]
        assert got_rows.tobytes() == want_rows.tobytes(), rep
        assert got_total.tobytes() == want_total.tobytes(), rep
    # a producer kernel between two sums: the second sum must see its output
    e = dev.compile(lambda v: v.mul(2.0), N.F32)
    q = dev.alloc(n * 4)
    for rep in range(10):
        dev.sum_into(N.F32, p, n, total)
        dev.apply(e, p + 4 * n * (rep % rows), q, n)
        dev.sum_into(N.F32, q, n, total)
        want = orc.sum_two_pass(orc.F32, (data[rep % rows] * np.float32(2)), plan["blocks"], plan["chunk"], plan["threads"], plan["vec"],
                                plan["threads2"])
        assert dev.d2h(total, 1, N.F32)[0].tobytes() == want.tobytes(), rep
    for ptr in (p, sums, total, q):
        dev.free(ptr)
