"""Golden fixtures (tests/golden/): the reference's own known-answer tests as data, and seeded
vectors produced by the pinned oracle.  CPU part: the oracle and the product's host logic against the
fixtures.  GPU part: the CUDA path against the same fixtures, without calling the oracle."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from custos_b200 import _native as N
from custos_b200 import expr as E
from custos_b200.expr import Combiner
from custos_b200.workloads import CHAIN8, CHAIN8_GRADS, CHEAP8, CONFIG1
from oracle import oracle as orc
from tests.helpers import assert_bit_exact

GOLDEN = Path(__file__).resolve().parent / "golden"
KATS = json.loads((GOLDEN / "reference_kats.json").read_text())
DT = {"f32": N.F32, "f64": N.F64, "f16": N.F16, "i32": N.I32}
NPDT = {"f32": np.float32, "f64": np.float64, "f16": np.float16, "i32": np.int32}


def closure(steps):
    """[["mul", 2.0], ["sin"], ["mul", "x"]] -> lambda x: x.mul(2.0).sin().mul(x)"""
    def f(x):
        cur = x
        for st in steps:
            name, args = st[0], st[1:]
            if name == "const":
                cur = Combiner._wrap(args[0])
            elif name == "add_xmul":
                cur = cur.add(x.mul(args[0]))
            else:
                cur = getattr(cur, name)(*[x if a == "x" else a for a in args])
        return cur
    return f


APPLY2 = {"pow_x3_y1": lambda x, y: x.mul(3.).pow(y.add(1.)), "eq": lambda x, y: x.eq(y),
          "add_mul36_sub": lambda x, y: x.add(y).mul(3.6).sub(y)}
LIBM = {"ln(abs(sin(x)))": lambda v: math.log(abs(math.sin(v))), "cos(sin(x))": lambda v: math.cos(math.sin(v)),
        "ln(cos(sin(x)))": lambda v: math.log(math.cos(math.sin(v))) if math.cos(math.sin(v)) > 0 else float("nan")}


def check(got, kat):
    want = np.array(kat["want"], dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    if kat.get("exact"):
        assert got.tolist() == want.tolist(), kat["src"]
    else:
        assert np.all(np.abs(got - want) <= kat["tol"]), (kat["src"], got, want)


# ------------------------------------------------------------------ CPU: oracle + host logic vs fixtures
@pytest.mark.parametrize("kat", KATS["apply"], ids=lambda k: k["src"])
def test_oracle_apply_kats(kat):
    check(orc.apply_fn(closure(kat["expr"]), DT[kat["dtype"]], np.array(kat["x"], NPDT[kat["dtype"]])), kat)


@pytest.mark.parametrize("kat", KATS["apply2"], ids=lambda k: k["src"])
def test_oracle_apply2_kats(kat):
    t = NPDT[kat["dtype"]]
    check(orc.apply2(APPLY2[kat["expr"]], DT[kat["dtype"]], np.array(kat["x"], t), np.array(kat["y"], t)), kat)


@pytest.mark.parametrize("kat", KATS["unary_grad"], ids=lambda k: k["src"])
def test_oracle_unary_grad_kats(kat):
    t = NPDT[kat["dtype"]]
    check(orc.add_unary_grad(closure(kat["grad"]), DT[kat["dtype"]], np.array(kat["lhs"], t), np.array(kat["lhs_grad"], t),
                             np.array(kat["out_grad"], t)), kat)


@pytest.mark.parametrize("kat", KATS["chain"], ids=lambda k: k["src"] + k["want_fn"])
def test_oracle_chain_kats(kat):
    t = NPDT[kat["dtype"]]
    got = orc.apply_chain([closure(op) for op in kat["ops"]], DT[kat["dtype"]], np.array(kat["x"], t))
    want = np.array([LIBM[kat["want_fn"]](v) for v in kat["x"]])
    if kat.get("exact_vs_libm"):
        assert got.tolist() == want.tolist()  # assert_eq! in the reference: same libm, same order
    else:
        ok = ~np.isnan(want)
        assert np.all(np.abs(got[ok] - want[ok]) < kat["tol"])


@pytest.mark.parametrize("kat", KATS["source"], ids=lambda k: k["src"])
def test_source_kats(kat):
    if "fused" in kat:
        assert E.ops_to_fused_src([closure(op) for op in kat["fused"]], DT[kat["dtype"]]) == kat["want"]
    else:
        assert E.to_cl_source(closure(kat["expr"]), DT[kat["dtype"]], marker_x=kat["marker"]) == kat["want"]


def test_oracle_reproduces_its_committed_vectors():
    g = np.load(GOLDEN / "oracle_vectors.npz")
    x = g["chain8_x_f32"]
    assert_bit_exact(orc.apply_chain(CHAIN8, orc.F32, x), g["chain8_y_f32"], "chain8 f32")
    assert_bit_exact(orc.apply_chain(CHEAP8, orc.F32, x), g["cheap8_y_f32"], "cheap8 f32")
    xh = g["chain8_x_f16"].view(np.float16)
    assert np.array_equal(orc.apply_chain(CHAIN8, orc.F16, xh).view(np.uint16), g["chain8_y_f16"])
    assert_bit_exact(orc.apply_chain(CONFIG1, orc.F32, g["config1_x_f32"]), g["config1_y_f32"], "config1")
    assert_bit_exact(orc.binary(0, orc.F32, g["binary_lhs_f32"], g["binary_rhs_f32"]), g["binary_add_f32"], "add")
    assert orc.sum_two_pass(orc.F32, g["sum_x_f32"], 4, 1024, 256, 4, 256) == g["sum_two_pass_f32"][0]
    assert orc.sum_seq(orc.F32, g["sum_x_f32"]) == g["sum_seq_f32"][0]


# ------------------------------------------------------------------ GPU: the CUDA path vs the same fixtures
def gpu_apply(dev, fs, dt, x, kind=N.KERNEL_APPLY):
    e = dev.compile(fs, dt)
    p, q = dev.upload(x), dev.alloc(x.nbytes)
    dev.apply(e, p, q, x.size)
    out = dev.d2h(q, x.size, dt)
    dev.free(p)
    dev.free(q)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("kat", KATS["apply"], ids=lambda k: k["src"])
def test_gpu_apply_kats(raw_device, kat):
    got = gpu_apply(raw_device, closure(kat["expr"]), DT[kat["dtype"]], np.array(kat["x"], NPDT[kat["dtype"]]))
    if kat.get("exact") and kat["src"].endswith("64-68"):  # exp(1) == E is exact only with glibc's expf; ours is within 2 ulp
        assert abs(float(got[0]) - kat["want"][0]) <= 2 * 2.4e-7
    else:
        check(got, kat)


@pytest.mark.gpu
@pytest.mark.parametrize("kat", KATS["apply2"], ids=lambda k: k["src"])
def test_gpu_apply2_kats(raw_device, kat):
    dev, t, dt = raw_device, NPDT[kat["dtype"]], DT[kat["dtype"]]
    x, y = np.array(kat["x"], t), np.array(kat["y"], t)
    e = dev.compile(APPLY2[kat["expr"]], dt, N.KERNEL_BINARY)
    px, py, po = dev.upload(x), dev.upload(y), dev.alloc(x.nbytes)
    dev.apply2(e, px, py, po, x.size)
    check(dev.d2h(po, x.size, dt), kat)
    for p in (px, py, po):
        dev.free(p)


@pytest.mark.gpu
@pytest.mark.parametrize("kat", KATS["unary_grad"], ids=lambda k: k["src"])
def test_gpu_unary_grad_kats(raw_device, kat):
    dev, t, dt = raw_device, NPDT[kat["dtype"]], DT[kat["dtype"]]
    pl, pg, po = dev.upload(np.array(kat["lhs"], t)), dev.upload(np.array(kat["lhs_grad"], t)), dev.upload(np.array(kat["out_grad"], t))
    dev.unary_grad(dev.compile(closure(kat["grad"]), dt, N.KERNEL_UNARY_GRAD), pl, pg, po, len(kat["lhs"]))
    check(dev.d2h(pg, len(kat["lhs"]), dt), kat)
    for p in (pl, pg, po):
        dev.free(p)


@pytest.mark.gpu
def test_gpu_against_committed_vectors(raw_device):
    dev = raw_device
    g = np.load(GOLDEN / "oracle_vectors.npz")
    x = g["chain8_x_f32"]
    # bit-exact class
    assert_bit_exact(gpu_apply(dev, CHEAP8, N.F32, x), g["cheap8_y_f32"], "cheap8 f32 vs golden")
    xh = g["chain8_x_f16"].view(np.float16)
    assert np.array_equal(gpu_apply(dev, CHEAP8, N.F16, xh).view(np.uint16), g["cheap8_y_f16"])
    for op, key in ((N.BIN_ADD, "binary_add_f32"), (N.BIN_MUL, "binary_mul_f32")):
        a, b = dev.upload(g["binary_lhs_f32"]), dev.upload(g["binary_rhs_f32"])
        o = dev.alloc(x.nbytes)
        dev.binary(N.F32, op, a, b, o, x.size)
        assert_bit_exact(dev.d2h(o, x.size, N.F32), g[key], key)
        for p in (a, b, o):
            dev.free(p)
    # transcendental chains: absolute bars (outputs of CHAIN8 live in (-1, 1), CONFIG1 in [-1, 3])
    y = gpu_apply(dev, CHAIN8, N.F32, x)
    assert float(np.max(np.abs(y.astype(np.float64) - g["chain8_y_f32"]))) < 1e-5
    assert float(np.mean(y == g["chain8_y_f32"])) > 0.6  # most results are bit-identical to the CPU device
    yh = gpu_apply(dev, CHAIN8, N.F16, xh)
    assert float(np.max(np.abs(yh.astype(np.float64) - g["chain8_y_f16"].view(np.float16).astype(np.float64)))) < 4e-3
    assert float(np.mean(yh.view(np.uint16) == g["chain8_y_f16"])) > 0.97
    y1 = gpu_apply(dev, CONFIG1, N.F32, g["config1_x_f32"])
    assert float(np.max(np.abs(y1.astype(np.float64) - g["config1_y_f32"]))) < 2e-6
    # backward of CHAIN8 (seed ones) through eight unary_grad launches over the device's own activations
    n = x.size
    acts = [x]
    for f in CHAIN8:
        acts.append(gpu_apply(dev, f, N.F32, acts[-1]))
    grad = dev.alloc(n * 4)
    dev.fill(N.F32, grad, n, 1.0)
    for k in reversed(range(8)):
        nxt, pl = dev.alloc(n * 4), dev.upload(acts[k])
        dev.unary_grad(dev.compile(CHAIN8_GRADS[k], N.F32, N.KERNEL_UNARY_GRAD), pl, nxt, grad, n)
        dev.free(pl)
        dev.free(grad)
        grad = nxt
    got = dev.d2h(grad, n, N.F32).astype(np.float64)
    dev.free(grad)
    want = g["chain8_grad_f32"].astype(np.float64)
    assert np.all(np.abs(got - want) <= 2e-5 + 1e-4 * np.abs(want))
    # sum: the 4-block plan of the fixture is the device's plan for 4096 elements
    from custos_b200.raw import sum_plan
    plan = sum_plan(N.F32, 4096)
    assert (plan["blocks"], plan["chunk"]) == (4, 1024)
    p = dev.upload(g["sum_x_f32"])
    assert dev.sum(N.F32, p, 4096) == g["sum_two_pass_f32"][0]
    assert abs(float(dev.sum(N.F32, p, 4096)) - g["sum_f64"][0]) <= 1e-6 * g["sum_f64"][0]
    dev.free(p)


# ------------------------------------------------------------------ the dtypes added later (second fixture file)
def _dtype_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", GOLDEN / "make_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.INT_EXPR, mod.INT_TYPES


def test_oracle_reproduces_the_dtype_vectors():
    g = np.load(GOLDEN / "oracle_vectors_dtypes.npz")
    xb, yb = g["x_bf16"], g["y_bf16"]
    assert np.array_equal(orc.apply_chain(CHAIN8, orc.BF16, xb), g["chain8_y_bf16"])
    assert np.array_equal(orc.apply_chain(CHEAP8, orc.BF16, xb), g["cheap8_y_bf16"])
    for k, name in enumerate(("add", "mul", "sub", "div")):
        assert np.array_equal(orc.binary(k, orc.BF16, xb, yb), g[f"binary_{name}_bf16"])
    assert np.array_equal(orc.apply2(lambda a, b: a.max(b).min(a.mul(b)), orc.BF16, xb, yb), g["max_min_bf16"])
    assert g["max_ties_f16"].tolist() == [0x0000, 0x8000, 0x8000, 0x0000]  # half's f16::max keeps `self` on +0 / -0 ties
    expr, types = _dtype_cases()
    for name, (dt, t) in types.items():
        assert np.array_equal(orc.apply_fn(expr, dt, g[f"x_{name}"]), g[f"expr_y_{name}"]), name
        assert orc.sum_seq(dt, g[f"x_{name}"]) == g[f"sum_{name}"][0]


@pytest.mark.gpu
def test_gpu_against_the_dtype_vectors(raw_device):
    from tests.helpers import assert_bf16_bit_exact, bf16_ulp_distance
    dev = raw_device
    g = np.load(GOLDEN / "oracle_vectors_dtypes.npz")
    xb, yb = g["x_bf16"], g["y_bf16"]
    assert_bf16_bit_exact(gpu_apply(dev, CHEAP8, N.BF16, xb), g["cheap8_y_bf16"], "cheap8 bf16 vs golden")
    pa, pb, po = dev.upload(xb), dev.upload(yb), dev.alloc(xb.nbytes)
    for k, name in enumerate(("add", "mul", "sub", "div")):
        dev.binary(N.BF16, k, pa, pb, po, xb.size)
        assert_bf16_bit_exact(dev.d2h(po, xb.size, N.BF16), g[f"binary_{name}_bf16"], f"bf16 {name} vs golden")
    dev.apply2(dev.compile(lambda a, b: a.max(b).min(a.mul(b)), N.BF16, N.KERNEL_BINARY), pa, pb, po, xb.size)
    assert_bf16_bit_exact(dev.d2h(po, xb.size, N.BF16), g["max_min_bf16"], "bf16 max/min vs golden")
    for p in (pa, pb, po):
        dev.free(p)
    y = gpu_apply(dev, CHAIN8, N.BF16, xb)
    assert float(np.mean(y == g["chain8_y_bf16"])) > 0.97  # composition of three transcendentals, each <= 1 ulp(bf16)
    assert float(np.max(bf16_ulp_distance(y, g["chain8_y_bf16"]))) <= 16
    zeros = np.array([0.0, -0.0, -0.0, 0.0], np.float16)
    pz, pr, pq = dev.upload(zeros), dev.upload(zeros[::-1].copy()), dev.alloc(8)
    dev.apply2(dev.compile(lambda a, b: a.max(b), N.F16, N.KERNEL_BINARY), pz, pr, pq, 4)
    assert dev.d2h(pq, 4, N.F16).view(np.uint16).tolist() == g["max_ties_f16"].tolist()
    for p in (pz, pr, pq):
        dev.free(p)
    expr, types = _dtype_cases()
    for name, (dt, t) in types.items():
        x = g[f"x_{name}"]
        assert np.array_equal(gpu_apply(dev, expr, dt, x), g[f"expr_y_{name}"]), name
        p = dev.upload(x)
        assert dev.sum(dt, p, x.size) == g[f"sum_{name}"][0]
        dev.free(p)


# ------------------------------------------------------------------ f16 / bf16 min / max over {+-0, +-1, +-inf, NaN}
def _minmax_cases():
    import json
    kats = json.loads((GOLDEN / "half_minmax_kats.json").read_text())
    for ty, dt in (("f16", orc.F16), ("bf16", orc.BF16)):
        a = np.array([r["self_bits"] for r in kats[ty]], np.uint16)
        b = np.array([r["rhs_bits"] for r in kats[ty]], np.uint16)
        mx = np.array([r["max_bits"] for r in kats[ty]], np.uint16)
        mn = np.array([r["min_bits"] for r in kats[ty]], np.uint16)
        yield ty, dt, a, b, mx, mn


def _as_dtype(bits, dt):
    return bits.view(np.float16) if dt == orc.F16 else bits


def test_oracle_half_min_max_known_answers():
    """src/number.rs:507-510 and :536-539 (max -> half's inherent max) and :207-209 (min): 49 operand pairs each;
    the expected bits come from the reference text (tests/golden/make_golden.py: half_minmax_kats), not the oracle."""
    for ty, dt, a, b, mx, mn in _minmax_cases():
        got_max = orc.apply2(lambda p, q: p.max(q), dt, _as_dtype(a, dt), _as_dtype(b, dt)).view(np.uint16)
        got_min = orc.apply2(lambda p, q: p.min(q), dt, _as_dtype(a, dt), _as_dtype(b, dt)).view(np.uint16)
        assert got_max.tolist() == mx.tolist(), ty
        assert got_min.tolist() == mn.tolist(), ty


@pytest.mark.gpu
def test_gpu_half_min_max_known_answers(raw_device):
    dev = raw_device
    for ty, dt, a, b, mx, mn in _minmax_cases():
        ndt = N.F16 if dt == orc.F16 else N.BF16
        pa, pb, po = dev.upload(_as_dtype(a, dt)), dev.upload(_as_dtype(b, dt)), dev.alloc(a.nbytes)
        for f, want in ((lambda p, q: p.max(q), mx), (lambda p, q: p.min(q), mn)):
            dev.apply2(dev.compile(f, ndt, N.KERNEL_BINARY), pa, pb, po, a.size)
            got = np.asarray(dev.d2h(po, a.size, ndt)).view(np.uint16)
            assert got.tolist() == want.tolist(), ty
        for p in (pa, pb, po):
            dev.free(p)
