"""Host logic without a GPU: the expression IR, the reference-format source strings
(`to_cl_source`, `operations_to_fused_src`), the generated CUDA, and NVRTC for sm_100a."""
import numpy as np
import pytest

from custos_b200 import _native as N
from custos_b200 import CustosError
from custos_b200 import expr as E
from custos_b200.workloads import CHAIN8, CHAIN8_GRADS, CHEAP8, CONFIG1


def test_to_cl_source_kats():
    # src/two_way_ops/mod.rs:64-202 — the exact strings the reference asserts
    assert E.to_cl_source(lambda x: x.exp()) == "exp(x)"
    assert E.to_cl_source(lambda x: x.tan().neg(), marker_x="val") == "-(tan(val))"
    assert E.to_cl_source(lambda x, y: x.mul(3.).pow(y.add(1.)), n_args=2) == "pow((x * 3.0), (y + 1.0))"
    assert E.to_cl_source(lambda x, y: x.eq(y), N.I32, "var_x", "other", 2) == "(var_x == other)"
    assert E.to_cl_source(lambda x: x.geq(0).mul(x), N.I32, "var_x") == "((var_x >= 0) * var_x)"
    assert E.to_cl_source(lambda x: x.add(3), N.I32, "var_x") == "(var_x + 3)"
    assert E.to_cl_source(lambda x: x.geq(4), N.I32, "var_x") == "(var_x >= 4)"
    assert E.to_cl_source(lambda x, y: x.add(y), n_args=2) == "(x + y)"
    assert E.to_cl_source(lambda x, y: x.add(y).mul(3.6).sub(y), n_args=2) == "(((x + y) * 3.6) - y)"
    assert E.to_cl_source(lambda x: x.add(2.).mul(x).add(x.mul(8.)).mul(5.)) == "((((x + 2.0) * x) + (x * 8.0)) * 5.0)"
    # src/two_way_ops/resolve.rs:9-22 doc tests
    assert E.to_cl_source(lambda x: x.mul(x).add(2.)) == "((x * x) + 2.0)"
    assert E.to_cl_source(lambda x: x.add(x).mul(2.)) == "((x + x) * 2.0)"
    # src/devices/cuda/ops.rs:165,220: the markers the reference's CUDA kernels use
    assert E.to_cl_source(lambda x: x.add(1.0), marker_x="x[idx]") == "(x[idx] + 1.0)"
    assert E.to_cl_source(lambda x: x.mul(2).add(1), N.I32, "lhs[idx]") == "((lhs[idx] * 2) + 1)"


def test_remaining_op_format_strings():
    # src/two_way_ops/ops.rs and ops/unary.rs format strings
    assert E.to_cl_source(lambda x: x.sub(1.).div(2.)) == "((x - 1.0) / 2.0)"
    assert E.to_cl_source(lambda x: x.min(3.).max(5.)) == "max(min(x, 3.0), 5.0)"
    assert E.to_cl_source(lambda x: x.sin().cos().tanh().ln().abs()) == "abs(log(tanh(cos(sin(x)))))"
    assert E.to_cl_source(lambda x: x.identity().leq(2.)) == "(x <= 2.0)"


def test_rust_debug_float_formatting():
    # to_cl_source.rs:7-12 renders literals with `{:?}`
    cases = {0.5: "0.5", 2.0: "2.0", 3.6: "3.6", 1e16: "1e16", 1.5e-7: "1.5e-7", 0.0001: "0.0001", 1e-5: "1e-5",
             123456.0: "123456.0", -2.0: "-2.0", 1e15: "1000000000000000.0", float("inf"): "inf"}
    for v, s in cases.items():
        assert E.to_cl_source(lambda x, v=v: x.add(v), N.F64) == f"(x + {s})", v
    assert E.to_cl_source(lambda x: x.mul(0.1), N.F32) == "(x * 0.1)"       # shortest f32 repr, not 0.10000000149
    assert E.to_cl_source(lambda x: x.mul(0.1), N.F16) == "(x * 0.099975586)"  # half: Debug goes through f32


def test_operations_to_fused_src():
    # src/devices/fusing.rs:99-121
    assert E.ops_to_fused_src([lambda x: x.sin(), lambda x: x.neg(), lambda x: x.cos()]) == "x = sin(x);\nx = -(x);\nx = cos(x);\n"


def test_generated_cuda_uses_typed_literals():
    src = E.cuda_source([lambda x: x.mul(2.0).add(1.0)], N.F32)
    assert "__uint_as_float(0x40000000u)" in src and "2.0" in src  # exact f32 bits, decimal only in a comment
    assert "cb_mul(" in src and "cb_add(" in src
    assert "#define CB_DTYPE 0" in src and "#define CB_KIND 0" in src
    src16 = E.cuda_source([lambda x: x.mul(2.0)], N.F16)
    assert "((T)0x4000u)" in src16
    src_i = E.cuda_source([lambda x: x.add(3)], N.I32)
    assert "((T)0x00000003u)" in src_i


def test_shared_subtrees_are_shared_nodes():
    arr, n = E.flatten(lambda x: x.mul(x).add(x), N.F32)
    assert n == 3 and [arr[i].op for i in range(n)] == [E.OP["x"], E.OP["mul"], E.OP["add"]]


def test_invalid_programs_are_rejected():
    with pytest.raises(CustosError) as ei:
        E.compile_check([lambda x: x.sin()], N.I32)
    assert ei.value.code == N.CB_ERR_UNSUPPORTED
    with pytest.raises(CustosError) as ei:
        E.compile_check([lambda x: x.neg()], N.U32)
    assert ei.value.code == N.CB_ERR_UNSUPPORTED
    with pytest.raises(CustosError) as ei:
        E.compile_check([lambda x, y: x.add(y)], N.F32, N.KERNEL_APPLY, n_args=2)  # y in a unary kernel
    assert ei.value.code == N.CB_ERR_EXPR
    bad = (N.cb_node * 1)()
    bad[0].op, bad[0].a, bad[0].b = E.OP["add"], 0, 0
    buf = __import__("ctypes").create_string_buffer(64)
    assert N.load().cb_expr_to_cl_source(N.F32, bad, 1, b"x", b"y", buf, 64) == N.CB_ERR_EXPR


@pytest.mark.parametrize("dt", range(12))
def test_every_kernel_kind_compiles_for_sm_100a(dt):
    f = (lambda x: x.mul(2.0).add(1.0).sin().exp().ln().tanh().abs().neg()) if dt in E.FLOAT_DTYPES else (lambda x: x.mul(2).add(1))
    assert E.compile_check([f], dt, N.KERNEL_APPLY) > 1000
    assert E.compile_check([f], dt, N.KERNEL_UNARY_GRAD) > 1000
    assert E.compile_check([lambda x, y: x.mul(y).sub(x)], dt, N.KERNEL_BINARY, n_args=2) > 1000


def test_benchmark_chains_compile():
    for chain in (CHAIN8, CHEAP8, CONFIG1):
        assert E.compile_check(chain, N.F32) > 1000
    assert E.compile_check(CHAIN8, N.F16) > 1000
    for g in CHAIN8_GRADS:
        assert E.compile_check([g], N.F32, N.KERNEL_UNARY_GRAD) > 1000


def test_bool_is_storage_only():
    # bool is a CDatatype (cdatatype.rs:7-9) but not a Number: no expression can be built over it
    with pytest.raises(CustosError) as ei:
        E.compile_check([lambda x: x.add(1)], N.BOOL)
    assert ei.value.code == N.CB_ERR_UNSUPPORTED


def test_new_dtype_literals_and_sources():
    # half's Debug for bf16 prints the f32 value of the ROUNDED literal
    assert E.to_cl_source(lambda x: x.add(1.3).mul(2.0), N.BF16) == "((x + 1.296875) * 2.0)"
    assert E.to_cl_source(lambda x: x.add(-3), N.I8) == "(x + -3)"
    assert E.to_cl_source(lambda x: x.mul(65535), N.U16) == "(x * 65535)"
    assert E.to_cl_source(lambda x: x.add(18446744073709551615), N.U64) == "(x + 18446744073709551615)"
    src = E.cuda_source([lambda x: x.mul(1.5)], N.BF16)
    assert "cbw_mul_c(t0, (T)0x3fc0u)" in src and "((T)0x3fc0u)" in src
    with pytest.raises(CustosError):  # unsigned types have no Neg in Rust
        E.compile_check([lambda x: x.neg()], N.U16)
    assert E.compile_check([lambda x: x.neg()], N.I16) > 1000


def test_literals_round_to_the_dtype():
    arr, n = E.flatten(lambda x: x.mul(0.1), N.BF16)
    assert arr[1].fimm == 0.10009765625
    arr, n = E.flatten(lambda x: x.mul(0.1), N.F16)
    assert arr[1].fimm == float(np.float16(0.1))
    arr, n = E.flatten(lambda x: x.mul(0.1), N.F32)
    assert arr[1].fimm == float(np.float32(0.1))


@pytest.mark.parametrize("dt,np_t", [(N.F32, np.float32), (N.F64, np.float64)])
def test_literal_formatting_is_rusts_debug_on_random_values(dt, np_t):
    # Rust `{:?}` of a float (library/core/src/fmt/float.rs): the shortest digits that read back to the same value, in
    # decimal notation with at least one fractional digit when 1e-4 <= |v| < 1e16 (or v == 0), otherwise `d.ddde<exp>`.
    # NumPy's unique (Dragon4) formatting yields the same digits independently.
    rng = np.random.default_rng(21)
    span = 36 if np_t is np.float32 else 300
    vals = (rng.standard_normal(1500) * 10.0 ** rng.uniform(-span, span, 1500)).astype(np_t)
    vals = np.concatenate([vals[np.isfinite(vals) & (vals != 0)], np.array([1e-4, 9.999e-5, 1e16, 9.99e15, 1.0, 0.1, 123456.0], np_t)])
    for v in vals:
        src = E.to_cl_source(lambda x: x.add(v), dt)
        assert src.startswith("(x + ") and src.endswith(")")
        lit = src[5:-1]
        sci = np.format_float_scientific(v, unique=True, trim="-")
        digits, exp = sci.lstrip("-").split("e")
        digits, exp = digits.replace(".", "").rstrip("0") or "0", int(exp)
        if 1e-4 <= abs(float(v)) < 1e16:
            want = np.format_float_positional(v, unique=True, trim="0")
            want = want + "0" if want.endswith(".") else want
        else:
            want = ("-" if v < 0 else "") + digits[0] + ("." + digits[1:] if len(digits) > 1 else "") + f"e{exp}"
        assert lit == want, (float(v), lit, want)
        assert np_t(lit) == v  # and it reads back to the same value
