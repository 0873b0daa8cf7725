"""Pins the CPU oracle with the reference's own known-answer tests.

Every case cites the reference test it restates.  The oracle is the checker of the CUDA
path (tests/test_gpu_*.py); this file is what makes it trustworthy.
"""
import math

import numpy as np
import pytest

from oracle import oracle as orc

F32, F64, F16, I32 = orc.F32, orc.F64, orc.F16, orc.I32


def test_exp_of_one_is_e_bit_exact():
    # src/two_way_ops/mod.rs:64-68: assert_eq!(res, core::f32::consts::E)
    res = orc.eval_scalar(lambda x: x.exp(), F32, 1.0)
    assert res.view(np.uint32) == np.uint32(0x402DF854)


def test_neg_tan():
    # src/two_way_ops/mod.rs:79-85
    res = orc.eval_scalar(lambda x: x.tan().neg(), F32, 2.0)
    assert abs(float(res) - 2.1850398) < 1e-6
    assert res == np.float32(-math.tan(2.0))


def test_pow():
    # src/two_way_ops/mod.rs:96-101: (3*3)^(2+1) == 729
    assert orc.eval_scalar(lambda x, y: x.mul(3.).pow(y.add(1.)), F32, 3.0, 2.0) == np.float32(729.0)


def test_int_comparisons_and_arithmetic():
    # src/two_way_ops/mod.rs:107-168
    assert orc.eval_scalar(lambda x, y: x.eq(y), I32, 3, 3) == 1
    assert orc.eval_scalar(lambda x: x.geq(0).mul(x), I32, 3) == 3
    assert orc.eval_scalar(lambda x: x.add(3), I32, 3) == 6
    assert orc.eval_scalar(lambda x: x.geq(4), I32, 3) == 0
    assert orc.eval_scalar(lambda x: x.add(2).add(x.mul(8)), I32, 4) == 38


def test_eq_evaluates_le_like_the_reference():
    # src/two_way_ops/ops/cmps.rs:135: Eq::eval calls `.le()`
    assert orc.eval_scalar(lambda x, y: x.eq(y), I32, 2, 3) == 1
    assert orc.eval_scalar(lambda x, y: x.eq(y), I32, 4, 3) == 0


def test_two_arg_expression():
    # src/two_way_ops/mod.rs:181-190: ((x + y) * 3.6) - y at (4, 3) ~ 22.2
    res = orc.eval_scalar(lambda x, y: x.add(y).mul(3.6).sub(y), F32, 4.0, 3.0)
    assert res == np.float32(np.float32(np.float32(7.0) * np.float32(3.6)) - np.float32(3.0))
    assert abs(float(res) - 22.2) < 1e-5


def test_clip_min_max():
    # src/two_way_ops/mod.rs:206-219
    out = orc.apply_fn(lambda x: x.max(3.).min(5.), F64, [1., 3., 4., 6., 3., 2.])
    assert out.tolist() == [3., 3., 4., 5., 3., 3.]


def test_min_max_follow_number_trait_not_fmin():
    # src/number.rs:202-209: `if self < rhs {self} else {rhs}` — NaN in the lhs yields rhs
    nan = float("nan")
    assert orc.eval_scalar(lambda x, y: x.min(y), F32, nan, 1.0) == np.float32(1.0)
    assert math.isnan(orc.eval_scalar(lambda x, y: x.min(y), F32, 1.0, nan))
    assert math.isnan(orc.eval_scalar(lambda x, y: x.max(y), F32, 1.0, nan))
    # -0.0 < 0.0 is false -> rhs
    assert math.copysign(1.0, float(orc.eval_scalar(lambda x, y: x.min(y), F32, -0.0, 0.0))) == 1.0


def test_apply_fn_int_add():
    # src/two_way_ops/mod.rs:237-246
    assert orc.apply_fn(lambda x: x.add(3), I32, [3, 3, 4, 5, 3, 2]).tolist() == [6, 6, 7, 8, 6, 5]


def test_apply_fn_more_complex():
    # src/two_way_ops/mod.rs:265-283 (f64 on the CPU device)
    out = orc.apply_fn(lambda x: x.mul(2.).add(4.).sin().mul(x).add(1.), F64, [3., 3., 4., 5., 3., 2.])
    want = [-0.6320633326681093, -0.6320633326681093, -1.1462916720017398, 5.953036778474352,
            -0.6320633326681093, 2.978716493246764]
    np.testing.assert_allclose(out, want, rtol=0, atol=1e-15)


def test_apply_fn_more_complex_f32():
    # src/two_way_ops/mod.rs:320-333 (f32 values quoted for the Vulkan variant)
    out = orc.apply_fn(lambda x: x.mul(2.).add(4.).sin().mul(x).add(1.), F32, [3., 3., 4., 5., 3., 2.])
    want = np.array([-0.632_063_3, -0.632_063_3, -1.146_291_6, 5.953_037, -0.632_063_3, 2.978_716_6], np.float32)
    np.testing.assert_allclose(out, want, rtol=0, atol=1e-6)


def test_chained_eval_ln_cos_sin():
    # src/two_way_ops/mod.rs:360-388: val_out == (3.4f32).ln().cos().sin()
    out = orc.apply_chain([lambda x: x.ln(), lambda x: x.cos(), lambda x: x.sin()], F32, [3.4])
    want = np.sin(np.cos(np.log(np.float32(3.4))))
    assert abs(float(out[0]) - float(want)) <= 2e-7
    # and exactly the composition of the oracle's own scalar functions
    a = orc.eval_scalar(lambda x: x.ln(), F32, 3.4)
    b = orc.eval_scalar(lambda x: x.cos(), F32, a)
    c = orc.eval_scalar(lambda x: x.sin(), F32, b)
    assert out[0] == c


def test_unary_ew_sin_and_its_gradient():
    # src/unary.rs:141-200: sin([1,2,3,4]) and grad = cos
    x = [1., 2., 3., 4.]
    out = orc.apply_fn(lambda x: x.sin(), F64, x)
    np.testing.assert_allclose(out, [0.8414709848078965, 0.9092974268256817, 0.1411200080598672,
                                     -0.7568024953079282], rtol=0, atol=1e-16)
    grad = orc.add_unary_grad(lambda x: x.cos(), F64, x, np.zeros(4), np.ones(4))
    np.testing.assert_allclose(grad, [0.5403023058681398, -0.4161468365471424, -0.9899924966004454,
                                      -0.6536436208636119], rtol=0, atol=1e-16)


def test_gradient_accumulates_over_backward_passes():
    # src/unary.rs:277-300 (run_several_times!): after i backward passes grad == i * cos(x)
    x = np.array([1., 2., 3., 4.])
    g = np.zeros(4)
    for i in range(1, 10):
        g = orc.add_unary_grad(lambda x: x.cos(), F64, x, g, np.ones(4))
        np.testing.assert_allclose(g, i * np.cos(x), rtol=0, atol=1e-14)


def test_doc_examples_apply_fn_and_unary_grad():
    # src/unary.rs:12-18 and :36-47
    assert orc.apply_fn(lambda x: x.mul(2.), F64, [1., 2., 3., 3., 2., 1.]).tolist() == [2., 4., 6., 6., 4., 2.]
    g = orc.add_unary_grad(lambda x: 2.0, F64, [1., 2., 3., 3., 2., 1.], np.zeros(6), np.ones(6))
    assert g.tolist() == [2.] * 6


def test_cuda_kats_of_the_reference():
    # src/devices/cuda/ops.rs:252-258: x + 1.0 on [1..6]
    assert orc.apply_fn(lambda x: x.add(1.0), F32, [1, 2, 3, 4, 5, 6]).tolist() == [2, 3, 4, 5, 6, 7]
    # src/devices/cuda/ops.rs:261-275: i32, grad fn 2x+1
    g = orc.add_unary_grad(lambda x: x.mul(2).add(1), I32, [1, 2, 3, 4, 5, 6], [1, 2, 3, 4, 5, 6], [1] * 6)
    assert g.tolist() == [4, 7, 10, 13, 16, 19]
    # src/devices/cuda/ops.rs:279-294: lazy variant, grad fn x+2
    g = orc.add_unary_grad(lambda x: x.add(2), I32, [1, 2, 3, 4, 5, 6], [1, 2, 3, 4, 5, 6], [1] * 6)
    assert g.tolist() == [4, 6, 8, 10, 12, 14]


def test_fused_sin_cos_ln():
    # src/op_hint.rs:121-140: within 1e-3 of buf.sin().cos().ln()
    x = np.array([1., 2., 3., 4., 5.], np.float32)
    out = orc.apply_chain([lambda x: x.sin(), lambda x: x.cos(), lambda x: x.ln()], F32, x)
    with np.errstate(invalid="ignore"):
        want = np.log(np.cos(np.sin(x)))
    np.testing.assert_allclose(out, want, atol=1e-3)


def test_fused_complex_case_is_exact():
    # src/op_hint.rs:225-251: two independent chains, assert_eq! on the CPU
    buf = np.array([1., 2., 3., 4., 5.])
    rhs = np.array([8., 2., 3., 4., 5.])
    out = orc.apply_chain([lambda x: x.sin(), lambda x: x.abs(), lambda x: x.ln()], F64, buf)
    assert out.tolist() == [math.log(abs(math.sin(v))) for v in buf]
    out2 = orc.apply_chain([lambda x: x.sin(), lambda x: x.cos()], F64, rhs)
    assert out2.tolist() == [math.cos(math.sin(v)) for v in rhs]


def test_binary_demo_impl():
    # tests/demo_impl/cuda/mod.rs:40-66: 655 360 x (1 + 4) == 5
    n = 655_360
    out = orc.binary(0, F32, np.full(n, 1, np.float32), np.full(n, 4, np.float32))
    assert np.all(out == 5.0)
    # README.md:96-122 MulBuf
    assert orc.binary(1, F32, [1, 2, 3], [4, 5, 6]).tolist() == [4, 10, 18]


def test_clear():
    # src/devices/cuda/ops.rs:244-249, tests/clear.rs:31-43
    buf = np.array([1, 2, 3, 4, 5, 6], np.uint32)
    assert orc.clear(orc.U32, buf).tolist() == [0] * 6


def test_int_sum_of_range():
    # tests/cuda/gpu_or_cpu.rs:34-64: sum of 0..20000 (i32) = 199 990 000
    assert orc.sum_seq(I32, np.arange(20000, dtype=np.int32)) == 199_990_000


# ------------------------------------------------------------------ f16 (parity unpinned in the reference)
def test_f16_conversions_are_ieee_rne():
    # independent check against NumPy's binary16 (IEEE round-to-nearest-even), every half value
    allh = np.arange(65536, dtype=np.uint16)
    f = allh.view(np.float16).astype(np.float32)
    back = np.array([orc.f32_to_f16_bits(float(v)) for v in f[::7]], dtype=np.uint16)
    ref = allh[::7]
    nan = np.isnan(f[::7])
    assert np.array_equal(back[~nan], ref[~nan])
    got = np.array([orc.f16_bits_to_f32(int(h)) for h in allh[::7]], dtype=np.float32)
    assert np.array_equal(got[~nan], f[::7][~nan])
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.standard_normal(20000).astype(np.float32) * 100, rng.uniform(-7e-5, 7e-5, 20000).astype(np.float32),
                           np.array([65504, 65519.99, 65520, 1e9, -1e9, 5.96e-8, 2.98e-8, 2.9802322e-8, 0.0, -0.0], np.float32)])
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    got = np.array([orc.f32_to_f16_bits(float(v)) for v in vals], dtype=np.uint16)
    assert np.array_equal(got, want)


def test_f16_rounds_after_every_op():
    # src/number.rs:543-608: each op goes f16 -> f32 -> f16
    x = np.array([0.1, 1.5, 3.25, -2.0], np.float16)
    out = orc.apply_fn(lambda v: v.mul(3.0).add(0.1).exp(), F16, x)
    want = np.exp(((x * np.float16(3.0)).astype(np.float16) + np.float16(0.1)).astype(np.float16).astype(np.float32)).astype(np.float16)
    assert np.array_equal(out.view(np.uint16), want.view(np.uint16))


def test_f16_tan_is_cos_like_the_reference():
    # src/number.rs:575-577: `fn tan(&self) -> Self { Self::from_f32(self.to_f32().cos()) }`
    x = np.array([0.5, 1.0, 2.0], np.float16)
    assert np.array_equal(orc.apply_fn(lambda v: v.tan(), F16, x), orc.apply_fn(lambda v: v.cos(), F16, x))


def test_f16_max_is_halfs_inherent_max():
    # src/number.rs:507-510 (f16) and :536-539 (bf16): Number::max forwards to half's inherent max, which keeps `self`
    # unless `other > self`; the trait default (f32, f64; number.rs:202-204) returns `rhs` unless `self > rhs`.
    # The two only differ for equal values with different bits (+0 / -0) and for NaN operands.
    pz, nz = np.float16(0.0), np.float16(-0.0)
    assert orc.eval_scalar(lambda x, y: x.max(y), F16, pz, nz).view(np.uint16) == 0x0000
    assert orc.eval_scalar(lambda x, y: x.max(y), F16, nz, pz).view(np.uint16) == 0x8000
    assert orc.eval_scalar(lambda x, y: x.max(y), F32, 0.0, -0.0).view(np.uint32) == 0x80000000
    assert orc.eval_scalar(lambda x, y: x.max(y), orc.BF16, 0x0000, 0x8000) == 0x0000
    assert orc.eval_scalar(lambda x, y: x.max(y), orc.BF16, 0x8000, 0x0000) == 0x8000
    assert orc.eval_scalar(lambda x, y: x.min(y), F16, pz, nz).view(np.uint16) == 0x8000  # default min: rhs


# ------------------------------------------------------------------ bf16 (parity unpinned in the reference)
def test_bf16_conversions_are_rne_with_halfs_nan_rule():
    cases = [(1.0, 0x3F80), (-2.0, 0xC000), (0.0, 0x0000), (-0.0, 0x8000), (float("inf"), 0x7F80), (3.4e38, 0x7F80),
             (3.3895314e38, 0x7F7F), (1e-40, 0x0001)]
    for v, bits in cases:
        assert orc.f32_to_bf16_bits(v) == bits, v
    f = lambda u: float(np.array([u], np.uint32).view(np.float32)[0])
    assert orc.f32_to_bf16_bits(f(0x3F808000)) == 0x3F80   # tie -> even (down)
    assert orc.f32_to_bf16_bits(f(0x3F818000)) == 0x3F82   # tie -> even (up)
    assert orc.f32_to_bf16_bits(f(0x3F808001)) == 0x3F81   # above the tie
    assert orc.f32_to_bf16_bits(f(0x7F7FFFFF)) == 0x7F80   # f32::MAX rounds to inf
    assert orc.f32_to_bf16_bits(float("nan")) & 0x7FC0 == 0x7FC0
    assert orc.bf16_bits_to_f32(0x3FC0) == 1.5
    # the package's host helper (used to prepare bf16 inputs) agrees with the oracle on random values
    from custos_b200.expr import bf16_from_f32, bf16_to_f32
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.standard_normal(5000).astype(np.float32) * 1e3, rng.uniform(-1e-38, 1e-38, 2000).astype(np.float32)])
    want = np.array([orc.f32_to_bf16_bits(float(v)) for v in vals], np.uint16)
    assert np.array_equal(bf16_from_f32(vals), want)
    assert np.array_equal(bf16_to_f32(want), np.array([orc.bf16_bits_to_f32(int(b)) for b in want], np.float32))


def test_bf16_rounds_after_every_op_and_tan_is_cos():
    # src/number.rs:611-676: each op goes bf16 -> f32 -> bf16; `tan` calls cos (:643-645)
    from custos_b200.expr import bf16_from_f32, bf16_to_f32
    x = bf16_from_f32(np.array([0.1, 1.5, 3.25, -2.0], np.float32))
    out = orc.apply_fn(lambda v: v.mul(3.0).add(0.1).exp(), orc.BF16, x)
    r = lambda a: bf16_to_f32(bf16_from_f32(a))
    step = r(r(bf16_to_f32(x) * np.float32(3.0)) + r(np.float32([0.1]))[0])
    assert np.array_equal(out, bf16_from_f32(np.exp(step)))
    assert np.array_equal(orc.apply_fn(lambda v: v.tan(), orc.BF16, x), orc.apply_fn(lambda v: v.cos(), orc.BF16, x))


def test_narrow_and_wide_integers_wrap():
    # release-mode Rust arithmetic wraps; i8 / i16 / u16 / u64 are CDatatypes (cdatatype.rs:28-51)
    assert orc.eval_scalar(lambda x: x.add(100), orc.I8, 100) == np.int8(-56)
    assert orc.eval_scalar(lambda x: x.mul(3), orc.I16, 20000) == np.int16(60000 - 65536)
    assert orc.eval_scalar(lambda x: x.sub(1), orc.U16, 0) == 65535
    assert orc.eval_scalar(lambda x: x.add(2), orc.U64, 2 ** 64 - 1) == 1
    assert orc.eval_scalar(lambda x: x.neg(), orc.I8, -128) == np.int8(-128)
    assert orc.eval_scalar(lambda x: x.div(-1), orc.I8, -128) == np.int8(-128)   # Rust panics; defined as wrap
    assert orc.eval_scalar(lambda x: x.div(0), orc.I16, 5) == 0                  # Rust panics; defined as 0
    assert orc.sum_seq(orc.I8, np.full(1000, -128, np.int8)) == -128000
    assert orc.sum_seq(orc.U16, np.full(1000, 65535, np.uint16)) == 65_535_000
    with pytest.raises(orc.OracleError):
        orc.apply_fn(lambda x: x.neg(), orc.U16, [1])
    with pytest.raises(orc.OracleError):
        orc.apply_fn(lambda x: x.add(1), orc.BOOL, [True])


# ------------------------------------------------------------------ sums (order defined by this project)
def test_two_pass_sum_matches_f64_ground_truth():
    rng = np.random.default_rng(5)
    x = rng.random(1 << 18, dtype=np.float32)
    two = orc.sum_two_pass(F32, x, 64, 4096, 256, 4, 256)
    truth = orc.sum_f64(F32, x)
    assert abs(float(two) - truth) / truth < 1e-6
    assert abs(float(orc.sum_seq(F32, x)) - truth) / truth < 1e-4


def test_unsupported_int_ops_are_rejected():
    # Float-bounded ops do not exist for integers in the reference (ops/unary.rs: `T: Float`)
    with pytest.raises(orc.OracleError):
        orc.apply_fn(lambda x: x.sin(), I32, [1, 2])


def test_the_boxed_and_the_unfused_cpu_baselines_give_the_bits_of_the_fused_interpreter():
    # src/devices/cpu/cpu_device.rs:217-229 + src/op_hint.rs:30-33 (a boxed dyn op per element per op) and
    # src/devices/cpu_stack_ops.rs:7-15 (one loop per op): different costs, same arithmetic
    from custos_b200.workloads import CHAIN8, CONFIG1
    from tests.helpers import edge_values
    x = np.concatenate([np.random.default_rng(4).uniform(-4, 4, 5000).astype(np.float32), edge_values(np.float32)])
    for chain in (CHAIN8, CONFIG1, CHAIN8[:1]):
        want = orc.apply_chain(chain, F32, x).view(np.uint32)
        assert np.array_equal(orc.apply_chain_boxed(chain, x).view(np.uint32), want)
        assert np.array_equal(orc.apply_chain_unfused(chain, x).view(np.uint32), want)
