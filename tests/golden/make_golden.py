#!/usr/bin/env python
"""Regenerates tests/golden/oracle_vectors.npz: seeded inputs and the oracle's outputs for the
benchmark chains.  The reference (a Rust crate) cannot be executed in this image, so these vectors are
produced by the C restatement in oracle/ AFTER it has been pinned by the reference's own known-answer
tests (tests/golden/reference_kats.json, tests/test_oracle_kat.py).  They serve two purposes: the GPU
tests can compare against committed vectors, and a silent change of the oracle (or of the libm it links)
shows up as a diff of this file.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from custos_b200.workloads import CHAIN8, CHAIN8_GRADS, CHEAP8, CONFIG1  # noqa: E402
from oracle import oracle as orc  # noqa: E402

N = 4096


def main():
    out = {}
    x = np.random.default_rng(4).uniform(-4, 4, N).astype(np.float32)
    out["chain8_x_f32"] = x
    out["chain8_y_f32"] = orc.apply_chain(CHAIN8, orc.F32, x)
    out["cheap8_y_f32"] = orc.apply_chain(CHEAP8, orc.F32, x)
    xh = x.astype(np.float16)
    out["chain8_x_f16"] = xh.view(np.uint16)
    out["chain8_y_f16"] = orc.apply_chain(CHAIN8, orc.F16, xh).view(np.uint16)
    out["cheap8_y_f16"] = orc.apply_chain(CHEAP8, orc.F16, xh).view(np.uint16)
    x1 = np.random.default_rng(1).uniform(-2, 2, N).astype(np.float32)
    out["config1_x_f32"] = x1
    out["config1_y_f32"] = orc.apply_chain(CONFIG1, orc.F32, x1)
    # backward of CHAIN8 with seed ones (activations from the oracle)
    acts = [x]
    for f in CHAIN8:
        acts.append(orc.apply_fn(f, orc.F32, acts[-1]))
    g = np.ones(N, np.float32)
    for k in reversed(range(8)):
        g = orc.add_unary_grad(CHAIN8_GRADS[k], orc.F32, acts[k], np.zeros(N, np.float32), g)
    out["chain8_grad_f32"] = g
    lhs = np.random.default_rng(2).uniform(-1, 1, N).astype(np.float32)
    rhs = np.random.default_rng(3).uniform(-1, 1, N).astype(np.float32)
    out["binary_lhs_f32"], out["binary_rhs_f32"] = lhs, rhs
    out["binary_add_f32"] = orc.binary(0, orc.F32, lhs, rhs)
    out["binary_mul_f32"] = orc.binary(1, orc.F32, lhs, rhs)
    s = np.random.default_rng(5).random(N, dtype=np.float32)
    out["sum_x_f32"] = s
    out["sum_seq_f32"] = np.array([orc.sum_seq(orc.F32, s)], np.float32)
    out["sum_two_pass_f32"] = np.array([orc.sum_two_pass(orc.F32, s, 4, 1024, 256, 4, 256)], np.float32)
    out["sum_f64"] = np.array([orc.sum_f64(orc.F32, s)], np.float64)
    np.savez_compressed(Path(__file__).resolve().parent / "oracle_vectors.npz", **out)
    print("wrote", len(out), "arrays")
    more = dtype_vectors()
    np.savez_compressed(Path(__file__).resolve().parent / "oracle_vectors_dtypes.npz", **more)
    print("wrote", len(more), "arrays (bf16 and the narrow / wide integers)")


INT_EXPR = lambda v: v.mul(3).add(7).sub(v.div(2))  # noqa: E731
INT_TYPES = {"i8": (orc.I8, np.int8), "i16": (orc.I16, np.int16), "u16": (orc.U16, np.uint16), "u64": (orc.U64, np.uint64)}


def dtype_vectors():
    """The dtypes added after the first fixture file: bf16 (bit patterns) and i8 / i16 / u16 / u64."""
    from custos_b200.expr import bf16_from_f32
    out = {}
    x = np.random.default_rng(4).uniform(-4, 4, N).astype(np.float32)
    xb = bf16_from_f32(x)
    xb[:6] = [0x0000, 0x8000, 0x7f80, 0xff80, 0x7fc0, 0x0001]  # +-0, +-inf, NaN, smallest subnormal
    out["x_bf16"] = xb
    out["chain8_y_bf16"] = orc.apply_chain(CHAIN8, orc.BF16, xb)
    out["cheap8_y_bf16"] = orc.apply_chain(CHEAP8, orc.BF16, xb)
    yb = bf16_from_f32(np.random.default_rng(3).uniform(-1, 1, N).astype(np.float32))
    out["y_bf16"] = yb
    for k, name in enumerate(("add", "mul", "sub", "div")):
        out[f"binary_{name}_bf16"] = orc.binary(k, orc.BF16, xb, yb)
    out["max_min_bf16"] = orc.apply2(lambda a, b: a.max(b).min(a.mul(b)), orc.BF16, xb, yb)
    zeros = np.array([0.0, -0.0, -0.0, 0.0], np.float16)
    out["max_ties_f16"] = orc.apply2(lambda a, b: a.max(b), orc.F16, zeros, zeros[::-1].copy()).view(np.uint16)
    for name, (dt, t) in INT_TYPES.items():
        info = np.iinfo(t)
        v = np.random.default_rng(7).integers(info.min, min(info.max, 2 ** 62), N, dtype=np.int64).astype(t)
        v[:4] = [info.min, info.max, 0, 1]
        out[f"x_{name}"] = v
        out[f"expr_y_{name}"] = orc.apply_fn(INT_EXPR, dt, v)
        out[f"sum_{name}"] = np.array([orc.sum_seq(dt, v)], np.int64)
    return out


def half_minmax_kats():
    """Known answers for Number::min / Number::max on f16 and bf16 over {+-0, +-1, +-inf, NaN} x both operand
    orders, derived here from the reference text alone (NOT from the oracle):
      max: `self.max(rhs)` -> half's inherent max (src/number.rs:507-510 for f16, :536-539 for bf16):
           `if other > self && !other.is_nan() { other } else { self }`
      min: the Number default (src/number.rs:207-209): `if self < rhs { self } else { rhs }`
    Comparisons are IEEE (NaN compares false, +0 == -0)."""
    import json
    import math
    vals = {"f16": {"+0": 0x0000, "-0": 0x8000, "1": 0x3c00, "-1": 0xbc00, "inf": 0x7c00, "-inf": 0xfc00, "nan": 0x7e00},
            "bf16": {"+0": 0x0000, "-0": 0x8000, "1": 0x3f80, "-1": 0xbf80, "inf": 0x7f80, "-inf": 0xff80, "nan": 0x7fc0}}
    num = {"+0": 0.0, "-0": -0.0, "1": 1.0, "-1": -1.0, "inf": math.inf, "-inf": -math.inf, "nan": math.nan}
    out = {}
    for ty, bits in vals.items():
        rows = []
        for a in bits:
            for b in bits:
                mx = b if (num[b] > num[a] and not math.isnan(num[b])) else a
                mn = a if num[a] < num[b] else b
                rows.append({"self": a, "rhs": b, "self_bits": bits[a], "rhs_bits": bits[b],
                             "max_bits": bits[mx], "min_bits": bits[mn]})
        out[ty] = rows
    path = Path(__file__).resolve().parent / "half_minmax_kats.json"
    path.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", path.name, sum(len(v) for v in out.values()), "cases")


if __name__ == "__main__":
    main()
    half_minmax_kats()
