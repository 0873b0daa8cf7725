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


if __name__ == "__main__":
    main()
