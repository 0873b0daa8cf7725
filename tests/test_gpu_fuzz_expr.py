"""Differential fuzzing of the expression code generator: random `Combiner` trees over the exactly-rounded ops
(add, sub, mul, div, neg, abs, min, max, identity, the comparisons, literals on either side, shared sub-trees)
are compiled by the device and evaluated by the CPU oracle on the same inputs — results must be bit-identical for
every dtype and every kernel kind (apply, fused chain, two-marker, unary_grad).  Seeds are fixed: a failure
reproduces, and prints the `to_cl_source` string of the offending tree.

The transcendental ops are left out on purpose: their bar is an ulp bound per function (test_gpu_kernels.py),
and a composition of them has no meaningful bit-level expectation.
"""
import os
import random

import numpy as np
import pytest

from custos_b200 import _native as N
from custos_b200.expr import Combiner, Resolve, bf16_from_f32, to_cl_source
from oracle import oracle as orc
from tests.helpers import NP, assert_same, edge_values, random_inputs

pytestmark = pytest.mark.gpu

FLOAT_BIN = ["add", "sub", "mul", "div", "min", "max", "geq", "leq", "eq"]
FLOAT_UN = ["neg", "abs", "identity"]
INT_BIN = ["add", "sub", "mul", "div", "geq", "leq", "eq"]
FLOAT_LITS = [0.5, 2.0, -1.5, 3.0, 0.25, 1.0, -0.0, 8.0, -0.75]
INT_LITS = [0, 1, 2, 3, 7, 100]


def rand_tree(rng: random.Random, depth: int, leaves, bins, uns, lits) -> Combiner:
    roll = rng.random()
    if depth == 0 or roll < 0.2:
        return rng.choice(leaves) if rng.random() < 0.75 else Combiner._wrap(rng.choice(lits))
    if uns and roll < 0.4:
        return getattr(rand_tree(rng, depth - 1, leaves, bins, uns, lits), rng.choice(uns))()
    a = rand_tree(rng, depth - 1, leaves, bins, uns, lits)
    b = a if rng.random() < 0.15 else rand_tree(rng, depth - 1, leaves, bins, uns, lits)  # shared sub-tree sometimes
    return getattr(a, rng.choice(bins))(b)


def uses(tree: Combiner, marker: str) -> bool:
    if tree is None:
        return False
    return tree.op == marker or uses(tree.lhs, marker) or uses(tree.rhs, marker)


def inputs_for(dt, n, seed):
    if dt == N.BF16:
        return np.concatenate([random_inputs(dt, n, seed), bf16_from_f32(edge_values(np.float32))])
    if dt in (N.F32, N.F64, N.F16):
        return np.concatenate([random_inputs(dt, n, seed), edge_values(NP[dt])])
    info = np.iinfo(NP[dt])
    x = random_inputs(dt, n, seed)
    x[:4] = [info.min, info.max, 0, 1]
    return x


def ops_for(dt):
    if dt in (N.F32, N.F64, N.F16, N.BF16):
        return FLOAT_BIN, FLOAT_UN, FLOAT_LITS
    signed = dt in (N.I8, N.I16, N.I32, N.I64)
    return INT_BIN, (["neg"] if signed else []), INT_LITS


# CB_FUZZ_ROUNDS=k repeats every test with k different seed sets (the committed default is one, ~30 s on a B200)
ROUNDS = int(os.environ.get("CB_FUZZ_ROUNDS", "1"))
ALL_NUMBERS = [N.F32, N.F64, N.F16, N.BF16, N.I32, N.I64, N.U32, N.U8, N.I8, N.I16, N.U16, N.U64]


@pytest.mark.parametrize("rnd", range(ROUNDS))
@pytest.mark.parametrize("dt", ALL_NUMBERS)
def test_random_unary_expressions_and_chains(raw_device, dt, rnd):
    dev = raw_device
    rng = random.Random(1000 + dt + 7919 * rnd)
    bins, uns, lits = ops_for(dt)
    x = inputs_for(dt, 4099, 50 + dt)
    px, po = dev.upload(x), dev.alloc(x.nbytes)
    trees = []
    for _ in range(14):
        t = rand_tree(rng, rng.randint(1, 4), [Resolve("x")], bins, uns, lits)
        trees.append(t)
        dev.apply(dev.compile(t, dt), px, po, x.size)
        assert_same(dt, dev.d2h(po, x.size, dt), orc.apply_fn(t, dt, x), f"dtype {dt}: {to_cl_source(t, dt)}")
    for k in range(0, 12, 4):  # fused chains of four random ops
        chain = trees[k:k + 4]
        dev.apply(dev.compile(chain, dt), px, po, x.size)
        what = " ; ".join(to_cl_source(t, dt) for t in chain)
        assert_same(dt, dev.d2h(po, x.size, dt), orc.apply_chain(chain, dt, x), f"dtype {dt} chain: {what}")
    dev.free(px)
    dev.free(po)


@pytest.mark.parametrize("rnd", range(ROUNDS))
@pytest.mark.parametrize("dt", ALL_NUMBERS)
def test_random_two_marker_expressions_and_grads(raw_device, dt, rnd):
    dev = raw_device
    rng = random.Random(2000 + dt + 7919 * rnd)
    bins, uns, lits = ops_for(dt)
    x, y = inputs_for(dt, 3001, 60 + dt), inputs_for(dt, 3001, 70 + dt)
    g0 = inputs_for(dt, 3001, 80 + dt)
    px, py, po, pg = dev.upload(x), dev.upload(y), dev.alloc(x.nbytes), dev.alloc(x.nbytes)
    done = 0
    while done < 8:
        t = rand_tree(rng, rng.randint(1, 4), [Resolve("x"), Resolve("y")], bins, uns, lits)
        if not uses(t, "y"):
            continue
        done += 1
        dev.apply2(dev.compile(t, dt, N.KERNEL_BINARY), px, py, po, x.size)
        assert_same(dt, dev.d2h(po, x.size, dt), orc.apply2(t, dt, x, y), f"dtype {dt}: {to_cl_source(t, dt, n_args=2)}")
    for _ in range(6):  # lhs_grad += out_grad * g(lhs): multiply, then add
        t = rand_tree(rng, rng.randint(0, 3), [Resolve("x")], bins, uns, lits)
        dev.h2d(pg, g0)
        dev.unary_grad(dev.compile(t, dt, N.KERNEL_UNARY_GRAD), px, pg, py, x.size)
        assert_same(dt, dev.d2h(pg, x.size, dt), orc.add_unary_grad(t, dt, x, g0, y), f"grad dtype {dt}: {to_cl_source(t, dt)}")
    for p in (px, py, po, pg):
        dev.free(p)


@pytest.mark.parametrize("rnd", range(ROUNDS))
@pytest.mark.parametrize("dt", ALL_NUMBERS)
def test_random_chain_grads_equal_the_replayed_tape(raw_device, dt, rnd):
    """CB_KERNEL_CHAIN_GRAD on random chains of exactly-rounded closures, every dtype: the one recomputing kernel
    against the oracle's restatement of what the tape does — activations op by op, then K `add_unary_grad` calls in
    reverse over zero-initialised intermediate gradients (src/unary.rs:118-128, src/modules/autograd/tape.rs:39-47).
    Seeded (the kernel writes the ones) and with a general out_grad; bit-exact."""
    dev = raw_device
    rng = random.Random(3000 + dt + 7919 * rnd)
    bins, uns, lits = ops_for(dt)
    x, og, g0 = inputs_for(dt, 2053, 90 + dt), inputs_for(dt, 2053, 91 + dt), inputs_for(dt, 2053, 92 + dt)
    px, pog, pg = dev.upload(x), dev.upload(og), dev.upload(g0)
    one = np.ones(1, NP[dt])[0] if dt != N.BF16 else np.uint16(0x3f80)
    for _ in range(6):
        K = rng.randint(1, 5)
        fwd = [rand_tree(rng, rng.randint(0, 2), [Resolve("x")], bins, uns, lits) for _ in range(K)]
        grads = [rand_tree(rng, rng.randint(0, 2), [Resolve("x")], bins, uns, lits) for _ in range(K)]
        e = dev.compile(fwd + grads, dt, N.KERNEL_CHAIN_GRAD)
        acts = [x]
        for f in fwd[:-1]:
            acts.append(orc.apply_fn(f, dt, acts[-1]))
        what = " ; ".join(to_cl_source(t, dt) for t in fwd) + " | " + " ; ".join(to_cl_source(t, dt) for t in grads)
        for seeded in (False, True):
            seed = np.full(x.size, one, NP[dt]) if seeded else og
            want = seed
            for k in reversed(range(K)):
                into = g0 if k == 0 else np.zeros_like(x)
                want = orc.add_unary_grad(grads[k], dt, acts[k], into, want)
            dev.h2d(pg, g0)
            dev.h2d(pog, og)
            dev.unary_grad_ex(e, px, pg, pog, x.size, N.GRAD_SEED_ONES if seeded else 0)
            assert_same(dt, dev.d2h(pg, x.size, dt), want, f"chain grad dtype {dt} seeded={seeded}: {what}")
            if seeded:
                assert np.all(dev.d2h(pog, x.size, dt) == one), "the kernel did not write the seed"
    for p in (px, pog, pg):
        dev.free(p)
