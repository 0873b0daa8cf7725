"""Differential fuzzing of the module stack: a random program of apply_fn / binary / checkpoint steps is run
eagerly on `CUDA<Base>` and then on the lazy stacks with every rewrite the backend has — deferred allocation,
`optimize_mem_graph` aliasing, `unary_fusing`, `elementwise_fusing`, CUDA-graph replay, Cached — and every buffer
that no later step consumes must hold the same bits.  (Same device functions on both sides, so transcendental
steps are bit-identical too; what is under test is the recording, aliasing and fusing logic.)
"""
import os
import random

import numpy as np
import pytest

from custos_b200.device import CUDA

pytestmark = pytest.mark.gpu

ROUNDS = int(os.environ.get("CB_FUZZ_ROUNDS", "1"))
UNARY = [lambda x: x.add(1.5), lambda x: x.mul(0.5), lambda x: x.neg(), lambda x: x.abs(), lambda x: x.sin(),
         lambda x: x.mul(x).add(0.25), lambda x: x.tanh(), lambda x: x.max(-0.5).min(0.75), lambda x: x.sub(2.0).mul(3.0),
         lambda x: x.exp().mul(0.125), lambda x: x.cos(), lambda x: x.geq(0.0).mul(x)]
BINARY = ["add", "mul", "sub"]


def make_program(rng: random.Random):
    """-> list of steps over buffer slots; slots 0..2 are inputs (slot 2 has another length)."""
    lens = {0: 1000, 1: 1000, 2: 37}
    steps = []
    n_slots = 3
    for _ in range(rng.randint(5, 14)):
        roll = rng.random()
        if roll < 0.55:
            src = rng.randrange(n_slots)
            steps.append(("unary", src, rng.randrange(len(UNARY))))
            lens[n_slots] = lens[src]
        elif roll < 0.9:
            a = rng.randrange(n_slots)
            same = [s for s in range(n_slots) if lens[s] == lens[a] and s != a]
            if not same:
                continue
            steps.append(("binary", a, rng.choice(same), rng.choice(BINARY)))
            lens[n_slots] = lens[a]
        else:
            steps.append(("checkpoint", rng.randrange(n_slots)))
            continue
        n_slots += 1
    return steps


def run_program(dev, steps, inputs, prepare=None, runs=1):
    bufs = [dev.buffer(x) for x in inputs]
    consumed = set()
    for st in steps:
        if st[0] == "unary":
            bufs.append(dev.apply_fn(bufs[st[1]], UNARY[st[2]]))
            consumed.add(st[1])
        elif st[0] == "binary":
            bufs.append(getattr(dev, st[3])(bufs[st[1]], bufs[st[2]]))
            consumed.update((st[1], st[2]))
        else:
            bufs[st[1]].checkpoint()
    if prepare:
        prepare(dev)
    for _ in range(runs):
        if "Lazy" in dev.modules:
            dev.run()
    sinks = [i for i in range(len(bufs)) if i not in consumed]
    return {i: bufs[i].replace().read() for i in sinks}


CONFIGS = {
    "lazy": (("Lazy", "Base"), None, 1),
    "lazy_cached": (("Lazy", "Cached", "Base"), None, 2),
    "graph_lazy": (("Graph", "Lazy", "Base"), None, 1),
    "mem_graph": (("Graph", "Lazy", "Base"), lambda d: d.optimize_mem_graph(), 1),
    "unary_fusing": (("Graph", "Lazy", "Base"), lambda d: d.unary_fusing(), 1),
    "mem_graph+unary_fusing": (("Graph", "Lazy", "Base"), lambda d: (d.optimize_mem_graph(), d.unary_fusing()), 1),
    "elementwise_fusing": (("Graph", "Lazy", "Base"), lambda d: d.elementwise_fusing(), 1),
    "elementwise_fusing+replay": (("Graph", "Lazy", "Base"), lambda d: (d.elementwise_fusing(), d.set_graph_replay(True)), 3),
    "replay": (("Lazy", "Base"), lambda d: d.set_graph_replay(True), 3),
    "graph_lazy_cached+mem_graph": (("Graph", "Lazy", "Cached", "Base"), lambda d: d.optimize_mem_graph(), 2),
    "autograd_lazy": (("Autograd", "Lazy", "Base"), None, 1),
    "autograd_graph_lazy+unary_fusing": (("Autograd", "Graph", "Lazy", "Base"), lambda d: d.unary_fusing(), 1),
    "cached": (("Cached", "Base"), None, 1),
}


@pytest.mark.parametrize("seed", range(12 * ROUNDS))
def test_random_program_every_stack_agrees_with_eager(seed):
    rng = random.Random(4242 + seed)
    steps = make_program(rng)
    data = np.random.default_rng(seed)
    inputs = [data.uniform(-2, 2, 1000).astype(np.float32), data.uniform(-2, 2, 1000).astype(np.float32),
              data.uniform(-2, 2, 37).astype(np.float32)]
    with CUDA("Base") as dev:
        want = run_program(dev, steps, inputs)
    assert want, "every program has at least one sink"
    for name, (mods, prepare, runs) in CONFIGS.items():
        with CUDA(*mods) as dev:
            got = run_program(dev, steps, inputs, prepare, runs)
        assert got.keys() == want.keys()
        for i in want:
            assert got[i].tobytes() == want[i].tobytes(), f"seed {seed}, stack {name}: buffer {i} differs; program {steps}"


# ------------------------------------------------------------------ gradients: tape planning under fusing / aliasing
# (forward closure, grad closure) as a custos user would pass them to unary_ew; the grad closures only have to be
# deterministic for a differential test, but they are the analytic derivatives anyway
UNARY_EW = [(lambda x: x.add(1.5), lambda x: 1.0), (lambda x: x.mul(0.5), lambda x: 0.5), (lambda x: x.neg(), lambda x: -1.0),
            (lambda x: x.abs(), lambda x: x.geq(0.0).mul(2.0).sub(1.0)), (lambda x: x.sin(), lambda x: x.cos()),
            (lambda x: x.mul(x).add(0.25), lambda x: x.mul(2.0)),
            (lambda x: x.tanh(), lambda x: x.tanh().mul(x.tanh()).neg().add(1.0)),
            (lambda x: x.sub(2.0).mul(3.0), lambda x: 3.0), (lambda x: x.exp().mul(0.125), lambda x: x.exp().mul(0.125)),
            (lambda x: x.cos(), lambda x: x.sin().neg())]


def make_grad_program(rng: random.Random):
    """Chains of unary_ew with the occasional apply_fn (no grad function: breaks the tape), binary op (fan-in) and
    fan-out; -> steps over slots (0 and 1 are inputs that require gradients)."""
    steps, n_slots = [], 2
    for _ in range(rng.randint(4, 14)):
        roll = rng.random()
        src = n_slots - 1 if rng.random() < 0.75 else rng.randrange(n_slots)  # mostly extend the last chain
        if roll < 0.7:
            steps.append(("ew", src, rng.randrange(len(UNARY_EW))))
        elif roll < 0.82:
            steps.append(("fn", src, rng.randrange(len(UNARY))))
        elif roll < 0.95 and n_slots > 2:
            other = rng.choice([s for s in range(n_slots) if s != src])
            steps.append(("binary", src, other, rng.choice(BINARY)))
        else:
            steps.append(("checkpoint", rng.randrange(n_slots)))
            continue
        n_slots += 1
    return steps


def run_grad_program(dev, steps, inputs, seed_grad, prepare=None):
    bufs = [dev.buffer(x).require_grad() for x in inputs]
    last_ew = None
    for st in steps:
        if st[0] == "ew":
            f, g = UNARY_EW[st[2]]
            bufs.append(dev.unary_ew(bufs[st[1]], f, g))
            last_ew = len(bufs) - 1
        elif st[0] == "fn":
            bufs.append(dev.apply_fn(bufs[st[1]], UNARY[st[2]]))
        elif st[0] == "binary":
            bufs.append(getattr(dev, st[3])(bufs[st[1]], bufs[st[2]]))
        else:
            bufs[st[1]].checkpoint()
    if prepare:
        prepare(dev)
    if "Lazy" in dev.modules:
        dev.run()
    out = {"last": bufs[-1].replace().read()}
    if last_ew is not None:
        if seed_grad is None:
            bufs[last_ew].backward()
        else:
            bufs[last_ew].backward_with(seed_grad)
        out["sink"] = bufs[last_ew].replace().read()
        out["g0"], out["g1"] = bufs[0].grad().read(), bufs[1].grad().read()
    return out


GRAD_CONFIGS = {
    "lazy": (("Lazy", "Graph", "Autograd", "Base"), None),
    "unary_fusing": (("Lazy", "Graph", "Autograd", "Base"), lambda d: d.unary_fusing()),
    "mem_graph": (("Lazy", "Graph", "Autograd", "Base"), lambda d: d.optimize_mem_graph()),
    "mem_graph+unary_fusing": (("Lazy", "Graph", "Autograd", "Base"), lambda d: (d.optimize_mem_graph(), d.unary_fusing())),
    "unary_fusing+mem_graph": (("Lazy", "Graph", "Autograd", "Base"), lambda d: (d.unary_fusing(), d.optimize_mem_graph())),
    "unary_fusing+replay": (("Lazy", "Graph", "Autograd", "Base"), lambda d: (d.unary_fusing(), d.set_graph_replay(True))),
}


@pytest.mark.parametrize("seed", range(16 * ROUNDS))
def test_random_autograd_program_gradients_survive_fusing_and_aliasing(seed):
    """The tape reads the inputs of the ops that unary_fusing stops writing and optimize_mem_graph overwrites: whatever
    the passes decide (fuse the grad functions into one chain-grad kernel, or leave the chain alone), the gradients of
    the inputs must have the bits of the eager `CUDA<Autograd<Base>>` run."""
    rng = random.Random(777 + seed)
    steps = make_grad_program(rng)
    data = np.random.default_rng(100 + seed)
    n = 1003
    inputs = [data.uniform(-2, 2, n).astype(np.float32), data.uniform(-2, 2, n).astype(np.float32)]
    inputs[0][:6] = [0.0, -0.0, 1e-40, -3.5e5, np.inf, 88.0]
    seed_grad = None if seed % 2 == 0 else data.uniform(-1, 1, n).astype(np.float32)
    with CUDA("Autograd", "Base") as dev:
        want = run_grad_program(dev, steps, inputs, seed_grad)
    for name, (mods, prepare) in GRAD_CONFIGS.items():
        with CUDA(*mods) as dev:
            got = run_grad_program(dev, steps, inputs, seed_grad, prepare)
        assert got.keys() == want.keys()
        for k in want:
            same = (got[k].view(np.uint32) == want[k].view(np.uint32)) | (np.isnan(got[k]) & np.isnan(want[k]))
            assert np.all(same), (f"seed {seed}, stack {name}: {k} differs at {np.flatnonzero(~same)[:5]} "
                                  f"({got[k][~same][:3]} vs {want[k][~same][:3]}); program {steps}")
