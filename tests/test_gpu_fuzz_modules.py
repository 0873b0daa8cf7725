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
