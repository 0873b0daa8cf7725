"""The synthetic workloads of BASELINE.json / SURVEY.md §8(d), shared by bench.py and tests.

Each chain is a list of closures over `Resolve`, exactly as a custos user would record them
with successive `apply_fn` calls on a `Graph<Lazy<..>>` device.
"""
from .expr import Combiner

# config[2]: the 8-op fused unary chain (3 transcendentals)
CHAIN8 = [lambda x: x.add(1.0), lambda x: x.mul(0.5), lambda x: x.exp(), lambda x: x.sin(),
          lambda x: x.mul(2.0), lambda x: x.add(1.0), lambda x: x.tanh(), lambda x: x.neg()]

# analytic gradients of CHAIN8, op by op (unary_ew's grad closures)
CHAIN8_GRADS = [lambda x: 1.0, lambda x: 0.5, lambda x: x.exp(), lambda x: x.cos(),
                lambda x: 2.0, lambda x: 1.0,
                lambda x: Combiner._wrap(1.0).sub(x.tanh().mul(x.tanh())), lambda x: -1.0]

# the pure-bandwidth ceiling: 8 ops without transcendentals
CHEAP8 = [lambda x: x.add(1.0), lambda x: x.mul(0.5), lambda x: x.neg(), lambda x: x.abs(),
          lambda x: x.add(2.0), lambda x: x.mul(3.0), lambda x: x.neg(), lambda x: x.add(1.0)]

# config[0]: the reference's own CPU-runnable case, exp().sin()*2+1 on 1M f32
CONFIG1 = [lambda x: x.exp(), lambda x: x.sin(), lambda x: x.mul(2.0), lambda x: x.add(1.0)]

# SURVEY §8(d) input distributions: (low, high, seed)
INPUTS = {"chain8": (-4.0, 4.0, 4), "config1": (-2.0, 2.0, 1), "binary_lhs": (-1.0, 1.0, 2),
          "binary_rhs": (-1.0, 1.0, 3), "sum": (0.0, 1.0, 5), "sum_signed": (-1.0, 1.0, 6)}
