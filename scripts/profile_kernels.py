#!/usr/bin/env python
"""Launches every hot kernel of the path twice on 2^28-element buffers (for `ncu --set full`)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from custos_b200 import _native as N  # noqa: E402
from custos_b200.raw import RawDevice  # noqa: E402
from custos_b200.workloads import CHAIN8, CHAIN8_GRADS, CHEAP8  # noqa: E402

n = 1 << 28
dev = RawDevice(0)
a, b, c = dev.alloc(n * 4), dev.alloc(n * 4), dev.alloc(n * 4)
blk = np.random.default_rng(4).uniform(-4, 4, 1 << 24).astype(np.float32)
pb = dev.upload(blk)
ph = dev.upload(blk.astype(np.float16))
for off in range(0, n, 1 << 24):
    dev.copy(N.F32, a, off, pb, 0, 1 << 24)
    dev.copy(N.F16, b, off, ph, 0, 1 << 24)   # first half of b: binary16 inputs
s = dev.alloc(64)
work = [
    ("chain8_f32", lambda e=dev.compile(CHAIN8, N.F32): dev.apply(e, a, c, n)),
    ("cheap8_f32", lambda e=dev.compile(CHEAP8, N.F32): dev.apply(e, a, c, n)),
    ("chain8_f16", lambda e=dev.compile(CHAIN8, N.F16): dev.apply(e, b, c, n)),
    ("unary_grad_cos", lambda e=dev.compile(CHAIN8_GRADS[3], N.F32, N.KERNEL_UNARY_GRAD): dev.unary_grad(e, a, c, a, n)),
    ("binary_add", lambda: dev.binary(N.F32, N.BIN_ADD, a, a, c, n)),
    ("clear", lambda: dev.clear(N.F32, c, n)),
    ("copy", lambda: dev.copy(N.F32, c, 0, a, 0, n)),
    ("sum", lambda: dev.sum_into(N.F32, a, n, s)),
]
for name, fn in work:
    fn()
    fn()
dev.sync()
print("done")
