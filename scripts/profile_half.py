#!/usr/bin/env python
"""Launches the fused CHAIN8 on 2^28 f16 and bf16 elements (for `ncu --set full -k regex:cb_apply_vec`)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from custos_b200 import _native as N  # noqa: E402
from custos_b200.expr import bf16_from_f32  # noqa: E402
from custos_b200.raw import RawDevice  # noqa: E402
from custos_b200.workloads import CHAIN8  # noqa: E402

n = 1 << 28
dev = RawDevice(0)
src, dst = dev.alloc(n * 2, zero=False), dev.alloc(n * 2, zero=False)
blk = np.random.default_rng(4).uniform(-4, 4, 1 << 24).astype(np.float32)
for dt, host in ((N.F16, blk.astype(np.float16)), (N.BF16, bf16_from_f32(blk))):
    p = dev.upload(host)
    for off in range(0, n, 1 << 24):
        dev.copy(dt, src, off, p, 0, 1 << 24)
    e = dev.compile(CHAIN8, dt)
    for _ in range(2):  # launch 1, 2 = f16; 3, 4 = bf16
        dev.apply(e, src, dst, n)
    dev.sync()
    dev.free(p)
print("done")
