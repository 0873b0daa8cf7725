#!/usr/bin/env python
"""f16 / bf16 CHAIN8 on 2^28 elements: the table-lookup kernel (launch shape from CB_LUT_SHAPE) against the arithmetic
kernel.  One line of JSON."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import fill_tiled, make_input, timeit  # noqa: E402
from custos_b200 import _native as N  # noqa: E402
from custos_b200.expr import bf16_from_f32  # noqa: E402
from custos_b200.raw import RawDevice  # noqa: E402
from custos_b200.workloads import CHAIN8, CHEAP8  # noqa: E402

n = 1 << 28
raw = RawDevice(0)
a, c = raw.alloc(n * 2, zero=False), raw.alloc(n * 2, zero=False)
out = {"shape": os.environ.get("CB_LUT_SHAPE", "0")}
for name, dt, block in (("f16", N.F16, make_input(1 << 24).astype(np.float16)), ("bf16", N.BF16, bf16_from_f32(make_input(1 << 24)))):
    fill_tiled(raw, N, dt, a, n, block)
    for cname, chain in (("chain8", CHAIN8), ("cheap8", CHEAP8)):
        e = raw.compile(chain, dt)
        ms = timeit(raw, lambda: raw.apply(e, a, c, n), reps=30, warm=5)
        reps = max(30, int(1.0 / (ms * 1e-3)))
        ms_s = timeit(raw, lambda: raw.apply(e, a, c, n), reps=reps, warm=reps // 2)
        raw.set_lut(e, False)
        ms_a = timeit(raw, lambda: raw.apply(e, a, c, n), reps=10, warm=3)
        out[f"{cname}_{name}"] = {"lut_GB/s": round(n * 4 / ms / 1e6, 1), "lut_sustained_GB/s": round(n * 4 / ms_s / 1e6, 1),
                                  "arithmetic_GB/s": round(n * 4 / ms_a / 1e6, 1)}
print(json.dumps(out))
