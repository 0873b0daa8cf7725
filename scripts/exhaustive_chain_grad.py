#!/usr/bin/env python
"""The fused backward against the tape it replaces, over ALL 2^32 f32 inputs.

For every f32 bit pattern x the gradient of CHAIN8 at x is computed twice on `CUDA<Lazy<Graph<Autograd<Base>>>>`:
by the reference's schedule (8 forward kernels, 8 add_unary_grad kernels over materialised intermediates) and by the
fused schedule (`unary_fusing`: one forward kernel, ONE recomputing chain-grad kernel).  Forward outputs and gradients
must agree bit for bit (NaN == NaN).  Even chunks seed with ones (`backward()`), odd chunks with an explicit seed
(`backward_with`: the chunk's own inputs reversed — every kind of value meets every other).  Prints one JSON object.

    python scripts/exhaustive_chain_grad.py [chunks]      # 16 chunks of 2^28 inputs = everything
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from custos_b200.device import CUDA  # noqa: E402
from custos_b200.workloads import CHAIN8, CHAIN8_GRADS  # noqa: E402

CHUNK = 1 << 28


def run(x, seed, fuse):
    with CUDA("Lazy", "Graph", "Autograd", "Base") as dev:
        buf = dev.buffer(x).require_grad()
        cur = buf
        for f, g in zip(CHAIN8, CHAIN8_GRADS):
            cur = dev.unary_ew(cur, f, g)
        if fuse:
            dev.unary_fusing()
        dev.run()
        if seed is None:
            cur.backward()
        else:
            cur.backward_with(seed)
        return cur.replace().read(), buf.grad().read()


def same(a, b):
    ua, ub = a.view(np.uint32), b.view(np.uint32)
    bad = (ua != ub) & ~(np.isnan(a) & np.isnan(b))
    return int(np.count_nonzero(bad)), (int(np.flatnonzero(bad)[0]) if bad.any() else None)


def main():
    chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    out = {"chunks": chunks, "inputs": chunks * CHUNK, "forward_mismatches": 0, "gradient_mismatches": 0, "first": None}
    t0 = time.time()
    for c in range(chunks):
        x = np.arange(c * CHUNK, (c + 1) * CHUNK, dtype=np.uint32).view(np.float32)
        seed = None if c % 2 == 0 else np.ascontiguousarray(x[::-1])
        y0, g0 = run(x, seed, False)
        y1, g1 = run(x, seed, True)
        fy, wy = same(y0, y1)
        fg, wg = same(g0, g1)
        out["forward_mismatches"] += fy
        out["gradient_mismatches"] += fg
        if out["first"] is None and (fy or fg):
            w = wy if fy else wg
            out["first"] = {"chunk": c, "index": w, "x_bits": hex(c * CHUNK + w), "unfused": float(g0[w]), "fused": float(g1[w])}
        print(f"chunk {c}: forward {fy}, gradient {fg} mismatches ({time.time() - t0:.0f} s)", file=sys.stderr, flush=True)
    out["seconds"] = round(time.time() - t0, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
