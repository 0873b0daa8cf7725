#!/usr/bin/env python
"""Backward of the fused CHAIN8 (BASELINE configs[2]) on 2^28 f32: ONE recomputing chain-grad kernel
(CB_KERNEL_CHAIN_GRAD, 16 B/element) against the eight add_unary_grad kernels of the unfused tape (8 x 16 B/element
plus the seed fill), at the raw C ABI and through the module stack CUDA<Lazy<Graph<Autograd<Base>>>>.
Prints one JSON object."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from custos_b200 import _native as N  # noqa: E402
from custos_b200.device import CUDA  # noqa: E402
from custos_b200.raw import RawDevice  # noqa: E402
from custos_b200.workloads import CHAIN8, CHAIN8_GRADS  # noqa: E402

PEAK = 6547.5
if (ROOT / "MEASURED_PEAKS.json").exists():
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])


def timeit(dev, fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    dev.sync()
    e0, e1 = dev.event(), dev.event()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.sync()
    return e0.elapsed_ms(e1) / reps


def row(ms, n, bpe):
    gbs = n * bpe / (ms * 1e-3) / 1e9
    return {"ms": round(ms, 5), "GB/s": round(gbs, 1), "frac": round(gbs / PEAK, 4)}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
    out = {"elements": n}
    dev = RawDevice(0)
    x, g, og = dev.alloc(n * 4, zero=False), dev.alloc(n * 4), dev.alloc(n * 4, zero=False)
    block = np.random.default_rng(4).uniform(-4, 4, 1 << 24).astype(np.float32)
    pb = dev.upload(block)
    for off in range(0, n, 1 << 24):
        dev.copy(N.F32, x, off, pb, 0, min(1 << 24, n - off))
    dev.fill(N.F32, og, n, 1.0)
    e = dev.compile(CHAIN8 + CHAIN8_GRADS, N.F32, N.KERNEL_CHAIN_GRAD)
    out["chain8_bwd_fused"] = row(timeit(dev, lambda: dev.unary_grad_ex(e, x, g, og, n, 0)), n, 16)
    out["chain8_bwd_fused_seeded"] = row(timeit(dev, lambda: dev.unary_grad_ex(e, x, g, og, n, N.GRAD_SEED_ONES)), n, 16)
    reps = max(20, int(1.0 / (out["chain8_bwd_fused_seeded"]["ms"] * 1e-3)))
    out["chain8_bwd_fused_seeded_sustained"] = row(
        timeit(dev, lambda: dev.unary_grad_ex(e, x, g, og, n, N.GRAD_SEED_ONES), reps=reps, warm=reps // 2), n, 16)
    fwd = dev.compile(CHAIN8, N.F32)
    out["chain8_fwd_fused"] = row(timeit(dev, lambda: dev.apply(fwd, x, og, n)), n, 8)
    dev.fill(N.F32, og, n, 1.0)
    grads = [dev.compile(gr, N.F32, N.KERNEL_UNARY_GRAD) for gr in CHAIN8_GRADS]

    def unfused():
        for gr in grads:
            dev.unary_grad(gr, x, g, og, n)
    out["chain8_bwd_8_unary_grads"] = row(timeit(dev, unfused, reps=5, warm=2), n, 16 * 8)
    for p in (x, g, og, pb):
        dev.free(p)
    dev.close()

    # the same through the module stack: record 8 unary_ew, fuse, run() + backward()
    host = np.tile(block, max(1, n // block.size))[:n]
    with CUDA("Lazy", "Graph", "Autograd", "Base") as d:
        buf = d.buffer(host).require_grad()
        cur = buf
        for f, gr in zip(CHAIN8, CHAIN8_GRADS):
            cur = d.unary_ew(cur, f, gr)
        d.optimize_mem_graph()
        d.unary_fusing()
        d.set_graph_replay(True)
        d.run()
        cur.backward()
        d.sync()
        l0 = d.raw.launches
        ms_f = timeit(d.raw, d.run)
        l1 = d.raw.launches
        ms_b = timeit(d.raw, cur.backward)
        l2 = d.raw.launches
        out["stack_forward_run"] = dict(row(ms_f, n, 8), launches_per_call=(l1 - l0) / 25)
        out["stack_backward"] = dict(row(ms_b, n, 16), launches_per_call=(l2 - l1) / 25)

        def step():
            d.run()
            cur.backward()
        ms = timeit(d.raw, step)
        out["stack_forward_plus_backward"] = dict(row(ms, n, 24), launches_per_step=2)
        t = time.perf_counter()
        steps = max(20, int(1.0 / (ms * 1e-3)))
        ms_s = timeit(d.raw, step, reps=steps, warm=steps // 2)
        out["stack_forward_plus_backward_sustained"] = dict(row(ms_s, n, 24), wall_s=time.perf_counter() - t)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
