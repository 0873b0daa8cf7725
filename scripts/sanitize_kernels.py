#!/usr/bin/env python
"""Drives every kernel of the library once at small, ragged sizes — meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python scripts/sanitize_kernels.py
    compute-sanitizer --tool racecheck python scripts/sanitize_kernels.py
    compute-sanitizer --tool initcheck python scripts/sanitize_kernels.py

Covers: the NVRTC skeleton (apply / unary_grad / two-marker, vector and scalar variants, every dtype), the AOT
kernels (binary, fill, copy, two-pass sum), unaligned sub-slices, in-place application, graph capture/replay,
the host pipeline (cb_apply_host) and the module stack (Lazy + Graph fusing, Autograd backward).
Exits non-zero on a wrong result; the sanitizer reports memory errors itself.
"""
import os
import sys
from pathlib import Path

import numpy as np

os.environ.setdefault("CB_LUT16_MIN_ELEMS", "2048")  # let small buffers take the table-lookup kernels too

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from custos_b200 import _native as N  # noqa: E402
from custos_b200.device import CUDA  # noqa: E402
from custos_b200.expr import bf16_from_f32  # noqa: E402
from custos_b200.raw import RawDevice  # noqa: E402
from custos_b200.workloads import CHAIN8, CHAIN8_GRADS, CHEAP8  # noqa: E402

FLOATS = [N.F32, N.F64, N.F16, N.BF16]
INTS = [N.I8, N.I16, N.I32, N.I64, N.U8, N.U16, N.U32, N.U64]
NP = {N.F32: np.float32, N.F64: np.float64, N.F16: np.float16, N.BF16: np.uint16, N.I8: np.int8, N.I16: np.int16,
      N.I32: np.int32, N.I64: np.int64, N.U8: np.uint8, N.U16: np.uint16, N.U32: np.uint32, N.U64: np.uint64}


def inputs(dt, n, seed):
    rng = np.random.default_rng(seed)
    if dt == N.BF16:
        return bf16_from_f32(rng.uniform(-4, 4, n).astype(np.float32))
    if dt in FLOATS:
        return rng.uniform(-4, 4, n).astype(NP[dt])
    return rng.integers(0, 100, n).astype(NP[dt])


def main():
    dev = RawDevice(0)
    launches0 = dev.launches
    for n in (1, 37, 5000 + 3):
        for dt in FLOATS + INTS:
            isf = dt in FLOATS
            x, y = inputs(dt, n + 4, 1), inputs(dt, n + 4, 2)
            px, py, po = dev.upload(x), dev.upload(y), dev.alloc(x.nbytes)
            sz = x.itemsize
            chain = (CHAIN8 if n > 1000 else CHEAP8) if isf else [lambda v: v.mul(3).add(1), lambda v: v.sub(2)]
            e = dev.compile(chain, dt)
            dev.apply(e, px, po, n)                       # vector kernel + tail
            dev.apply(e, px + sz, po + sz, n)             # unaligned: scalar kernel (vector again for 16-byte types)
            dev.apply(e, px, px, n)                       # in place
            g = dev.compile((lambda v: v.mul(2.0).cos()) if isf else (lambda v: v.mul(2)), dt, N.KERNEL_UNARY_GRAD)
            dev.unary_grad(g, py, po, px, n)
            dev.unary_grad(g, py + sz, po + sz, px + sz, n)
            dev.unary_grad_ex(g, py, po, px, n, N.GRAD_SEED_ONES)   # seed written by the kernel (lookup kernel for large 16-bit n)
            # the backward of a whole chain in one kernel
            fw = (CHAIN8 if n > 1000 else CHEAP8) if isf else [lambda v: v.mul(3).add(1), lambda v: v.sub(2)]
            gr = (CHAIN8_GRADS if n > 1000 else [lambda v: 1.0] * 8) if isf else [lambda v: 3, lambda v: 1]
            cg = dev.compile(list(fw) + list(gr), dt, N.KERNEL_CHAIN_GRAD)
            dev.unary_grad_ex(cg, py, po, px, n, 0)
            dev.unary_grad_ex(cg, py, po, px, n, N.GRAD_SEED_ONES)
            dev.unary_grad_ex(cg, py + sz, po + sz, px + sz, n, N.GRAD_SEED_ONES)
            if dt in (N.F16, N.BF16):
                dev.set_lut(e, False)
                dev.apply(e, py, po, n)                   # the arithmetic kernel where the lookup kernel ran above
                dev.set_lut(e, True)
            b2 = dev.compile((lambda a, b: a.mul(b).max(0.5)) if isf else (lambda a, b: a.mul(b).add(a)), dt, N.KERNEL_BINARY)
            dev.apply2(b2, px, py, po, n)
            dev.apply2(b2, px + sz, py + sz, po + sz, n)
            for op in (N.BIN_ADD, N.BIN_MUL, N.BIN_SUB, N.BIN_DIV):
                dev.binary(dt, op, px, py, po, n)
                dev.binary(dt, op, px + sz, py + sz, po + sz, n)
            dev.fill(dt, po, n, 1)
            dev.fill(dt, po + sz, n, 1)
            dev.clear(dt, po, n + 4)
            dev.copy(dt, po, 1, px, 2, n)
            dev.copy(dt, po, 0, px, 0, n + 4)
            dev.sum(dt, px, n)
            dev.sum(dt, px + sz, n)
            dev.mean(dt, px, n)
            for p in (px, py, po):
                dev.free(p)
    # one larger sum (several pass-1 blocks) and its check
    x = np.random.default_rng(3).random(300_001, dtype=np.float32)
    p = dev.upload(x)
    s = float(dev.sum(N.F32, p, x.size))
    assert abs(s - float(x.astype(np.float64).sum())) < 1e-5 * x.size, s
    dev.free(p)
    # graph capture / replay
    n = 4096
    a, b = dev.upload(inputs(N.F32, n, 4)), dev.alloc(n * 4)
    exprs = [dev.compile(f, N.F32) for f in CHEAP8]
    dev.sync()
    dev.graph_begin()
    src, dst = a, b
    for e in exprs:
        dev.apply(e, src, dst, n)
        src, dst = dst, src
    gr = dev.graph_end()
    for _ in range(3):
        dev.graph_launch(gr)
    dev.sync()
    dev.graph_destroy(gr)
    # host pipeline (pageable operands: staged path) — small chunk count
    hx = inputs(N.F32, (1 << 20) + 7, 5)
    out = np.zeros_like(hx)
    e = dev.compile(CHEAP8, N.F32)
    dev.apply_host(e, hx.ctypes.data, out.ctypes.data, hx.size)
    assert np.isfinite(out).all() and np.any(out != 0)
    print("raw device: launches", dev.launches - launches0)
    dev.close()
    # module stack
    x = inputs(N.F32, 10_001, 6)
    with CUDA("Graph", "Lazy", "Base") as d:
        cur = d.buffer(x)
        for f in CHAIN8:
            cur = d.apply_fn(cur, f)
        d.optimize_mem_graph()
        d.unary_fusing()
        d.run()
        assert np.isfinite(cur.replace().read()).all()
    for dt, xs in ((np.float32, x), (np.float16, x.astype(np.float16))):  # fused forward + fused (f16: looked-up) backward
        with CUDA("Lazy", "Graph", "Autograd", "Base", dtype=dt) as d:
            buf = d.buffer(xs).require_grad()
            cur = buf
            for f, g in zip(CHAIN8, CHAIN8_GRADS):
                cur = d.unary_ew(cur, f, g)
            d.optimize_mem_graph()
            d.unary_fusing()
            d.set_graph_replay(True)
            for _ in range(2):
                d.run()
                cur.backward()
            assert np.isfinite(buf.grad().read().astype(np.float32)).all()
            d.zero_grad()
            assert d.sum(cur.replace()) == d.sum(cur.replace())
    with CUDA("Autograd", "Cached", "Base") as d:
        buf = d.buffer(x).require_grad()
        cur = buf
        for f, g in zip(CHAIN8, CHAIN8_GRADS):
            cur = d.unary_ew(cur, f, g)
        cur.backward()
        assert np.isfinite(buf.grad().read()).all()
        text = cur.serialize()
        assert d.deserialize(text, np.float32).read().tobytes() == cur.read().tobytes()
    print("sanitize_kernels: ok")


if __name__ == "__main__":
    main()
