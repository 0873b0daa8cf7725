#!/usr/bin/env python
"""Measures every row of BASELINE.md §4 (the BASELINE.json configs) on the GPU(s) of this box.

    python scripts/bench_configs.py                      # 1 GPU rows
    python -m torch.distributed.run --nproc-per-node N ... scripts/bench_configs.py --sum-only   # sharded sum

Not the driver's bench contract (that is bench.py); this fills the result table and profiles/.
Prints one JSON object.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from custos_b200 import _native as N  # noqa: E402
from custos_b200.device import CUDA  # noqa: E402
from custos_b200.raw import Comm, RawDevice, shard_range  # noqa: E402
from custos_b200.workloads import CHAIN8, CHAIN8_GRADS, CHEAP8, CONFIG1  # noqa: E402

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


def row(ms, elems, bytes_per_elem):
    gbs = elems * bytes_per_elem / (ms * 1e-3) / 1e9
    return {"ms": round(ms, 5), "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 4),
            "elements_per_s": elems / (ms * 1e-3)}


def single_gpu(n):
    out = {}
    dev = RawDevice(0)
    a, b, c = dev.alloc(n * 4), dev.alloc(n * 4), dev.alloc(n * 4)
    rng = np.random.default_rng(2)
    seed = rng.uniform(-1, 1, 1 << 24).astype(np.float32)
    ps = dev.upload(seed)
    for off in range(0, n, 1 << 24):  # tile a 2^24 random block over the big buffers
        m = min(1 << 24, n - off)
        dev.copy(N.F32, a, off, ps, 0, m)
        dev.copy(N.F32, b, off, ps, (1 << 23) % m if m == 1 << 24 else 0, m - ((1 << 23) % m if m == 1 << 24 else 0))
    # configs[1]: binary add / mul on 2^28 f32
    out["binary_add_f32"] = row(timeit(dev, lambda: dev.binary(N.F32, N.BIN_ADD, a, b, c, n)), n, 12)
    out["binary_mul_f32"] = row(timeit(dev, lambda: dev.binary(N.F32, N.BIN_MUL, a, b, c, n)), n, 12)
    # configs[2]: fused chain forward, f32 and f16
    e32 = dev.compile(CHAIN8, N.F32)
    out["chain8_fwd_f32"] = row(timeit(dev, lambda: dev.apply(e32, a, c, n)), n, 8)
    # f16: proper binary16 inputs U[-4,4) (a 2^24 block tiled over the first half of `b`)
    h = np.random.default_rng(4).uniform(-4, 4, 1 << 24).astype(np.float16)
    ph = dev.upload(h)
    for off in range(0, n, 1 << 24):
        dev.copy(N.F16, b, off, ph, 0, min(1 << 24, n - off))
    e16 = dev.compile(CHAIN8, N.F16)
    out["chain8_fwd_f16"] = row(timeit(dev, lambda: dev.apply(e16, b, c, n)), n, 4)
    c16 = dev.compile(CHEAP8, N.F16)
    out["cheap8_fwd_f16"] = row(timeit(dev, lambda: dev.apply(c16, b, c, n)), n, 4)
    dev.free(ph)
    # bf16: same inputs rounded to bfloat16
    from custos_b200.expr import bf16_from_f32
    pb = dev.upload(bf16_from_f32(np.random.default_rng(4).uniform(-4, 4, 1 << 24).astype(np.float32)))
    for off in range(0, n, 1 << 24):
        dev.copy(N.BF16, b, off, pb, 0, min(1 << 24, n - off))
    eb = dev.compile(CHAIN8, N.BF16)
    out["chain8_fwd_bf16"] = row(timeit(dev, lambda: dev.apply(eb, b, c, n)), n, 4)
    cb = dev.compile(CHEAP8, N.BF16)
    out["cheap8_fwd_bf16"] = row(timeit(dev, lambda: dev.apply(cb, b, c, n)), n, 4)
    dev.free(pb)
    for off in range(0, n, 1 << 24):  # restore b
        dev.copy(N.F32, b, off, ps, 0, min(1 << 24, n - off))
    cheap = dev.compile(CHEAP8, N.F32)
    out["cheap8_fwd_f32"] = row(timeit(dev, lambda: dev.apply(cheap, a, c, n)), n, 8)
    # configs[2] backward: eight add_unary_grad launches (lhs, out_grad -> lhs_grad), 16 B per element per op
    grads = [dev.compile(g, N.F32, N.KERNEL_UNARY_GRAD) for g in CHAIN8_GRADS]

    def backward():
        for g in grads:
            dev.unary_grad(g, a, c, b, n)

    ms = timeit(dev, backward, reps=10, warm=3)
    out["chain8_bwd_f32_8_unary_grads"] = row(ms, n, 16 * 8)
    out["unary_grad_each"] = {f"op{k}": row(timeit(dev, lambda g=g: dev.unary_grad(g, a, c, b, n), reps=10, warm=2), n, 16)
                              for k, g in enumerate(grads)}
    out["clear_f32"] = row(timeit(dev, lambda: dev.clear(N.F32, c, n)), n, 4)
    out["copy_f32"] = row(timeit(dev, lambda: dev.copy(N.F32, c, 0, a, 0, n)), n, 8)
    for p in (a, b, c, ps):
        dev.free(p)
    dev.close()
    return out


def replay_config():
    """configs[4]: Cached+Lazy, 20 recorded ops on 4096-element f32 buffers: graph replay vs eager launches."""
    n = 4096
    x = np.random.default_rng(70).uniform(-1, 1, n).astype(np.float32)
    res = {}
    for mode in ("eager_launches", "graph_replay"):
        with CUDA("Lazy", "Cached", "Base") as dev:
            dev.set_graph_replay(mode == "graph_replay")
            a, b = dev.buffer(x), dev.buffer(x)
            cur = a
            for k in range(10):
                cur = dev.apply_fn(cur, CHAIN8[k % 8])
                cur = dev.add(cur, b)
            dev.run()
            dev.sync()
            reps = 10000
            t = time.perf_counter()
            for _ in range(reps):
                dev.run()
            dev.sync()
            dt = time.perf_counter() - t
            res[mode] = {"us_per_run": dt / reps * 1e6, "us_per_op": dt / reps * 1e6 / 20, "runs": reps}
    res["speedup"] = res["eager_launches"]["us_per_run"] / res["graph_replay"]["us_per_run"]
    # beyond the reference: Graph + Lazy with element-wise fusing collapses the 20 ops into one kernel
    with CUDA("Graph", "Lazy", "Base") as dev:
        a, b = dev.buffer(x), dev.buffer(x)
        cur = a
        for k in range(10):
            cur = dev.apply_fn(cur, CHAIN8[k % 8])
            cur = dev.add(cur, b)
        dev.elementwise_fusing()
        dev.set_graph_replay(True)
        dev.run()
        dev.sync()
        reps = 10000
        t = time.perf_counter()
        for _ in range(reps):
            dev.run()
        dev.sync()
        dt = time.perf_counter() - t
        res["elementwise_fused_graph_replay"] = {"us_per_run": dt / reps * 1e6, "kernels": dev.replay_kernel_nodes(), "runs": reps}
    return res


def sum_config(total):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = RawDevice(local_rank)
    b, e = shard_range(total, 4, world, rank)
    n = e - b
    # x ~ U[0,1) seed 5: a 2^24 block tiled over the slice (the truth is then blocks * block_sum in fp64)
    block = np.random.default_rng(5).random(1 << 24, dtype=np.float32)
    pb = dev.upload(block)
    buf = dev.alloc(n * 4, zero=False)
    for off in range(0, n, 1 << 24):
        dev.copy(N.F32, buf, off, pb, 0, min(1 << 24, n - off))
    out = dev.alloc(64)
    comm = None
    if world > 1:
        uid = [Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = Comm(dev, world, rank, uid[0])

    def step():
        if comm:
            comm.sum_into(N.F32, buf, n, out)
        else:
            dev.sum_into(N.F32, buf, n, out)

    for _ in range(5):
        step()
    dev.sync()
    if world > 1:
        dist.barrier()
    e0, e1 = dev.event(), dev.event()
    reps = 50
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    e1.sync()
    ms = e0.elapsed_ms(e1) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    got = float(dev.d2h(out, 1, N.F32)[0])
    full, rem = divmod(total, 1 << 24)
    truth = full * float(np.sum(block.astype(np.float64))) + float(np.sum(block[:rem].astype(np.float64)))
    res = {"n_gpus": world, "elements": total, "ms": ms, "GB/s": total * 4 / (ms * 1e-3) / 1e9,
           "frac_of_measured_peak_x_gpus": total * 4 / (ms * 1e-3) / 1e9 / (PEAK * world),
           "rel_err_vs_fp64": abs(got - truth) / truth, "sum": got,
           "exchange": ("peer memory (fused kernel)" if comm and comm.uses_peer_memory else ("nccl all-gather" if comm else "none"))}
    if comm:
        comm.close()
    dev.close()
    if world > 1:
        dist.destroy_process_group()
    return res if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sum-only", action="store_true")
    ap.add_argument("--elems", type=int, default=1 << 28)
    ap.add_argument("--sum-elems", type=int, default=1 << 30)
    args = ap.parse_args()
    if args.sum_only:
        r = sum_config(args.sum_elems)
        if r:
            print(json.dumps({"sum_f32_2^30": r}))
        return
    out = single_gpu(args.elems)
    out["replay_20op_4096"] = replay_config()
    out["sum_f32_2^30"] = sum_config(args.sum_elems)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
