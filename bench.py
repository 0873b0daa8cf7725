#!/usr/bin/env python
"""bench.py — the headline benchmark of BASELINE.json on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-configs]

Metric: achieved HBM GB/s of the fused 8-op unary chain on 2^28 f32 per GPU.

The step is what a custos user runs: on the module stack `CUDA<Lazy<Graph<Autograd<Base>>>>` eight `apply_fn` calls
are RECORDED, `optimize_mem_graph` aliases the trace, `unary_fusing` turns the eight ops into one kernel, and every
step is one `run()` — the replay of the captured CUDA graph (one kernel node).  The buffers are 1 GiB each, far larger
than the 126 MB L2, so no flush is needed between steps.  `value` is device-timed (CUDA events on the library's
stream) with inputs resident in HBM; `e2e` is the same fused kernel through the C ABI with HOST buffers (pinned), H2D
and D2H inside the timed region.  Weak scaling: every rank owns its own 2^28-element slice, no collective on the
data path (max-over-ranks timing through torch.distributed/NCCL).

The same line carries the other BASELINE.json configs (`configs`: binary add/mul, f16/bf16 chains, the fused backward,
sum/mean 2^30 with a bit-exact parity check against the oracle, the 20-op replay, config 1) and, at N > 1, a
strong-scaling figure (one 2^28 buffer and one 2^30 sum split N ways) and the timed collective (`cb_comm_sum`: peer
memory vs NCCL) — `multi_gpu`.

`--impl reference` times the reference's CPU device (restated in oracle/, the Rust crate cannot be built here) on the
host cores for the same metric/config; every step is a bounded sample and `ms_per_step` is the measured time of it.
Prints exactly one JSON line on stdout.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_ELEMS = 1 << 28
SUM_ELEMS = 1 << 30
BYTES_PER_ELEM = 8  # f32: 1 read + 1 write per element, whatever the chain length
METRIC = "achieved HBM GB/s, fused 8-op unary chain, 2^28 f32 per GPU"
REF_SAMPLE = 1 << 26  # elements per step of the reference arm


def bench_config(world: int, n: int) -> dict:
    """Identical for both arms (the driver compares them)."""
    return {"workload": "chain8_f32_2^28_per_gpu" if n == N_ELEMS else f"chain8_f32_{n}_per_gpu",
            "chain": "add(1) mul(0.5) exp sin mul(2) add(1) tanh neg", "elements_per_gpu": n,
            "l2": "inputs (1 GiB in + 1 GiB out) larger than L2", "parallelism": f"slice{world}"}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture of this bench command (a number
    only a profiler can produce; the file names the capture it came from)."""
    p = ROOT / "profiles" / "roofline_traffic.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return d.get("chain8_f32_traffic_bytes"), d.get("source")
        except Exception:
            return None, None
    return None, None


class ClockSampler:
    """SM clock, power and throttle reasons polled through NVML (about 1 kHz) from a helper thread,
    so that even a 20 ms timed region gets its own samples; `nvidia-smi` exposes the same counters
    (B200_PROFILING.md recipe) but cannot be sampled faster than every ~50 ms."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index: int):
        self.samples, self.ok, self._stop = [], False, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            if visible:
                try:
                    gpu_index = int(visible.split(",")[gpu_index])
                except ValueError:
                    pass
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _poll(self):
        nv, h = self.nv, self.h
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append((time.time(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                     nv.nvmlDeviceGetPowerUsage(h) / 1000.0, int(reasons_fn(h))))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.0005)

    def window(self, t0: float, t1: float, how: str):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {getattr(self, 'err', '')}"]}
        rows = [r for r in self.samples if t0 <= r[0] <= t1]
        if not rows:  # region shorter than one poll: take the nearest samples
            rows = sorted(self.samples, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))[:3]
        if not rows:  # NVML refused to answer (e.g. the process runs under a profiler)
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no NVML samples"], "samples": 0, "how": how}
        mask = 0
        for r in rows:
            mask |= r[3]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max_sm,
                "reasons": sorted(name for bit, name in self.REASONS.items() if mask & bit),
                "power_w_max": max(r[2] for r in rows), "samples": len(rows), "how": how}

    def wait_ready(self, timeout: float = 5.0):
        """Block until the poller has delivered its first sample (nvmlInit and the first query can take
        longer than a whole 20 ms timed region on a fresh box)."""
        t_end = time.time() + timeout
        while self.ok and not self.samples and time.time() < t_end:
            time.sleep(0.001)

    def stop(self):
        self._stop.set()


def make_input(n: int, seed: int = 4, lo: float = -4.0, hi: float = 4.0, out: np.ndarray | None = None) -> np.ndarray:
    """x ~ U[lo, hi) f32, generated in chunks to bound host memory."""
    rng = np.random.default_rng(seed)
    if out is None:
        out = np.empty(n, np.float32)
    step = 1 << 24
    for i in range(0, n, step):
        m = min(step, n - i)
        out[i:i + m] = rng.uniform(lo, hi, m).astype(np.float32)
    return out


def pcie_ceiling(torch, barrier, max_over_ranks, nbytes: int = 1 << 28, reps: int = 4):
    """What the host link of this box moves with plain pinned-memory copies WHILE EVERY RANK DOES THE SAME — the
    ceiling of the end-to-end number: H2D alone, D2H alone, and both directions at once (GB/s per GPU, CUDA events,
    the slowest rank's time)."""
    h_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    cur = torch.cuda.current_stream()

    def timed(h2d: bool, d2h: bool) -> float:
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)
        e1.record()
        e1.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1))
        return (int(h2d) + int(d2h)) * reps * nbytes / (ms * 1e-3) / 1e9

    timed(True, True)  # warm-up
    return {"h2d_gbs": timed(True, False), "d2h_gbs": timed(False, True), "bidir_gbs": timed(True, True),
            "how": f"torch pinned copies of {nbytes >> 20} MiB x {reps}, one stream per direction, all ranks at once, "
                   f"CUDA events, slowest rank; GB/s per GPU"}


# ============================================================================================ CPU arms
def cpu_baselines(sample_single: int = 1 << 25, sample_boxed: int = 1 << 24):
    """The oracle (kind "port") on the host cores.  `value` is the single-threaded fused path, because the reference
    CPU device is single threaded (no rayon / threads / SIMD in src/); beside it the two variants SURVEY §8(d) names
    and the all-threads split used by the reference arm."""
    from custos_b200.workloads import CHAIN8
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    x = make_input(sample_single)
    orc.apply_chain(CHAIN8, orc.F32, x[:1 << 16])

    def rate(fn, n):
        t = time.perf_counter()
        fn(x[:n])
        return n / (time.perf_counter() - t)
    single = rate(lambda v: orc.apply_chain(CHAIN8, orc.F32, v), sample_single)
    boxed = rate(lambda v: orc.apply_chain_boxed(CHAIN8, v), sample_boxed)
    unfused = rate(lambda v: orc.apply_chain_unfused(CHAIN8, v), sample_boxed)
    allc = rate(lambda v: orc.apply_chain(CHAIN8, orc.F32, v, threads=cores), sample_single)
    gbs = lambda r: r * BYTES_PER_ELEM / 1e9  # noqa: E731
    return {"value": gbs(single), "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": f"fused CHAIN8 over 2^{int(np.log2(sample_single))} f32 (U[-4,4), seed 4), oracle/ C port of the "
                      f"reference CPU device, ONE thread as in the reference; {single / 1e6:.1f} M elem/s",
            "variants": {
                "fused_interpreter_1_thread": {"GB/s": gbs(single), "elements_per_s": single},
                "faithful_boxed_dyn_op_per_element_per_op_1_thread": {
                    "GB/s": gbs(boxed), "elements_per_s": boxed, "sample_elems": sample_boxed,
                    "ref": "src/devices/cpu/cpu_device.rs:217-229, src/op_hint.rs:30-33"},
                "monomorphised_unfused_8_passes_1_thread": {
                    "GB/s": gbs(unfused), "elements_per_s": unfused, "sample_elems": sample_boxed,
                    "ref": "src/devices/cpu_stack_ops.rs:7-15"},
                f"fused_interpreter_{cores}_threads_not_what_the_reference_does": {"GB/s": gbs(allc), "elements_per_s": allc}}}


def config1_cpu():
    """BASELINE configs[0]: CPU<Lazy<Graph<Base>>> exp().sin()*2+1 on 1M f32 — the reference's own CPU-runnable case."""
    from custos_b200.workloads import CONFIG1
    from oracle import oracle as orc
    x = make_input(1 << 20, seed=1, lo=-2.0, hi=2.0)
    orc.apply_chain(CONFIG1, orc.F32, x)
    res = {}
    for name, fn in (("fused_interpreter", lambda: orc.apply_chain(CONFIG1, orc.F32, x)),
                     ("faithful_boxed_dyn", lambda: orc.apply_chain_boxed(CONFIG1, x))):
        t = time.perf_counter()
        for _ in range(3):
            fn()
        res[name + "_ms"] = (time.perf_counter() - t) / 3 * 1e3
    return res


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path on the host cores, all the threads it can use;
    each step = the fused chain over a 2^26-element sample, and ms_per_step is the MEASURED time of that step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from custos_b200.workloads import CHAIN8
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    x = make_input(REF_SAMPLE)
    for _ in range(args.warmup):
        orc.apply_chain(CHAIN8, orc.F32, x, threads=cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        orc.apply_chain(CHAIN8, orc.F32, x, threads=cores)
    dt = (time.perf_counter() - t) / args.steps
    val = REF_SAMPLE * BYTES_PER_ELEM / dt / 1e9
    # the faithful single-threaded forms beside it (small samples: they are 10-100x slower)
    t = time.perf_counter()
    orc.apply_chain(CHAIN8, orc.F32, x[:1 << 23])
    single = (1 << 23) / (time.perf_counter() - t)
    t = time.perf_counter()
    orc.apply_chain_boxed(CHAIN8, x[:1 << 22])
    boxed = (1 << 22) / (time.perf_counter() - t)
    desc = (f"each step = fused CHAIN8 over a 2^26-element sample of the 2^28-element f32 workload (ms_per_step is the "
            f"measured time of that sample, not a projection), oracle/ C port of the reference CPU device (the Rust crate "
            f"cannot be built here), split over {cores} host threads; the reference itself is single threaded: "
            f"{single * BYTES_PER_ELEM / 1e9:.3f} GB/s on one thread, {boxed * BYTES_PER_ELEM / 1e9:.3f} GB/s with its "
            f"boxed dyn op per element per op (cpu_device.rs:217-229)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.gpus, args.elems),
        "elements_per_step": REF_SAMPLE, "elements_per_s": REF_SAMPLE / dt,
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port", "sample": desc,
                         "single_thread_gbs": single * BYTES_PER_ELEM / 1e9,
                         "single_thread_boxed_dyn_gbs": boxed * BYTES_PER_ELEM / 1e9},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ============================================================================================ GPU arm
def timeit(raw, fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    raw.sync()
    e0, e1 = raw.event(), raw.event()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.sync()
    return e0.elapsed_ms(e1) / reps


def fill_tiled(raw, N, dtype, dst, n, block: np.ndarray):
    """tiles a host block over a big device buffer (seeded inputs without a 1 GiB host array)"""
    pb = raw.upload(block)
    for off in range(0, n, block.size):
        raw.copy(dtype, dst, off, pb, 0, min(block.size, n - off))
    raw.free(pb)


def other_configs(local_rank: int, n: int, peak: float):
    """The other BASELINE.json configs on one GPU.  Every row: ms per launch (CUDA events, 20 launches after
    5 warm-ups), algorithmic GB/s and the fraction of the measured HBM peak."""
    from custos_b200 import _native as N
    from custos_b200.device import CUDA
    from custos_b200.expr import bf16_from_f32
    from custos_b200.raw import RawDevice, sum_plan
    from custos_b200.workloads import CHAIN8, CHAIN8_GRADS, CHEAP8, CONFIG1
    from oracle import oracle as orc
    res = {}

    def row(ms, elems, bpe, **kw):
        gbs = elems * bpe / (ms * 1e-3) / 1e9
        return dict({"ms": round(ms, 5), "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 4),
                     "elements_per_s": elems / (ms * 1e-3), "bytes_per_element": bpe}, **kw)

    raw = RawDevice(local_rank)
    a, b, c = raw.alloc(n * 4, zero=False), raw.alloc(n * 4, zero=False), raw.alloc(n * 4, zero=False)
    blk_a, blk_b = make_input(1 << 24, seed=2, lo=-1, hi=1), make_input(1 << 24, seed=3, lo=-1, hi=1)
    fill_tiled(raw, N, N.F32, a, n, blk_a)
    fill_tiled(raw, N, N.F32, b, n, blk_b)
    # configs[1]: binary add / mul, CUDA<Base>
    res["binary_add_f32_2^28"] = row(timeit(raw, lambda: raw.binary(N.F32, N.BIN_ADD, a, b, c, n)), n, 12)
    res["binary_mul_f32_2^28"] = row(timeit(raw, lambda: raw.binary(N.F32, N.BIN_MUL, a, b, c, n)), n, 12)
    # bit-exact against the oracle on sampled positions of the tiled inputs
    off = ((1 << 24) * 3 + 12345) if n >= (1 << 26) else 0  # (small --elems runs sample the start of the buffer)
    m = min(1 << 16, n)
    got = raw.d2h(c, m, N.F32, offset_bytes=4 * off)
    la, lb = blk_a[off % (1 << 24):off % (1 << 24) + m], blk_b[off % (1 << 24):off % (1 << 24) + m]
    res["binary_mul_f32_2^28"]["bit_exact_vs_oracle_sample"] = bool(np.array_equal(got, orc.binary(1, orc.F32, la, lb)))
    res["clear_f32_2^28"] = row(timeit(raw, lambda: raw.clear(N.F32, c, n)), n, 4)
    res["copy_f32_2^28"] = row(timeit(raw, lambda: raw.copy(N.F32, c, 0, a, 0, n)), n, 8)
    cheap = raw.compile(CHEAP8, N.F32)
    res["cheap8_f32_2^28"] = row(timeit(raw, lambda: raw.apply(cheap, a, c, n)), n, 8)
    g = raw.compile(CHAIN8_GRADS[3], N.F32, N.KERNEL_UNARY_GRAD)
    res["unary_grad_cos_f32_2^28"] = row(timeit(raw, lambda: raw.unary_grad(g, a, c, b, n)), n, 16)
    # configs[2], 16-bit: the fused chain on f16 / bf16 (large buffers take the table-lookup kernel)
    h16 = make_input(1 << 24, seed=4).astype(np.float16)
    fill_tiled(raw, N, N.F16, b, n, h16)
    e16 = raw.compile(CHAIN8, N.F16)
    res["chain8_f16_2^28"] = row(timeit(raw, lambda: raw.apply(e16, b, c, n)), n, 4, kernel="lut16_kernel (table filled by the arithmetic kernel)")
    got16 = raw.d2h(c, 1 << 16, N.F16)
    want16 = orc.apply_chain(CHAIN8, orc.F16, h16[:1 << 16])
    res["chain8_f16_2^28"]["max_abs_err_vs_oracle_sample"] = float(np.max(np.abs(got16.astype(np.float64) - want16.astype(np.float64))))
    raw.set_lut(e16, False)  # the arithmetic kernel (f32 math, a rounding after every op), for comparison
    res["chain8_f16_2^28_arithmetic_kernel"] = row(timeit(raw, lambda: raw.apply(e16, b, c, n)), n, 4)
    same = raw.d2h(c, 1 << 16, N.F16)
    res["chain8_f16_2^28"]["lookup_equals_arithmetic_kernel_sample"] = bool(same.tobytes() == got16.tobytes())
    raw.set_lut(e16, True)
    fill_tiled(raw, N, N.BF16, b, n, bf16_from_f32(make_input(1 << 24, seed=4)))
    eb = raw.compile(CHAIN8, N.BF16)
    res["chain8_bf16_2^28"] = row(timeit(raw, lambda: raw.apply(eb, b, c, n)), n, 4, kernel="lut16_kernel")
    # configs[0] on the GPU (1M f32, launch-latency bound) and on the CPU port
    c1 = raw.compile(CONFIG1, N.F32)
    ms = timeit(raw, lambda: raw.apply(c1, a, c, 1 << 20), reps=200, warm=20)
    res["config1_exp_sin_mul_add_1M_f32"] = dict(row(ms, 1 << 20, 8), us_per_launch=ms * 1e3, cpu_port=config1_cpu())
    for p in (a, b, c):
        raw.free(p)

    # configs[3]: sum / mean over 2^30 f32 (4 GiB), deterministic one-launch reduction; parity inside the run
    ns = SUM_ELEMS
    s_in, s_out = raw.alloc(ns * 4, zero=False), raw.alloc(64)
    block = np.random.default_rng(5).random(1 << 24, dtype=np.float32)
    fill_tiled(raw, N, N.F32, s_in, ns, block)
    ms = timeit(raw, lambda: raw.sum_into(N.F32, s_in, ns, s_out))
    got = raw.sum(N.F32, s_in, ns)
    plan = sum_plan(N.F32, ns)
    host = np.tile(block, ns // block.size)
    want = orc.sum_two_pass(orc.F32, host, plan["blocks"], plan["chunk"], plan["threads"], plan["vec"], plan["threads2"])
    truth = float(np.sum(block.astype(np.float64))) * (ns // block.size)
    res["sum_f32_2^30"] = dict(row(ms, ns, 4), sum=float(got), sum_parity=bool(got.tobytes() == want.tobytes()),
                               rel_err_vs_fp64=abs(float(got) - truth) / truth, launches_per_sum=1)
    del host
    mean = raw.mean(N.F32, s_in, ns)
    res["mean_f32_2^30"] = {"mean": float(mean), "mean_parity": bool(mean == np.float32(want / np.float32(ns)))}
    raw.free(s_in)
    raw.free(s_out)
    raw.close()

    # configs[2] with Autograd: forward + backward of the fused chain on the north-star stack — 2 kernels
    with CUDA("Lazy", "Graph", "Autograd", "Base", ordinal=local_rank) as d:
        buf = d.new_buffer(np.float32, n).require_grad()
        blk_x = make_input(1 << 24, seed=4)
        fill_tiled(d.raw, N, N.F32, buf.ptr(), n, blk_x)
        cur = buf
        for f, gr in zip(CHAIN8, CHAIN8_GRADS):
            cur = d.unary_ew(cur, f, gr)
        d.optimize_mem_graph()
        d.unary_fusing()
        d.set_graph_replay(True)
        d.run()
        cur.backward()
        l0 = d.raw.launches
        ms_b = timeit(d.raw, cur.backward)
        per_call = (d.raw.launches - l0) / 25
        res["chain8_bwd_fused_f32_2^28"] = row(ms_b, n, 16, launches_per_backward=per_call,
                                               kernel="cb_chain_grad_vec (recomputes the 7 intermediates, seed folded in)",
                                               replaces="8 add_unary_grad kernels + a seed fill: 8 x 16 + 4 B/element")

        def step():
            d.run()
            cur.backward()
        res["chain8_fwd_plus_bwd_f32_2^28"] = row(timeit(d.raw, step), n, 24, launches_per_step=2)
        # sampled gradient check against the oracle's op-by-op backward
        d.zero_grad()
        cur.backward()
        x0 = blk_x[:1 << 14]
        acts = [x0]
        for f in CHAIN8:
            acts.append(orc.apply_fn(f, orc.F32, acts[-1]))
        gg = np.ones(x0.size, np.float32)
        for k in reversed(range(8)):
            gg = orc.add_unary_grad(CHAIN8_GRADS[k], orc.F32, acts[k], np.zeros_like(x0), gg)
        got_g = d.raw.d2h(buf.grad().ptr(), x0.size, N.F32)
        err = np.abs(got_g.astype(np.float64) - gg.astype(np.float64))
        res["chain8_bwd_fused_f32_2^28"]["max_err_vs_oracle_sample"] = float(np.max(err / (1e-4 * np.abs(gg) + 2e-5)))

    # configs[2] on f16: the same stack typed by f16 (Lazy<Mods, f16>); forward and seeded backward are one table
    # lookup per element each (the tables hold what the arithmetic kernels compute, a rounding after every op)
    with CUDA("Lazy", "Graph", "Autograd", "Base", ordinal=local_rank, dtype=np.float16) as d:
        buf = d.new_buffer(np.float16, n).require_grad()
        fill_tiled(d.raw, N, N.F16, buf.ptr(), n, blk_x.astype(np.float16))
        cur = buf
        for f, gr in zip(CHAIN8, CHAIN8_GRADS):
            cur = d.unary_ew(cur, f, gr)
        d.optimize_mem_graph()
        d.unary_fusing()
        d.set_graph_replay(True)
        d.run()
        cur.backward()
        res["chain8_fwd_f16_2^28_module_stack"] = row(timeit(d.raw, d.run), n, 4)
        res["chain8_bwd_fused_f16_2^28"] = row(timeit(d.raw, cur.backward), n, 8, launches_per_backward=1,
                                               kernel="lut16_kernel<grad>: seeded with ones the backward term is a function of x alone "
                                                      "(table filled by cb_chain_grad_vec itself)")

    # configs[4]: Cached+Lazy CUDA-graph replay of a 20-op sequence on 4K-element buffers (launch-latency bound)
    x4k = make_input(4096, seed=70, lo=-1, hi=1)
    rep = {}
    for mode in ("eager_launches", "graph_replay"):
        with CUDA("Lazy", "Cached", "Base", ordinal=local_rank) as d:
            d.set_graph_replay(mode == "graph_replay")
            pa, pb_ = d.buffer(x4k), d.buffer(x4k)
            cur = pa
            for k in range(10):
                cur = d.apply_fn(cur, CHAIN8[k % 8])
                cur = d.add(cur, pb_)
            d.run()
            d.sync()
            reps = 3000
            t = time.perf_counter()
            for _ in range(reps):
                d.run()
            d.sync()
            rep[mode + "_us_per_run"] = (time.perf_counter() - t) / reps * 1e6
    rep["speedup"] = rep["eager_launches_us_per_run"] / rep["graph_replay_us_per_run"]
    res["replay_20op_4096_f32"] = rep
    return res


def multi_gpu_rows(dev, world, rank, local_rank, n, barrier, max_over_ranks, dist, peak):
    """N > 1: strong scaling (ONE 2^28 buffer / ONE 2^30 sum split N ways) and the timed collective."""
    from custos_b200 import _native as N
    from custos_b200.raw import Comm, shard_range, sum_plan
    from custos_b200.workloads import CHAIN8
    from oracle import oracle as orc
    raw = dev.raw
    out = {}
    # ---- strong scaling of the fused chain: rank 0 also times the whole buffer alone (T_1 of the same run)
    chain = raw.compile(CHAIN8, N.F32)
    a, c = raw.alloc(n * 4, zero=False), raw.alloc(n * 4, zero=False)
    fill_tiled(raw, N, N.F32, a, n, make_input(1 << 24, seed=4))
    t1 = timeit(raw, lambda: raw.apply(chain, a, c, n)) if rank == 0 else 0.0
    b, e = shard_range(n, 4, world, rank)
    barrier()
    tn = max_over_ranks(timeit(raw, lambda: raw.apply(chain, a + 4 * b, c + 4 * b, e - b)))
    t1 = max_over_ranks(t1)
    out["strong_chain8_f32_2^28_total"] = {"ms_1gpu": t1, "ms_Ngpus": tn, "speedup": t1 / tn, "efficiency": t1 / tn / world,
                                           "GB/s_aggregate": n * 8 / (tn * 1e-3) / 1e9, "collective": "none (slices)"}
    raw.free(a)
    raw.free(c)
    # ---- sharded sum of 2^30 f32: one fused reduce + exchange kernel per rank (peer memory) vs NCCL all-gather + fold
    ns = SUM_ELEMS
    sb, se = shard_range(ns, 4, world, rank)
    nl = se - sb
    block = np.random.default_rng(5).random(1 << 24, dtype=np.float32)
    s_in, s_out = raw.alloc(nl * 4, zero=False), raw.alloc(64)
    # the slice starts at a multiple of the block for every N in {2, 4, 8}: tile the block over it
    fill_tiled(raw, N, N.F32, s_in, nl, np.roll(block, -(sb % block.size)))
    t1s = 0.0
    if rank == 0:  # T_1: the whole 4 GiB buffer on one GPU
        full = raw.alloc(ns * 4, zero=False)
        fill_tiled(raw, N, N.F32, full, ns, block)
        t1s = timeit(raw, lambda: raw.sum_into(N.F32, full, ns, s_out))
        raw.free(full)
    t1s = max_over_ranks(t1s)
    rows = {}
    for name, env in (("peer_memory", None), ("nccl_all_gather", "0")):
        if env is not None:
            os.environ["CB_COMM_P2P"] = env
        uid = [Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = Comm(raw, world, rank, uid[0])
        os.environ.pop("CB_COMM_P2P", None)
        barrier()
        ms = max_over_ranks(timeit(raw, lambda: comm.sum_into(N.F32, s_in, nl, s_out), reps=50, warm=10))
        got = comm.sum(N.F32, s_in, nl)  # synchronises and checks the exchange status
        rows[name] = {"ms": ms, "GB/s_aggregate": ns * 4 / (ms * 1e-3) / 1e9, "speedup_vs_1gpu": t1s / ms,
                      "frac_of_N_x_measured_peak": ns * 4 / (ms * 1e-3) / 1e9 / (peak * world),
                      "uses_peer_memory": bool(comm.uses_peer_memory), "sum": float(got), "bits": got.tobytes().hex()}
        comm.close()
    # parity inside the run: every rank restates its slice's partial with the oracle, the partials are gathered on the
    # host and folded in rank order — the device result must have exactly those bits, on every rank
    host = np.tile(np.roll(block, -(sb % block.size)), nl // block.size + 1)[:nl]
    plan = sum_plan(N.F32, nl)
    part = orc.sum_two_pass(orc.F32, host, plan["blocks"], plan["chunk"], plan["threads"], plan["vec"], plan["threads2"])
    parts = [None] * world
    dist.all_gather_object(parts, (float(part), rows["peer_memory"]["bits"], rows["nccl_all_gather"]["bits"]))
    want = np.float32(parts[0][0])
    for p in parts[1:]:
        want = np.float32(want + np.float32(p[0]))
    parity = all(p[1] == want.tobytes().hex() and p[2] == want.tobytes().hex() for p in parts)
    for r in rows.values():
        r.pop("bits")
    out["sharded_sum_f32_2^30"] = {"ms_1gpu": t1s, "elements_per_gpu": nl, **rows, "sum_parity": bool(parity),
                                   "parity_how": "device result == rank-ordered f32 fold of the oracle's two-pass partial of "
                                                 "every slice, bit for bit, on every rank, through both exchange paths"}
    raw.free(s_in)
    raw.free(s_out)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from custos_b200 import _native as N
    from custos_b200.build import build
    from custos_b200.device import CUDA
    from custos_b200.workloads import CHAIN8

    build()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.elems
    nbytes = n * 4
    # ------------------------------------------------------------ record the chain on the north-star module stack
    dev = CUDA("Lazy", "Graph", "Autograd", "Base", ordinal=local_rank)
    raw = dev.raw
    # pinned: the user's data lives on the host (e2e); CB_BENCH_WC=1 makes the input write-combined (A/B)
    h_in, h_out = raw.host_alloc(nbytes, os.environ.get("CB_BENCH_WC") == "1"), raw.host_alloc(nbytes)
    host_in = np.ctypeslib.as_array((ctypes.c_float * n).from_address(h_in))
    host_out = np.ctypeslib.as_array((ctypes.c_float * n).from_address(h_out))
    make_input(n, seed=4 + rank, out=host_in)
    x = dev.new_buffer(np.float32, n)
    raw.h2d_async(x.ptr(), h_in, nbytes)
    raw.sync()
    cur = x
    for f in CHAIN8:
        cur = dev.apply_fn(cur, f)          # eight recorded ops (Lazy), eight graph nodes (Graph)
    assert dev.ops_count() == 8
    dev.optimize_mem_graph()                 # the trace x1..x8 shares ONE allocation
    dev.unary_fusing()                       # ... and becomes ONE kernel
    dev.set_graph_replay(True)               # run() = one cudaGraphLaunch
    dev.run()                                # allocates, captures, launches
    raw.sync()
    assert dev.replay_kernel_nodes() == 1, "the recorded chain did not fuse into one kernel"
    fused_index = next(i for i in range(8) if dev.op_expr(i) is not None)
    chain = dev.op_expr(fused_index)         # the very kernel run() replays, for the host-operand path

    # ---------------------------------------------------------------- device-resident timing
    sampler = ClockSampler(local_rank)
    sampler.wait_ready()
    for _ in range(args.warmup):
        dev.run()
    raw.sync()
    barrier()
    launches0 = raw.launches
    ev0, ev1 = raw.event(), raw.event()
    t0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        dev.run()
    ev1.record()
    ev1.sync()
    raw.sync()
    t1 = time.time()
    launches = raw.launches - launches0
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_ms(ev1))
    clocks = sampler.window(t0, t1, "NVML polled at ~1 kHz during the timed steps")
    ms_per_step = ms_total / args.steps
    value = world * n * BYTES_PER_ELEM / (ms_per_step * 1e-3) / 1e9
    fused_out = raw.d2h(cur.replace().ptr(), 4096, N.F32)

    # the same step back to back for >= 1 s: what a long-running job sees once the 1 kW power cap has pulled the SM
    # clock down (the target says "sustains"; reported in `roofline` beside the burst figure)
    sus_steps = max(args.steps, int(1.0 / (ms_per_step * 1e-3)))
    for _ in range(sus_steps // 2):
        dev.run()
    e0, e1 = raw.event(), raw.event()
    ts0 = time.time()
    e0.record()
    for _ in range(sus_steps):
        dev.run()
    e1.record()
    e1.sync()
    ts1 = time.time()
    sus_ms = max_over_ranks(e0.elapsed_ms(e1)) / sus_steps
    sus_clocks = sampler.window(ts0, ts1, "NVML during the sustained loop (after 0.5 s of the same load)")
    sampler.stop()

    # ---------------------------------------------------------------- end to end (host buffers)
    # the call a user with host data makes: cb_apply_host(expr, host_in, host_out, n) — H2D of the
    # input and D2H of the result are inside the timed region (chunked, overlapped with the kernel);
    # wall clock around the synchronous call, max over ranks
    e2e_steps = max(3, min(args.steps, 10))
    host_out[:] = 0
    for _ in range(2):
        raw.apply_host(chain, h_in, h_out, n)
    barrier()
    launches_e2e0 = raw.launches
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        raw.apply_host(chain, h_in, h_out, n)
    raw.sync()
    te1 = time.perf_counter()
    launches_e2e = raw.launches - launches_e2e0
    barrier()
    e2e_ms = max_over_ranks((te1 - te0) * 1e3) / e2e_steps
    e2e_value = world * n * BYTES_PER_ELEM / (e2e_ms * 1e-3) / 1e9
    # the ceiling of that number is the host link, not HBM: measure it with plain pinned copies, all ranks at once
    try:
        pcie = pcie_ceiling(torch, barrier, max_over_ranks)
        pcie["e2e_frac_of_bidir"] = (e2e_value / world) / pcie["bidir_gbs"]
        pcie["aggregate_bidir_gbs"] = pcie["bidir_gbs"] * world
    except Exception as e:  # noqa: BLE001
        pcie = {"error": repr(e)}

    # sanity: the timed kernel really computed the chain (sampled check against the oracle), and the module stack's
    # device-resident result has the bits of the host-operand path
    check = None
    if rank == 0:
        from oracle import oracle as orc
        idx = np.random.default_rng(0).integers(0, n, 4096)
        want = orc.apply_chain(CHAIN8, orc.F32, host_in[idx])
        check = float(np.max(np.abs(host_out[idx].astype(np.float64) - want.astype(np.float64))))
        assert np.array_equal(fused_out.view(np.uint32), host_out[:4096].view(np.uint32)), "run() and apply_host disagree"

    peak, peak_src = measured_peak()
    achieved = n * BYTES_PER_ELEM / (ms_per_step * 1e-3) / 1e9  # per GPU, the dominant (only) kernel
    sustained = n * BYTES_PER_ELEM / (sus_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic()
    out = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(world, n),
        "step": "CUDA<Lazy<Graph<Autograd<Base>>>>: 8 recorded apply_fn -> optimize_mem_graph -> unary_fusing -> run() = "
                "cudaGraphLaunch of ONE kernel node",
        "elements_per_s": world * n / (ms_per_step * 1e-3), "sampled_check_max_abs_err": check,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "frac_of_nominal_8000": achieved / 8000.0,
                     "achieved_sustained": sustained, "frac_sustained": sustained / peak,
                     "sustained_steps": sus_steps, "sustained_ms_per_step": sus_ms, "sustained_clocks": sus_clocks,
                     "kernel": "cb_apply_vec (NVRTC, fused CHAIN8), replayed from the Lazy module's CUDA graph",
                     "algorithmic_bytes_per_launch": n * BYTES_PER_ELEM},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                "ms_per_step": e2e_ms, "steps": e2e_steps, "gpu_launches": int(launches_e2e),
                "api": "cb_apply_host(the fused expression of the module stack): pinned host buffers, 16 MiB chunks, "
                       "H2D / kernel / D2H on three streams",
                "host_link": pcie},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    raw.host_free(h_in)
    raw.host_free(h_out)
    if world > 1 and not args.no_configs:
        try:
            out["multi_gpu"] = multi_gpu_rows(dev, world, rank, local_rank, n, barrier, max_over_ranks, dist, peak)
        except Exception as e:  # noqa: BLE001
            out["multi_gpu"] = {"error": repr(e)}
    dev.close()
    if rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_baselines()
        if not args.no_configs:
            try:
                out["configs"] = other_configs(local_rank, n, peak)
            except Exception as e:  # noqa: BLE001
                out["configs"] = {"error": repr(e)}
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elems", type=int, default=N_ELEMS, help="elements per GPU (default 2^28, the BASELINE config)")
    ap.add_argument("--no-configs", action="store_true", help="only the headline (skip the other BASELINE configs / multi-GPU rows)")
    ap.add_argument("--extra", action="store_true", help=argparse.SUPPRESS)  # kept for older scripts: the rows are on by default
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
