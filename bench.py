#!/usr/bin/env python
"""bench.py — the headline benchmark of BASELINE.json on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--extra]

Metric: achieved HBM GB/s of the fused 8-op unary chain on 2^28 f32
per GPU.  One "step" = one pass of the fused chain kernel over the resident buffer (the
buffers are 1 GiB each, far larger than the 126 MB L2, so no flush is needed between steps).
`value` is device-timed (CUDA events on the library's stream) with inputs resident in HBM;
`e2e` is the same work through the C-ABI with HOST buffers (pinned), H2D and D2H inside the
timed region.  Weak scaling: every rank owns its own 2^28-element slice, no collective on the
data path (max-over-ranks timing through torch.distributed/NCCL).

`--impl reference` times the reference's CPU device (restated in oracle/, the Rust crate cannot
be built here) on the host cores for the same metric/config on a bounded sample.
Prints exactly one JSON line on stdout.
"""
from __future__ import annotations

import argparse
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
BYTES_PER_ELEM = 8  # f32: 1 read + 1 write per element, whatever the chain length
METRIC = "achieved HBM GB/s, fused 8-op unary chain, 2^28 f32 per GPU"


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel, from the committed ncu capture (or None)."""
    p = ROOT / "profiles" / "roofline_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("chain8_f32_traffic_bytes")
        except Exception:
            return None
    return None


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


def make_input(n: int, seed: int = 4) -> np.ndarray:
    """x ~ U[-4, 4) f32, generated in chunks to bound host memory."""
    rng = np.random.default_rng(seed)
    out = np.empty(n, np.float32)
    step = 1 << 24
    for i in range(0, n, step):
        m = min(step, n - i)
        out[i:i + m] = rng.uniform(-4.0, 4.0, m).astype(np.float32)
    return out


def pcie_ceiling(torch, nbytes: int = 1 << 28, reps: int = 4):
    """What the host link of this box moves with plain pinned-memory copies — the ceiling of the end-to-end
    number: H2D alone, D2H alone, and both directions at once (GB/s, CUDA events)."""
    h_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    cur = torch.cuda.current_stream()

    def timed(h2d: bool, d2h: bool) -> float:
        torch.cuda.synchronize()
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
        return (int(h2d) + int(d2h)) * reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9

    timed(True, True)  # warm-up
    return {"h2d_gbs": timed(True, False), "d2h_gbs": timed(False, True), "bidir_gbs": timed(True, True),
            "how": f"torch pinned copies of {nbytes >> 20} MiB x {reps}, one stream per direction, CUDA events"}


def cpu_baseline_single(sample: int):
    """The oracle (kind "port") on ONE host core: the reference CPU device is single threaded."""
    from custos_b200.workloads import CHAIN8
    from oracle import oracle as orc
    x = make_input(sample)
    orc.apply_chain(CHAIN8, orc.F32, x[:1 << 16])
    t = time.perf_counter()
    orc.apply_chain(CHAIN8, orc.F32, x)
    dt = time.perf_counter() - t
    return {"value": sample * BYTES_PER_ELEM / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": f"fused CHAIN8 over 2^{int(np.log2(sample))} f32 (U[-4,4), seed 4), oracle/ C port of the reference "
                      f"CPU device, single thread as in the reference; {sample / dt / 1e6:.1f} M elem/s"}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from custos_b200.workloads import CHAIN8
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    sample = 1 << 26
    x = make_input(sample)
    for _ in range(max(args.warmup, 1)):
        orc.apply_chain(CHAIN8, orc.F32, x[:1 << 22], threads=cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        orc.apply_chain(CHAIN8, orc.F32, x, threads=cores)
    dt = (time.perf_counter() - t) / args.steps
    val = sample * BYTES_PER_ELEM / dt / 1e9
    desc = (f"each step = fused CHAIN8 over a 2^26-element sample of the 2^28 f32 workload, oracle/ C port of the "
            f"reference CPU device (the Rust crate cannot be built here), split over {cores} host threads "
            f"(the reference itself is single threaded)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3 * (N_ELEMS / sample), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "chain8_f32_2^28_per_gpu", "sample_elems": sample, "elements_per_s": sample / dt},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist

    from custos_b200 import _native as N
    from custos_b200.build import build
    from custos_b200.raw import RawDevice
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

    dev = RawDevice(local_rank)
    n = args.elems
    nbytes = n * 4
    chain = dev.compile(CHAIN8, N.F32)

    # pinned host buffers: the user's data lives on the host in the e2e path
    h_in, h_out = dev.host_alloc(nbytes), dev.host_alloc(nbytes)
    import ctypes
    host_in = np.ctypeslib.as_array((ctypes.c_float * n).from_address(h_in))
    host_out = np.ctypeslib.as_array((ctypes.c_float * n).from_address(h_out))
    host_in[:] = make_input(n, seed=4 + rank)
    d_in, d_out = dev.alloc(nbytes, zero=False), dev.alloc(nbytes, zero=False)
    dev.h2d_async(d_in, h_in, nbytes)
    dev.sync()

    # ---------------------------------------------------------------- device-resident timing
    sampler = ClockSampler(local_rank)
    sampler.wait_ready()
    for _ in range(args.warmup):
        dev.apply(chain, d_in, d_out, n)
    dev.sync()
    barrier()
    launches0 = dev.launches
    ev0, ev1 = dev.event(), dev.event()
    t0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        dev.apply(chain, d_in, d_out, n)
    ev1.record()
    ev1.sync()
    dev.sync()
    t1 = time.time()
    launches = dev.launches - launches0
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_ms(ev1))
    clocks = sampler.window(t0, t1, "NVML polled at ~1 kHz during the timed steps")
    ms_per_step = ms_total / args.steps
    value = world * n * BYTES_PER_ELEM / (ms_per_step * 1e-3) / 1e9

    # the same kernel back to back for ~1 s: what a long-running job sees once the 1 kW power cap
    # has pulled the SM clock down (reported beside the headline, never instead of it)
    sus_steps = max(args.steps, int(1.0 / (ms_per_step * 1e-3)))
    for _ in range(sus_steps // 2):
        dev.apply(chain, d_in, d_out, n)
    e0, e1 = dev.event(), dev.event()
    ts0 = time.time()
    e0.record()
    for _ in range(sus_steps):
        dev.apply(chain, d_in, d_out, n)
    e1.record()
    e1.sync()
    ts1 = time.time()
    sus_ms = max_over_ranks(e0.elapsed_ms(e1)) / sus_steps
    sustained = {"value": world * n * BYTES_PER_ELEM / (sus_ms * 1e-3) / 1e9, "unit": "GB/s", "steps": sus_steps,
                 "ms_per_step": sus_ms,
                 "clocks": sampler.window(ts0, ts1, "NVML during the sustained loop (after 0.5 s of the same load)")}
    sampler.stop()

    # ---------------------------------------------------------------- end to end (host buffers)
    # the call a user with host data makes: cb_apply_host(expr, host_in, host_out, n) — H2D of the
    # input and D2H of the result are inside the timed region (chunked, overlapped with the kernel);
    # wall clock around the synchronous call, max over ranks
    e2e_steps = max(3, min(args.steps, 10))
    host_out[:] = 0
    for _ in range(2):
        dev.apply_host(chain, h_in, h_out, n)
    barrier()
    launches_e2e0 = dev.launches
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        dev.apply_host(chain, h_in, h_out, n)
    dev.sync()
    te1 = time.perf_counter()
    launches_e2e = dev.launches - launches_e2e0
    barrier()
    e2e_ms = max_over_ranks((te1 - te0) * 1e3) / e2e_steps
    e2e_value = world * n * BYTES_PER_ELEM / (e2e_ms * 1e-3) / 1e9
    # the ceiling of that number is the host link, not HBM: measure it with plain pinned copies
    try:
        barrier()
        pcie = pcie_ceiling(torch)
        pcie["e2e_frac_of_bidir"] = (e2e_value / world) / pcie["bidir_gbs"]
    except Exception as e:  # noqa: BLE001
        pcie = {"error": repr(e)}

    # sanity: the timed kernel really computed the chain (sampled check against the oracle)
    check = None
    if rank == 0:
        from oracle import oracle as orc
        idx = np.random.default_rng(0).integers(0, n, 4096)
        want = orc.apply_chain(CHAIN8, orc.F32, host_in[idx])
        got = host_out[idx]
        check = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))))

    peak, peak_src = measured_peak()
    achieved = n * BYTES_PER_ELEM / (ms_per_step * 1e-3) / 1e9  # per GPU, the dominant (only) kernel
    out = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "chain8_f32_2^28_per_gpu" if n == N_ELEMS else f"chain8_f32_{n}_per_gpu",
                   "chain": "add(1) mul(0.5) exp sin mul(2) add(1) tanh neg", "elements_per_gpu": n,
                   "elements_per_s": world * n / (ms_per_step * 1e-3), "l2": "inputs (1 GiB in + 1 GiB out) larger than L2",
                   "parallelism": f"slice{world}", "sampled_check_max_abs_err": check},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(), "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
                     "kernel": "cb_apply_vec (NVRTC, fused CHAIN8)", "algorithmic_bytes_per_launch": n * BYTES_PER_ELEM},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                "ms_per_step": e2e_ms, "steps": e2e_steps, "gpu_launches": int(launches_e2e),
                "api": "cb_apply_host: pinned host buffers, 16 MiB chunks, H2D / kernel / D2H on three streams",
                "host_link": pcie},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "sustained": sustained,
    }
    if rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_baseline_single(1 << 27)
    if rank == 0 and args.extra:
        out["extra"] = extra_workloads(dev, n)
    dev.host_free(h_in)
    dev.host_free(h_out)
    dev.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def extra_workloads(dev, n):
    """The other rows of BASELINE.md §4 (not bench lines of the contract; for DESIGN.md / profiles)."""
    from custos_b200 import _native as N
    from custos_b200.workloads import CHAIN8, CHAIN8_GRADS, CHEAP8
    peak, _ = measured_peak()
    res = {}

    def timeit(fn, reps=20, warm=5):
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

    a, b, c = dev.alloc(n * 4), dev.alloc(n * 4), dev.alloc(n * 4)
    dev.fill(N.F32, a, n, 0.5)
    dev.fill(N.F32, b, n, 0.25)

    def row(name, ms, bytes_per_elem, elems=n):
        gbs = elems * bytes_per_elem / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "GB/s": gbs, "frac_of_measured_peak": gbs / peak, "elements_per_s": elems / (ms * 1e-3)}

    cheap = dev.compile(CHEAP8, N.F32)
    row("cheap8_f32", timeit(lambda: dev.apply(cheap, a, c, n)), 8)
    chain = dev.compile(CHAIN8, N.F32)
    row("chain8_f32", timeit(lambda: dev.apply(chain, a, c, n)), 8)
    h16 = np.random.default_rng(4).uniform(-4, 4, 1 << 24).astype(np.float16)
    ph = dev.upload(h16)
    for off in range(0, n, 1 << 24):  # proper binary16 inputs in `c`'s first half, results into `b`
        dev.copy(N.F16, c, off, ph, 0, min(1 << 24, n - off))
    dev.free(ph)
    chain16 = dev.compile(CHAIN8, N.F16)
    row("chain8_f16", timeit(lambda: dev.apply(chain16, c, b, n)), 4)
    dev.fill(N.F32, b, n, 0.25)
    row("binary_add_f32", timeit(lambda: dev.binary(N.F32, N.BIN_ADD, a, b, c, n)), 12)
    row("binary_mul_f32", timeit(lambda: dev.binary(N.F32, N.BIN_MUL, a, b, c, n)), 12)
    g = dev.compile(CHAIN8_GRADS[3], N.F32, N.KERNEL_UNARY_GRAD)
    row("unary_grad_cos_f32", timeit(lambda: dev.unary_grad(g, a, c, b, n)), 16)
    row("clear_f32", timeit(lambda: dev.clear(N.F32, c, n)), 4)
    row("copy_f32", timeit(lambda: dev.copy(N.F32, c, 0, a, 0, n)), 8)
    s = dev.alloc(64)
    row("sum_f32", timeit(lambda: dev.sum_into(N.F32, a, n, s)), 4)
    for p in (a, b, c, s):
        dev.free(p)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elems", type=int, default=N_ELEMS, help="elements per GPU (default 2^28, the BASELINE config)")
    ap.add_argument("--extra", action="store_true", help="also time the other BASELINE.md rows (adds an 'extra' object)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
