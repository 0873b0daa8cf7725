#!/usr/bin/env python
"""Every one of the 2^32 f32 bit patterns through a transcendental kernel, against the oracle (glibc) on the host cores.

    python scripts/exhaustive_f32.py [exp ln sin cos tanh tan] > profiles/r1_exhaustive_f32.json

Per function: the maximum ulp distance over ALL inputs, where it occurs, the histogram of distances, and the same
restricted to the domain the <= 4 ulp bar is stated on (sin / cos / tan: |x| <= 1e9; everything else: all inputs).
The oracle runs multi-threaded here (it is the checker, not the thing measured).
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from custos_b200 import _native as N  # noqa: E402
from custos_b200.raw import RawDevice  # noqa: E402
from oracle import oracle as orc  # noqa: E402

FUNCS = {"exp": lambda x: x.exp(), "ln": lambda x: x.ln(), "sin": lambda x: x.sin(), "cos": lambda x: x.cos(),
         "tanh": lambda x: x.tanh(), "tan": lambda x: x.tan()}
CHUNK = 1 << 26


def main():
    names = sys.argv[1:] or ["exp", "ln", "sin", "cos", "tanh"]
    threads = os.cpu_count() or 1
    dev = RawDevice(0)
    p_in, p_out = dev.alloc(CHUNK * 4, zero=False), dev.alloc(CHUNK * 4, zero=False)
    report = {"threads": threads, "chunk": CHUNK, "functions": {}}
    for name in names:
        f = FUNCS[name]
        e = dev.compile(f, N.F32)
        t0 = time.time()
        worst_all, where_all, hist_all = 0, 0, [0] * 6
        worst_dom, where_dom, hist_dom = 0, 0, [0] * 6
        zero_sign_mismatch = 0
        for start in range(0, 1 << 32, CHUNK):
            bits = np.arange(start, start + CHUNK, dtype=np.uint64).astype(np.uint32)
            x = bits.view(np.float32)
            dev.h2d(p_in, x)
            dev.apply(e, p_in, p_out, CHUNK)
            got = dev.d2h(p_out, CHUNK, N.F32)
            want = orc.apply_chain([f], orc.F32, x, threads=threads)
            w, i, h = orc.ulp_stats_f32(got, want)
            hist_all = [a + b for a, b in zip(hist_all, h)]
            if w > worst_all:
                worst_all, where_all = w, start + i
            zero_sign_mismatch += int(np.count_nonzero((got == 0) & (want == 0) & (np.signbit(got) != np.signbit(want))))
            if name in ("sin", "cos", "tan"):
                ax = np.abs(x)
                if float(ax.min()) > 1e9 and not np.isnan(ax.min()):
                    continue  # the whole chunk is outside the stated domain
                keep = ~(ax > np.float32(1e9))  # keeps NaN inputs (NaN in, NaN out)
                if not keep.all():
                    got, want, idx = got[keep], want[keep], np.flatnonzero(keep)
                    w, i, h = orc.ulp_stats_f32(got, want)
                    i = int(idx[i]) if idx.size else 0
            hist_dom = [a + b for a, b in zip(hist_dom, h)]
            if w > worst_dom:
                worst_dom, where_dom = w, start + i
        report["functions"][name] = {
            "inputs": 1 << 32, "max_ulp_all_inputs": worst_all, "worst_input_bits_all": f"{where_all:#010x}",
            "hist_0_1_2_3_4_more_all": hist_all,
            "domain": "|x| <= 1e9 (and NaN)" if name in ("sin", "cos", "tan") else "all inputs",
            "max_ulp_domain": worst_dom, "worst_input_bits_domain": f"{where_dom:#010x}", "hist_0_1_2_3_4_more_domain": hist_dom,
            "signed_zero_mismatches": zero_sign_mismatch, "seconds": round(time.time() - t0, 1)}
        print(name, json.dumps(report["functions"][name]), file=sys.stderr, flush=True)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
