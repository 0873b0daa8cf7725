#!/usr/bin/env python
"""Markdown table of the per-kernel roofline metrics in an .ncu-rep (last launch of each kernel/grid)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
col = h.index
SC = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}
TS = {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
last = {}
for r in rows[2:]:
    name = r[col("Kernel Name")].split("(")[0].replace("void ", "").replace("cb::<unnamed>::", "").replace("unnamed>::", "")
    last[(name, r[col("launch__grid_size")])] = r
print("| kernel | grid | regs | duration µs | DRAM read GB | DRAM write GB | DRAM GB/s | warp instr. (M) | issue active % | fmaheavy % | XU % | warps active % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for (name, grid), r in last.items():
    d = float(r[col("gpu__time_duration.sum")]) * TS[units[col("gpu__time_duration.sum")]]
    rd = float(r[col("dram__bytes_read.sum")]) * SC[units[col("dram__bytes_read.sum")]]
    wr = float(r[col("dram__bytes_write.sum")]) * SC[units[col("dram__bytes_write.sum")]]
    g = lambda n: float(r[col(n)])
    print(f"| {name[:58]} | {grid} | {r[col('launch__registers_per_thread')]} | {d:.1f} | {rd:.4f} | {wr:.4f} | {(rd + wr) / (d * 1e-6):.0f} | "
          f"{g('smsp__inst_executed.sum') / 1e6:.1f} | {g('sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{g('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | {g('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} |")
