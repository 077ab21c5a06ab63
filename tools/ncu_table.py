#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into one line per launch (markdown table).
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_table.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = [("ms", "gpu__time_duration.sum", 1e-3), ("regs", "launch__registers_per_thread", 1),
        ("dram_rd_MB", "dram__bytes_read.sum", 1), ("dram_wr_MB", "dram__bytes_write.sum", 1),
        ("dram%", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
        ("fp64%", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1),
        ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
        ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
        ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("Minst", "smsp__inst_executed.sum", 1e-6),
        ("l2hit%", "lts__t_sector_hit_rate.pct", 1)]
units = rows[1]
print("| kernel | grid | " + " | ".join(w[0] for w in want) + " |")
print("|---|---|" + "---|" * len(want))
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("void ", "").replace("genpf::", "").split("(")[0][:48]
    grid = r[col.get("Grid Size", col.get("launch__grid_size", 0))]
    vals = []
    for label, key, scale in want:
        if key not in col:
            vals.append("-")
            continue
        v = r[col[key]].replace(",", "")
        try:
            x = float(v) * scale
            u = units[col[key]]
            if label.endswith("_MB"):
                x = x * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
            if label == "ms":
                x = float(v) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(u, 1e-3)
            vals.append(f"{x:.3f}" if abs(x) < 100 else f"{x:.0f}")
        except ValueError:
            vals.append(v)
    print(f"| {name} | {grid} | " + " | ".join(vals) + " |")
