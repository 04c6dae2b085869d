#!/usr/bin/env python
"""Markdown summary of an ncu report (one row per profiled launch): tools/ncu_summary.py report.ncu-rep > profiles/x.md"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
        ("smsp__inst_executed.sum", "warp inst"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
ki = hdr.index("Kernel Name")
print("| kernel | " + " | ".join("%s (%s)" % (n, units[i]) if units[i] else n for i, n in idx) + " |")
print("|---|" + "---|" * len(idx))
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    vals = []
    for i, n in idx:
        try:
            v = float(r[i].replace(",", ""))
            vals.append("%.4g" % v)
        except ValueError:
            vals.append(r[i])
    print("| %s | " % name + " | ".join(vals) + " |")
