#!/usr/bin/env python
"""Aggregate an ncu report's source page per CUDA source line: tools/ncu_lines.py report.ncu-rep kernel_regex [launch_skip]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern,
                      "--launch-skip", skip, "--launch-count", "1"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
agg = []
hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] in ('Function Name',) or hdr is None: continue
    if r[0] == '': continue
    try:
        ie = hdr.index('Instructions Executed'); sa = hdr.index('# Samples')
        agg.append((int(r[ie]), int(r[sa]), cur_file, r[0], r[1].strip()[:120]))
    except (ValueError, IndexError):
        pass
tot = sum(a[0] for a in agg) or 1; ts = sum(a[1] for a in agg) or 1
print("total warp-instructions %d, samples %d" % (tot, ts))
for n, s, f, l, src in sorted(agg, reverse=True)[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print("%5.1f%% inst %5.1f%% samp  %s:%s  %s" % (100.0 * n / tot, 100.0 * s / ts, f, l, src))
