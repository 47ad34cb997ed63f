#!/usr/bin/env python
"""Per-SASS-instruction executed counts (and stall samples) of the first kernel in an .ncu-rep, in address order.
usage: ncu_sass_hot.py rep [min_count]"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
for i, r in enumerate(rows):
    if 'Source' in r and 'Instructions Executed' in r:
        hdr, start = r, i + 1
        break
si, ei = hdr.index('Source'), hdr.index('Instructions Executed')
sa = hdr.index('Warp Stall Sampling (All Samples)') if 'Warp Stall Sampling (All Samples)' in hdr else None
mn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for n, r in enumerate(rows[start:]):
    if len(r) <= ei:
        continue
    try:
        c = int(r[ei].replace(',', ''))
    except ValueError:
        continue
    if c >= mn:
        print(f"{n:5d} {c:10d} {r[sa] if sa is not None else '':>8s}  {r[si].strip()}")
