#!/usr/bin/env python
"""Executed-instruction mix of the first kernel in an .ncu-rep (needs --import-source on / --set full)."""
import csv, collections, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
for i, r in enumerate(rows):
    if 'Source' in r and 'Instructions Executed' in r:
        hdr, start = r, i + 1
        break
si, ei = hdr.index('Source'), hdr.index('Instructions Executed')
agg, tot = collections.Counter(), 0
top = []
for r in rows[start:]:
    if len(r) <= ei:
        continue
    try:
        n = int(r[ei].replace(',', ''))
    except ValueError:
        continue
    op = r[si].strip().split()
    if not op:
        continue
    o = op[1] if op[0].startswith('@') and len(op) > 1 else op[0]
    full = o
    o = o.split('.')[0]
    if o in ('IMAD', 'MOV') and ('.MOV' in full or o == 'MOV'):
        o = 'MOV*'
    agg[o] += n
    tot += n
print('total warp instructions', tot)
for o, n in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
    print(f"{o:12s} {n:12d} {100 * n / tot:5.1f}%")
