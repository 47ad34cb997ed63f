#!/usr/bin/env python
"""Where a kernel's instructions and stall samples go: the SASS of the first kernel in an .ncu-rep (--set full,
--import-source on) cut into runs of lines with the same execution count (= loop nests), with each run's share of the
executed warp-instructions, of the stall samples, and its opcode mix.  No GPU needed.
    python scripts/ncu_segments.py gpurun_out/x.ncu-rep [items]      # items: divide counts by this (e.g. 38400 work items)"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
items = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
for i, r in enumerate(rows):
    if "Source" in r and "Instructions Executed" in r:
        hdr, start = r, i + 1
        break
si, ei, st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
out = []
for j, r in enumerate(rows[start:]):
    try:
        out.append((j, int(r[ei].replace(",", "")), int(r[st].replace(",", "") or 0), r[si].strip()))
    except (ValueError, IndexError):
        continue
op = lambda t: (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
tot, tots = sum(x[1] for x in out), max(1, sum(x[2] for x in out))
print(f"executed warp-instructions {tot}  ({tot / items:.0f} per item)   stall samples {tots}")
segs, seg, prev = [], [], None
for x in out:
    if prev is not None and abs(x[1] - prev) > 0.03 * max(x[1], prev, 1):
        segs.append(seg)
        seg = []
    seg.append(x)
    prev = x[1]
segs.append(seg)
for seg in segs:
    ti, ts = sum(x[1] for x in seg), sum(x[2] for x in seg)
    if ti / tot > 0.012 or ts / tots > 0.012:
        mix = collections.Counter(op(x[3]) for x in seg).most_common(7)
        print(f"SASS lines {seg[0][0]:5d}-{seg[-1][0]:5d}  runs {seg[0][1] / items:6.1f}/item  {len(seg):4d} instr  "
              f"{100 * ti / tot:5.1f} % of instructions  {100 * ts / tots:5.1f} % of stall samples   {dict(mix)}")
agg = collections.Counter()
for x in out:
    agg[op(x[3])] += x[1]
print("opcode mix:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in agg.most_common(18)))
