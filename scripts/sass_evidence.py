#!/usr/bin/env python
"""Blackwell / TMA evidence from the built library (no GPU needed): per kernel, how many SASS instructions of the kinds that
prove the design -- TMA tensor loads (UTMALDG), TMA reduce-add stores (UTMAREDG), bulk copies (UBLKCP), mbarrier traffic
(SYNCS.*), packed fp32 (FFMA2 / FMUL2 / FADD2), warp-uniform reductions (REDUX) -- plus a few literal lines of each.
    python scripts/sass_evidence.py > profiles/r2_sass_tma.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "2d_object_detection_b200", "libfrcnn_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kinds = ["UTMALDG", "UTMAREDG", "UBLKCP", "UTMACMDFLUSH", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "ELECT", "REDUX",
         "FFMA2", "FMUL2", "FADD2", "LDS.128", "LDS.64", "STS.64", "STG.E.EF.64", "ATOMG", "REDG", "LDL", "STL"]
per, samples, cur = collections.OrderedDict(), collections.defaultdict(list), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::|frcnn::|void ", "", cur).split("(")[0]
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
    if not m or cur is None:
        continue
    ins = m.group(1)
    per[cur]["_total"] += 1
    for k in kinds:
        if re.search(r"(^|\s)" + re.escape(k) + r"(\.|\s|$)", ins):
            per[cur][k] += 1
            if len(samples[(cur, k)]) < 2:
                samples[(cur, k)].append(ins.strip())
print("arch: sm_100a   library:", os.path.relpath(so, ROOT), "  (cuobjdump -sass; counts are static SASS instructions)\n")
for kname, c in per.items():
    if not any(c[k] for k in kinds[:6]):
        continue
    print(f"== {kname}  ({c['_total']} instructions)")
    print("   " + "  ".join(f"{k}={c[k]}" for k in kinds if c[k]))
    for k in ("UTMALDG", "UTMAREDG", "UBLKCP", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "FFMA2", "REDUX"):
        for sline in samples.get((kname, k), [])[:1]:
            print(f"      {sline}")
    print()
