#!/bin/bash
# round 2: launch list of the training-path step (train_b16)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-others --min-seconds 0.01 --workload train_b16"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/tl_launches_train.csv $CMD > gpurun_out/tl_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/tl_launches_train.csv')))
hdr=None; d=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        rec=dict(zip(hdr,r))
        try: d.setdefault(rec['Kernel Name'][:70],[]).append(float(rec['Metric Value'].replace(',','')))
        except: pass
tot=0
for k,v in d.items():
    med=sorted(v)[len(v)//2]/1000
    print(f"{k:72s} n={len(v):3d} median {med:8.1f} us")
PY
echo "trainlist done"
