#!/bin/bash
# round 2, call Y: second work item of a warp from the counter when a launch has few items per warp (batch 1)
mkdir -p gpurun_out
FRCNN_DEFINES=FRCNN_DEBUG_SPIN python 2d_object_detection_b200/build.py --force > gpurun_out/y_build_dbg.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "roi_pool or roi_stream or engine or full_size or pipeline or train_step or sm_budget or backward" > gpurun_out/y_pytest_dbg.log 2>&1; echo "debug-build pytest rc=$?"; tail -2 gpurun_out/y_pytest_dbg.log
python 2d_object_detection_b200/build.py --force > gpurun_out/y_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest.log; tail -3 gpurun_out/y_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/y_bench_n1.json 2> gpurun_out/y_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f | alone burst %.4f ms frac %.3f | latency %.4f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r['alone_burst']['kernel_ms'], r['alone_burst']['frac'], d['step_latency_ms'], d['clocks']['sm_mhz']))
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-others --min-seconds 0.01 --workload kitti_b1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/y_launches_b1.csv $CMD > gpurun_out/y_ncu_launches_b1.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/y_launches_b1.csv')))
hdr=None; d=collections.defaultdict(list)
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        rec=dict(zip(hdr,r))
        try: d[rec['Kernel Name'][:50]].append(float(rec['Metric Value'].replace(',','')))
        except: pass
for k,v in d.items(): print(f"{k:52s} n={len(v):3d} median {sorted(v)[len(v)//2]/1000:8.1f} us  min {min(v)/1000:.1f} max {max(v)/1000:.1f}")
PY
echo "call Y done"
