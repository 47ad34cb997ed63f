#!/bin/bash
# round 2, call M: NMS with the division-free IoU decision, 32-bit resolve, shuffle sort: phase timing, tests, launch list, bench
mkdir -p gpurun_out
FRCNN_DEFINES=FRCNN_NMS_TIMING python 2d_object_detection_b200/build.py --force > gpurun_out/m_build.log 2>&1
timeout 120 python scripts/run_kernel_once.py post > gpurun_out/m_nms_timing.log 2>&1; grep "nms timing" gpurun_out/m_nms_timing.log | tail -2
python 2d_object_detection_b200/build.py --force >> gpurun_out/m_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -3 gpurun_out/m_pytest.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-others --min-seconds 0.01"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/m_launches.csv $CMD > gpurun_out/m_ncu_launches.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/m_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); agg={}
for r in rows[1:]:
    agg.setdefault(r[ki].split('(')[0][-40:],[]).append(float(r[vi].replace(',','')))
for k,v in agg.items(): print('  %-42s n=%3d  median %.1f us'%(k,len(v),sorted(v)[len(v)//2]/1e3))
PY
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/m_bench_n1.json 2> gpurun_out/m_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/m_bench_n1.json').read().strip().splitlines()[-1])
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f latency %.4f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['step_latency_ms'], d['clocks']['sm_mhz']))
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
echo "call M done"
