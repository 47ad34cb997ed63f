#!/bin/bash
# round 2, call S32: NMS sorts short lists as 32-bit words (tie fallback to the 64-bit network)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s32_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s32_pytest.log; tail -3 gpurun_out/s32_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/s32_bench_n1.json 2> gpurun_out/s32_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s32_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f | alone burst %.4f ms frac %.3f | latency %.4f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r['alone_burst']['kernel_ms'], r['alone_burst']['frac'], d['step_latency_ms'], d['clocks']['sm_mhz']))
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-others --min-seconds 0.01 --workload kitti_b1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 -k regex:nms_select --csv --log-file gpurun_out/s32_launches_b1.csv $CMD > gpurun_out/s32_ncu_launches_b1.log 2>&1
grep nms_select gpurun_out/s32_launches_b1.csv | tail -3 | cut -d, -f5,12-
FRCNN_DEFINES=FRCNN_NMS_TIMING python 2d_object_detection_b200/build.py --force > gpurun_out/s32_build.log 2>&1
timeout 120 python scripts/run_kernel_once.py post > gpurun_out/s32_nms_timing.log 2>&1; grep "nms timing block 0" gpurun_out/s32_nms_timing.log | tail -2
python 2d_object_detection_b200/build.py --force >> gpurun_out/s32_build.log 2>&1
echo "call S32 done"
