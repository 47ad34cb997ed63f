#!/bin/bash
# round 2, call U: NMS -- keys relative to the largest key (spread first digit), warp-level slot reservation in the collect sweep,
# sanitised-area IoU test with the common outcome first, kept-list loop unrolled by four
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest.log; tail -3 gpurun_out/u_pytest.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nms_select --csv --log-file gpurun_out/u_nms_launches.csv python scripts/run_kernel_once.py post > gpurun_out/u_ncu_nms.log 2>&1
grep nms_select gpurun_out/u_nms_launches.csv | tail -3 | cut -d, -f5,12-
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nms_select --csv --log-file gpurun_out/u_nms_launches_dense.csv python scripts/run_kernel_once.py post dense_b16 > gpurun_out/u_ncu_nms_dense.log 2>&1
grep nms_select gpurun_out/u_nms_launches_dense.csv | tail -2 | cut -d, -f5,12-
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/u_bench_n1.json 2> gpurun_out/u_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/u_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f | alone burst %.4f ms frac %.3f | latency %.4f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r['alone_burst']['kernel_ms'], r['alone_burst']['frac'], d['step_latency_ms'], d['clocks']['sm_mhz']))
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
FRCNN_DEFINES=FRCNN_NMS_TIMING python 2d_object_detection_b200/build.py --force > gpurun_out/u_build.log 2>&1
timeout 120 python scripts/run_kernel_once.py post > gpurun_out/u_nms_timing.log 2>&1; grep "nms timing" gpurun_out/u_nms_timing.log | tail -3
python 2d_object_detection_b200/build.py --force >> gpurun_out/u_build.log 2>&1
echo "call U done"
