#!/bin/bash
# round 2, call W: NMS sweeps decided on the 32-bit key half; reserve-SM sweep with the faster proposal path
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w_pytest.log; tail -3 gpurun_out/w_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/w_bench_n1.json 2> gpurun_out/w_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f | alone burst %.4f ms frac %.3f | latency %.4f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r['alone_burst']['kernel_ms'], r['alone_burst']['frac'], d['step_latency_ms'], d['clocks']['sm_mhz']))
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
for R in 2 4 6; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-others --reserve-sms $R > gpurun_out/w_bench_r$R.json 2> gpurun_out/w_bench_r$R.err
python - <<PY
import json
d=json.loads(open('gpurun_out/w_bench_r$R.json').read().strip().splitlines()[-1])
r=d['roofline']
print('reserve $R: value %.0f ms/step %.4f kernel_ms %.4f frac %.3f share %.2f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r['share_of_step'], d['clocks']['sm_mhz']))
PY
done
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-others --min-seconds 0.01 --workload kitti_b1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 -k regex:nms_select --csv --log-file gpurun_out/w_launches_b1.csv $CMD > gpurun_out/w_ncu_launches_b1.log 2>&1
grep nms_select gpurun_out/w_launches_b1.csv | tail -3 | cut -d, -f5,12-
FRCNN_DEFINES=FRCNN_NMS_TIMING python 2d_object_detection_b200/build.py --force > gpurun_out/w_build.log 2>&1
timeout 120 python scripts/run_kernel_once.py post > gpurun_out/w_nms_timing.log 2>&1; grep "nms timing" gpurun_out/w_nms_timing.log | tail -3
python 2d_object_detection_b200/build.py --force >> gpurun_out/w_build.log 2>&1
echo "call W done"
