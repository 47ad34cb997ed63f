#!/bin/bash
# round 2, call XR: backward left weights per item; train engine on prepared records; A/B of the backward
mkdir -p gpurun_out
FRCNN_DEFINES=FRCNN_DEBUG_SPIN python 2d_object_detection_b200/build.py --force > gpurun_out/xr_build_dbg.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "roi_pool or roi_stream or engine or full_size or pipeline or train_step or sm_budget or backward or prepared" > gpurun_out/xr_pytest_dbg.log 2>&1; echo "debug-build pytest rc=$?"; tail -2 gpurun_out/xr_pytest_dbg.log
python 2d_object_detection_b200/build.py --force > gpurun_out/xr_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/xr_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/xr_pytest.log; tail -3 gpurun_out/xr_pytest.log
cp 2d_object_detection_b200/libfrcnn_b200.so /tmp/lib_keep.so
for i in 1 2; do
  for v in base xr; do
    cp scripts/ab/libfrcnn_$v.so 2d_object_detection_b200/libfrcnn_b200.so
    echo "$v $(timeout 300 python scripts/time_roi_kernels.py kitti_b8 4 2>&1 | tail -1)" | tee -a gpurun_out/xr_ab.txt
  done
done
cp /tmp/lib_keep.so 2d_object_detection_b200/libfrcnn_b200.so
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/xr_bench_n1.json 2> gpurun_out/xr_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/xr_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f | alone burst %.4f ms frac %.3f | latency %.4f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r['alone_burst']['kernel_ms'], r['alone_burst']['frac'], d['step_latency_ms'], d['clocks']['sm_mhz']))
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
echo "call XR done"
