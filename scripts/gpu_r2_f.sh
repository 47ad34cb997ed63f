#!/bin/bash
# round 2, call F: backward with the carry scatter (pixels finalised in registers), NMS back to adaptive chunks, new API tests
mkdir -p gpurun_out
FRCNN_DEFINES=FRCNN_DEBUG_SPIN python 2d_object_detection_b200/build.py --force > gpurun_out/f_build_dbg.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "roi_pool or full_size or train_step or nms or postprocess or pipeline or sm_budget" > gpurun_out/f_pytest_dbg.log 2>&1; echo "debug-build pytest rc=$?" | tee -a gpurun_out/f_pytest_dbg.log
tail -4 gpurun_out/f_pytest_dbg.log
python 2d_object_detection_b200/build.py --force > gpurun_out/f_build.log 2>&1
if grep -q "rc=0" gpurun_out/f_pytest_dbg.log; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
  tail -4 gpurun_out/f_pytest.log
  timeout 120 python scripts/run_kernel_once.py bwdcodes > gpurun_out/f_plain_bwd.log 2>&1 &&
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:roi_pool_bwd_persist -s 1 -c 1 -f -o gpurun_out/f_prof_bwd python scripts/run_kernel_once.py bwdcodes > gpurun_out/f_ncu_bwd.log 2>&1
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --reserve-sms 4 --no-e2e > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo "bench rc=$?"
  python - <<'PY'
import json
d=json.loads(open('gpurun_out/f_bench_n1.json').read().strip().splitlines()[-1])
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f latency %.4f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['step_latency_ms'], d['clocks']['sm_mhz']))
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
fi
echo "call F done"
