#!/bin/bash
# round 2, call H: longest-first front list in the forward; train bench headline at N=1 (step-count fix)
mkdir -p gpurun_out
FRCNN_DEFINES=FRCNN_DEBUG_SPIN python 2d_object_detection_b200/build.py --force > gpurun_out/h_build_dbg.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "roi_pool or roi_stream or engine or full_size or pipeline or train_step or sm_budget" > gpurun_out/h_pytest_dbg.log 2>&1; echo "debug-build pytest rc=$?" | tee -a gpurun_out/h_pytest_dbg.log
tail -4 gpurun_out/h_pytest_dbg.log
python 2d_object_detection_b200/build.py --force > gpurun_out/h_build.log 2>&1
if grep -q "rc=0" gpurun_out/h_pytest_dbg.log; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
  tail -3 gpurun_out/h_pytest.log
  for r in 0 4; do
    timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-others --no-e2e --reserve-sms $r > gpurun_out/h_bench_r$r.json 2> gpurun_out/h_bench_r$r.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/h_bench_r$r.json').read().strip().splitlines()[-1])
print('reserve $r', 'value %.0f ms/step %.4f kernel_ms %.4f frac %.3f share %.2f pool_sms %s clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['share_of_step'], d.get('pool_sms'), d['clocks']['sm_mhz']))
PY
  done
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/h_bench_n1.json 2> gpurun_out/h_bench_n1.err
  python - <<'PY'
import json
d=json.loads(open('gpurun_out/h_bench_n1.json').read().strip().splitlines()[-1])
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f latency %.4f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['step_latency_ms'], d['clocks']['sm_mhz']))
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
  CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-others --min-seconds 0.01 --reserve-sms 0"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:roi_pool_fwd_persist -s 6 -c 1 -f -o gpurun_out/h_prof_fwd $CMD > gpurun_out/h_ncu_full.log 2>&1
fi
echo "call H done"
