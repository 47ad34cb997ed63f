#!/bin/bash
# round 2, call O: float work-item counter (no warp-aggregated atom)
mkdir -p gpurun_out
FRCNN_DEFINES=FRCNN_DEBUG_SPIN python 2d_object_detection_b200/build.py --force > gpurun_out/o_build_dbg.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "roi_pool or roi_stream or engine or full_size or pipeline or train_step or sm_budget" > gpurun_out/o_pytest_dbg.log 2>&1; echo "debug-build pytest rc=$?" | tee -a gpurun_out/o_pytest_dbg.log
python 2d_object_detection_b200/build.py --force > gpurun_out/o_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log; tail -2 gpurun_out/o_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/o_bench_n1.json 2> gpurun_out/o_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/o_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f | alone burst %.4f ms frac %.3f | latency %.4f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r['alone_burst']['kernel_ms'], r['alone_burst']['frac'], d['step_latency_ms'], d['clocks']['sm_mhz']))
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-others --min-seconds 0.01 --reserve-sms 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:roi_pool_fwd_persist -s 6 -c 1 -f -o gpurun_out/o_prof_fwd $CMD > gpurun_out/o_ncu_fwd.log 2>&1
timeout 120 python scripts/run_kernel_once.py bwdcodes > gpurun_out/o_plain_bwd.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:roi_pool_bwd_persist -s 1 -c 1 -f -o gpurun_out/o_prof_bwd python scripts/run_kernel_once.py bwdcodes > gpurun_out/o_ncu_bwd.log 2>&1
echo "call O done"
