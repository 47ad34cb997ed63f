#!/bin/bash
# round 2, call Z2: records pre-kernel built with the proposals (prepare / forward_prepared), grid-independent work counters,
# default reserve of ceil(B/2) SMs, new NMS key-layout tests
mkdir -p gpurun_out
FRCNN_DEFINES=FRCNN_DEBUG_SPIN python 2d_object_detection_b200/build.py --force > gpurun_out/z2_build_dbg.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "roi_pool or roi_stream or engine or full_size or pipeline or train_step or sm_budget or backward or prepared" > gpurun_out/z2_pytest_dbg.log 2>&1; echo "debug-build pytest rc=$?"; tail -2 gpurun_out/z2_pytest_dbg.log
python 2d_object_detection_b200/build.py --force > gpurun_out/z2_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/z2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z2_pytest.log; tail -3 gpurun_out/z2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/z2_bench_n1.json 2> gpurun_out/z2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/z2_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f share %.3f | alone burst %.4f ms frac %.3f | latency %.4f clocks %s pool_sms %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r['share_of_step'], r['alone_burst']['kernel_ms'], r['alone_burst']['frac'], d['step_latency_ms'], d['clocks']['sm_mhz'], d['pool_sms']))
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
tail -3 gpurun_out/z2_bench_n1.err
echo "call Z2 done"
