#!/bin/bash
# round 2, call PS2: default bench line with the two pooling streams (>= 4 images per batch), full GPU tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/ps2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ps2_pytest.log; tail -3 gpurun_out/ps2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ps2_smoke.log 2>&1; tail -1 gpurun_out/ps2_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/ps2_bench_n1.json 2> gpurun_out/ps2_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/ps2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ps2_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f (bracket %.4f) frac %.3f share %.3f | alone burst %.4f ms frac %.3f | latency %.4f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['kernel_ms_bracketed'], r['frac'], r['share_of_step'], r['alone_burst']['kernel_ms'], r['alone_burst']['frac'], d['step_latency_ms'], d['clocks']))
print('e2e', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','pcie_gbs','pinned_d2h_copy_gbs','frac_of_pinned_d2h_copy')}, 'cpu', d['cpu_baseline']['value'])
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
echo "call PS2 done"
