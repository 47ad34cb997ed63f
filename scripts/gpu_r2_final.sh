#!/bin/bash
# round 2: the driver's sequence on one GPU (tests, smoke, both bench arms) + the profile captures committed under profiles/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f2_pytest.log; tail -2 gpurun_out/f2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f2_smoke.log 2>&1; tail -1 gpurun_out/f2_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/f2_bench_ref.json 2> gpurun_out/f2_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f2_bench_n1.json 2> gpurun_out/f2_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/f2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f2_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f | alone burst %.4f ms frac %.3f | latency %.4f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r['alone_burst']['kernel_ms'], r['alone_burst']['frac'], d['step_latency_ms'], d['clocks']))
print('e2e', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','pcie_gbs','pinned_d2h_copy_gbs','frac_of_pinned_d2h_copy')}, 'cpu', d['cpu_baseline']['value'])
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-others --min-seconds 0.01"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/f2_launches.csv $CMD > gpurun_out/f2_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:roi_pool_fwd_persist -s 6 -c 1 -f -o gpurun_out/f2_prof_fwd $CMD > gpurun_out/f2_ncu_fwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nms_select -s 4 -c 1 -f -o gpurun_out/f2_prof_nms $CMD > gpurun_out/f2_ncu_nms.log 2>&1
timeout 120 python scripts/run_kernel_once.py bwdcodes > gpurun_out/f2_plain_bwd.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:roi_pool_ -s 2 -c 4 -f -o gpurun_out/f2_prof_train python scripts/run_kernel_once.py bwdcodes > gpurun_out/f2_ncu_train.log 2>&1
CMDD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-others --min-seconds 0.01 --workload dense_b16"
timeout 400 ncu --set full --clock-control none -k regex:roi_pool_fwd_persist -s 6 -c 1 -f -o gpurun_out/f2_prof_fwd_dense $CMDD > gpurun_out/f2_ncu_fwd_dense.log 2>&1
ls -la gpurun_out/f2_prof_*.ncu-rep
echo "final2 done"
