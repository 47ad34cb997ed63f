#!/bin/bash
# round 2: the driver's two bench arms on one GPU with the final build
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/lb_bench_ref.json 2> gpurun_out/lb_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/lb_bench_n1.json 2> gpurun_out/lb_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/lb_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/lb_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f (bracket %.4f eager span %.4f graph span %s) frac %.3f share %.3f | burst %.4f frac %.3f | clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['kernel_ms_bracketed'], r['kernel_ms_eager_span'], r['kernel_ms_graph_span'], r['frac'], r['share_of_step'], r['alone_burst']['kernel_ms'], r['alone_burst']['frac'], d['clocks']))
print('e2e', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','pcie_gbs','pinned_d2h_copy_gbs','frac_of_pinned_d2h_copy')}, 'cpu', d['cpu_baseline']['value'])
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','step_latency_ms','clocks_sm_mhz','error')})
PY
echo "lastbench done"
