#!/bin/bash
# round 2: the bench line at N = 4
mkdir -p gpurun_out
timeout 540 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err; echo "bench n4 rc=$?"
tail -c 300 gpurun_out/n4_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n4_bench.json').read().strip().splitlines()[-1])
print('N=4 kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f clocks %s e2e %.0f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['e2e']['value']))
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','allreduce','error')})
PY
echo "n4 done"
