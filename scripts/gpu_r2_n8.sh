#!/bin/bash
# round 2: the bench line at N = 8 (one rank per GPU over NCCL): inference headline + other workloads (train_b16 with the all-reduce)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n8_gpus.txt; nvidia-smi topo -m > gpurun_out/n8_topo.txt 2>&1; lscpu | head -25 > gpurun_out/n8_lscpu.txt
timeout 540 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err; echo "bench n8 rc=$?"
tail -c 400 gpurun_out/n8_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n8_bench.json').read().strip().splitlines()[-1])
print('N=8 kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']['sm_mhz']))
print('e2e', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','ms_per_step','pcie_gbs','d2h_gbs','pinned_d2h_copy_gbs','frac_of_pinned_d2h_copy','numa')})
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','allreduce','error')})
PY
echo "n8 done"
