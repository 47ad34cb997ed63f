#!/bin/bash
# round 2: the bench line at N = 2 (one rank per GPU over NCCL), default workload + train_b16 as headline
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/n2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n2_bench.json').read().strip().splitlines()[-1])
print('N=2 kitti_b8 value %.0f ms/step %.4f kernel_ms %.4f frac %.3f clocks %s e2e %.0f pcie %.1f pinned %.1f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['e2e']['value'], d['e2e']['pcie_gbs'], d['e2e']['pinned_d2h_copy_gbs']))
for o in d['other_workloads']: print(o['workload'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k in ('value','ms_per_step','kernel_ms','roofline_frac','allreduce','error')})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --workload train_b16 --no-others > gpurun_out/n2_bench_train.json 2> gpurun_out/n2_bench_train.err; echo "train n2 rc=$?"
tail -c 1200 gpurun_out/n2_bench_train.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > gpurun_out/n2_ref.json 2> gpurun_out/n2_ref.err; echo "ref n2 rc=$?"; tail -c 300 gpurun_out/n2_ref.json
echo "n2 done"
