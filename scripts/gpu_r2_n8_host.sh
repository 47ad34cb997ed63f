#!/bin/bash
# round 2: host-side ceiling of the end-to-end path with 1 / 2 / 4 / 8 ranks active on one 8-GPU box
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scripts/e2e_host_ceiling.py 0.8 > gpurun_out/n8_host_ceiling.jsonl 2> gpurun_out/n8_host_ceiling.err; echo "rc=$?"
cat gpurun_out/n8_host_ceiling.jsonl; tail -c 600 gpurun_out/n8_host_ceiling.err
free -g | head -2 > gpurun_out/n8_free.txt
