#!/bin/bash
# Stage timings, the training-path step and one ncu --set full capture of the RoI backward kernel in the build.
mkdir -p gpurun_out
timeout 200 python scripts/bench_stages.py > gpurun_out/stages.jsonl 2> gpurun_out/stages.err
timeout 200 python scripts/bench_train_path.py > gpurun_out/train_n1.json 2> gpurun_out/train_n1.err; tail -n 1 gpurun_out/train_n1.json | cut -c1-300
timeout 60 python scripts/run_kernel_once.py bwdcodes > gpurun_out/plain_bwd.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_pool_bwd_stream -s 1 -c 1 -f -o gpurun_out/prof_final_bwd python scripts/run_kernel_once.py bwdcodes > gpurun_out/ncu_bwd.log 2>&1
echo "profile done $?"
