#!/bin/bash
# Everything the round's numbers come from, in one GPU call: parity tests, smoke, bench, per-stage timings, the
# training-path step, the ncu launch list of the bench command and one --set full capture of the two RoI kernels.
mkdir -p gpurun_out
bash scripts/gpu_check.sh
timeout 200 python scripts/bench_stages.py > gpurun_out/stages.jsonl 2> gpurun_out/stages.err
timeout 200 python scripts/bench_train_path.py > gpurun_out/train_n1.json 2> gpurun_out/train_n1.err; tail -n 2 gpurun_out/train_n1.json
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 120 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_pool_fwd_stream -s 4 -c 1 -f -o gpurun_out/prof_final_fwd $CMD > gpurun_out/ncu_full.log 2>&1
timeout 60 python scripts/run_kernel_once.py bwdcodes > gpurun_out/plain_bwd.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_pool_bwd_stream -s 1 -c 1 -f -o gpurun_out/prof_final_bwd python scripts/run_kernel_once.py bwdcodes > gpurun_out/ncu_bwd.log 2>&1
echo "measure done"
