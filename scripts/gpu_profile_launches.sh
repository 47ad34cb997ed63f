#!/bin/bash
# Per-launch device times of one short bench run (ncu, cold-cache serialised: compare SHARES not absolutes).
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu exit: $?"
tail -n 3 gpurun_out/ncu_launches.log
