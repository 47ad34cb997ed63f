#!/bin/bash
# ncu --set full on selected kernels of a short bench run.  usage: gpu_profile_full.sh <kernel-regex> <out-name> [skip] [count]
mkdir -p gpurun_out
REGEX=${1:-roi_pool_fwd}
OUT=${2:-prof}
SKIP=${3:-4}
COUNT=${4:-2}
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$REGEX -s $SKIP -c $COUNT -f -o gpurun_out/$OUT $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit: $?"
tail -n 2 gpurun_out/ncu_full.log | cut -c1-300
