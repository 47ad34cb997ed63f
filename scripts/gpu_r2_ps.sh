#!/bin/bash
# round 2, call PS: two pooling streams (consecutive persistent pooling kernels overlap at their ends) vs one
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "engine or pipelined or pipeline" > gpurun_out/ps_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/ps_pytest.log
for i in 1 2; do
for P in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-others --pool-streams $P > gpurun_out/ps_bench_p$P.json 2> gpurun_out/ps_bench_p$P.err
python - <<PY
import json
d=json.loads(open('gpurun_out/ps_bench_p$P.json').read().strip().splitlines()[-1])
r=d['roofline']
print('pool streams $P: value %.0f ms/step %.4f kernel_ms(bracket) %.4f share %.2f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['share_of_step'], d['clocks']['sm_mhz']))
PY
done
done
for W in kitti_b1 dense_b16; do
for P in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-others --workload $W --pool-streams $P > gpurun_out/ps_bench_${W}_p$P.json 2> gpurun_out/ps_bench_${W}_p$P.err
python - <<PY
import json
d=json.loads(open('gpurun_out/ps_bench_${W}_p$P.json').read().strip().splitlines()[-1])
r=d['roofline']
print('$W pool streams $P: value %.0f ms/step %.4f kernel_ms(bracket) %.4f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], d['clocks']['sm_mhz']))
PY
done
done
echo "call PS done"
