#!/bin/bash
# round 2, call C: 16-warp persistent CTAs (one per SM) + SMs reserved for the proposal path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -3 gpurun_out/c_pytest.log
for r in 0 4 8 12; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-others --no-e2e --reserve-sms $r > gpurun_out/c_bench_r$r.json 2> gpurun_out/c_bench_r$r.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/c_bench_r$r.json').read().strip().splitlines()[-1])
print('reserve $r', 'value %.0f ms/step %.4f kernel_ms %.4f frac %.3f share %.2f pool_sms %s clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['share_of_step'], d.get('pool_sms'), d['clocks']['sm_mhz']))
PY
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench_n1.json 2> gpurun_out/c_bench_n1.err; echo "bench rc=$?"
tail -c 1800 gpurun_out/c_bench_n1.json
echo "call C done"
