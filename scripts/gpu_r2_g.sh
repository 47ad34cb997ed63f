#!/bin/bash
# round 2, call G: product build without the legacy per-CTA kernels; backward flush with CS2R carries
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
tail -4 gpurun_out/g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g_smoke.log 2>&1; tail -1 gpurun_out/g_smoke.log
timeout 120 python scripts/run_kernel_once.py bwdcodes > gpurun_out/g_plain_bwd.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:roi_pool_bwd_persist -s 1 -c 1 -f -o gpurun_out/g_prof_bwd python scripts/run_kernel_once.py bwdcodes > gpurun_out/g_ncu_bwd.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --workload train_b16 --no-others > gpurun_out/g_bench_train.json 2> gpurun_out/g_bench_train.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g_bench_train.json').read().strip().splitlines()[-1])
print('train_b16 value %.0f ms/step %.4f bwd_ms %.4f frac %.3f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']['sm_mhz']))
PY
echo "call G done"
