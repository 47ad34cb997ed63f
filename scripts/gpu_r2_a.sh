#!/bin/bash
# round 2, call A: GPU tests + smoke + the bench line with the new methodology (baseline kernels)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
tail -2 gpurun_out/a_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/a_bench_n1.json; tail -5 gpurun_out/a_bench_n1.err
