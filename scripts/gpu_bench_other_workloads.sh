#!/bin/bash
# The two other bench workloads (BASELINE.json configs[0] and configs[3]) with the kernels in the build.
mkdir -p gpurun_out
timeout 300 python bench.py --workload kitti_b1 --steps 30 --warmup 5 > gpurun_out/bench_kitti_b1.json 2> gpurun_out/bench_kitti_b1.err; echo "b1 exit $?"
timeout 400 python bench.py --workload dense_b16 --steps 20 --warmup 5 > gpurun_out/bench_dense_b16.json 2> gpurun_out/bench_dense_b16.err; echo "dense exit $?"
python - <<'PY'
import json
for n in ("kitti_b1", "dense_b16"):
    d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d.get("step_latency_ms"))
PY
