#!/bin/bash
# round 2: the driver's GPU checks on the final commit
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/lc_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/lc_pytest.log; tail -3 gpurun_out/lc_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/lc_smoke.log 2>&1; tail -1 gpurun_out/lc_smoke.log
echo "lastcheck done"
