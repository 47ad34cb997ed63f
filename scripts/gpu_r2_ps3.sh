#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/ps3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ps3_pytest.log; tail -3 gpurun_out/ps3_pytest.log
echo "call PS3 done"
