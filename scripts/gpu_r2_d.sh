#!/bin/bash
# round 2, call D: persistent backward -- bounded-wait debug build first, then product build: tests, train bench, ncu
mkdir -p gpurun_out
ROI_TESTS="roi_pool or roi_stream or full_size or train_step or smoke"
FRCNN_DEFINES=FRCNN_DEBUG_SPIN python 2d_object_detection_b200/build.py --force > gpurun_out/d_build_dbg.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "$ROI_TESTS" > gpurun_out/d_pytest_dbg.log 2>&1; echo "debug-build pytest rc=$?" | tee -a gpurun_out/d_pytest_dbg.log
tail -4 gpurun_out/d_pytest_dbg.log
python 2d_object_detection_b200/build.py --force > gpurun_out/d_build.log 2>&1
if grep -q "rc=0" gpurun_out/d_pytest_dbg.log; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
  tail -4 gpurun_out/d_pytest.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d_smoke.log 2>&1; tail -1 gpurun_out/d_smoke.log
  timeout 600 python bench.py --steps 20 --warmup 5 --workload train_b16 --no-others > gpurun_out/d_bench_train.json 2> gpurun_out/d_bench_train.err; echo "bench rc=$?"
  tail -c 1500 gpurun_out/d_bench_train.json; tail -3 gpurun_out/d_bench_train.err
  timeout 120 python scripts/run_kernel_once.py bwdcodes > gpurun_out/d_plain_bwd.log 2>&1 &&
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:roi_pool_bwd_persist -s 1 -c 1 -f -o gpurun_out/d_prof_bwd python scripts/run_kernel_once.py bwdcodes > gpurun_out/d_ncu_bwd.log 2>&1
  ls -la gpurun_out/d_prof_bwd.ncu-rep
fi
echo "call D done"
