#!/bin/bash
# round 2, call B: the persistent forward kernel -- bounded-wait debug build first (a stuck mbarrier traps instead of
# hanging the GPU), then the product build: all GPU tests, bench, ncu launch list + full capture of the forward kernel.
mkdir -p gpurun_out
ROI_TESTS="roi_pool or roi_stream or engine or full_size or host_pipeline or train_step"
FRCNN_DEFINES=FRCNN_DEBUG_SPIN python 2d_object_detection_b200/build.py --force > gpurun_out/b_build_dbg.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "$ROI_TESTS" > gpurun_out/b_pytest_dbg.log 2>&1; echo "debug-build pytest rc=$?" | tee -a gpurun_out/b_pytest_dbg.log
tail -4 gpurun_out/b_pytest_dbg.log
python 2d_object_detection_b200/build.py --force > gpurun_out/b_build.log 2>&1
if grep -q "rc=0" gpurun_out/b_pytest_dbg.log; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
  tail -4 gpurun_out/b_pytest.log
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench_n1.json 2> gpurun_out/b_bench_n1.err; echo "bench rc=$?"
  tail -c 2500 gpurun_out/b_bench_n1.json; tail -3 gpurun_out/b_bench_n1.err
  CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-others --min-seconds 0.01"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/b_launches.csv $CMD > gpurun_out/b_ncu_launches.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:roi_pool_fwd_persist -s 6 -c 1 -f -o gpurun_out/b_prof_fwd $CMD > gpurun_out/b_ncu_full.log 2>&1
  ls -la gpurun_out/b_prof_fwd.ncu-rep
fi
echo "call B done"
