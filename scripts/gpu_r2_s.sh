#!/bin/bash
# round 2, call S: wide RoIs (27..52 columns, reversed boxes) scatter sample by sample with 2-pixel TMA reduce-adds instead of atomics
mkdir -p gpurun_out
FRCNN_DEFINES=FRCNN_DEBUG_SPIN python 2d_object_detection_b200/build.py --force > gpurun_out/s_build_dbg.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "roi_pool or roi_stream or train_step or full_size or backward" > gpurun_out/s_pytest_dbg.log 2>&1; echo "debug-build pytest rc=$?"; tail -2 gpurun_out/s_pytest_dbg.log
python 2d_object_detection_b200/build.py --force > gpurun_out/s_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s_pytest.log
cp 2d_object_detection_b200/libfrcnn_b200.so /tmp/lib_keep.so
for i in 1 2; do
  for v in pred wide; do
    cp scripts/ab/libfrcnn_$v.so 2d_object_detection_b200/libfrcnn_b200.so
    echo "$v $(timeout 300 python scripts/time_roi_kernels.py kitti_b8 4 2>&1 | tail -1)" | tee -a gpurun_out/s_ab.txt
    echo "$v dense $(timeout 300 python scripts/time_roi_kernels.py dense_b16 3 2>&1 | tail -1)" | tee -a gpurun_out/s_ab.txt
  done
done
cp /tmp/lib_keep.so 2d_object_detection_b200/libfrcnn_b200.so
echo "call S done"
