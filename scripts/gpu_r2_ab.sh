#!/bin/bash
# round 2, call AB: forward x-interpolation that reuses the pixel pair of the sample before (FRCNN_SHARE_PAIRS_MIN = 15: never, 1, 4, 7)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "roi_pool or roi_stream or engine or full_size or pipeline or train_step or sm_budget or prepared" > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/ab_pytest.log
cp 2d_object_detection_b200/libfrcnn_b200.so /tmp/lib_keep.so
for i in 1 2; do
  for v in share15 share1 share4 share7; do
    cp scripts/ab/libfrcnn_$v.so 2d_object_detection_b200/libfrcnn_b200.so
    echo "$v $(timeout 300 python scripts/time_roi_kernels.py kitti_b8 4 2>&1 | tail -1)" | tee -a gpurun_out/ab_ab.txt
    echo "$v dense $(timeout 300 python scripts/time_roi_kernels.py dense_b16 3 2>&1 | tail -1)" | tee -a gpurun_out/ab_ab.txt
  done
done
cp /tmp/lib_keep.so 2d_object_detection_b200/libfrcnn_b200.so
echo "call AB done"
