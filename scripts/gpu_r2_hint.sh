#!/bin/bash
# round 2, call HINT: backward dY / codes TMA loads with an L2 evict-first policy (A/B)
mkdir -p gpurun_out
cp 2d_object_detection_b200/libfrcnn_b200.so /tmp/lib_keep.so
cp scripts/ab/libfrcnn_hint.so 2d_object_detection_b200/libfrcnn_b200.so
timeout 600 python -m pytest tests -m gpu -x -q -k "backward or train_step or prepared or full_size" > gpurun_out/hint_pytest.log 2>&1; echo "hint-build pytest rc=$?"; tail -2 gpurun_out/hint_pytest.log
for i in 1 2 3; do
  for v in base hint; do
    cp scripts/ab/libfrcnn_$v.so 2d_object_detection_b200/libfrcnn_b200.so
    echo "$v $(timeout 300 python scripts/time_roi_kernels.py kitti_b8 4 2>&1 | tail -1)" | tee -a gpurun_out/hint_ab.txt
  done
done
for v in base hint; do
  cp scripts/ab/libfrcnn_$v.so 2d_object_detection_b200/libfrcnn_b200.so
  echo "$v dense $(timeout 300 python scripts/time_roi_kernels.py dense_b16 3 2>&1 | tail -1)" | tee -a gpurun_out/hint_ab.txt
done
cp /tmp/lib_keep.so 2d_object_detection_b200/libfrcnn_b200.so
echo "call HINT done"
