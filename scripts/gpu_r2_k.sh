#!/bin/bash
# round 2, call K: NMS phase timing (debug build with clock64 probes), one proposal call on KITTI batch 8
mkdir -p gpurun_out
FRCNN_DEFINES=FRCNN_NMS_TIMING python 2d_object_detection_b200/build.py --force > gpurun_out/k_build.log 2>&1
timeout 120 python scripts/run_kernel_once.py post > gpurun_out/k_nms_timing.log 2>&1; tail -6 gpurun_out/k_nms_timing.log
python 2d_object_detection_b200/build.py --force >> gpurun_out/k_build.log 2>&1
echo "call K done"
