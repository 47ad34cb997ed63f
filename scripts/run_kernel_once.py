#!/usr/bin/env python
"""Run one stage a few times on the KITTI batch-8 workload (for ncu captures): bwd | bwdcodes | targets | post"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
det = importlib.import_module("2d_object_detection_b200"); syn = det.synthetic
what = sys.argv[1] if len(sys.argv) > 1 else "bwd"
w = syn.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "kitti_b8"]
rng = np.random.default_rng(5); dev = torch.device("cuda")
anchors = det.clip_to_window(det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape), [0, 0, w.image_shape[1], w.image_shape[0]])
scores, deltas = syn.rpn_head_outputs(rng, w.batch, w.num_anchors)
scores, deltas = torch.as_tensor(scores).to(dev), torch.as_tensor(deltas).to(dev)
feat = torch.as_tensor(syn.feature_maps(rng, w.batch, w.grid, w.channels)).to(dev)
out = det.postprocess_output(w.image_shape, anchors, scores, deltas, **w.nms_kwargs(), pre_nms_top_n=w.pre_nms_top_n)
rois = out["pred_boxes"]
for _ in range(3):
    if what == "bwd":
        dy = torch.randn((w.batch, w.max_total_size, 7, 7, w.channels), device=dev)
        det.roi_pool_backward(feat, rois, dy, 7, 2)
    elif what == "bwdcodes":
        dy = torch.randn((w.batch, w.max_total_size, 7, 7, w.channels), device=dev)
        _, codes = det.roi_pool_forward_argmax(feat, rois, 7, 2)
        det.roi_pool_backward_argmax(tuple(feat.shape), rois, dy, codes, 7, 2)
    elif what == "post":
        det.postprocess_output(w.image_shape, anchors, scores, deltas, **w.nms_kwargs(), pre_nms_top_n=w.pre_nms_top_n)
    elif what == "targets":
        a_all = det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
        valid = a_all[det.valid_anchor_indices(a_all, w.image_shape)].contiguous()
        gl, gb = syn.ground_truth(rng, w.batch)
        det.generate_targets_batched(det.rpn_objectness_labels(torch.as_tensor(gl).to(dev)), torch.as_tensor(gb).to(dev), valid, w.image_shape, [0.7, 1.0], [0.0, 0.3])
torch.cuda.synchronize()
print("done", what)
