#!/usr/bin/env python
"""A/B helper: time the RoI-pool forward and backward-from-codes alone (KITTI batch 8), CUDA events around blocks of launches."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
det = importlib.import_module("2d_object_detection_b200"); syn = det.synthetic
w = syn.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "kitti_b8"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rng = np.random.default_rng(5); dev = torch.device("cuda")
anchors = det.clip_to_window(det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape), [0, 0, w.image_shape[1], w.image_shape[0]])
scores, deltas = syn.rpn_head_outputs(rng, w.batch, w.num_anchors)
scores, deltas = torch.as_tensor(scores).to(dev), torch.as_tensor(deltas).to(dev)
feat = torch.as_tensor(syn.feature_maps(rng, w.batch, w.grid, w.channels)).to(dev)
rois = det.postprocess_output(w.image_shape, anchors, scores, deltas, **w.nms_kwargs(), pre_nms_top_n=w.pre_nms_top_n)["pred_boxes"]
dy = torch.randn((w.batch, w.max_total_size, 7, 7, w.channels), device=dev)
_, codes = det.roi_pool_forward_argmax(feat, rois, 7, 2)
def timed(fn, n):
    for _ in range(20): fn()
    best = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        best.append(a.elapsed_time(b) / n)
    return best
f = timed(lambda: det.roi_pool_forward(feat, rois, 7, 2), 400)
g = timed(lambda: det.roi_pool_backward_argmax(tuple(feat.shape), rois, dy, codes, 7, 2), 200)
print("fwd ms/launch", " ".join("%.4f" % x for x in f), "| bwd(codes, incl. zero-init) ms/launch", " ".join("%.4f" % x for x in g))
