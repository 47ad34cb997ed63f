#!/usr/bin/env python
"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck): tiny shapes, all code paths."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
det = importlib.import_module("2d_object_detection_b200"); syn = det.synthetic
rng = np.random.default_rng(1)
image_shape, grid = (160, 256, 3), (10, 16)
anchors = det.clip_to_window(det.generate_anchors(grid, [0.25, 0.5, 1.0], [0.5, 1.0, 2.0], [128, 128]), [0, 0, 256, 160])
N = anchors.shape[0]
scores, deltas = syn.rpn_head_outputs(rng, 2, N)
for top_n in (None, 200):
    out = det.postprocess_output(image_shape, anchors, scores, deltas, 0.0, 0.7, 50, 50, pre_nms_top_n=top_n, return_extra=True)
# multi-class form
rois_abs = det.to_absolute(out["pred_boxes"], image_shape)
sc7 = torch.softmax(torch.randn(2, 50, 8, device="cuda"), -1).contiguous()
dl7 = (torch.randn(2, 50, 7, 4, device="cuda") * 0.1).contiguous()
det.postprocess_output(image_shape, rois_abs, sc7, dl7, 0.0, 0.6, 10, 30)
special = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.9, 0.8, 0.1, 0.2], [-0.2, 0.3, 0.4, 1.3], [-0.5, -0.6, -0.1, -0.2], [0.5, 0.5, 0.5, 0.5]],
                       device="cuda")
out["pred_boxes"][1, :5] = special  # whole map, reversed, partly / completely outside, degenerate: every RoI-pool path
for C in (1024, 256, 6):
    feat = torch.as_tensor(syn.feature_maps(rng, 2, grid, C)).cuda().requires_grad_(True)
    pooled = det.ROIPooling(7, 2)(feat, out["pred_boxes"], True, True)
    pooled.backward(torch.randn_like(pooled))
    det.roi_pool_backward(feat.detach(), out["pred_boxes"], torch.randn(2, 50, 7, 7, C, device="cuda"), 7, 2)
    det.roi_pool_forward(feat.detach(), out["pred_boxes"], 4, 1)
gl, gb = syn.ground_truth(rng, 2)
det.generate_targets_batched(det.rpn_objectness_labels(gl), gb, anchors, image_shape, [0.7, 1.0], [0.0, 0.3])
det.generate_targets_batched(gl, gb, rois_abs, image_shape, [0.5, 1.0], [0.0, 0.5])
idx = det.valid_anchor_indices(anchors, image_shape)
det.iou(anchors[:64], anchors[64:128], pairwise=True)
det.encode(anchors, anchors); det.decode(torch.zeros_like(anchors), anchors)
pipe = det.ProposalRoIPipeline(image_shape, 2, grid, 64, N, 0.0, 0.7, 50)
pipe(anchors.cpu().numpy(), scores, deltas, syn.feature_maps(rng, 2, grid, 64)); pipe.close()
torch.cuda.synchronize()
print("sanitize run ok")
