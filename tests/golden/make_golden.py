#!/usr/bin/env python
"""Generate tests/golden/hotpath_small.npz.

The reference cannot be executed here (it needs TensorFlow 2, which is not installable), so these golden vectors
are produced by the *independent NumPy restatement* (oracle/reference_np.py -- pure NumPy + Python loops, written
separately from the C oracle and the CUDA kernels) on small seeded inputs, and committed.  Both the C oracle
(CPU test) and the CUDA path (GPU test) are then compared against the same file, so a slip in any one of the three
implementations shows up as a disagreement.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_np as rnp  # noqa: E402

F = np.float32


def main():
    rng = np.random.default_rng(424242)
    g = {}
    # --- anchors + RPN-form postprocess (nc = 1)
    grid, image_shape = (5, 8), (80, 128, 3)
    anchors = rnp.generate_anchors(grid, [0.25, 0.5], [0.5, 1.0, 2.0], [128, 128])
    N = anchors.shape[0]
    logits = rng.normal(0, 1.5, (2, N, 2)).astype(F)
    sc = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
    sc = sc.astype(F)
    sc[..., 1] += (np.arange(N, dtype=F) * F(1e-6))[None]  # break ties deterministically
    sc[..., 0] = F(1) - sc[..., 1]
    dl = np.concatenate([rng.normal(0, 0.1, (2, N, 1, 2)), rng.normal(0, 0.2, (2, N, 1, 2))], -1).astype(F)
    clipped = rnp.clip_to_window(anchors, [0, 0, image_shape[1], image_shape[0]])
    post = rnp.postprocess_output(image_shape, clipped, sc, dl, 0.0, 0.7, 25, 25)
    g.update(rpn_grid=np.array(grid), rpn_image_shape=np.array(image_shape), rpn_anchors=anchors, rpn_regions=clipped,
             rpn_scores=sc, rpn_deltas=dl, rpn_out_boxes=post["pred_boxes"], rpn_out_scores=post["pred_scores"],
             rpn_out_valid=post["num_valid_detections"], rpn_out_idx=post["kept_indices"],
             rpn_decoded=post["decoded_boxes"])
    # --- RCNN-form postprocess (nc = 3, per-image regions, cross-class merge)
    B, R, nc = 2, 30, 3
    lo = rng.uniform(0, 0.6, (B, R, 2))
    rois = np.concatenate([lo, lo + rng.uniform(0.05, 0.4, (B, R, 2))], -1).astype(F) * F([128, 80, 128, 80])
    lg = rng.normal(0, 1, (B, R, nc + 1)).astype(F)
    sc2 = (np.exp(lg) / np.exp(lg).sum(-1, keepdims=True)).astype(F)
    dl2 = rng.normal(0, 0.1, (B, R, nc, 4)).astype(F)
    post2 = rnp.postprocess_output(image_shape, rois, sc2, dl2, 0.05, 0.6, 6, 14)
    g.update(rcnn_regions=rois, rcnn_scores=sc2, rcnn_deltas=dl2, rcnn_out_boxes=post2["pred_boxes"],
             rcnn_out_scores=post2["pred_scores"], rcnn_out_classes=post2["pred_classes"],
             rcnn_out_valid=post2["num_valid_detections"], rcnn_out_idx=post2["kept_indices"])
    # --- RoI pooling forward / backward (post-ReLU features: exact-zero ties; padded + partly-outside RoIs)
    feat = np.maximum(rng.standard_normal((2, 6, 9, 8)).astype(F), 0)
    prois = np.clip(np.concatenate([lo[:, :6], lo[:, :6] + rng.uniform(0.1, 0.5, (B, 6, 2))], -1), 0, 1).astype(F)
    prois[0, 4] = [0, 0, 0, 0]
    prois[1, 5] = [0.25, 0.1, 1.0, 1.0]
    prois[1, 4] = [-0.2, 0.3, 0.4, 1.3]
    dy = rng.standard_normal((2, 6, 7, 7, 8)).astype(F)
    g.update(pool_feat=feat, pool_rois=prois, pool_out=rnp.roi_pool_forward(feat, prois, 7, 2), pool_dy=dy,
             pool_dfeat=rnp.roi_pool_backward(feat, prois, dy, 7, 2))
    # --- target matching (RCNN labels with padding; RPN objectness form)
    gl = np.zeros((5, 4), F)
    gb = np.zeros((5, 4), F)
    gb[:3] = [[0.1, 0.1, 0.5, 0.6], [0.4, 0.3, 0.9, 0.9], [0.05, 0.5, 0.3, 0.95]]
    gl[0, 1] = gl[1, 3] = gl[2, 2] = 1
    regions = clipped[::3]
    tl, tb = rnp.generate_targets(gl, gb, regions, image_shape, [0.5, 1.0], [0.0, 0.5])
    obj = np.eye(2, dtype=F)[gl.sum(-1).astype(int)]
    tl2, tb2 = rnp.generate_targets(obj, gb, regions, image_shape, [0.7, 1.0], [0.0, 0.3])
    g.update(tgt_gt_labels=gl, tgt_gt_boxes=gb, tgt_regions=regions, tgt_labels=tl, tgt_boxes=tb, tgt_obj=obj,
             tgt_rpn_labels=tl2, tgt_rpn_boxes=tb2)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_small.npz")
    np.savez_compressed(out, **g)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
