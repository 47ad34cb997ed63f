#!/usr/bin/env python
"""Generate tests/golden/reference_exec.npz by EXECUTING THE REFERENCE'S OWN PYTHON SOURCE on small seeded inputs.

STATUS: NOT YET RUN (attempted in round 1 and in both sessions of round 2; the build sandbox's security classifier refuses to execute modules from
the external reference tree without a human's authorisation -- DESIGN.md section 2 quotes the refusal).  This script is
the prepared recipe for a maintainer who may run it.  tests/test_reference_golden.py skips
until reference_exec.npz exists.

The reference (TensorFlow 2) cannot run as is in the build container (no TensorFlow, no network).  Its hot-path modules,
however, only touch ~45 TensorFlow symbols; tests/golden/tf_numpy_shim.py stands in for them with NumPy float32
arithmetic, which lets this script import utils/boxes.py, utils/metrics.py, utils/training.py, utils/post_processing.py
and models/detectors/{rpn,fast_rcnn}_detector.py unmodified from where they lie and call

    RPNDetector._generate_anchors, RPNDetector._remove_invalid_anchors_and_predictions, clip_to_window, encode, decode,
    to_absolute, to_relative, area, intersection, iou, generate_targets (+ _get_labels_masks), postprocess_output,
    ROIPooling.call

Everything those functions do in Python is then the reference's code, not a restatement.  The three TF kernels they call
(CombinedNMS, crop_and_resize, MaxPool2D) are the shim's NumPy restatements -- see the shim's header for what that does
and does not pin.  The reference tree is only read; it does not exist on the GPU box, so the outputs are committed.
Run (build container, once authorised):  python tests/golden/make_reference_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

F = np.float32


def rand_boxes(rng, shape, lo, hi):
    p = np.sort(rng.uniform(lo, hi, shape + (2, 2)).astype(F), axis=-2)
    return (p.reshape(shape + (4,)) + F([0, 0, 1, 1])).astype(F)


def main():
    tf_numpy_shim.install()
    sys.path.insert(0, os.environ.get("FRCNN_REFERENCE_DIR", "/root/reference"))
    from models.detectors.fast_rcnn_detector import ROIPooling
    from models.detectors.rpn_detector import RPNDetector
    from utils import boxes as rb, metrics as rm, post_processing as rp, training as rt

    rng = np.random.default_rng(20261018)
    g = {}
    # ---- anchors: rpn_detector.py:162-199 on two small grids and the KITTI grid
    cases = [((3, 4), [0.5, 1.0], [1.0], [64, 64]), ((5, 8), [0.25, 0.5], [0.5, 1.0, 2.0], [128, 128]),
             ((24, 78), [0.125, 0.25, 0.5, 1.0], [0.5, 1.0, 2.0], [256, 256])]
    for i, (grid, scales, ratios, base) in enumerate(cases):
        g[f"anchors{i}_args"] = np.asarray(list(grid) + base, np.int32)
        g[f"anchors{i}_scales"], g[f"anchors{i}_ratios"] = np.asarray(scales, F), np.asarray(ratios, F)
        g[f"anchors{i}"] = RPNDetector._generate_anchors(None, grid, scales, ratios, base)
    image_shape = (80, 128, 3)
    anchors = g["anchors1"]
    N = anchors.shape[0]
    # ---- training-mode filter: rpn_detector.py:201-230
    me = types.SimpleNamespace(_image_shape=image_shape, _anchors=anchors)
    sc = rng.uniform(0, 1, (2, N, 2)).astype(F)
    dl = rng.normal(0, 0.2, (2, N, 1, 4)).astype(F)
    va, vs, vd = RPNDetector._remove_invalid_anchors_and_predictions(me, sc, dl)
    g.update(filt_scores_in=sc, filt_deltas_in=dl, filt_anchors=va, filt_scores=vs, filt_deltas=vd)
    # ---- utils/boxes.py
    b = rand_boxes(rng, (3, 40), 0, 120)
    r = rand_boxes(rng, (40,), 0, 120)
    t = rng.normal(0, 0.3, (3, 40, 4)).astype(F)
    g.update(box_b=b, box_r=r, box_t=t, box_clip=rb.clip_to_window(b, [10, 20, 100, 60]), box_encode=rb.encode(b, r),
             box_decode=rb.decode(t, r), box_rel=rb.to_relative(b, image_shape), box_abs=rb.to_absolute(b / F(128), image_shape))
    # ---- utils/metrics.py:136-208
    b1, b2 = rand_boxes(rng, (30,), 0, 100), rand_boxes(rng, (17,), 0, 100)
    b2[3] = b1[5]                      # IoU exactly 1
    b2[4] = [300, 300, 310, 310]       # disjoint: intersection 0 -> IoU 0 through the tf.where
    g.update(iou_b1=b1, iou_b2=b2, iou_area=rm.area(b1), iou_inter_pair=rm.intersection(b1, b2, pairwise=True),
             iou_pair=rm.iou(b1, b2, pairwise=True), iou_elem=rm.iou(b1[:17], b2), iou_inter_elem=rm.intersection(b1[:17], b2))
    # ---- utils/training.py:7-77, 123-143: RCNN form (7 classes), RPN form (objectness), forced-foreground case
    G, nc = 12, 7
    gt_boxes = np.zeros((G, 4), F)
    gt_labels = np.zeros((G, nc + 1), F)
    gt_boxes[:5, :2] = rng.uniform(0, 0.6, (5, 2)).astype(F)
    gt_boxes[:5, 2:] = np.minimum(gt_boxes[:5, :2] + rng.uniform(0.1, 0.4, (5, 2)).astype(F), F(1))
    gt_labels[np.arange(5), rng.integers(1, nc + 1, 5)] = 1
    regions = np.concatenate([anchors, rb.to_absolute(gt_boxes[:5], image_shape) + rng.normal(0, 2, (5, 4)).astype(F)]).astype(F)
    tl, tb = rt.generate_targets(gt_labels, gt_boxes, regions, image_shape, [0.5, 1.0], [0.0, 0.5])
    g.update(tgt_gt_labels=gt_labels, tgt_gt_boxes=gt_boxes, tgt_regions=regions, tgt_labels=tl, tgt_boxes=tb)
    obj = np.eye(2, dtype=F)[gt_labels.sum(-1).astype(np.int32)]  # rpn_detector.py:141 (padding rows become [1, 0])
    tl, tb = rt.generate_targets(obj, gt_boxes, regions, image_shape, [0.7, 1.0], [0.0, 0.3])
    g.update(tgt_obj=obj, tgt_rpn_labels=tl, tgt_rpn_boxes=tb)
    tl, tb = rt.generate_targets(gt_labels, gt_boxes, anchors, image_shape, [0.95, 1.0], [0.0, 0.1])  # no IoU in the interval
    g.update(tgt_forced_labels=tl, tgt_forced_boxes=tb)
    # ---- utils/post_processing.py:6-63 around the (restated) CombinedNMS kernel: RPN form and RCNN form, batched regions
    regs = rb.clip_to_window(anchors, [0, 0, image_shape[1], image_shape[0]])
    lg = rng.normal(0, 1.5, (2, N, 2)).astype(F)
    ps = (np.exp(lg) / np.exp(lg).sum(-1, keepdims=True)).astype(F)
    out = rp.postprocess_output(image_shape, regs, ps, dl, 0.0, 0.7, 25, 25)
    g.update(post_regions=regs, post_scores=ps, post_deltas=dl, **{"post_" + k: v for k, v in out.items()})
    Rr, B = 30, 2
    rois = rand_boxes(rng, (B, Rr), 0, 100)
    lg = rng.normal(0, 1.5, (B, Rr, nc + 1)).astype(F)
    ps7 = (np.exp(lg) / np.exp(lg).sum(-1, keepdims=True)).astype(F)
    dl7 = rng.normal(0, 0.1, (B, Rr, nc, 4)).astype(F)
    out = rp.postprocess_output(image_shape, rois, ps7, dl7, 0.05, 0.6, 6, 14)
    g.update(post7_regions=rois, post7_scores=ps7, post7_deltas=dl7, **{"post7_" + k: v for k, v in out.items()})
    # ---- ROIPooling.call, fast_rcnn_detector.py:142-177 around the (restated) crop_and_resize / MaxPool2D kernels
    feat = rng.normal(0, 1, (2, 6, 9, 8)).astype(F)
    prois = np.clip(rand_boxes(rng, (2, 5), 0, 1) * F(0.9), 0, 1).astype(F)
    prois[1, 4] = 0                      # zero-padded RoI
    prois[0, 0] = [-0.2, 0.1, 1.3, 0.9]  # partly outside: extrapolation value 0
    layer = ROIPooling(pooled_size=7, kernel_size=2, name="regions_of_interest_pooling")
    g.update(pool_feat=feat, pool_rois=prois, pool_flat_batched=layer.call(feat, prois, True, True),
             pool_unbatched=layer.call(feat, prois, False, False))
    g = {k: np.asarray(v) for k, v in g.items()}
    for k, v in g.items():
        assert v.dtype != np.float64, (k, v.dtype)
    np.savez_compressed(os.path.join(HERE, "reference_exec.npz"), **g)
    print("wrote reference_exec.npz:", {k: (v.shape, str(v.dtype)) for k, v in g.items()})


if __name__ == "__main__":
    main()
