"""The anchor part of the reference's models/detectors/rpn_detector.py (the conv heads are out of scope).

Free functions ``generate_anchors`` / ``valid_anchor_indices`` / ``remove_invalid_anchors_and_predictions``
replace ``RPNDetector._generate_anchors`` (:162-199) and ``._remove_invalid_anchors_and_predictions``
(:201-230); ``RPNAnchors`` bundles them with the eval-mode clip (:93-94) the way ``RPNDetector.call`` uses
them, returning the same ``{"regions", "pred_scores", "pred_boxes"}`` dict (:96).
"""
from __future__ import annotations

import ctypes as C

import torch

from ..._tensor import as_device, check, lib, ptr, require_cuda, stream
from ...utils.boxes import clip_to_window
from ...utils.training import get_training_samples as _get_training_samples


def generate_anchors(grid_shape, scales, aspect_ratios, base_anchor_shape, stride_shape=(16, 16)):
    """rpn_detector.py:162-199 -> [grid_h*grid_w*len(scales)*len(aspect_ratios), 4] abs.
    Anchor index = (i*grid_w + j)*A + (r*len(scales) + s); centres at multiples of the stride (no offset)."""
    dev = require_cuda()
    sc = (C.c_float * len(scales))(*[float(v) for v in scales])
    ar = (C.c_float * len(aspect_ratios))(*[float(v) for v in aspect_ratios])
    gh, gw = int(grid_shape[0]), int(grid_shape[1])
    out = torch.empty((gh * gw * len(scales) * len(aspect_ratios), 4), dtype=torch.float32, device=dev)
    check(lib().frcnn_generate_anchors(gh, gw, sc, len(scales), ar, len(aspect_ratios), float(base_anchor_shape[0]),
                                       float(base_anchor_shape[1]), float(stride_shape[0]), float(stride_shape[1]),
                                       ptr(out), stream()), "generate_anchors")
    return out


def valid_anchor_indices(anchors, image_shape):
    """rpn_detector.py:215-224: ascending int64 indices of anchors fully inside the image (x0>=0, y0>=0, x1<=W, y1<=H).
    Static per configuration; this one-off call synchronises to size its result."""
    a = as_device(anchors).reshape(-1, 4)
    idx = torch.empty((a.shape[0],), dtype=torch.int64, device=a.device)
    cnt = torch.zeros((1,), dtype=torch.int32, device=a.device)
    check(lib().frcnn_valid_anchor_indices(ptr(a), a.shape[0], float(image_shape[0]), float(image_shape[1]), ptr(idx),
                                           ptr(cnt), stream()), "valid_anchor_indices")
    return idx[: int(cnt.item())]


def _gather_rows(x, idx):
    """tf.gather(x, idx, axis=1) for x [B, N, ...] (rpn_detector.py:227-228)."""
    x = as_device(x)
    B, N = x.shape[0], x.shape[1]
    row = x.numel() // max(B * N, 1)
    out = torch.empty((B, idx.numel()) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    check(lib().frcnn_gather_rows(ptr(x), B, N, row, ptr(idx), idx.numel(), ptr(out), stream()), "gather_rows")
    return out


def remove_invalid_anchors_and_predictions(anchors, image_shape, pred_scores, pred_boxes_encoded, valid_indices=None):
    """rpn_detector.py:201-230: drop anchors crossing the image border and the matching predictions (order kept)."""
    a = as_device(anchors)
    idx = valid_indices if valid_indices is not None else valid_anchor_indices(a, image_shape)
    return (_gather_rows(a[None], idx)[0], _gather_rows(pred_scores, idx), _gather_rows(pred_boxes_encoded, idx))


class RPNAnchors:
    """Anchor state of RPNDetector (rpn_detector.py:20-22): built once, cached on the device."""

    def __init__(self, image_shape, feature_maps_shape, config):
        self._image_shape = image_shape
        _, grid_height, grid_width, _ = feature_maps_shape
        self._anchors = generate_anchors(grid_shape=(grid_height, grid_width), **config["anchors"])
        image_height, image_width = image_shape[0], image_shape[1]
        self._clipped = clip_to_window(self._anchors, [0, 0, image_width, image_height])  # :93-94
        self._valid_idx = valid_anchor_indices(self._anchors, image_shape)                 # :216-224
        self._valid_anchors = _gather_rows(self._anchors[None], self._valid_idx)[0]               # :226

    @property
    def anchors(self):
        return self._anchors

    def get_training_samples(self, gt_labels, gt_boxes, regions, pred_scores, pred_boxes, foreground_iou_interval,
                             background_iou_interval, num_samples, foreground_proportion, **sampler):
        """RPNDetector.get_training_samples, rpn_detector.py:98-160 -- same signature, called as
        ``get_training_samples(gt_labels, gt_boxes, **rpn_output, **config["rpn"]["sampling"])`` (models/faster_rcnn.py:77-79).
        gt_labels [B,G,nc+1], gt_boxes [B,G,4] rel, regions [N,4] abs, pred_scores [B,N,2], pred_boxes [B,N,1,4] ->
        {"target_labels", "pred_scores": [B,S,2]; "target_boxes", "pred_boxes": [B,S,1,4]}.  The objectness one-hot (:141)
        is folded into the matching kernels.  ``sampler``: generator= / seed= / check_background= (the reference draws from
        TF's unseeded generator)."""
        return _get_training_samples(gt_labels, gt_boxes, regions, pred_scores, pred_boxes, self._image_shape,
                                     foreground_iou_interval, background_iou_interval, num_samples, foreground_proportion,
                                     objectness=True, **sampler)

    def __call__(self, pred_scores, pred_boxes_encoded, training=False):
        """The tail of RPNDetector.call (rpn_detector.py:88-96) given the conv heads' outputs."""
        if training:
            return {"regions": self._valid_anchors, "pred_scores": _gather_rows(pred_scores, self._valid_idx),
                    "pred_boxes": _gather_rows(pred_boxes_encoded, self._valid_idx)}
        return {"regions": self._clipped, "pred_scores": as_device(pred_scores),
                "pred_boxes": as_device(pred_boxes_encoded)}


RPNDetector = RPNAnchors  # the reference's class name; the conv heads stay with the caller (out of scope)
