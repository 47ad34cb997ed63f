"""``ROIPooling`` of the reference's models/detectors/fast_rcnn_detector.py:133-177, differentiable w.r.t. the
feature maps (the gradient TF's tape derives at models/faster_rcnn.py:103).  The Dense heads are out of scope.
"""
from __future__ import annotations

import torch

from ..._tensor import as_device, check, lib, ptr, stream, workspace
from ...utils.boxes import to_absolute
from ...utils.training import get_training_samples as _get_training_samples


def roi_pool_forward(feature_maps, rois, pooled_size, kernel_size):
    """crop_and_resize(S=P*k) + MaxPool2D(k) fused: feature_maps [B,Hf,Wf,C], rois [B,R,4] rel x1,y1,x2,y2
    -> [B,R,P,P,C]."""
    f, r = as_device(feature_maps), as_device(rois)
    if f.dim() != 4 or r.dim() != 3 or r.shape[-1] != 4 or r.shape[0] != f.shape[0]:
        raise ValueError(f"feature_maps [B,H,W,C] / rois [B,R,4] expected, got {tuple(f.shape)} / {tuple(r.shape)}")
    B, Hf, Wf, C = f.shape
    R = r.shape[1]
    P, k = int(pooled_size), int(kernel_size)
    out = torch.empty((B, R, P, P, C), dtype=torch.float32, device=f.device)
    ws_bytes = lib().frcnn_roi_pool_workspace_bytes(B, R, C, P, k)
    ws = workspace(ws_bytes)
    check(lib().frcnn_roi_pool_forward(ptr(f), B, Hf, Wf, C, ptr(r), R, P, k, ptr(out), ptr(ws), ws_bytes, stream()),
          "roi_pool_forward")
    return out


def roi_pool_backward(feature_maps, rois, dy, pooled_size, kernel_size):
    """Gradient of roi_pool_forward w.r.t. feature_maps: dy [B,R,P,P,C] (any shape with that many elements)."""
    f, r = as_device(feature_maps), as_device(rois)
    B, Hf, Wf, C = f.shape
    R = r.shape[1]
    P, k = int(pooled_size), int(kernel_size)
    g = as_device(dy)
    if g.numel() != B * R * P * P * C:
        raise ValueError(f"dy has {g.numel()} elements, expected {B * R * P * P * C}")
    df = torch.empty_like(f)
    check(lib().frcnn_roi_pool_backward(ptr(f), B, Hf, Wf, C, ptr(r), R, P, k, ptr(g), ptr(df), stream()),
          "roi_pool_backward")
    return df


def roi_pool_forward_argmax(feature_maps, rois, pooled_size, kernel_size):
    """Training-mode forward: (pooled [B,R,P,P,C], argmax codes [B,R,P,P,C/4] uint8 -- four 2-bit window positions per
    float4 of channels).  The codes let ``roi_pool_backward_argmax`` skip the second interpolation pass."""
    f, r = as_device(feature_maps), as_device(rois)
    B, Hf, Wf, C = f.shape
    R = r.shape[1]
    P, k = int(pooled_size), int(kernel_size)
    out = torch.empty((B, R, P, P, C), dtype=torch.float32, device=f.device)
    codes = torch.empty((B, R, P, P, C // 4), dtype=torch.uint8, device=f.device)
    ws_bytes = lib().frcnn_roi_pool_workspace_bytes(B, R, C, P, k)
    ws = workspace(ws_bytes)
    check(lib().frcnn_roi_pool_forward_argmax(ptr(f), B, Hf, Wf, C, ptr(r), R, P, k, ptr(out), ptr(codes), ptr(ws), ws_bytes,
                                              stream()), "roi_pool_forward_argmax")
    return out, codes


def roi_pool_backward_argmax(feature_shape, rois, dy, argmax_codes, pooled_size, kernel_size):
    """Gradient w.r.t. the feature maps from saved arg-max codes (no feature maps needed)."""
    r, g = as_device(rois), as_device(dy)
    B, Hf, Wf, C = (int(v) for v in feature_shape)
    R = r.shape[1]
    P, k = int(pooled_size), int(kernel_size)
    if g.numel() != B * R * P * P * C:
        raise ValueError(f"dy has {g.numel()} elements, expected {B * R * P * P * C}")
    df = torch.empty((B, Hf, Wf, C), dtype=torch.float32, device=g.device)
    ws_bytes = lib().frcnn_roi_pool_workspace_bytes(B, R, C, P, k)
    ws = workspace(ws_bytes)
    check(lib().frcnn_roi_pool_backward_argmax(B, Hf, Wf, C, ptr(r), R, P, k, ptr(g), ptr(argmax_codes), ptr(df), ptr(ws), ws_bytes,
                                               stream()), "roi_pool_backward_argmax")
    return df


class _ROIPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feature_maps, rois, pooled_size, kernel_size):
        ctx.pk = (pooled_size, kernel_size)
        ctx.fshape = tuple(feature_maps.shape)
        ctx.with_codes = bool(feature_maps.requires_grad and
                              lib().frcnn_roi_pool_argmax_supported(feature_maps.shape[-1], pooled_size, kernel_size))
        if ctx.with_codes:  # training: remember which crop sample won each window (1 byte per float4)
            out, codes = roi_pool_forward_argmax(feature_maps, rois, pooled_size, kernel_size)
            ctx.save_for_backward(rois, codes)
            return out
        ctx.save_for_backward(feature_maps, rois)
        return roi_pool_forward(feature_maps, rois, pooled_size, kernel_size)

    @staticmethod
    def backward(ctx, dy):
        # no box gradient: the proposals are stop_gradient-ed in the reference (post_processing.py:58)
        if ctx.with_codes:
            rois, codes = ctx.saved_tensors
            return roi_pool_backward_argmax(ctx.fshape, rois, dy.contiguous(), codes, *ctx.pk), None, None, None
        feature_maps, rois = ctx.saved_tensors
        return roi_pool_backward(feature_maps, rois, dy.contiguous(), *ctx.pk), None, None, None


class ROIPooling:
    """fast_rcnn_detector.py:133-177.  ``ROIPooling(pooled_size, kernel_size)``; ``call(feature_maps, rois,
    flatten=False, keep_batch_dim=False)`` (also ``__call__``), as ``FastRCNNDetector`` uses it (:23, :60)."""

    def __init__(self, pooled_size, kernel_size, **kwargs):
        self._pooled_size = int(pooled_size)
        self._kernel_size = int(kernel_size)
        self.name = kwargs.get("name", "regions_of_interest_pooling")

    def call(self, feature_maps, rois, flatten=False, keep_batch_dim=False):
        f, r = as_device(feature_maps), as_device(rois)
        if isinstance(feature_maps, torch.Tensor) and feature_maps.requires_grad:
            f = feature_maps if feature_maps.is_cuda and feature_maps.is_contiguous() else f
        pooled = _ROIPoolFn.apply(f, r, self._pooled_size, self._kernel_size)  # [B,R,P,P,C]
        B, R = r.shape[0], r.shape[1]
        if flatten:  # Flatten in (py, px, c) order (:170-171) -- a view, nothing is copied
            pooled = pooled.reshape(B, R, -1)
        if not keep_batch_dim:  # the reference regroups with split/stack (:173-175); here it is a reshape
            pooled = pooled.reshape((B * R,) + tuple(pooled.shape[2:]))
        return pooled

    __call__ = call


class FastRCNNDetector:
    """The hot-path part of fast_rcnn_detector.py:7-130: RoI pooling, the relative -> absolute regions of ``call`` (:66)
    and ``get_training_samples`` (:71-130).  The Dense heads stay with the caller (cuBLAS territory, out of scope):
    ``pool`` returns what they consume, ``regions`` what ``call`` returns next to their outputs."""

    def __init__(self, image_shape, num_classes, config, name="fast_rcnn_detector"):
        self._image_shape = image_shape
        self.num_classes = int(num_classes)
        self.name = name
        self._roi_pooling = ROIPooling(name="regions_of_interest_pooling", **config["roi_pooling"])  # :23

    def pool(self, feature_maps, rois):
        """:60 ``self._roi_pooling(feature_maps, rois, True, True)`` -> [B,R,P*P*C] (differentiable w.r.t. feature_maps)."""
        return self._roi_pooling(feature_maps, rois, True, True)

    def regions(self, rois):
        """:66 ``to_absolute(rois, self._image_shape)``."""
        return to_absolute(rois, self._image_shape)

    def get_training_samples(self, gt_labels, gt_boxes, regions, pred_scores, pred_boxes, foreground_iou_interval,
                             background_iou_interval, num_samples, foreground_proportion, **sampler):
        """FastRCNNDetector.get_training_samples, fast_rcnn_detector.py:71-130 -- same signature, called as
        ``get_training_samples(gt_labels, gt_boxes, **rcnn_output, **config["rcnn"]["sampling"])``
        (models/faster_rcnn.py:82-84).  regions [B,R,4] abs, pred_scores [B,R,nc+1], pred_boxes [B,R,nc,4] ->
        {"target_labels", "pred_scores": [B,S,nc+1]; "target_boxes", "pred_boxes": [B,S,nc,4]}."""
        return _get_training_samples(gt_labels, gt_boxes, regions, pred_scores, pred_boxes, self._image_shape,
                                     foreground_iou_interval, background_iou_interval, num_samples, foreground_proportion,
                                     objectness=False, **sampler)
