"""Drop-in for the reference's utils/losses.py (ClassificationLoss, RegressionLoss: utils/losses.py:4-43), evaluated
on the sampled rows at models/faster_rcnn.py:87-101.  Both losses and their gradients come from ONE hand-written kernel
launch (``frcnn_sampled_losses``); ``sampled_losses`` returns the pair the train step sums, the two classes keep the
reference's call signature ``loss(target, pred)`` and are differentiable with respect to ``pred``."""
from __future__ import annotations

import torch

from .._tensor import as_device, check, lib, ptr, stream


def _run(target_labels, pred_scores, target_boxes, pred_boxes, want_grads):
    dev = (target_labels if target_labels is not None else target_boxes).device
    losses = torch.empty(3, dtype=torch.float32, device=dev)
    rows = L = 0
    box_rows = 0
    d_scores = d_boxes = None
    if target_labels is not None:
        if target_labels.shape != pred_scores.shape or target_labels.dim() < 2:
            raise ValueError(f"target_labels / pred_scores must have the same [..., nc+1] shape, got "
                             f"{tuple(target_labels.shape)} / {tuple(pred_scores.shape)}")
        L = target_labels.shape[-1]
        rows = target_labels.numel() // L
        d_scores = torch.empty_like(pred_scores) if want_grads else None
    if target_boxes is not None:
        if target_boxes.shape != pred_boxes.shape or target_boxes.shape[-1] != 4:
            raise ValueError(f"target_boxes / pred_boxes must have the same [..., 4] shape, got "
                             f"{tuple(target_boxes.shape)} / {tuple(pred_boxes.shape)}")
        box_rows = target_boxes.numel() // 4
        d_boxes = torch.empty_like(pred_boxes) if want_grads else None
    check(lib().frcnn_sampled_losses(ptr(target_labels), ptr(pred_scores), rows, max(L, 1), ptr(target_boxes), ptr(pred_boxes),
                                     box_rows, 1.0, 1.0, ptr(losses), ptr(d_scores), ptr(d_boxes), stream()), "sampled_losses")
    return losses, d_scores, d_boxes


class _SampledLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, target_labels, pred_scores, target_boxes, pred_boxes):
        want = (pred_scores is not None and pred_scores.requires_grad) or (pred_boxes is not None and pred_boxes.requires_grad)
        losses, d_scores, d_boxes = _run(target_labels, pred_scores, target_boxes, pred_boxes, want)
        ctx.save_for_backward(*(t for t in (d_scores, d_boxes) if t is not None))
        ctx.has = (d_scores is not None, d_boxes is not None)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_cls, g_reg):
        saved = list(ctx.saved_tensors)
        d_scores = saved.pop(0) * g_cls if ctx.has[0] else None  # the kernel stored the gradients for unit upstream grads
        d_boxes = saved.pop(0) * g_reg if ctx.has[1] else None
        return None, d_scores, None, d_boxes


def _prep(t):
    return None if t is None else as_device(t)


def sampled_losses(target_labels, pred_scores, target_boxes, pred_boxes):
    """(classification loss, regression loss) of one detector's training samples, e.g. the dict returned by
    ``sample_and_gather_batched``: one kernel launch, differentiable w.r.t. ``pred_scores`` and ``pred_boxes``."""
    return _SampledLosses.apply(_prep(target_labels), _prep(pred_scores), _prep(target_boxes), _prep(pred_boxes))


class ClassificationLoss:
    """utils/losses.py:4-19: ``loss(target_class_labels, pred_class_scores)`` with [B,S,nc+1] tensors."""

    def __init__(self, name="classification_loss"):
        self.name = name

    def __call__(self, target_class_labels, pred_class_scores):
        return _SampledLosses.apply(_prep(target_class_labels), _prep(pred_class_scores), None, None)[0]


class RegressionLoss:
    """utils/losses.py:22-43: ``loss(target_boxes_encoded, pred_boxes_encoded)`` with [B,S,nc,4] tensors."""

    def __init__(self, name="regression_loss"):
        self.name = name

    def __call__(self, target_boxes_encoded, pred_boxes_encoded):
        return _SampledLosses.apply(None, None, _prep(target_boxes_encoded), _prep(pred_boxes_encoded))[1]
