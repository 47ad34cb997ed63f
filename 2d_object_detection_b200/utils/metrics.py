"""Drop-in for the IoU primitives of the reference's utils/metrics.py:136-208."""
from __future__ import annotations

import torch

from .._tensor import as_device, check, lib, ptr, stream


def _boxes2d(x, name):
    t = as_device(x)
    if t.dim() != 2 or t.shape[1] != 4:
        raise ValueError(f"{name} must have shape [N, 4], got {tuple(t.shape)}")
    return t


def area(boxes):
    """utils/metrics.py:136-147 -> [N]."""
    b = _boxes2d(boxes, "boxes")
    out = torch.empty(b.shape[0], dtype=torch.float32, device=b.device)
    check(lib().frcnn_area(ptr(b), b.shape[0], ptr(out), stream()), "area")
    return out


def intersection(boxes_1, boxes_2, pairwise=False):
    """utils/metrics.py:150-180 -> [N, M] if pairwise else [N]."""
    b1, b2 = _boxes2d(boxes_1, "boxes_1"), _boxes2d(boxes_2, "boxes_2")
    if not pairwise:
        if b1.shape != b2.shape:
            raise ValueError("elementwise intersection needs equally shaped inputs")
        out = torch.empty(b1.shape[0], dtype=torch.float32, device=b1.device)
        check(lib().frcnn_intersection_elementwise(ptr(b1), ptr(b2), b1.shape[0], ptr(out), stream()), "intersection")
        return out
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    check(lib().frcnn_intersection_pairwise(ptr(b1), b1.shape[0], ptr(b2), b2.shape[0], ptr(out), stream()),
          "intersection")
    return out


def iou(boxes_1, boxes_2, pairwise=False):
    """utils/metrics.py:183-208: inter / (area_1 + area_2 - inter), 0 where inter == 0."""
    b1, b2 = _boxes2d(boxes_1, "boxes_1"), _boxes2d(boxes_2, "boxes_2")
    if pairwise:
        out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
        check(lib().frcnn_iou_pairwise(ptr(b1), b1.shape[0], ptr(b2), b2.shape[0], ptr(out), stream()), "iou")
        return out
    if b1.shape != b2.shape:
        raise ValueError("elementwise iou needs equally shaped inputs")
    out = torch.empty(b1.shape[0], dtype=torch.float32, device=b1.device)
    check(lib().frcnn_iou_elementwise(ptr(b1), ptr(b2), b1.shape[0], ptr(out), stream()), "iou")
    return out
