"""Drop-in for the reference's utils/boxes.py (same names, argument order and layouts).

Boxes are ``[..., 4]`` fp32 ``[x_min, y_min, x_max, y_max]``; every function runs one sm_100a kernel
through the C ABI (frcnn_clip_to_window / frcnn_decode / frcnn_encode / frcnn_scale_boxes).
"""
from __future__ import annotations

import torch

from .._tensor import as_device, check, lib, ptr, stream


def clip_to_window(boxes, window):
    """utils/boxes.py:4-17.  ``window`` is ``[x_min, y_min, x_max, y_max]`` (what the code at :13 unpacks;
    the reference docstring's y-first order is wrong -- SURVEY trap T10)."""
    b = as_device(boxes)
    if b.shape[-1] != 4:
        raise ValueError(f"boxes must have a trailing dimension of 4, got {tuple(b.shape)}")
    w = [float(v) for v in (window.tolist() if hasattr(window, "tolist") else window)]
    out = torch.empty_like(b)
    check(lib().frcnn_clip_to_window(ptr(b), b.numel() // 4, w[0], w[1], w[2], w[3], ptr(out), stream()),
          "clip_to_window")
    return out


def _codec(fn, name, boxes, reference_boxes):
    b = as_device(boxes)
    r = as_device(reference_boxes)
    if b.shape[-1] != 4 or r.shape[-1] != 4:
        raise ValueError(f"{name}: boxes and reference_boxes need a trailing dimension of 4")
    n = b.numel() // 4
    if r.numel() == b.numel():
        inner, period = 1, max(n, 1)
    else:
        try:
            r = r.expand(b.shape).contiguous()
        except RuntimeError as e:
            raise ValueError(f"{name}: reference_boxes {tuple(r.shape)} not broadcastable with {tuple(b.shape)}") from e
        inner, period = 1, max(n, 1)
    out = torch.empty_like(b)
    check(fn(ptr(b), ptr(r), n, inner, period, ptr(out), stream()), name)
    return out


def decode(boxes, reference_boxes):
    """utils/boxes.py:20-41 (reverse of encode; no delta clamp, separate mul/add, IEEE exp)."""
    return _codec(lib().frcnn_decode, "decode", boxes, reference_boxes)


def encode(boxes, reference_boxes):
    """utils/boxes.py:44-73: t = [(x-xr)/wr, (y-yr)/hr, log(w/wr), log(h/hr)]."""
    return _codec(lib().frcnn_encode, "encode", boxes, reference_boxes)


def _scale(boxes, image_shape, divide, name):
    b = as_device(boxes)
    if b.shape[-1] != 4:
        raise ValueError(f"{name}: boxes must have a trailing dimension of 4")
    image_height, image_width = float(image_shape[0]), float(image_shape[1])
    out = torch.empty_like(b)
    check(lib().frcnn_scale_boxes(ptr(b), b.numel() // 4, image_height, image_width, divide, ptr(out), stream()), name)
    return out


def to_absolute(boxes, image_shape):
    """utils/boxes.py:76-83: multiply by [W, H, W, H]."""
    return _scale(boxes, image_shape, 0, "to_absolute")


def to_relative(boxes, image_shape):
    """utils/boxes.py:86-93: true division by [W, H, W, H]."""
    return _scale(boxes, image_shape, 1, "to_relative")
