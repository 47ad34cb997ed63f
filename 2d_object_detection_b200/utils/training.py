"""Drop-in for the reference's utils/training.py: ``generate_targets`` (single image, the signature the
reference calls inside tf.map_fn) and ``generate_targets_batched`` (the natural GPU form: all images in
two kernel launches).  ``get_sample_indices`` follows the reference's sampling rule with torch's Philox
generator (the reference uses unseeded tf.random, so parity is statistical -- SURVEY 8(f) rank 2).
"""
from __future__ import annotations

import torch

from .._tensor import as_device, check, lib, ptr, stream, workspace


def generate_targets_batched(gt_labels, gt_boxes, regions, image_shape, foreground_iou_interval,
                             background_iou_interval, return_extra=False):
    """Batched utils/training.py:7-77.  gt_labels [B,G,nc+1]; gt_boxes [B,G,4] rel; regions [N,4] or [B,N,4] abs.

    Returns (target_labels [B,N,nc+1], target_boxes_encoded [B,N,nc,4]) (+ max_iou [B,N], argmax [B,N] int32)."""
    gl, gb, reg = as_device(gt_labels), as_device(gt_boxes), as_device(regions)
    if gl.dim() != 3 or gb.dim() != 3 or gb.shape[-1] != 4 or gl.shape[:2] != gb.shape[:2]:
        raise ValueError(f"gt_labels [B,G,nc+1] / gt_boxes [B,G,4] expected, got {tuple(gl.shape)} / {tuple(gb.shape)}")
    B, G, L = gl.shape
    nc = L - 1
    batched = reg.dim() == 3
    N = reg.shape[-2]
    if reg.shape[-1] != 4 or (batched and reg.shape[0] != B):
        raise ValueError(f"regions must be [N,4] or [{B},N,4], got {tuple(reg.shape)}")
    dev = gl.device
    tl = torch.empty((B, N, L), dtype=torch.float32, device=dev)
    tb = torch.empty((B, N, nc, 4), dtype=torch.float32, device=dev)
    mi = torch.empty((B, N), dtype=torch.float32, device=dev) if return_extra else None
    am = torch.empty((B, N), dtype=torch.int32, device=dev) if return_extra else None
    Lb = lib()
    ws_bytes = Lb.frcnn_generate_targets_workspace_bytes(B, N)
    ws = workspace(ws_bytes)
    (fl, fh), (bl, bh) = foreground_iou_interval, background_iou_interval
    check(Lb.frcnn_generate_targets(ptr(gl), ptr(gb), B, G, nc, ptr(reg), int(batched), N, float(image_shape[0]),
                                    float(image_shape[1]), float(fl), float(fh), float(bl), float(bh), ptr(tl),
                                    ptr(tb), ptr(mi), ptr(am), ptr(ws), ws_bytes, stream()), "generate_targets")
    return (tl, tb, mi, am) if return_extra else (tl, tb)


def generate_targets(gt_labels, gt_boxes, regions, image_shape, foreground_iou_interval, background_iou_interval):
    """utils/training.py:7-77, single image: gt_labels [G,nc+1], gt_boxes [G,4] rel, regions [N,4] abs
    -> (target_labels [N,nc+1], target_boxes_encoded [N,nc,4])."""
    gl, gb = as_device(gt_labels), as_device(gt_boxes)
    tl, tb = generate_targets_batched(gl[None], gb[None], regions, image_shape, foreground_iou_interval,
                                      background_iou_interval)
    return tl[0], tb[0]


def rpn_objectness_labels(gt_labels):
    """rpn_detector.py:141: one_hot(int(sum(gt_labels, -1)), depth=2); padded rows become [1, 0] (trap T11)."""
    gl = as_device(gt_labels)
    s = gl.sum(-1).to(torch.int64)
    inside = ((s >= 0) & (s <= 1)).unsqueeze(-1)
    return torch.nn.functional.one_hot(s.clamp(0, 1), 2).to(torch.float32) * inside


def get_sample_indices(target_labels, num_samples, foreground_proportion, generator=None):
    """utils/training.py:80-120 for one image: up to round(num_samples*foreground_proportion) foregrounds without
    replacement (shuffle), the rest backgrounds WITH replacement (uniform draws), concatenated fg first."""
    tl = as_device(target_labels)
    nonzero = tl.sum(-1) != 0.0
    fg = torch.nonzero(nonzero & (tl[:, 0] == 0.0)).reshape(-1)
    bg = torch.nonzero(nonzero & (tl[:, 0] == 1.0)).reshape(-1)
    n_fg = min(int(fg.numel()), int(round(num_samples * foreground_proportion)))
    n_bg = int(num_samples) - n_fg
    fg = fg[torch.randperm(fg.numel(), device=tl.device, generator=generator)][:n_fg]
    if bg.numel() == 0 and n_bg > 0:
        raise ValueError("no background region to sample from")  # the reference's tf.random.uniform fails here too
    draw = torch.randint(0, max(int(bg.numel()), 1), (n_bg,), device=tl.device, generator=generator)
    return torch.cat([fg, bg[draw]], 0)
