"""Drop-in for the reference's utils/training.py: ``generate_targets`` (single image, the signature the
reference calls inside tf.map_fn) and ``generate_targets_batched`` (the natural GPU form: all images in
two kernel launches).  ``get_sample_indices`` / ``sample_and_gather_batched`` follow the reference's sampling rule in one
hand-written kernel with counter-based Philox draws (the reference uses unseeded tf.random, so parity is
distributional -- SURVEY 8(f) rank 2).
"""
from __future__ import annotations

import torch

from .._tensor import as_device, check, lib, ptr, stream, workspace


def generate_targets_batched(gt_labels, gt_boxes, regions, image_shape, foreground_iou_interval,
                             background_iou_interval, return_extra=False, objectness=False):
    """Batched utils/training.py:7-77.  gt_labels [B,G,nc+1]; gt_boxes [B,G,4] rel; regions [N,4] or [B,N,4] abs.

    Returns (target_labels [B,N,nc+1], target_boxes_encoded [B,N,nc,4]) (+ max_iou [B,N], argmax [B,N] int32).
    ``objectness=True`` is the RPN's call (rpn_detector.py:141-149): the labels generate_targets sees are
    one_hot(int(sum(gt_labels, -1)), 2), computed inside the matching kernels; outputs are [B,N,2] / [B,N,1,4]."""
    gl, gb, reg = as_device(gt_labels), as_device(gt_boxes), as_device(regions)
    if gl.dim() != 3 or gb.dim() != 3 or gb.shape[-1] != 4 or gl.shape[:2] != gb.shape[:2]:
        raise ValueError(f"gt_labels [B,G,nc+1] / gt_boxes [B,G,4] expected, got {tuple(gl.shape)} / {tuple(gb.shape)}")
    B, G, Lin = gl.shape
    L = 2 if objectness else Lin
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
    fn = Lb.frcnn_generate_targets_objectness if objectness else Lb.frcnn_generate_targets
    check(fn(ptr(gl), ptr(gb), B, G, Lin if objectness else nc, ptr(reg), int(batched), N, float(image_shape[0]),
             float(image_shape[1]), float(fl), float(fh), float(bl), float(bh), ptr(tl), ptr(tb), ptr(mi), ptr(am), ptr(ws),
             ws_bytes, stream()), "generate_targets")
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
    out = torch.empty(tuple(gl.shape[:-1]) + (2,), dtype=torch.float32, device=gl.device)
    check(lib().frcnn_objectness_labels(ptr(gl), out.numel() // 2, gl.shape[-1], ptr(out), stream()), "objectness_labels")
    return out


def get_training_samples(gt_labels, gt_boxes, regions, pred_scores, pred_boxes, image_shape, foreground_iou_interval,
                         background_iou_interval, num_samples, foreground_proportion, objectness=False, generator=None,
                         seed=None, check_background=True):
    """The body both detectors share (rpn_detector.py:98-160 with ``objectness=True``, fast_rcnn_detector.py:71-130):
    generate_targets per image -> get_sample_indices per image -> the four tf.gather(batch_dims=1).  Three kernel launches
    (match, labels, sample + gather).  Returns the reference's dict: target_labels, pred_scores [B,S,nc+1];
    target_boxes, pred_boxes [B,S,nc,4]."""
    tl, tb = generate_targets_batched(gt_labels, gt_boxes, regions, image_shape, foreground_iou_interval,
                                      background_iou_interval, objectness=objectness)
    out = sample_and_gather_batched(tl, num_samples, foreground_proportion, pred_scores, tb, pred_boxes, generator=generator,
                                    seed=seed, check_background=check_background)
    return {k: out[k] for k in ("target_labels", "pred_scores", "target_boxes", "pred_boxes")}


_seed_counter = [0]


def _next_seed(generator=None, seed=None) -> int:
    """A fresh 64-bit seed per call: explicit ``seed``, else derived from the generator's (or torch's) initial seed and a
    call counter, so that successive calls draw different samples while a re-seeded run repeats itself."""
    if seed is not None:
        return int(seed) & 0xFFFFFFFFFFFFFFFF
    base = int(generator.initial_seed()) if generator is not None else int(torch.initial_seed())
    _seed_counter[0] += 1
    return (base * 0x9E3779B97F4A7C15 + _seed_counter[0] * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF


def sample_and_gather_batched(target_labels, num_samples, foreground_proportion, pred_scores=None, target_boxes=None,
                              pred_boxes=None, generator=None, seed=None, check_background=True):
    """utils/training.py:80-120 for all images at once, fused with the four tf.gather(batch_dims=1) of
    rpn_detector.py:151-160 / fast_rcnn_detector.py:121-130: ONE kernel launch (block-wide compaction of the foreground /
    background lists, Philox keys, bitonic selection, row gathers).

    target_labels [B,N,nc+1] (+ optionally pred_scores [B,N,nc+1], target_boxes / pred_boxes [B,N,nc,4]).  Returns a dict
    with "indices" [B,num_samples] int64, "counts" [B,2] int32 (#foreground, #background) and, for every tensor given, its
    sampled rows under the reference's key ("target_labels", "pred_scores", "target_boxes", "pred_boxes").
    ``check_background`` reads the counts back (one synchronisation) and raises where the reference would."""
    tl = as_device(target_labels)
    if tl.dim() != 3:
        raise ValueError(f"target_labels must be [B,N,nc+1], got {tuple(tl.shape)}")
    B, N, L = tl.shape
    S = int(num_samples)
    if L < 2 or S <= 0 or not (0.0 <= float(foreground_proportion) <= 1.0):
        raise ValueError("need nc >= 1, num_samples > 0 and 0 <= foreground_proportion <= 1")
    dev = tl.device
    srcs, outs = {}, {}
    for name, t, shape in (("pred_scores", pred_scores, (B, N, L)), ("target_boxes", target_boxes, (B, N, L - 1, 4)),
                           ("pred_boxes", pred_boxes, (B, N, L - 1, 4))):
        if t is not None:
            t = as_device(t)
            if tuple(t.shape) != shape:
                raise ValueError(f"{name} must be {shape}, got {tuple(t.shape)}")
            srcs[name] = t
            outs[name] = torch.empty((B, S) + shape[2:], dtype=torch.float32, device=dev)
    outs["target_labels"] = torch.empty((B, S, L), dtype=torch.float32, device=dev)
    idx = torch.empty((B, S), dtype=torch.int64, device=dev)
    counts = torch.empty((B, 2), dtype=torch.int32, device=dev)
    Lb = lib()
    ws_bytes = Lb.frcnn_sample_indices_workspace_bytes(B, N)
    ws = workspace(ws_bytes)
    check(Lb.frcnn_sample_and_gather(ptr(tl), B, N, L, S, float(foreground_proportion), _next_seed(generator, seed), ptr(idx),
                                     ptr(counts), ptr(srcs.get("pred_scores")), ptr(srcs.get("target_boxes")),
                                     ptr(srcs.get("pred_boxes")), ptr(outs["target_labels"]), ptr(outs.get("pred_scores")),
                                     ptr(outs.get("target_boxes")), ptr(outs.get("pred_boxes")), ptr(ws), ws_bytes, stream()),
          "sample_and_gather")
    if check_background:
        c = counts.cpu()
        quota = int(round(S * float(foreground_proportion)))
        need_bg = S - torch.clamp(c[:, 0], max=quota)
        if bool(((c[:, 1] == 0) & (need_bg > 0)).any()):
            raise ValueError("no background region to sample from")  # the reference's tf.random.uniform fails here too
    outs["indices"] = idx
    outs["counts"] = counts
    return outs


def get_sample_indices(target_labels, num_samples, foreground_proportion, generator=None, seed=None):
    """utils/training.py:80-120 for one image ([N,nc+1] -> [num_samples] int64): up to round(num_samples *
    foreground_proportion) foregrounds without replacement, the rest backgrounds WITH replacement, foregrounds first."""
    tl = as_device(target_labels)
    if tl.dim() != 2:
        raise ValueError(f"target_labels must be [N,nc+1], got {tuple(tl.shape)}")
    return sample_and_gather_batched(tl[None], num_samples, foreground_proportion, generator=generator, seed=seed)["indices"][0]
