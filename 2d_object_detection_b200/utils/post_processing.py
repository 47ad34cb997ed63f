"""Drop-in for the reference's utils/post_processing.py: ``postprocess_output`` with the same
signature, keyword names (the reference splats ``config["nms"]`` and the detector output dict into it,
models/faster_rcnn.py:53,106,107) and return dict.  One C-ABI call (frcnn_postprocess) = three kernels.
"""
from __future__ import annotations

import torch

from .._tensor import as_device, check, lib, ptr, stream, workspace


def postprocess_output(image_shape, regions, pred_scores, pred_boxes, score_threshold, iou_threshold,
                       max_output_size_per_class, max_total_size, pre_nms_top_n=None, return_extra=False,
                       from_logits=False, return_softmax=False):
    """utils/post_processing.py:6-63.

    Args (as the reference): image_shape (H, W, _); regions [N,4] or [B,N,4] abs; pred_scores [B,N,nc+1];
    pred_boxes [B,N,nc,4] encoded; the four ``config["nms"]`` values.
    ``pre_nms_top_n`` is an extension (None = all candidates = reference behaviour, SURVEY trap T1).

    Returns {"pred_boxes" [B,R,4] rel clipped to [0,1], "pred_scores" [B,R], "pred_classes" [B,R] int32,
    "num_valid_detections" [B] int32}, zero padded; with ``return_extra`` also "kept_indices" [B,R] int32
    (-1 padded) and "exhausted" [B] int32 (1 = pre_nms_top_n cut the list before the per-class cap).
    Outputs carry no gradient (the reference wraps them in tf.stop_gradient, :57-63).

    ``from_logits`` (extension): ``pred_scores`` holds the classification head's reshaped pre-activation outputs and
    the Softmax of models/detectors/rpn_detector.py:84 is fused into the decode kernel (frcnn_postprocess_logits);
    with ``return_softmax`` the probabilities come back as "softmax_scores" [B,N,nc+1] (what the detector returns as
    ``pred_scores`` and the losses read).  The selection is bit-identical to calling this function on those
    probabilities.
    """
    reg, sc, pb = as_device(regions), as_device(pred_scores), as_device(pred_boxes)
    if pb.dim() != 4 or pb.shape[-1] != 4:
        raise ValueError(f"pred_boxes must be [batch, num_regions, num_classes, 4], got {tuple(pb.shape)}")
    B, N, nc, _ = pb.shape
    if tuple(sc.shape) != (B, N, nc + 1):
        raise ValueError(f"pred_scores must be [{B}, {N}, {nc + 1}], got {tuple(sc.shape)}")
    batched = reg.dim() == 3
    if tuple(reg.shape) != ((B, N, 4) if batched else (N, 4)):
        raise ValueError(f"regions must be [{N}, 4] or [{B}, {N}, 4], got {tuple(reg.shape)}")
    R = int(max_total_size)
    per_class = int(max_output_size_per_class)
    dev = pb.device
    out_boxes = torch.empty((B, R, 4), dtype=torch.float32, device=dev)
    out_scores = torch.empty((B, R), dtype=torch.float32, device=dev)
    out_classes = torch.empty((B, R), dtype=torch.int32, device=dev)
    out_valid = torch.empty((B,), dtype=torch.int32, device=dev)
    out_idx = torch.empty((B, R), dtype=torch.int32, device=dev) if return_extra else None
    out_exh = torch.empty((B,), dtype=torch.int32, device=dev) if return_extra else None
    L = lib()
    ws_bytes = L.frcnn_postprocess_workspace_bytes(B, N, nc, per_class, R)
    ws = workspace(ws_bytes)
    if return_softmax and not from_logits:
        raise ValueError("return_softmax needs from_logits=True")
    probs = torch.empty_like(sc) if return_softmax else None
    if from_logits:
        check(L.frcnn_postprocess_logits(float(image_shape[0]), float(image_shape[1]), ptr(reg), int(batched), ptr(sc),
                                         ptr(probs), ptr(pb), B, N, nc, float(score_threshold), float(iou_threshold),
                                         per_class, R, int(pre_nms_top_n or 0), ptr(out_boxes), ptr(out_scores),
                                         ptr(out_classes), ptr(out_valid), ptr(out_idx), ptr(out_exh), ptr(ws),
                                         ws_bytes, stream()), "postprocess_output")
    else:
        check(L.frcnn_postprocess(float(image_shape[0]), float(image_shape[1]), ptr(reg), int(batched), ptr(sc),
                                  ptr(pb), B, N, nc, float(score_threshold), float(iou_threshold), per_class, R,
                                  int(pre_nms_top_n or 0), ptr(out_boxes), ptr(out_scores), ptr(out_classes),
                                  ptr(out_valid), ptr(out_idx), ptr(out_exh), ptr(ws), ws_bytes, stream()),
              "postprocess_output")
    out = {"pred_boxes": out_boxes, "pred_scores": out_scores, "pred_classes": out_classes,
           "num_valid_detections": out_valid}
    if return_softmax:
        out["softmax_scores"] = probs
    if return_extra:
        out["kept_indices"] = out_idx
        out["exhausted"] = out_exh
    return out


def decode_filter(image_shape, regions, pred_scores, pred_boxes, score_threshold):
    """Stage 1 alone (frcnn_decode_filter): returns (boxes_rel [B,nc,N,4] unclipped, keys [B,nc,N] int32 view)."""
    reg, sc, pb = as_device(regions), as_device(pred_scores), as_device(pred_boxes)
    B, N, nc, _ = pb.shape
    boxes = torch.empty((B, nc, N, 4), dtype=torch.float32, device=pb.device)
    keys = torch.empty((B, nc, N), dtype=torch.int32, device=pb.device)
    check(lib().frcnn_decode_filter(float(image_shape[0]), float(image_shape[1]), ptr(reg), int(reg.dim() == 3),
                                    ptr(sc), ptr(pb), B, N, nc, float(score_threshold), ptr(boxes), ptr(keys),
                                    stream()), "decode_filter")
    return boxes, keys


def non_max_suppression(boxes, scores, score_threshold, iou_threshold, max_output_size, pre_nms_top_n=None):
    """Stage 2 alone (frcnn_nms): greedy NMS in TF CombinedNMS order on explicit boxes.

    boxes [S,N,4] or [N,4]; scores [S,N] or [N].  Returns (kept_idx [S,cap] int32 -1 padded, kept_count [S] int32,
    exhausted [S] int32); the leading S is dropped for 2-D input."""
    b, s = as_device(boxes), as_device(scores)
    squeeze = b.dim() == 2
    if squeeze:
        b, s = b[None], s[None]
    S, N, _ = b.shape
    if tuple(s.shape) != (S, N):
        raise ValueError(f"scores must be [{S}, {N}], got {tuple(s.shape)}")
    cap = min(int(max_output_size), N)
    kept = torch.empty((S, max(cap, 1)), dtype=torch.int32, device=b.device)
    cnt = torch.empty((S,), dtype=torch.int32, device=b.device)
    exh = torch.empty((S,), dtype=torch.int32, device=b.device)
    L = lib()
    ws_bytes = L.frcnn_nms_workspace_bytes(S, N)
    ws = workspace(ws_bytes)
    check(L.frcnn_nms(ptr(b), ptr(s), S, N, float(score_threshold), float(iou_threshold), int(max_output_size),
                      int(pre_nms_top_n or 0), ptr(kept), ptr(cnt), ptr(exh), ptr(ws), ws_bytes, stream()),
          "non_max_suppression")
    kept = kept[:, :cap]
    return (kept[0], cnt[0], exh[0]) if squeeze else (kept, cnt, exh)
