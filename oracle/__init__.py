"""CPU ORACLE -- test infrastructure, not product code.

ctypes/NumPy front-end of ``oracle/frcnn_oracle.c``: a CPU restatement of the
reference's region-proposal / RoI hot path *including* the TensorFlow kernels it
calls.  Function names, argument names/orders and tensor layouts follow the
reference modules (file:line cited per function).

PARITY UNPINNED: the reference has no tests/fixtures and TensorFlow cannot be
installed in this environment (see the header of frcnn_oracle.c and DESIGN.md).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(``2d_object_detection_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libfrcnn_oracle.so")
_lib = None

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (oracle/Makefile). Returns the .so path."""
    src = os.path.join(_HERE, "frcnn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_tf_nms_iou.restype = C.c_float
        _lib.orc_valid_anchor_indices.restype = C.c_int64
        _lib.orc_nms.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a: np.ndarray, t=_f32p):
    return a.ctypes.data_as(t)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(int(n)))


# ---------------------------------------------------------------- anchors
def generate_anchors(grid_shape, scales, aspect_ratios, base_anchor_shape, stride_shape=(16, 16)) -> np.ndarray:
    """models/detectors/rpn_detector.py:162-199."""
    sc, ar = _f32(scales), _f32(aspect_ratios)
    hf, wf = int(grid_shape[0]), int(grid_shape[1])
    out = np.empty((hf * wf * sc.size * ar.size, 4), np.float32)
    lib().orc_generate_anchors(hf, wf, _p(sc), int(sc.size), _p(ar), int(ar.size),
                               C.c_float(base_anchor_shape[0]), C.c_float(base_anchor_shape[1]),
                               C.c_float(stride_shape[0]), C.c_float(stride_shape[1]), _p(out))
    return out


def valid_anchor_indices(anchors, image_shape) -> np.ndarray:
    """models/detectors/rpn_detector.py:215-224 (index list, ascending)."""
    a = _f32(anchors).reshape(-1, 4)
    idx = np.empty(a.shape[0], np.int64)
    m = lib().orc_valid_anchor_indices(_p(a), C.c_int64(a.shape[0]), C.c_float(image_shape[1]),
                                       C.c_float(image_shape[0]), _p(idx, _i64p))
    return idx[:m].copy()


# ---------------------------------------------------------------- utils/boxes.py
def clip_to_window(boxes, window) -> np.ndarray:
    """utils/boxes.py:4-17 (window = [x_min, y_min, x_max, y_max], per the code at :13)."""
    b = _f32(boxes)
    w = _f32(window)
    out = np.empty_like(b)
    lib().orc_clip_to_window(_p(b), C.c_int64(b.size // 4), _p(w), _p(out))
    return out


def decode(boxes, reference_boxes) -> np.ndarray:
    """utils/boxes.py:20-41."""
    b = _f32(boxes)
    r = _f32(np.broadcast_to(np.asarray(reference_boxes, np.float32), b.shape))
    out = np.empty_like(b)
    lib().orc_decode(_p(b), _p(r), C.c_int64(b.size // 4), C.c_int64(b.size // 4), _p(out))
    return out


def encode(boxes, reference_boxes) -> np.ndarray:
    """utils/boxes.py:44-73."""
    b = _f32(boxes)
    r = _f32(np.broadcast_to(np.asarray(reference_boxes, np.float32), b.shape))
    out = np.empty_like(b)
    with np.errstate(all="ignore"):
        lib().orc_encode(_p(b), _p(r), C.c_int64(b.size // 4), C.c_int64(b.size // 4), _p(out))
    return out


def to_absolute(boxes, image_shape) -> np.ndarray:
    """utils/boxes.py:76-83."""
    b = _f32(boxes)
    out = np.empty_like(b)
    lib().orc_to_absolute(_p(b), C.c_int64(b.size // 4), C.c_float(image_shape[0]), C.c_float(image_shape[1]), _p(out))
    return out


def to_relative(boxes, image_shape) -> np.ndarray:
    """utils/boxes.py:86-93."""
    b = _f32(boxes)
    out = np.empty_like(b)
    lib().orc_to_relative(_p(b), C.c_int64(b.size // 4), C.c_float(image_shape[0]), C.c_float(image_shape[1]), _p(out))
    return out


# ---------------------------------------------------------------- utils/metrics.py
def area(boxes) -> np.ndarray:
    """utils/metrics.py:136-147."""
    b = _f32(boxes).reshape(-1, 4)
    out = np.empty(b.shape[0], np.float32)
    lib().orc_area(_p(b), C.c_int64(b.shape[0]), _p(out))
    return out


def intersection(boxes_1, boxes_2, pairwise=False) -> np.ndarray:
    """utils/metrics.py:150-180."""
    b1, b2 = _f32(boxes_1).reshape(-1, 4), _f32(boxes_2).reshape(-1, 4)
    out = np.empty((b1.shape[0], b2.shape[0]), np.float32)
    lib().orc_intersection_pairwise(_p(b1), C.c_int64(b1.shape[0]), _p(b2), C.c_int64(b2.shape[0]), _p(out))
    return out if pairwise else np.ascontiguousarray(np.diagonal(out))


def iou(boxes_1, boxes_2, pairwise=False) -> np.ndarray:
    """utils/metrics.py:183-208."""
    b1, b2 = _f32(boxes_1).reshape(-1, 4), _f32(boxes_2).reshape(-1, 4)
    if pairwise:
        out = np.empty((b1.shape[0], b2.shape[0]), np.float32)
        lib().orc_iou_pairwise(_p(b1), C.c_int64(b1.shape[0]), _p(b2), C.c_int64(b2.shape[0]), _p(out))
        return out
    assert b1.shape == b2.shape
    out = np.empty(b1.shape[0], np.float32)
    lib().orc_iou_elementwise(_p(b1), _p(b2), C.c_int64(b1.shape[0]), _p(out))
    return out


# ---------------------------------------------------------------- NMS / post-processing
def tf_nms_iou(a, b) -> float:
    """IOU of TF's non_max_suppression_op.cc."""
    a, b = _f32(a), _f32(b)
    return float(lib().orc_tf_nms_iou(_p(a), _p(b)))


def nms(boxes, scores, score_threshold, iou_threshold, max_output_size, pre_nms_top_n=None) -> np.ndarray:
    """Single-class greedy NMS in TF CombinedNMS order; returns kept indices in keep order."""
    b, s = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    kept = np.empty(max(1, min(int(max_output_size), b.shape[0])), np.int32)
    k = lib().orc_nms(_p(b), _p(s), int(b.shape[0]), C.c_float(score_threshold), C.c_float(iou_threshold),
                      int(max_output_size), int(pre_nms_top_n or 0), _p(kept, _i32p))
    return kept[:k].copy()


def combined_non_max_suppression(boxes, scores, max_output_size_per_class, max_total_size, iou_threshold,
                                 score_threshold, pre_nms_top_n=None):
    """tf.image.combined_non_max_suppression(pad_per_class=False, clip_boxes=True); SURVEY 8.2.

    boxes [B,N,q,4], scores [B,N,nc]. Returns (boxes[B,R,4], scores[B,R], classes[B,R] int32,
    valid[B] int32, kept_region_idx[B,R] int32 (-1 padded))."""
    b, s = _f32(boxes), _f32(scores)
    B, N, q, _ = b.shape
    nc = s.shape[2]
    R = int(max_total_size)
    ob = np.empty((B, R, 4), np.float32)
    os_ = np.empty((B, R), np.float32)
    oc = np.empty((B, R), np.int32)
    ov = np.empty((B,), np.int32)
    oi = np.empty((B, R), np.int32)
    lib().orc_combined_nms(_p(b), _p(s), B, N, q, nc, int(max_output_size_per_class), R, C.c_float(iou_threshold),
                           C.c_float(score_threshold), int(pre_nms_top_n or 0), _p(ob), _p(os_), _p(oc, _i32p),
                           _p(ov, _i32p), _p(oi, _i32p))
    return ob, os_, oc, ov, oi


def postprocess_output(image_shape, regions, pred_scores, pred_boxes, score_threshold, iou_threshold,
                       max_output_size_per_class, max_total_size, pre_nms_top_n=None, return_extra=False):
    """utils/post_processing.py:6-63. ``pre_nms_top_n`` is an extension (None = reference behaviour)."""
    reg, sc, pb = _f32(regions), _f32(pred_scores), _f32(pred_boxes)
    B, N, nc, _ = pb.shape
    assert sc.shape == (B, N, nc + 1), (sc.shape, pb.shape)
    batched = reg.ndim == 3
    assert reg.shape == ((B, N, 4) if batched else (N, 4))
    R = int(max_total_size)
    ob = np.empty((B, R, 4), np.float32)
    os_ = np.empty((B, R), np.float32)
    oc = np.empty((B, R), np.int32)
    ov = np.empty((B,), np.int32)
    oi = np.empty((B, R), np.int32)
    dec = np.empty((B, N, nc, 4), np.float32)
    with np.errstate(all="ignore"):
        lib().orc_postprocess_output(C.c_float(image_shape[0]), C.c_float(image_shape[1]), _p(reg), int(batched),
                                     _p(sc), _p(pb), B, N, nc, C.c_float(score_threshold), C.c_float(iou_threshold),
                                     int(max_output_size_per_class), R, int(pre_nms_top_n or 0), _p(ob), _p(os_),
                                     _p(oc, _i32p), _p(ov, _i32p), _p(oi, _i32p), _p(dec))
    out = {"pred_boxes": ob, "pred_scores": os_, "pred_classes": oc, "num_valid_detections": ov}
    if return_extra:
        out["kept_indices"] = oi
        out["decoded_boxes"] = dec
    return out


# ---------------------------------------------------------------- RoI pooling
def roi_pool_forward(feature_maps, rois, pooled_size, kernel_size, return_argmax=False):
    """models/detectors/fast_rcnn_detector.py:142-177 -> [B,R,P,P,C] (flatten/regroup are views)."""
    f, r = _f32(feature_maps), _f32(rois)
    B, Hf, Wf, Cc = f.shape
    R = r.shape[1]
    P, k = int(pooled_size), int(kernel_size)
    assert r.shape == (B, R, 4) and P * k <= 64
    out = np.empty((B, R, P, P, Cc), np.float32)
    am = np.empty((B, R, P, P, Cc), np.uint8) if return_argmax else None
    lib().orc_roi_pool_forward(_p(f), B, Hf, Wf, Cc, _p(r), R, P, k, _p(out), _p(am, _u8p) if am is not None else None)
    return (out, am) if return_argmax else out


def roi_pool_backward(feature_maps, rois, dy, pooled_size, kernel_size) -> np.ndarray:
    """Gradient of roi_pool_forward w.r.t. feature_maps (models/faster_rcnn.py:73,103)."""
    f, r = _f32(feature_maps), _f32(rois)
    B, Hf, Wf, Cc = f.shape
    R = r.shape[1]
    P, k = int(pooled_size), int(kernel_size)
    g = _f32(dy).reshape(B, R, P, P, Cc)
    df = np.empty_like(f)
    lib().orc_roi_pool_backward(_p(f), B, Hf, Wf, Cc, _p(r), R, P, k, _p(g), _p(df))
    return df


# ---------------------------------------------------------------- utils/training.py
def generate_targets(gt_labels, gt_boxes, regions, image_shape, foreground_iou_interval, background_iou_interval,
                     return_extra=False):
    """utils/training.py:7-77 (+ _get_labels_masks :123-143), single image."""
    gl, gb, reg = _f32(gt_labels), _f32(gt_boxes), _f32(regions)
    G, L = gl.shape
    N = reg.shape[0]
    nc = L - 1
    tl = np.empty((N, L), np.float32)
    tb = np.empty((N, nc, 4), np.float32)
    mi = np.empty((N,), np.float32)
    am = np.empty((N,), np.int32)
    with np.errstate(all="ignore"):
        lib().orc_generate_targets(_p(gl), _p(gb), G, nc, _p(reg), N, C.c_float(image_shape[0]),
                                   C.c_float(image_shape[1]), C.c_float(foreground_iou_interval[0]),
                                   C.c_float(foreground_iou_interval[1]), C.c_float(background_iou_interval[0]),
                                   C.c_float(background_iou_interval[1]), _p(tl), _p(tb), _p(mi), _p(am, _i32p))
    if return_extra:
        return tl, tb, mi, am
    return tl, tb


def generate_targets_batched(gt_labels, gt_boxes, regions, image_shape, foreground_iou_interval,
                             background_iou_interval):
    """tf.map_fn(generate_targets) over images (rpn_detector.py:143-149, fast_rcnn_detector.py:113-119)."""
    gl, gb, reg = _f32(gt_labels), _f32(gt_boxes), _f32(regions)
    B, G, L = gl.shape
    batched = reg.ndim == 3
    N = reg.shape[-2]
    nc = L - 1
    tl = np.empty((B, N, L), np.float32)
    tb = np.empty((B, N, nc, 4), np.float32)
    lib().orc_generate_targets_batched(_p(gl), _p(gb), B, G, nc, _p(reg), int(batched), N, C.c_float(image_shape[0]),
                                       C.c_float(image_shape[1]), C.c_float(foreground_iou_interval[0]),
                                       C.c_float(foreground_iou_interval[1]), C.c_float(background_iou_interval[0]),
                                       C.c_float(background_iou_interval[1]), _p(tl), _p(tb))
    return tl, tb


def rpn_objectness_labels(gt_labels) -> np.ndarray:
    """rpn_detector.py:141: one_hot(int(sum(gt_labels, -1)), 2) -- padded rows become [1, 0] (trap T11)."""
    s = np.asarray(gt_labels, np.float32).sum(-1).astype(np.int32)
    inside = ((s >= 0) & (s <= 1))[..., None]  # tf.one_hot yields an all-zero row for out-of-range indices
    return (np.eye(2, dtype=np.float32)[np.clip(s, 0, 1)] * inside).astype(np.float32)
