"""Second, independent CPU restatement (NumPy fp32 + Python loops) -- test infrastructure only.

Exists to cross-check ``frcnn_oracle.c``: the two were written separately from
the reference Python (file:line cited) and the TF kernel semantics in SURVEY 8.2,
so agreement between them guards against a transcription slip in either.
Small inputs only (the NMS / crop loops are pure Python).  PARITY UNPINNED, as
the C oracle.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def generate_anchors(grid_shape, scales, aspect_ratios, base_anchor_shape, stride_shape=(16, 16)):
    """rpn_detector.py:162-199 via numpy meshgrid (xy indexing, like tf.meshgrid)."""
    sc, ar = np.meshgrid(np.asarray(scales, F), np.asarray(aspect_ratios, F))
    sc, ar = sc.reshape(-1), ar.reshape(-1)
    rs = np.sqrt(ar)
    hs = sc / rs * F(base_anchor_shape[0])
    ws = sc * rs * F(base_anchor_shape[1])
    xc = np.arange(grid_shape[1], dtype=F) * F(stride_shape[1])
    yc = np.arange(grid_shape[0], dtype=F) * F(stride_shape[0])
    xc, yc = np.meshgrid(xc, yc)
    ws, xc = np.meshgrid(ws, xc.reshape(-1))
    hs, yc = np.meshgrid(hs, yc.reshape(-1))
    ctr = np.stack([xc, yc], axis=2).reshape(-1, 2)
    sz = np.stack([ws, hs], axis=2).reshape(-1, 2)
    return np.concatenate([ctr - F(0.5) * sz, ctr + F(0.5) * sz], axis=1).astype(F)


def clip_to_window(boxes, window):
    """boxes.py:4-17."""
    x0, y0, x1, y1 = [F(v) for v in window]
    hi = np.array([x1, y1, x1, y1], F)
    lo = np.array([x0, y0, x0, y0], F)
    return np.maximum(np.minimum(np.asarray(boxes, F), hi), lo)


def decode(boxes, reference_boxes):
    """boxes.py:20-41."""
    boxes = np.asarray(boxes, F)
    ref = np.asarray(reference_boxes, F)
    rmin, rmax = ref[..., :2], ref[..., 2:]
    rc = (rmax + rmin) / F(2.0)
    rs = rmax - rmin
    c = boxes[..., :2] * rs + rc
    with np.errstate(all="ignore"):
        s = np.exp(boxes[..., 2:]) * rs
    return np.concatenate([c - F(0.5) * s, c + F(0.5) * s], axis=-1).astype(F)


def encode(boxes, reference_boxes):
    """boxes.py:44-73."""
    boxes = np.asarray(boxes, F)
    ref = np.asarray(reference_boxes, F)
    rc = (ref[..., 2:] + ref[..., :2]) / F(2.0)
    rs = ref[..., 2:] - ref[..., :2]
    c = (boxes[..., 2:] + boxes[..., :2]) / F(2.0)
    s = boxes[..., 2:] - boxes[..., :2]
    with np.errstate(all="ignore"):
        return np.concatenate([(c - rc) / rs, np.log(s / rs)], axis=-1).astype(F)


def to_absolute(boxes, image_shape):
    h, w = F(image_shape[0]), F(image_shape[1])
    return np.asarray(boxes, F) * np.array([w, h, w, h], F)


def to_relative(boxes, image_shape):
    h, w = F(image_shape[0]), F(image_shape[1])
    return np.asarray(boxes, F) / np.array([w, h, w, h], F)


def iou(boxes_1, boxes_2, pairwise=False):
    """metrics.py:136-208."""
    a, b = np.asarray(boxes_1, F), np.asarray(boxes_2, F)
    if pairwise:
        a, b = a[:, None, :], b[None, :, :]
    dw = np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0])
    dh = np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1])
    inter = np.maximum(F(0), dw) * np.maximum(F(0), dh)
    ar_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    ar_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    with np.errstate(all="ignore"):
        return np.where(inter == 0, F(0), inter / (ar_a + ar_b - inter)).astype(F)


def tf_nms_iou(a, b):
    """IOU() of TF non_max_suppression_op.cc."""
    a, b = np.asarray(a, F), np.asarray(b, F)
    a0, a1, a2, a3 = min(a[0], a[2]), min(a[1], a[3]), max(a[0], a[2]), max(a[1], a[3])
    b0, b1, b2, b3 = min(b[0], b[2]), min(b[1], b[3]), max(b[0], b[2]), max(b[1], b[3])
    area_a = F(a2 - a0) * F(a3 - a1)
    area_b = F(b2 - b0) * F(b3 - b1)
    if area_a <= 0 or area_b <= 0:
        return F(0)
    ih = max(F(min(a2, b2) - max(a0, b0)), F(0))
    iw = max(F(min(a3, b3) - max(a1, b1)), F(0))
    inter = F(ih * iw)
    with np.errstate(all="ignore"):
        return F(inter / F(F(area_a + area_b) - inter))


def nms(boxes, scores, score_threshold, iou_threshold, max_output_size, pre_nms_top_n=None):
    """Greedy NMS of CombinedNMS, one class; ties: lower index first (T4)."""
    boxes, scores = np.asarray(boxes, F), np.asarray(scores, F)
    cand = [i for i in range(len(scores)) if scores[i] > F(score_threshold)]
    cand.sort(key=lambda i: (-float(scores[i]), i))
    if pre_nms_top_n:
        cand = cand[: int(pre_nms_top_n)]
    kept = []
    cap = min(int(max_output_size), len(scores))
    for i in cand:
        if len(kept) >= cap:
            break
        if all(not (tf_nms_iou(boxes[i], boxes[j]) > F(iou_threshold)) for j in reversed(kept)):
            kept.append(i)
    return np.asarray(kept, np.int32)


def combined_non_max_suppression(boxes, scores, max_output_size_per_class, max_total_size, iou_threshold,
                                 score_threshold, pre_nms_top_n=None):
    """SURVEY 8.2 CombinedNMS (pad_per_class=False, clip_boxes=True)."""
    boxes, scores = np.asarray(boxes, F), np.asarray(scores, F)
    B, N, q, _ = boxes.shape
    nc = scores.shape[2]
    R = int(max_total_size)
    ob = np.zeros((B, R, 4), F)
    os_ = np.zeros((B, R), F)
    oc = np.zeros((B, R), np.int32)
    ov = np.zeros((B,), np.int32)
    oi = np.full((B, R), -1, np.int32)
    for b in range(B):
        merged = []
        for c in range(nc):
            qi = c if q > 1 else 0
            kept = nms(boxes[b, :, qi], scores[b, :, c], score_threshold, iou_threshold, max_output_size_per_class,
                       pre_nms_top_n)
            merged += [(float(scores[b, i, c]), len(merged) + t, c, int(i)) for t, i in enumerate(kept)]
        merged.sort(key=lambda m: (-m[0], m[1]))
        merged = merged[:R]
        ov[b] = len(merged)
        for t, (s, _, c, i) in enumerate(merged):
            ob[b, t] = np.clip(boxes[b, i, c if q > 1 else 0], F(0), F(1))
            os_[b, t], oc[b, t], oi[b, t] = s, c, i
    return ob, os_, oc, ov, oi


def postprocess_output(image_shape, regions, pred_scores, pred_boxes, score_threshold, iou_threshold,
                       max_output_size_per_class, max_total_size, pre_nms_top_n=None):
    """post_processing.py:6-63."""
    pred_boxes = np.asarray(pred_boxes, F)
    regions = np.expand_dims(np.asarray(regions, F), -2)  # :40-43 (broadcast == tile)
    rel = to_relative(decode(pred_boxes, np.broadcast_to(regions, pred_boxes.shape)), image_shape)
    ob, os_, oc, ov, oi = combined_non_max_suppression(rel, np.asarray(pred_scores, F)[..., 1:],
                                                       max_output_size_per_class, max_total_size, iou_threshold,
                                                       score_threshold, pre_nms_top_n)
    return {"pred_boxes": ob, "pred_scores": os_, "pred_classes": oc, "num_valid_detections": ov,
            "kept_indices": oi, "decoded_boxes": rel}


def _axis_samples(c1, c2, extent, S):
    """crop_and_resize sampling positions along one axis."""
    c1, c2 = F(c1), F(c2)
    res = []
    for i in range(S):
        if S > 1:
            scale = F(F(c2 - c1) * F(extent - 1)) / F(S - 1)
            pos = F(F(c1 * F(extent - 1)) + F(F(i) * scale))
        else:
            pos = F(F(F(0.5) * F(c1 + c2)) * F(extent - 1))
        if not (pos >= 0 and pos <= F(extent - 1)):  # NaN -> not sampled (undefined in TF; defined choice, see frcnn_oracle.c)
            res.append(None)
        else:
            lo = int(np.floor(pos))
            res.append((lo, int(np.ceil(pos)), F(pos - F(lo))))
    return res


def crop_and_resize(image, boxes_yxyx, box_indices, crop_size):
    """tf.image.crop_and_resize, bilinear, extrapolation_value 0; image [B,H,W,C], boxes [n,4] y1,x1,y2,x2."""
    image = np.asarray(image, F)
    _, H, W, Cc = image.shape
    S_h, S_w = crop_size
    out = np.zeros((len(boxes_yxyx), S_h, S_w, Cc), F)
    for n, (y1, x1, y2, x2) in enumerate(np.asarray(boxes_yxyx, F)):
        ys = _axis_samples(y1, y2, H, S_h)
        xs = _axis_samples(x1, x2, W, S_w)
        img = image[box_indices[n]]
        for iy, sy in enumerate(ys):
            if sy is None:
                continue
            for ix, sx in enumerate(xs):
                if sx is None:
                    continue
                tl, tr = img[sy[0], sx[0]], img[sy[0], sx[1]]
                bl, br = img[sy[1], sx[0]], img[sy[1], sx[1]]
                top = tl + (tr - tl) * sx[2]
                bot = bl + (br - bl) * sx[2]
                out[n, iy, ix] = top + (bot - top) * sy[2]
    return out


def roi_pool_forward(feature_maps, rois, pooled_size, kernel_size):
    """fast_rcnn_detector.py:142-177 -> [B,R,P,P,C]."""
    feature_maps, rois = np.asarray(feature_maps, F), np.asarray(rois, F)
    B, R = rois.shape[:2]
    S = pooled_size * kernel_size
    flat = rois.reshape(-1, 4)[:, [1, 0, 3, 2]]  # :157-158
    crops = crop_and_resize(feature_maps, flat, np.repeat(np.arange(B), R), (S, S))  # :160-166
    k, P = kernel_size, pooled_size
    pooled = crops.reshape(B * R, P, k, P, k, -1).max(axis=(2, 4))  # :168 MaxPool2D(k), VALID
    return pooled.reshape(B, R, P, P, -1)


def roi_pool_backward(feature_maps, rois, dy, pooled_size, kernel_size):
    """MaxPoolGrad (first max, row-major window order) + CropAndResizeGradImage."""
    feature_maps, rois = np.asarray(feature_maps, F), np.asarray(rois, F)
    B, H, W, Cc = feature_maps.shape
    R = rois.shape[1]
    P, k = pooled_size, kernel_size
    S = P * k
    dy = np.asarray(dy, F).reshape(B, R, P, P, Cc)
    flat = rois.reshape(-1, 4)[:, [1, 0, 3, 2]]
    crops = crop_and_resize(feature_maps, flat, np.repeat(np.arange(B), R), (S, S)).reshape(B, R, P, k, P, k, Cc)
    win = crops.transpose(0, 1, 2, 4, 3, 5, 6).reshape(B, R, P, P, k * k, Cc)
    am = win.argmax(axis=4)  # numpy argmax = first maximum, row-major (dy,dx)
    df = np.zeros_like(feature_maps)
    ch = np.arange(Cc)
    for b in range(B):
        for r in range(R):
            x1, y1, x2, y2 = rois[b, r]
            ys = _axis_samples(y1, y2, H, S)
            xs = _axis_samples(x1, x2, W, S)
            for py in range(P):
                for px in range(P):
                    a = am[b, r, py, px]
                    for w_i in range(k * k):
                        sel = ch[a == w_i]
                        if sel.size == 0:
                            continue
                        sy, sx = ys[py * k + w_i // k], xs[px * k + w_i % k]
                        if sy is None or sx is None:
                            continue
                        g = dy[b, r, py, px, sel]
                        dtop = (F(1) - sy[2]) * g
                        dbot = sy[2] * g
                        df[b, sy[0], sx[0], sel] += (F(1) - sx[2]) * dtop
                        df[b, sy[0], sx[1], sel] += sx[2] * dtop
                        df[b, sy[1], sx[0], sel] += (F(1) - sx[2]) * dbot
                        df[b, sy[1], sx[1], sel] += sx[2] * dbot
    return df


def generate_targets(gt_labels, gt_boxes, regions, image_shape, foreground_iou_interval, background_iou_interval):
    """training.py:7-77 and :123-143."""
    gt_labels, gt_boxes, regions = np.asarray(gt_labels, F), np.asarray(gt_boxes, F), np.asarray(regions, F)
    nc = gt_labels.shape[1] - 1
    N = regions.shape[0]
    keep = gt_labels.sum(-1) != 0  # :43-45
    gt_boxes, gt_labels = gt_boxes[keep], gt_labels[keep]
    tl = np.zeros((N, nc + 1), F)
    tb = np.zeros((N, nc, 4), F)
    if gt_boxes.shape[0] == 0:  # T13: undefined in the reference; all-zero IoU, no forced foreground
        lo, hi = background_iou_interval
        if lo <= 0 < hi:
            tl[:, 0] = 1
        return tl, tb
    gabs = to_absolute(gt_boxes, image_shape)  # :48
    ious = iou(regions, gabs, pairwise=True)  # :51
    amax = ious.argmax(-1)
    mx = ious.max(-1)
    gmx = mx.max()
    (fl, fh), (bl, bh) = foreground_iou_interval, background_iou_interval
    bg = (mx >= F(bl)) & (mx < F(bh))
    fg = (mx >= F(fl)) & (mx < F(fh))
    fg[np.nonzero(mx == gmx)[0][0]] = True  # :137-140 forced foreground = first arg of global max
    tl[bg, 0] = 1
    tl[fg] = gt_labels[amax[fg]]
    enc = encode(gabs[amax], regions)
    mask = tl[:, 1:] != 0
    tb[mask] = np.broadcast_to(enc[:, None, :], (N, nc, 4))[mask]
    return tl, tb


# ---------------------------------------------------------------------------------------------------------------
# Classification-head epilogue, models/detectors/rpn_detector.py:81-84 (1x1 conv -> Reshape((-1, 2)) -> Softmax) and
# fast_rcnn_detector.py:60-61 (second half).  tf.keras.layers.Softmax over the last axis in the form of
# TF's GPU kernel: exp(l - max) / sum(exp(l - max)), every operation rounded to float32, sum left to right.
def softmax(logits):
    l = np.asarray(logits, np.float32)
    e = np.exp(l - l.max(-1, keepdims=True)).astype(np.float32)
    tot = np.zeros(l.shape[:-1], np.float32)
    for j in range(l.shape[-1]):
        tot = (tot + e[..., j]).astype(np.float32)
    return (e / tot[..., None]).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# utils/losses.py:4-43.  tf.keras.losses.CategoricalCrossentropy() on probabilities and
# tf.keras.losses.Huber(reduction=NONE) with the reference's row filter; float32 element arithmetic, float64 only for
# the final reductions (Eigen's reduction order is unspecified: compare within 1e-5).
def classification_loss(target_labels, pred_scores):
    t = np.asarray(target_labels, np.float32).reshape(-1, np.shape(target_labels)[-1])
    p = np.asarray(pred_scores, np.float32).reshape(t.shape)
    q = p / p.sum(-1, keepdims=True, dtype=np.float32)             # Keras rescales probabilities to sum to 1
    c = np.clip(q, np.float32(1e-7), np.float32(1.0) - np.float32(1e-7))  # backend.epsilon()
    rows = -(t * np.log(c)).sum(-1, dtype=np.float64)
    return float(rows.mean())                                       # Reduction.SUM_OVER_BATCH_SIZE


def regression_loss(target_boxes, pred_boxes):
    t = np.asarray(target_boxes, np.float32).reshape(-1, 4)
    p = np.asarray(pred_boxes, np.float32).reshape(-1, 4)
    keep = t.sum(-1, dtype=np.float32) != 0.0                      # utils/losses.py:36-38
    e = p[keep] - t[keep]
    a = np.abs(e)
    h = np.where(a <= 1.0, np.float32(0.5) * e * e, a - np.float32(0.5))  # Huber, delta = 1
    return float(h.mean(-1, dtype=np.float64).sum())                # mean over the 4 coordinates, SUM over the rows


# ---------------------------------------------------------------------------------------------------------------
# utils/metrics.py:4-84 AveragePrecision, restated for one stream of updates.
def _iou_ref(a, b):
    """utils/metrics.py:150-208 on two boxes (float32)."""
    F = np.float32
    iw = max(F(0), F(min(a[2], b[2]) - max(a[0], b[0])))
    ih = max(F(0), F(min(a[3], b[3]) - max(a[1], b[1])))
    inter = F(iw * ih)
    if inter == 0:
        return F(0)
    area_a = F(F(a[2] - a[0]) * F(a[3] - a[1]))
    area_b = F(F(b[2] - b[0]) * F(b[3] - b[1]))
    return F(inter / F(F(area_a + area_b) - inter))


def ap_true_positives(gt_boxes, pred_boxes, iou_threshold):
    """update_state_single_image (:57-84): -> (true_positives [M] int32, number of non-padding gt boxes)."""
    gt = np.asarray(gt_boxes, np.float32)
    pb = np.asarray(pred_boxes, np.float32)
    gt = gt[gt.sum(-1, dtype=np.float32) != 0.0]
    tp = np.zeros(len(pb), np.int32)
    for g in gt:
        ious = np.array([_iou_ref(g, p) for p in pb], np.float32)
        if len(ious):
            j = int(np.argmax(ious))                         # first maximum, like tf.math.argmax
            if ious[j] > np.float32(iou_threshold):
                tp[j] = 1
    return tp, len(gt)


def average_precision(true_pos, scores, true_count, num_points=11):
    """result() (:25-42) for the concatenated state."""
    true_pos, scores = np.asarray(true_pos), np.asarray(scores, np.float32)
    order = np.argsort(-scores, kind="stable")
    cum_tp = np.cumsum(true_pos[order]).astype(np.float32)
    cum_pos = (np.arange(len(scores)) + 1).astype(np.float32)
    precisions = np.concatenate([cum_tp / cum_pos, [np.float32(0.0)]]).astype(np.float32)
    recalls = np.concatenate([cum_tp / np.float32(true_count), [np.float32(2.0)]]).astype(np.float32)
    levels = np.arange(num_points, dtype=np.float32) / np.float32(num_points - 1)
    return float(np.mean([precisions[recalls >= r].max() for r in levels], dtype=np.float32))
