"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): bit-exact kept-index sets for NMS / top-k given identical fp32 boxes;
fp32 within 1e-5 relative for decoded boxes, RoI-pool outputs and gradients (the RoI-pool forward and the
IoU matrix are in fact bit-identical to the oracle, which the tests assert).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
F = np.float32
RTOL = 1e-5  # the tolerance north_star states for floating point


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def host(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------ elementwise box utils
def test_boxes_elementwise(det, orc, rng):
    b = np.sort(rng.uniform(0, 600, (3, 50, 2, 2)).astype(F), axis=2).reshape(3, 50, 4) + F([0, 0, 1, 1])
    r = np.sort(rng.uniform(0, 600, (50, 2, 2)).astype(F), axis=1).reshape(50, 4) + F([0, 0, 1, 1])
    t = rng.normal(0, 0.3, (3, 50, 4)).astype(F)
    np.testing.assert_array_equal(host(det.clip_to_window(dev(b), [10, 20, 500, 300])), orc.clip_to_window(b, [10, 20, 500, 300]))
    np.testing.assert_allclose(host(det.decode(dev(t), dev(r))), orc.decode(t, r), rtol=RTOL, atol=1e-3)
    np.testing.assert_allclose(host(det.encode(dev(b), dev(r))), orc.encode(b, r), rtol=RTOL, atol=1e-5)
    np.testing.assert_array_equal(host(det.to_relative(dev(b), (375, 1242, 3))), orc.to_relative(b, (375, 1242, 3)))
    np.testing.assert_array_equal(host(det.to_absolute(dev(b / 600), (375, 1242, 3))), orc.to_absolute(b / 600, (375, 1242, 3)))
    # exp(0) == 1 exactly, so zero deltas decode bit-identically to the oracle (== ref up to the centre/size rounding)
    np.testing.assert_array_equal(host(det.decode(dev(np.zeros((50, 4), F)), dev(r))), orc.decode(np.zeros((50, 4), F), r))
    # broadcast over classes as post_processing.py:40-46 does
    tt = rng.normal(0, 0.1, (2, 50, 3, 4)).astype(F)
    np.testing.assert_allclose(host(det.decode(dev(tt), dev(r[:, None, :]))), orc.decode(tt, r[:, None, :]), rtol=RTOL, atol=1e-3)
    # numpy inputs are uploaded, not computed on the host
    assert det.decode(t, r).is_cuda


def test_anchors_and_valid_indices_exact(det, orc, syn):
    for name in ("default_b2", "kitti_b8", "dense_b16"):
        w = syn.WORKLOADS[name]
        a = det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
        ref = orc.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
        np.testing.assert_array_equal(host(a), ref)
        idx = det.valid_anchor_indices(a, w.image_shape)
        assert idx.dtype == torch.int64
        np.testing.assert_array_equal(host(idx), orc.valid_anchor_indices(ref, w.image_shape))


def test_rpn_anchors_train_eval_modes(det, orc, syn, rng):
    w = syn.WORKLOADS["kitti_b8_ref12"]
    cfg = {"anchors": dict(scales=list(w.scales), aspect_ratios=list(w.aspect_ratios), base_anchor_shape=[256, 256])}
    ra = det.RPNAnchors(w.image_shape, (None, w.grid[0], w.grid[1], 1024), cfg)
    scores, deltas = syn.rpn_head_outputs(rng, 2, w.num_anchors)
    ref = orc.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
    ev = ra(dev(scores), dev(deltas), training=False)
    np.testing.assert_array_equal(host(ev["regions"]), orc.clip_to_window(ref, [0, 0, w.image_shape[1], w.image_shape[0]]))
    tr = ra(dev(scores), dev(deltas), training=True)
    idx = orc.valid_anchor_indices(ref, w.image_shape)
    assert len(idx) == 8768  # N_valid for KITTI with the reference's 12 anchors (T16)
    np.testing.assert_array_equal(host(tr["regions"]), ref[idx])
    np.testing.assert_array_equal(host(tr["pred_scores"]), scores[:, idx])
    np.testing.assert_array_equal(host(tr["pred_boxes"]), deltas[:, idx])


def test_iou_bit_exact(det, orc, syn, rng):
    a = syn.random_rois(rng, 1, 300)[0] * F([1242, 375, 1242, 375])
    b = np.concatenate([syn.random_rois(rng, 1, 40)[0] * F([1242, 375, 1242, 375]), np.zeros((10, 4), F)])
    np.testing.assert_array_equal(host(det.iou(dev(a), dev(b), pairwise=True)), orc.iou(a, b, pairwise=True))
    np.testing.assert_array_equal(host(det.iou(dev(a[:50]), dev(b), pairwise=False)), orc.iou(a[:50], b))
    np.testing.assert_array_equal(host(det.intersection(dev(a), dev(b), pairwise=True)), orc.intersection(a, b, pairwise=True))
    np.testing.assert_array_equal(host(det.intersection(dev(a[:50]), dev(b))), orc.intersection(a[:50], b))
    np.testing.assert_array_equal(host(det.area(dev(a))), orc.area(a))


# ------------------------------------------------------------------ NMS stage: bit-exact kept indices
def _decoded(orc, syn, wname, batch, seed):
    w = syn.WORKLOADS[wname]
    rng = np.random.default_rng(seed)
    anchors = orc.clip_to_window(orc.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape),
                                 [0, 0, w.image_shape[1], w.image_shape[0]])
    scores, deltas = syn.rpn_head_outputs(rng, batch, w.num_anchors)
    boxes = orc.to_relative(orc.decode(deltas[:, :, 0], anchors), w.image_shape)
    return w, anchors, scores, deltas, boxes


@pytest.mark.parametrize("wname,batch,top_n,cap", [("kitti_b8", 4, None, 300), ("kitti_b8", 2, 2000, 300),
                                                   ("kitti_b8_ref12", 2, None, 300), ("default_b2", 1, None, 300),
                                                   ("dense_b16", 1, 12000, 1000), ("kitti_b8", 1, 100, 300),
                                                   ("kitti_b8", 1, None, 2000)])
def test_nms_kept_indices_bit_exact(det, orc, syn, wname, batch, top_n, cap):
    w, anchors, scores, deltas, boxes = _decoded(orc, syn, wname, batch, 7)
    kept, cnt, exh = det.non_max_suppression(dev(boxes), dev(scores[..., 1]), 0.0, 0.7, cap, top_n)
    kept, cnt, exh = host(kept), host(cnt), host(exh)
    for b in range(batch):
        ref = orc.nms(boxes[b], scores[b, :, 1], 0.0, 0.7, cap, top_n)
        assert cnt[b] == len(ref)
        np.testing.assert_array_equal(kept[b, :cnt[b]], ref)
        assert (kept[b, cnt[b]:] == -1).all()
        if top_n is not None and len(ref) < cap:
            assert exh[b] == 1  # the cap cut the list short: flagged (T1)
        if top_n is None:
            assert exh[b] == 0


def test_nms_heavy_suppression_many_chunks(det, orc, rng):
    """Spatially clustered boxes: most candidates are suppressed, so the greedy loop walks many 512-candidate chunks."""
    n = 20000
    ctr = rng.uniform(0.1, 0.9, (40, 2))[rng.integers(0, 40, n)] + rng.normal(0, 0.01, (n, 2))
    wh = rng.uniform(0.08, 0.12, (n, 2))
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], -1).astype(F)
    scores = (rng.permutation(n).astype(F) + 1) / (n + 1)
    for thr, cap in ((0.5, 300), (0.3, 1000), (0.7, 50)):
        kept, cnt, _ = det.non_max_suppression(dev(boxes), dev(scores), 0.0, thr, cap)
        ref = orc.nms(boxes, scores, 0.0, thr, cap)
        np.testing.assert_array_equal(host(kept)[:int(cnt)], ref)


def test_nms_score_ties_lower_index_first(det, orc, rng):
    n = 3000
    xy = rng.uniform(0, 1, (n, 2))
    boxes = np.concatenate([xy, xy + rng.uniform(0.02, 0.2, (n, 2))], -1).astype(F)
    for scores in (np.full(n, 0.5, F), (rng.integers(0, 7, n) / 8 + 0.0625).astype(F),
                   np.repeat(rng.uniform(0.1, 0.9, n // 600).astype(F), 600)):
        kept, cnt, _ = det.non_max_suppression(dev(boxes), dev(scores), 0.0, 0.5, 700)
        np.testing.assert_array_equal(host(kept)[:int(cnt)], orc.nms(boxes, scores, 0.0, 0.5, 700))
        kept, cnt, _ = det.non_max_suppression(dev(boxes), dev(scores), 0.0, 0.5, 700, 900)
        np.testing.assert_array_equal(host(kept)[:int(cnt)], orc.nms(boxes, scores, 0.0, 0.5, 700, 900))


@pytest.mark.parametrize("n", [4097, 4098, 4099, 40000, 45001])
def test_nms_key_layouts_of_the_radix_select(det, orc, rng, n):
    """The radix select works on keys relative to the largest key and starts below the leading bits all candidates share:
    score sets whose keys span the whole float range, a few thousand consecutive floats (narrow last key digit), one
    binade, with the best candidate at index 0 / at the end, for segment lengths that leave an unaligned tail
    (n % 4 != 0) and on either side of the limit up to which a segment's keys stay in shared memory (~40 k)."""
    xy = rng.uniform(0, 1, (n, 2))
    boxes = np.concatenate([xy, xy + rng.uniform(0.02, 0.15, (n, 2))], -1).astype(F)
    perm = rng.permutation(n)
    consecutive = (np.float32(0.5).view(np.uint32) + perm.astype(np.uint32)).view(F)  # n consecutive floats above 0.5
    wide = np.exp(rng.uniform(-80.0, 0.0, n)).astype(F)                                # 1e-35 .. 1
    binade = (0.5 + 0.5 * rng.permutation(n) / n).astype(F)
    first = binade.copy(); first[0] = 0.999999
    last = binade.copy(); last[-1] = 0.999999
    for scores in (consecutive, wide, binade, first, last):
        for cap, top_n in ((300, None), (64, 1000)):
            kept, cnt, _ = det.non_max_suppression(dev(boxes), dev(scores), 0.0, 0.6, cap, top_n)
            np.testing.assert_array_equal(host(kept)[:int(cnt)], orc.nms(boxes, scores, 0.0, 0.6, cap, top_n))


def test_nms_edge_cases(det, orc, rng):
    boxes = np.array([[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30]], F)
    scores = np.array([0.9, 0.8, 0.7], F)
    for args in ((0.0, 0.5, 10), (0.0, 0.7, 10), (0.0, 0.5, 1), (0.75, 0.7, 10), (0.7, 0.7, 10), (0.95, 0.5, 10), (0.0, 0.5, 0)):
        kept, cnt, _ = det.non_max_suppression(dev(boxes), dev(scores), *args)
        np.testing.assert_array_equal(host(kept)[:int(cnt)], orc.nms(boxes, scores, *args))
    # IoU exactly at the threshold is kept (strict >), degenerate and flipped boxes, negative / NaN scores
    b2 = np.array([[0, 0, 2, 2], [0, 0, 2, 1], [5, 5, 5, 9], [10, 10, 0, 0], [0, 0, 10, 10]], F)
    s2 = np.array([0.9, 0.8, 0.7, 0.6, np.nan], F)
    for thr in (0.5, 0.49):
        kept, cnt, _ = det.non_max_suppression(dev(b2), dev(s2), 0.0, thr, 10)
        np.testing.assert_array_equal(host(kept)[:int(cnt)], orc.nms(b2, s2, 0.0, thr, 10))
    s3 = np.array([-0.5, -0.2, -0.9, 0.1, -0.3], F)
    kept, cnt, _ = det.non_max_suppression(dev(b2), dev(s3), -1.0, 0.5, 10)
    np.testing.assert_array_equal(host(kept)[:int(cnt)], orc.nms(b2, s3, -1.0, 0.5, 10))


# ------------------------------------------------------------------ postprocess_output end to end
def _check_post(det, orc, image_shape, regions, scores, deltas, kw, top_n=None):
    out = det.postprocess_output(image_shape, dev(regions), dev(scores), dev(deltas), **kw, pre_nms_top_n=top_n, return_extra=True)
    ref = orc.postprocess_output(image_shape, regions, scores, deltas, **kw, pre_nms_top_n=top_n, return_extra=True)
    assert out["pred_classes"].dtype == torch.int32 and out["num_valid_detections"].dtype == torch.int32  # T14
    np.testing.assert_array_equal(host(out["num_valid_detections"]), ref["num_valid_detections"])
    np.testing.assert_array_equal(host(out["kept_indices"]), ref["kept_indices"])  # bit-exact index sets
    np.testing.assert_array_equal(host(out["pred_classes"]), ref["pred_classes"])
    np.testing.assert_array_equal(host(out["pred_scores"]), ref["pred_scores"])
    np.testing.assert_allclose(host(out["pred_boxes"]), ref["pred_boxes"], rtol=RTOL, atol=1e-6)
    return out, ref


@pytest.mark.parametrize("wname,batch,top_n", [("kitti_b8", 8, None), ("kitti_b8", 8, 2000), ("kitti_b8_ref12", 3, None),
                                               ("default_b2", 2, None)])
def test_postprocess_rpn_form(det, orc, syn, wname, batch, top_n):
    w, anchors, scores, deltas, _ = _decoded(orc, syn, wname, batch, 20241)
    out, ref = _check_post(det, orc, w.image_shape, anchors, scores, deltas, w.nms_kwargs(), top_n)
    pb = host(out["pred_boxes"])
    assert pb.min() >= 0 and pb.max() <= 1  # clipped on output only (T2)
    assert (host(out["num_valid_detections"]) == 300).all()


def test_postprocess_dense_b16_end_to_end(det, orc, syn):
    """BASELINE configs[3] through the whole postprocess_output (decode + score keys + pre-NMS top 12000 + NMS + finalize,
    R = 1000, N = 117,600 anchors/image), two images of the dense workload: kept-index sets bit-exact."""
    w, anchors, scores, deltas, _ = _decoded(orc, syn, "dense_b16", 2, 20243)
    assert w.num_anchors == 117600 and w.max_total_size == 1000 and w.pre_nms_top_n == 12000
    out, ref = _check_post(det, orc, w.image_shape, anchors, scores, deltas, w.nms_kwargs(), w.pre_nms_top_n)
    assert (host(out["num_valid_detections"]) > 300).all()


def test_postprocess_decode_stage_and_unclipped_iou(det, orc, syn):
    w, anchors, scores, deltas, boxes = _decoded(orc, syn, "kitti_b8", 2, 5)
    brel, keys = det.decode_filter(w.image_shape, dev(anchors), dev(scores), dev(deltas), 0.0)
    brel = host(brel)[:, 0]
    np.testing.assert_allclose(brel, boxes, rtol=RTOL, atol=1e-6)
    assert brel.min() < 0 or brel.max() > 1  # unclipped boxes feed NMS (T2)
    # given the GPU's own decoded boxes, the kept set must be bit-exact
    out = det.postprocess_output(w.image_shape, dev(anchors), dev(scores), dev(deltas), **w.nms_kwargs(), return_extra=True)
    for b in range(2):
        ref = orc.nms(brel[b], scores[b, :, 1], 0.0, 0.7, 300)
        np.testing.assert_array_equal(host(out["kept_indices"])[b, :len(ref)], ref)


def test_postprocess_rcnn_form_multiclass(det, orc, syn, rng):
    B, R, nc = 4, 300, 7
    image_shape = (375, 1242, 3)
    rois = syn.random_rois(rng, B, R) * F([1242, 375, 1242, 375])
    logits = rng.normal(0, 1.5, (B, R, nc + 1)).astype(F)
    sc = (np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)).astype(F)
    dl = rng.normal(0, 0.1, (B, R, nc, 4)).astype(F)
    kw = dict(score_threshold=0.0, iou_threshold=0.6, max_output_size_per_class=100, max_total_size=300)
    out, ref = _check_post(det, orc, image_shape, rois, sc, dl, kw)
    assert (host(out["num_valid_detections"]) == 300).all()
    kw2 = dict(score_threshold=0.3, iou_threshold=0.5, max_output_size_per_class=10, max_total_size=300)
    out, ref = _check_post(det, orc, image_shape, rois, sc, dl, kw2)  # fewer than max_total valid -> zero padding
    nv = host(out["num_valid_detections"])
    assert (nv < 300).all()
    assert (host(out["pred_boxes"])[0, nv[0]:] == 0).all() and (host(out["pred_scores"])[0, nv[0]:] == 0).all()


def _check_fused_softmax(det, orc, image_shape, regions, logits, deltas, kw):
    """frcnn_postprocess_logits: probabilities within a few ulp of the oracle's softmax, selection bit-identical to the
    unfused call on those probabilities, and both equal to the oracle's postprocess_output on them."""
    from oracle import reference_np as ref
    fused = det.postprocess_output(image_shape, dev(regions), dev(logits), dev(deltas), **kw, return_extra=True,
                                   from_logits=True, return_softmax=True)
    probs = fused["softmax_scores"]
    np.testing.assert_allclose(host(probs), ref.softmax(logits), rtol=2e-6, atol=1e-30)  # expf: <= 2 ulp, then one division
    plain = det.postprocess_output(image_shape, dev(regions), probs, dev(deltas), **kw, return_extra=True)
    for k in ("pred_boxes", "pred_scores", "pred_classes", "num_valid_detections", "kept_indices", "exhausted"):
        assert torch.equal(fused[k], plain[k]), k
    _check_post(det, orc, image_shape, regions, host(probs), deltas, kw)
    nosm = det.postprocess_output(image_shape, dev(regions), dev(logits), dev(deltas), **kw, return_extra=True, from_logits=True)
    assert "softmax_scores" not in nosm and torch.equal(nosm["kept_indices"], fused["kept_indices"])
    return fused


def test_postprocess_from_logits_rpn_head_epilogue(det, orc, syn, rng):
    # rpn_detector.py:81-84: the cls head's reshaped outputs [B, N, 2] -> Softmax, fused into the decode kernel
    w, anchors, _, deltas, _ = _decoded(orc, syn, "kitti_b8", 4, 77)
    logits = rng.normal(0, 1.5, (4, w.num_anchors, 2)).astype(F)
    logits[0, :5] = [[90.0, -90.0], [-90.0, 90.0], [0.0, 0.0], [np.inf, 0.0], [np.nan, 1.0]]  # saturated, tie, inf, NaN rows
    with np.errstate(invalid="ignore"):
        out = _check_fused_softmax(det, orc, w.image_shape, anchors, logits, deltas, w.nms_kwargs())
    p = host(out["softmax_scores"])
    np.testing.assert_array_equal(p[0, :3], F([[1, 0], [0, 1], [0.5, 0.5]]))
    assert np.isnan(p[0, 3:5]).all() and not np.isin(host(out["kept_indices"])[0], [3, 4]).any()  # NaN never a candidate
    assert (host(out["num_valid_detections"]) == 300).all()
    with pytest.raises(ValueError):
        det.postprocess_output(w.image_shape, dev(anchors), dev(logits), dev(deltas), **w.nms_kwargs(), return_softmax=True)


def test_postprocess_from_logits_multiclass(det, orc, syn, rng):
    # fast_rcnn_detector.py:60-61 Dense -> Softmax over num_classes + 1
    B, R, nc = 3, 300, 7
    rois = syn.random_rois(rng, B, R) * F([1242, 375, 1242, 375])
    logits = rng.normal(0, 1.5, (B, R, nc + 1)).astype(F)
    dl = rng.normal(0, 0.1, (B, R, nc, 4)).astype(F)
    for kw in (dict(score_threshold=0.0, iou_threshold=0.6, max_output_size_per_class=100, max_total_size=300),
               dict(score_threshold=0.3, iou_threshold=0.5, max_output_size_per_class=10, max_total_size=300)):
        _check_fused_softmax(det, orc, (375, 1242, 3), rois, logits, dl, kw)


def test_postprocess_small_and_empty(det, orc, syn, rng):
    anchors = orc.generate_anchors((3, 4), [0.5, 1.0], [1.0], [64, 64])
    scores, deltas = syn.rpn_head_outputs(rng, 2, anchors.shape[0])
    kw = dict(score_threshold=0.0, iou_threshold=0.7, max_output_size_per_class=300, max_total_size=300)
    _check_post(det, orc, (48, 64, 3), anchors, scores, deltas, kw)  # N < max_total: cap = N (T17)
    kw["score_threshold"] = 2.0  # nothing passes the threshold
    out, _ = _check_post(det, orc, (48, 64, 3), anchors, scores, deltas, kw)
    assert (host(out["num_valid_detections"]) == 0).all() and (host(out["pred_boxes"]) == 0).all()
    with pytest.raises(ValueError):
        det.postprocess_output((48, 64, 3), dev(anchors), dev(scores[:, :-1]), dev(deltas), **kw)


# ------------------------------------------------------------------ RoI pooling
@pytest.mark.parametrize("grid,C,R,relu", [((24, 78), 1024, 40, True), ((6, 9), 64, 7, False), ((38, 125), 256, 12, True),
                                           ((5, 5), 6, 4, True)])
def test_roi_pool_forward_bit_exact(det, orc, syn, rng, grid, C, R, relu):
    f = syn.feature_maps(rng, 2, grid, C, relu=relu)
    rois = syn.random_rois(rng, 2, R, num_valid=[R - 2, R])  # padded [0,0,0,0] rows still pooled (T7)
    rois[1, 0] = [0.25, 0.1, 1.0, 1.0]  # y2 = x2 = 1.0 exactly (T6)
    out = det.roi_pool_forward(dev(f), dev(rois), 7, 2)
    np.testing.assert_array_equal(host(out), orc.roi_pool_forward(f, rois, 7, 2))
    pooled = det.ROIPooling(pooled_size=7, kernel_size=2)(dev(f), dev(rois), True, True)
    assert tuple(pooled.shape) == (2, R, 49 * C)
    np.testing.assert_array_equal(host(pooled).reshape(2, R, 7, 7, C), host(out))
    assert tuple(det.ROIPooling(7, 2)(dev(f), dev(rois)).shape) == (2 * R, 7, 7, C)


@pytest.mark.parametrize("grid,C", [((24, 78), 1024), ((50, 84), 512), ((9, 11), 512), ((3, 2), 1536), ((1, 1), 512)])
def test_roi_pool_forward_row_streaming_edge_cases(det, orc, syn, rng, grid, C):
    """C % 512 == 0 with the 7x7 / k=2 configuration takes the persistent row-streaming TMA kernel: every box width class, the zero
    pixel for samples outside the map, reversed boxes, RoIs wider than the ring (plain gather path), degenerate boxes,
    1-pixel maps -- all bit-exact against the oracle."""
    f = syn.feature_maps(rng, 2, grid, C, relu=False)
    R = 64
    rois = syn.random_rois(rng, 2, R, num_valid=[R - 3, R])
    w = np.linspace(0.0, 1.0, 24, dtype=F)  # one RoI per x-extent, sweeping all box widths up to the whole map
    rois[0, :24] = np.stack([0.5 - w / 2, np.full_like(w, 0.2), 0.5 + w / 2, np.full_like(w, 0.8)], -1)
    rois[1, 0] = [0.25, 0.1, 1.0, 1.0]       # x2 = y2 = 1.0 exactly: integer position on the last column / row (T6)
    rois[1, 1] = [0.0, 0.0, 1.0, 1.0]        # whole map: wider than the ring at 78 / 84 columns
    rois[1, 2] = [0.9, 0.8, 0.1, 0.2]        # reversed: columns and rows walk backwards
    rois[1, 3] = [-0.2, 0.3, 0.4, 1.3]       # partly outside: zero pixel + crop rows after the map
    rois[1, 4] = [-0.5, -0.6, -0.1, -0.2]    # completely outside: all zeros
    rois[1, 5] = [0.5, 0.5, 0.5, 0.5]        # degenerate point
    rois[1, 6] = [0.3, -0.4, 0.6, 0.5]       # crop rows before the map
    rois[1, 7] = [np.nan, 0.1, 0.5, 0.6]     # NaN coordinate: nothing sampled along x
    rois[1, 8] = [0.0, 0.45, 1.0, 0.55]      # whole width, two feature rows
    ref = orc.roi_pool_forward(f, rois, 7, 2)
    out = host(det.roi_pool_forward(dev(f), dev(rois), 7, 2))
    np.testing.assert_array_equal(out, ref)


def test_roi_pool_integer_positions_read_one_pixel_further_finite_inputs(det, orc, rng):
    """The streaming kernels read the right / bottom neighbour of a sample at lo + 1 with weight exactly 0 where TF reads
    ceil(in) == lo (integer sample positions).  For FINITE maps -- the documented precondition of the RoI kernels
    (include/frcnn_b200.h) -- (next - a) * 0 == (a - a) * 0 == 0, so results are bit-identical; this pins it on RoIs whose
    every sample sits on an integer position, including the last column / row where lo + 1 is the next feature row, the
    next image or the TMA zero fill.  (A non-finite neighbour would give NaN here and the sampled value in TF.)"""
    Hf, Wf, C = 14, 27, 512
    f = (rng.standard_normal((2, Hf, Wf, C)) * 1e30).astype(F)  # huge but finite neighbours
    rois = np.zeros((2, 4, 4), F)
    rois[:, 0] = [0.0, 0.0, 0.5, 1.0]                 # x samples on pixels 0, 1, ..., 13; y samples on rows 0 .. 13
    rois[:, 1] = [0.5, 0.0, 1.0, 1.0]                 # x samples on pixels 13 .. 26: the last one is the last column
    rois[:, 2] = [1.0, 1.0, 1.0, 1.0]                 # every sample on the last pixel of the map
    rois[:, 3] = [13 / 26, 0.0, 13 / 26, 1.0]         # one integer column
    ref = orc.roi_pool_forward(f, rois, 7, 2)
    np.testing.assert_array_equal(host(det.roi_pool_forward(dev(f), dev(rois), 7, 2)), ref)
    out, codes = det.roi_pool_forward_argmax(dev(f), dev(rois), 7, 2)
    np.testing.assert_array_equal(host(out), ref)
    dy = rng.standard_normal(ref.shape).astype(F)
    g = host(det.roi_pool_backward_argmax(f.shape, dev(rois), dev(dy), codes, 7, 2))
    gref = orc.roi_pool_backward(f, rois, dy, 7, 2)
    np.testing.assert_allclose(g, gref, rtol=1e-5, atol=1e-5 * np.abs(gref).max())


def test_roi_pool_other_window_sizes_and_out_of_range(det, orc, syn, rng):
    f = syn.feature_maps(rng, 1, (9, 11), 8, relu=False)
    rois = np.array([[[-0.2, -0.1, 0.6, 0.7], [0.3, 0.4, 1.3, 1.2], [0.5, 0.5, 0.5, 0.5], [0.9, 0.8, 0.1, 0.2]]], F)
    for P, k in ((7, 2), (4, 1), (3, 3), (1, 2), (1, 1), (8, 4)):
        np.testing.assert_array_equal(host(det.roi_pool_forward(dev(f), dev(rois), P, k)), orc.roi_pool_forward(f, rois, P, k))


@pytest.mark.parametrize("grid,C,R,relu", [((24, 78), 256, 30, True), ((6, 9), 64, 9, False), ((7, 5), 3, 5, True),
                                           ((12, 20), 1024, 12, True), ((9, 11), 512, 8, False), ((24, 78), 512, 48, True)])
def test_roi_pool_backward(det, orc, syn, rng, grid, C, R, relu):
    f = syn.feature_maps(rng, 2, grid, C, relu=relu)
    rois = syn.random_rois(rng, 2, R, num_valid=[R - 3, R])
    rois[1, 0] = [0.0, 0.0, 1.0, 1.0]          # whole map: 4-row bin-rows, right neighbour beyond the last column
    rois[1, 1] = [0.9, 0.8, 0.1, 0.2]          # flipped box: cells walk right to left
    rois[1, 2] = [-0.2, 0.3, 0.4, 1.3]         # partly outside: zero-filled samples can win the max
    if R >= 40:                                # every row-buffer width of the streaming backward, up to the atomic path
        w = np.linspace(0.0, 0.6, 24, dtype=F)
        rois[0, :24] = np.stack([0.45 - w / 2, np.full_like(w, 0.1), 0.45 + w / 2, np.full_like(w, 0.9)], -1)
    dy = rng.standard_normal((2, R, 7, 7, C)).astype(F)
    ref = orc.roi_pool_backward(f, rois, dy, 7, 2)
    got = host(det.roi_pool_backward(dev(f), dev(rois), dev(dy), 7, 2))
    # summation order differs (atomics vs the oracle's sequential adds): 1e-5 relative with an absolute floor
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())
    # autograd route (what the train step uses) gives the same gradient; no gradient flows to the boxes
    ft = dev(f).requires_grad_(True)
    rt = dev(rois).requires_grad_(True)
    det.ROIPooling(7, 2)(ft, rt, True, True).backward(dev(dy).reshape(2, R, -1))
    np.testing.assert_allclose(host(ft.grad), ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())
    assert rt.grad is None
    # explicit saved-argmax pair (what autograd uses when the shape supports it) == recompute path
    if C % 4 == 0:
        out_a, codes = det.roi_pool_forward_argmax(dev(f), dev(rois), 7, 2)
        np.testing.assert_array_equal(host(out_a), orc.roi_pool_forward(f, rois, 7, 2))
        _, am = orc.roi_pool_forward(f, rois, 7, 2, return_argmax=True)  # uint8 window position per channel
        packed = (am[..., 0::4] | (am[..., 1::4] << 2) | (am[..., 2::4] << 4) | (am[..., 3::4] << 6)).astype(np.uint8)
        np.testing.assert_array_equal(host(codes), packed)  # first-max tie rule (T8), bit for bit
        got2 = host(det.roi_pool_backward_argmax(f.shape, dev(rois), dev(dy), codes, 7, 2))
        np.testing.assert_allclose(got2, ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())


# ------------------------------------------------------------------ generate_targets
def test_generate_targets_rpn_and_rcnn(det, orc, syn, rng):
    w = syn.WORKLOADS["kitti_b8_ref12"]
    anchors = orc.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
    regions = anchors[orc.valid_anchor_indices(anchors, w.image_shape)]
    B = 5
    gl, gb = syn.ground_truth(rng, B, allow_empty=False)
    obj = orc.rpn_objectness_labels(gl)
    np.testing.assert_array_equal(host(det.rpn_objectness_labels(dev(gl))), obj)
    tl, tb, mi, am = det.generate_targets_batched(dev(obj), dev(gb), dev(regions), w.image_shape, [0.7, 1.0], [0.0, 0.3], return_extra=True)
    for b in range(B):
        rl, rb, rmi, ram = orc.generate_targets(obj[b], gb[b], regions, w.image_shape, [0.7, 1.0], [0.0, 0.3], return_extra=True)
        np.testing.assert_array_equal(host(mi)[b], rmi)  # IoU is bit-exact
        np.testing.assert_array_equal(host(am)[b], ram)
        np.testing.assert_array_equal(host(tl)[b], rl)
        np.testing.assert_allclose(host(tb)[b], rb, rtol=RTOL, atol=1e-6)
        assert host(tl)[b][:, 1].sum() >= 1  # at least one foreground (forced, T12)
    # RCNN form: per-image regions [B,300,4], 7 classes, padding really removed
    rois = syn.random_rois(rng, B, 300) * F([w.image_shape[1], w.image_shape[0]] * 2)
    tl, tb = det.generate_targets_batched(dev(gl), dev(gb), dev(rois), w.image_shape, [0.5, 1.0], [0.0, 0.5])
    for b in range(B):
        rl, rb = orc.generate_targets(gl[b], gb[b], rois[b], w.image_shape, [0.5, 1.0], [0.0, 0.5])
        np.testing.assert_array_equal(host(tl)[b], rl)
        np.testing.assert_allclose(host(tb)[b], rb, rtol=RTOL, atol=1e-6)
    # single-image signature of the reference
    a, c = det.generate_targets(dev(gl[0]), dev(gb[0]), dev(rois[0]), w.image_shape, [0.5, 1.0], [0.0, 0.5])
    np.testing.assert_array_equal(host(a), host(tl)[0])


def test_generate_targets_kat_and_no_gt(det, orc):
    image_shape = (100, 200, 3)
    regions = np.array([[0, 0, 50, 50], [0, 0, 50, 40], [0, 0, 50, 20], [100, 50, 150, 100], [0, 0, 50, 50]], F)
    gb = np.zeros((4, 4), F)
    gb[0] = [0, 0, 0.25, 0.5]
    gl = np.zeros((4, 3), F)
    gl[0, 2] = 1
    tl, tb = det.generate_targets(dev(gl), dev(gb), dev(regions), image_shape, [0.7, 1.0], [0.0, 0.3])
    np.testing.assert_array_equal(host(tl), [[0, 0, 1], [0, 0, 1], [0, 0, 0], [1, 0, 0], [0, 0, 0]])  # T12
    rl, rb = orc.generate_targets(gl, gb, regions, image_shape, [0.7, 1.0], [0.0, 0.3])
    np.testing.assert_allclose(host(tb), rb, rtol=RTOL, atol=1e-7)
    tl, tb = det.generate_targets(dev(np.zeros((4, 3), F)), dev(gb), dev(regions), image_shape, [0.5, 1.0], [0.0, 0.5])
    np.testing.assert_array_equal(host(tl), np.tile([1, 0, 0], (5, 1)))  # T13: documented choice
    np.testing.assert_array_equal(host(tb), 0)


def test_get_sample_indices_rule(det, rng):
    tl = np.zeros((1000, 2), F)
    tl[:40, 1] = 1
    tl[100:900, 0] = 1
    g = torch.Generator(device="cuda").manual_seed(0)
    idx = host(det.get_sample_indices(dev(tl), 256, 0.5, generator=g))
    assert idx.shape == (256,) and idx.dtype == np.int64  # T14
    assert (idx[:40] < 40).all() and len(set(idx[:40].tolist())) == 40  # all 40 foregrounds, no repeats
    assert ((idx[40:] >= 100) & (idx[40:] < 900)).all()


def test_sample_and_gather_batched_rule_and_distribution(det, rng):
    """utils/training.py:80-120 + the gathers of rpn_detector.py:155-160, batched.  The reference draws from an unseeded
    generator, so the checks are the rule (which regions may appear where, how many, no repeats among foregrounds), the
    gathers (bit-exact against torch.gather), reproducibility per seed, and the distributions of the draws."""
    B, N, nc, S = 3, 9000, 7, 64
    L = nc + 1
    tl = np.zeros((B, N, L), F)
    fg_sets, bg_sets = [], []
    for b, n_fg in enumerate((5, 16, 5000)):  # fewer than / exactly / far more than the quota of round(64 * 0.25) = 16
        perm = rng.permutation(N)
        fg, bg = np.sort(perm[:n_fg]), np.sort(perm[n_fg:n_fg + 3000])  # the remaining rows are all-zero: ignored
        tl[b, fg, 1 + rng.integers(0, nc, n_fg)] = 1
        tl[b, bg, 0] = 1
        fg_sets.append(set(fg.tolist()))
        bg_sets.append(set(bg.tolist()))
    ps = rng.standard_normal((B, N, L)).astype(F)
    tb = rng.standard_normal((B, N, nc, 4)).astype(F)
    pb = rng.standard_normal((B, N, nc, 4)).astype(F)
    out = det.sample_and_gather_batched(dev(tl), S, 0.25, dev(ps), dev(tb), dev(pb), seed=1234)
    idx, counts = host(out["indices"]), host(out["counts"])
    assert idx.shape == (B, S) and idx.dtype == np.int64
    for b, n_fg_all in enumerate((5, 16, 5000)):
        n_fg = min(n_fg_all, 16)
        assert counts[b].tolist() == [n_fg_all, 3000]
        assert set(idx[b, :n_fg].tolist()) <= fg_sets[b] and len(set(idx[b, :n_fg].tolist())) == n_fg
        assert set(idx[b, n_fg:].tolist()) <= bg_sets[b]
    gi = torch.as_tensor(idx).cuda()
    for key, src in (("target_labels", tl), ("pred_scores", ps)):
        ref = torch.gather(dev(src), 1, gi[..., None].expand(-1, -1, L))
        assert torch.equal(out[key], ref)
    for key, src in (("target_boxes", tb), ("pred_boxes", pb)):
        ref = torch.gather(dev(src), 1, gi[..., None, None].expand(-1, -1, nc, 4))
        assert torch.equal(out[key], ref)
    # same seed -> same sample; another seed -> another one
    again = det.sample_and_gather_batched(dev(tl), S, 0.25, seed=1234)
    other = det.sample_and_gather_batched(dev(tl), S, 0.25, seed=1235)
    assert torch.equal(again["indices"], out["indices"]) and not torch.equal(other["indices"], out["indices"])
    # distributions: 40 foregrounds, quota 8 -> each is chosen with probability 1/5; 50 backgrounds drawn uniformly
    N2, trials = 200, 400
    t2 = np.zeros((trials, N2, 2), F)
    t2[:, :40, 1] = 1
    t2[:, 100:150, 0] = 1
    d = host(det.sample_and_gather_batched(dev(t2), 32, 0.25, seed=99)["indices"])  # every image gets its own draws
    fg_hits = np.bincount(d[:, :8].ravel(), minlength=N2)[:40]
    bg_hits = np.bincount(d[:, 8:].ravel(), minlength=N2)[100:150]
    assert fg_hits.sum() == trials * 8 and bg_hits.sum() == trials * 24
    chi_fg = ((fg_hits - trials * 8 / 40) ** 2 / (trials * 8 / 40)).sum()  # 39 degrees of freedom: 99.9 % quantile 72
    chi_bg = ((bg_hits - trials * 24 / 50) ** 2 / (trials * 24 / 50)).sum()  # 49 degrees of freedom: 99.9 % quantile 85
    assert chi_fg < 80 and chi_bg < 95, (chi_fg, chi_bg)
    first = np.bincount(d[:, 0], minlength=N2)[:40]  # the FIRST foreground slot is uniform too (a shuffle, not a sorted subset)
    assert ((first - trials / 40) ** 2 / (trials / 40)).sum() < 80
    # an image that needs background draws but has none: the reference raises
    t3 = np.zeros((1, 10, 2), F)
    t3[0, :3, 1] = 1
    with pytest.raises(ValueError):
        det.sample_and_gather_batched(dev(t3), 8, 0.5)
    # more foregrounds than one sort pass holds (chunked selection): still distinct foregrounds only
    t4 = np.zeros((1, 20000, 2), F)
    t4[0, :, 1] = 1
    t4[0, ::7, :] = [1, 0]
    i4 = host(det.sample_and_gather_batched(dev(t4), 256, 0.5, seed=5)["indices"])[0]
    assert len(set(i4[:128].tolist())) == 128 and (i4[:128] % 7 != 0).all() and (i4[128:] % 7 == 0).all()


# ------------------------------------------------------------------ host-buffer pipeline == device API
def test_host_pipeline_matches_device_api(det, orc, syn):
    w, anchors, scores, deltas, _ = _decoded(orc, syn, "kitti_b8", 2, 11)
    f = syn.feature_maps(np.random.default_rng(3), 2, w.grid, 128)
    pipe = det.ProposalRoIPipeline(w.image_shape, 2, w.grid, 128, w.num_anchors, 0.0, 0.7, 300, pre_nms_top_n=2000)
    rois, rs, nv, pooled = pipe(anchors, scores, deltas, f)
    out = det.postprocess_output(w.image_shape, dev(anchors), dev(scores), dev(deltas), **w.nms_kwargs(), pre_nms_top_n=2000)
    np.testing.assert_array_equal(rois.numpy(), host(out["pred_boxes"]))
    np.testing.assert_array_equal(rs.numpy(), host(out["pred_scores"]))
    np.testing.assert_array_equal(nv.numpy(), host(out["num_valid_detections"]))
    np.testing.assert_array_equal(pooled.numpy(), host(det.roi_pool_forward(dev(f), out["pred_boxes"], 7, 2)))
    np.testing.assert_array_equal(pooled.numpy(), orc.roi_pool_forward(f, rois.numpy(), 7, 2))
    pipe.close()


# ------------------------------------------------------------------ full-size properties (BASELINE sizes)
def test_full_size_properties_kitti_b8(det, orc, syn):
    w, anchors, scores, deltas, boxes = _decoded(orc, syn, "kitti_b8", 8, 20241)
    out = det.postprocess_output(w.image_shape, dev(anchors), dev(scores), dev(deltas), **w.nms_kwargs(), return_extra=True)
    idx, sc = host(out["kept_indices"]), host(out["pred_scores"])
    for b in range(8):
        assert len(set(idx[b].tolist())) == 300 and (np.diff(sc[b]) < 0).all()  # distinct, strictly descending
        kb = boxes[b][idx[b]]
        iou = np.array([[orc.tf_nms_iou(kb[i], kb[j]) for j in range(0, 300, 7)] for i in range(0, 300, 5)])
        assert (iou[iou < 0.999] <= F(0.7)).all()  # independent set at the threshold
    f = syn.feature_maps(np.random.default_rng(1), 8, w.grid, 1024)
    pooled = det.roi_pool_forward(dev(f), out["pred_boxes"], 7, 2)
    assert tuple(pooled.shape) == (8, 300, 7, 7, 1024)
    assert float(pooled.min()) >= 0.0 and float(pooled.max()) <= float(f.max())  # max of convex combinations
    const = det.roi_pool_forward(torch.full((8, 24, 78, 1024), 2.5, device="cuda"), out["pred_boxes"], 7, 2)
    assert bool((const == 2.5).all())  # RoI pool of a constant map is the constant
    # idempotence / determinism of the whole forward
    again = det.postprocess_output(w.image_shape, dev(anchors), dev(scores), dev(deltas), **w.nms_kwargs(), return_extra=True)
    assert torch.equal(again["kept_indices"], out["kept_indices"]) and torch.equal(again["pred_boxes"], out["pred_boxes"])
    # bit-exact pooled features for one image against the oracle at full width
    np.testing.assert_array_equal(host(pooled[3]), orc.roi_pool_forward(f[3:4], host(out["pred_boxes"])[3:4], 7, 2)[0])


def test_full_size_backward_properties_kitti_b8(det, orc, syn):
    """BASELINE size (8 x 300 RoIs x 1024 channels): the streaming backward (arg-max codes, TMA reduce-add) against the
    independent recompute kernel, conservation of gradient mass per channel, linearity in dY, and the oracle on one image."""
    w, anchors, scores, deltas, _ = _decoded(orc, syn, "kitti_b8", 8, 777)
    out = det.postprocess_output(w.image_shape, dev(anchors), dev(scores), dev(deltas), **w.nms_kwargs())
    rois = out["pred_boxes"]
    f = syn.feature_maps(np.random.default_rng(2), 8, w.grid, 1024)
    ft = dev(f)
    dy = torch.randn((8, 300, 7, 7, 1024), device="cuda")
    pooled, codes = det.roi_pool_forward_argmax(ft, rois, 7, 2)
    assert torch.equal(pooled, det.roi_pool_forward(ft, rois, 7, 2))  # the two forward variants agree bit for bit
    g_codes = det.roi_pool_backward_argmax(tuple(ft.shape), rois, dy, codes, 7, 2)
    g_walk = det.roi_pool_backward(ft, rois, dy, 7, 2)
    scale = float(g_walk.abs().max())
    assert float((g_codes - g_walk).abs().max()) <= 2e-5 * scale
    # proposals are clipped to the image, every sample lies inside the map: all of dY arrives somewhere (weights sum to 1)
    mass_in = dy.double().sum(dim=(1, 2, 3))
    mass_out = g_codes.double().sum(dim=(1, 2))
    assert float((mass_in - mass_out).abs().max()) <= 1e-4 * float(dy.double().abs().sum(dim=(1, 2, 3)).max())
    # linear in dY
    g2 = det.roi_pool_backward_argmax(tuple(ft.shape), rois, 2.0 * dy, codes, 7, 2)
    assert float((g2 - 2.0 * g_codes).abs().max()) <= 2e-5 * 2.0 * scale
    # the oracle on one image
    ref = orc.roi_pool_backward(f[5:6], host(rois)[5:6], host(dy)[5:6], 7, 2)[0]
    np.testing.assert_allclose(host(g_codes[5]), ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())


def test_engine_step_is_cuda_graph_capturable(det, orc, syn):
    """The serial engine step (decode + NMS + finalize + row-streaming RoI pool: TMA descriptors are kernel parameters)
    can be captured once and replayed; replays give the bits of the eager run, also after the inputs change in place."""
    w, anchors, scores, deltas, _ = _decoded(orc, syn, "kitti_b8", 2, 31)
    f = syn.feature_maps(np.random.default_rng(4), 2, w.grid, 1024)
    eng = det.engine.ProposalRoIEngine(w.image_shape, 2, w.grid, 1024, w.num_anchors, 0.0, 0.7, 300, pre_nms_top_n=2000,
                                       overlap=False)
    a, s, d, ft = dev(anchors), dev(scores), dev(deltas), dev(f)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        eng.step(a, s, d, ft)  # warm-up outside the capture (function attributes, descriptor cache)
        eng.step(a, s, d, ft)  # twice: the engine alternates between two proposal buffers
    torch.cuda.synchronize()
    eager = [t.clone() for t in (eng.rois, eng.pooled)]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.step(a, s, d, ft)
        eng.step(a, s, d, ft)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(eng.rois, eager[0]) and torch.equal(eng.pooled, eager[1])
    ft.mul_(2.0)  # same buffers, new contents: the replay must see them
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(eng.pooled, eager[1] * 2.0)  # scaling by a power of two is exact through lerps and max


def test_sampled_losses_forward_and_gradients(det, rng):
    """utils/losses.py:4-43 on sampled rows: the kernel against the NumPy restatement (1e-5) and, for the gradients,
    against torch autograd through an independent composite of the same formulas."""
    from oracle import reference_np as ref
    B, S, nc = 4, 64, 7
    L = nc + 1
    t_lab = np.zeros((B, S, L), F)
    t_lab[np.arange(B)[:, None], np.arange(S)[None, :], rng.integers(0, L, (B, S))] = 1
    logits = rng.standard_normal((B, S, L)).astype(F) * 3
    p = np.exp(logits - logits.max(-1, keepdims=True))
    p = (p / p.sum(-1, keepdims=True)).astype(F)
    p[0, 0] = [1, 0, 0, 0, 0, 0, 0, 0]            # exact 0 / 1 probabilities: clipped at 1e-7, no gradient through the clip
    p[0, 1] *= 0.5                                 # a row that does not sum to 1: Keras renormalises
    t_box = np.zeros((B, S, nc, 4), F)
    fg = rng.random((B, S, nc)) < 0.1
    t_box[fg] = rng.standard_normal((int(fg.sum()), 4)).astype(F)
    p_box = (t_box + rng.standard_normal(t_box.shape).astype(F) * 1.5).astype(F)  # errors on both sides of delta = 1
    ps, pb = dev(p).requires_grad_(True), dev(p_box).requires_grad_(True)
    cls, reg = det.sampled_losses(dev(t_lab), ps, dev(t_box), pb)
    np.testing.assert_allclose(cls.item(), ref.classification_loss(t_lab, p), rtol=RTOL)
    np.testing.assert_allclose(reg.item(), ref.regression_loss(t_box, p_box), rtol=RTOL)
    (2.0 * cls + 0.5 * reg).backward()
    # independent composite in torch (float64 reductions)
    ps2, pb2 = dev(p).double().requires_grad_(True), dev(p_box).double().requires_grad_(True)
    q = ps2 / ps2.sum(-1, keepdim=True)
    c = torch.clamp(q, 1e-7, 1 - 1e-7)
    cls2 = -(dev(t_lab).double() * torch.log(c)).sum(-1).mean()
    keep = dev(t_box).sum(-1) != 0
    e = (pb2 - dev(t_box).double())[keep]
    h = torch.where(e.abs() <= 1, 0.5 * e * e, e.abs() - 0.5)
    reg2 = h.mean(-1).sum()
    (2.0 * cls2 + 0.5 * reg2).backward()
    np.testing.assert_allclose(cls.item(), cls2.item(), rtol=RTOL)
    np.testing.assert_allclose(host(ps.grad), host(ps2.grad.float()), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(host(pb.grad), host(pb2.grad.float()), rtol=1e-5, atol=1e-8)
    # the reference's class interface
    assert det.ClassificationLoss()(dev(t_lab), dev(p)).item() == cls.item()
    assert det.RegressionLoss()(dev(t_box), dev(p_box)).item() == reg.item()
    # no kept row: regression loss 0 (the reference's reduce_sum over an empty tensor)
    assert float(det.RegressionLoss()(torch.zeros(2, 8, 1, 4, device="cuda"), torch.ones(2, 8, 1, 4, device="cuda"))) == 0.0


def test_roi_stream_kernels_vs_walk_kernels_random_shapes(det):
    """Randomised cross-check of the two independent RoI implementations on shapes the oracle tests do not enumerate:
    C = 512 takes the row-streaming TMA kernels, the same tensor split into two 256-channel halves takes the bin-row walk
    kernels.  Forward and arg-max codes must agree bit for bit, the gradients to 1e-5 (different summation orders)."""
    g = torch.Generator(device="cuda").manual_seed(1234)
    rs = np.random.default_rng(99)
    for it in range(24):
        B, R = int(rs.integers(1, 4)), int(rs.integers(1, 40))
        Hf, Wf = int(rs.integers(1, 61)), int(rs.integers(1, 131))
        f = torch.randn((B, Hf, Wf, 512), device="cuda", generator=g)
        if it % 2:
            f = torch.relu(f)  # exact zeros: arg-max ties (T8)
        c = torch.rand((B, R, 2), device="cuda", generator=g)
        wh = torch.rand((B, R, 2), device="cuda", generator=g) ** 2 * (1.6 if it % 3 == 0 else 0.7)
        rois = torch.cat([c - wh / 2, c + wh / 2], -1)  # x1,y1,x2,y2; some reach outside [0,1]
        if it % 4 == 0:
            rois = rois.clamp(0.0, 1.0)
        if it % 5 == 0:
            rois[:, : max(1, R // 4)] = 0.0  # zero padding (T7)
        rois = rois.contiguous()
        halves = [f[..., :256].contiguous(), f[..., 256:].contiguous()]
        out = det.roi_pool_forward(f, rois, 7, 2)
        ref = torch.cat([det.roi_pool_forward(h, rois, 7, 2) for h in halves], -1)
        assert torch.equal(out, ref), (it, B, R, Hf, Wf)
        out_a, codes = det.roi_pool_forward_argmax(f, rois, 7, 2)
        assert torch.equal(out_a, out)
        codes_ref = torch.cat([det.roi_pool_forward_argmax(h, rois, 7, 2)[1] for h in halves], -1)
        assert torch.equal(codes, codes_ref), (it, B, R, Hf, Wf)
        dy = torch.randn(out.shape, device="cuda", generator=g)
        gs = det.roi_pool_backward_argmax(tuple(f.shape), rois, dy, codes, 7, 2)
        gw = torch.cat([det.roi_pool_backward(h, rois, dy[..., i * 256:(i + 1) * 256].contiguous(), 7, 2)
                        for i, h in enumerate(halves)], -1)
        tol = 2e-5 * max(float(gw.abs().max()), 1e-6)
        assert float((gs - gw).abs().max()) <= tol, (it, B, R, Hf, Wf)


def test_average_precision_matching_and_metric(det, rng):
    """utils/metrics.py:4-133: the batched matching kernel against the NumPy restatement (exact: IoU
    is bit-identical, arg-max takes the first maximum), then the AP / mAP numbers (1e-6: identical fp32 arithmetic)."""
    from oracle import reference_np as ref
    B, G, M = 4, 100, 300
    gt = np.zeros((B, G, 4), F)
    pred = np.zeros((B, M, 4), F)
    scores = np.zeros((B, M), F)
    for b in range(B):
        n = int(rng.integers(0 if b == 3 else 1, 20)) if b != 3 else 0   # image 3: no ground truth at all
        xy, wh = rng.uniform(0, 0.7, (n, 2)), rng.uniform(0.05, 0.3, (n, 2))
        gt[b, :n] = np.concatenate([xy, xy + wh], -1)
        k = int(rng.integers(50, 250))                                  # the rest stays zero padded (score 0)
        base = gt[b, rng.integers(0, max(n, 1), k)] if n else rng.uniform(0, 1, (k, 4))
        pred[b, :k] = (base + rng.normal(0, 0.03, (k, 4))).astype(F)
        pred[b, 5] = pred[b, 4]                                         # duplicate prediction: first maximum wins
        scores[b, :k] = np.sort(rng.uniform(0.01, 1, k))[::-1]
    ap = det.AveragePrecision(0.5)
    ap.update_state(dev(gt[:2]), dev(pred[:2]), dev(scores[:2]))
    ap.update_state(dev(gt[2:]), dev(pred[2:]), dev(scores[2:]))
    tps, total = [], 0
    for b in range(B):
        tp, n = ref.ap_true_positives(gt[b], pred[b], 0.5)
        tps.append(tp)
        total += n
    np.testing.assert_array_equal(host(torch.cat(ap._true_pos)), np.concatenate(tps))
    assert int(ap.true_count) == total and ap.pos_count == B * M
    np.testing.assert_allclose(float(ap.result()), ref.average_precision(np.concatenate(tps), scores.reshape(-1), total), rtol=1e-6)
    # mAP over 3 classes == mean of per-class APs fed with the masked tensors
    labels = np.zeros((B, G, 4), F)
    cls = rng.integers(0, 3, (B, G))
    valid = gt.sum(-1) != 0
    labels[np.arange(B)[:, None], np.arange(G)[None, :], 1 + cls] = 1
    labels[~valid] = 0
    pcls = rng.integers(0, 3, (B, M)).astype(np.int32)
    m = det.MeanAveragePrecision(3, 0.5)
    m.update_state(dev(gt), dev(labels), dev(pred), dev(scores), torch.as_tensor(pcls).cuda())
    want = []
    for c in range(3):
        tps, total = [], 0
        sc = np.where(pcls == c, scores, 0).astype(F)
        for b in range(B):
            tp, n = ref.ap_true_positives(np.where((labels[b, :, 1 + c] == 1)[:, None], gt[b], 0),
                                          np.where((pcls[b] == c)[:, None], pred[b], 0), 0.5)
            tps.append(tp)
            total += n
        want.append(ref.average_precision(np.concatenate(tps), sc.reshape(-1), total))
    np.testing.assert_allclose(float(m.result()), float(np.mean(want)), rtol=1e-5)


# ------------------------------------------------------------------ training boundary: get_training_samples + TrainStepEngine
def test_objectness_folded_into_matching_equals_two_step(det, orc, syn, rng):
    """rpn_detector.py:141 folded into the matching kernels == the explicit one-hot followed by generate_targets, and both
    == the oracle (padding rows become [1, 0] and stay in the gt list, T11); odd label sums (2 -> [0, 0]) included."""
    w = syn.WORKLOADS["kitti_b8"]
    anchors = orc.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
    regions = anchors[orc.valid_anchor_indices(anchors, w.image_shape)]
    gl, gb = syn.ground_truth(rng, 4)
    gl[0, 1, 3] = 1.0  # a row with two hot entries: objectness index 2 -> [0, 0] -> removed as padding
    obj = orc.rpn_objectness_labels(gl)
    tl, tb = det.generate_targets_batched(dev(gl), dev(gb), dev(regions), w.image_shape, [0.7, 1.0], [0.0, 0.3], objectness=True)
    tl2, tb2 = det.generate_targets_batched(det.rpn_objectness_labels(dev(gl)), dev(gb), dev(regions), w.image_shape,
                                            [0.7, 1.0], [0.0, 0.3])
    assert torch.equal(tl, tl2) and torch.equal(tb, tb2)
    rl, rb = orc.generate_targets_batched(obj, gb, regions, w.image_shape, [0.7, 1.0], [0.0, 0.3])
    np.testing.assert_array_equal(host(tl), rl)
    np.testing.assert_allclose(host(tb), rb, rtol=RTOL, atol=1e-6)


def test_get_training_samples_reference_signature(det, orc, syn, rng):
    """RPNDetector.get_training_samples / FastRCNNDetector.get_training_samples called the way models/faster_rcnn.py:77-84
    calls them (**detector_output, **config["sampling"]): shapes, the sampling rule, and every sampled row equal to the row
    the oracle's targets hold at the sampled index (the draw itself is random in the reference)."""
    cfg = syn.REFERENCE_CONFIG
    w = syn.WORKLOADS["kitti_b8_ref12"]
    B, R, nc = 3, 300, cfg["num_classes"]
    rpn = det.RPNDetector(w.image_shape, (None,) + tuple(w.grid) + (1024,), cfg["rpn"])
    scores, deltas = syn.rpn_head_outputs(rng, B, w.num_anchors)
    gl, gb = syn.ground_truth(rng, B)
    rpn_output = rpn(dev(scores), dev(deltas), training=True)
    smp = rpn.get_training_samples(dev(gl), dev(gb), **rpn_output, **cfg["rpn"]["sampling"], seed=5)
    S = cfg["rpn"]["sampling"]["num_samples"]
    assert {k: tuple(v.shape) for k, v in smp.items()} == {"target_labels": (B, S, 2), "pred_scores": (B, S, 2),
                                                          "target_boxes": (B, S, 1, 4), "pred_boxes": (B, S, 1, 4)}
    regions = host(rpn_output["regions"])
    rl, rb = orc.generate_targets_batched(orc.rpn_objectness_labels(gl), gb, regions, w.image_shape, [0.7, 1.0], [0.0, 0.3])
    sc_v, dl_v = host(rpn_output["pred_scores"]), host(rpn_output["pred_boxes"])
    tl, ps, tb, pb = (host(smp[k]) for k in ("target_labels", "pred_scores", "target_boxes", "pred_boxes"))
    for b in range(B):
        # every sampled (target, prediction) row is a row of the full tensors at ONE common index (the four gathers agree)
        rows = {tuple(np.concatenate([rl[b, i], sc_v[b, i], dl_v[b, i].ravel()]).tolist()) for i in range(regions.shape[0])}
        for s in range(S):
            assert tuple(np.concatenate([tl[b, s], ps[b, s], pb[b, s].ravel()]).tolist()) in rows
        n_fg = int((tl[b, :, 1] == 1).sum())
        assert n_fg <= S // 2 and (tl[b, :n_fg, 1] == 1).all() and (tl[b, n_fg:, 0] == 1).all()  # foregrounds first
        fg_rows = np.nonzero(tl[b, :, 1] == 1)[0]
        assert (np.abs(tb[b, fg_rows]).sum((-1, -2)) > 0).all() and (tb[b, n_fg:] == 0).all()
    # Fast R-CNN form: batched regions, 7 classes
    rcnn = det.FastRCNNDetector(w.image_shape, nc, cfg["rcnn"])
    rois = syn.random_rois(rng, B, R)
    rois[:, :20] = gb[:, :20]  # some RoIs on the ground truth: real foregrounds
    lg = rng.normal(0, 1, (B, R, nc + 1)).astype(F)
    ps7 = (np.exp(lg) / np.exp(lg).sum(-1, keepdims=True)).astype(F)
    dl7 = rng.normal(0, 0.1, (B, R, nc, 4)).astype(F)
    rcnn_output = {"regions": rcnn.regions(dev(rois)), "pred_scores": dev(ps7), "pred_boxes": dev(dl7)}
    np.testing.assert_array_equal(host(rcnn_output["regions"]), orc.to_absolute(rois, w.image_shape))
    smp = rcnn.get_training_samples(dev(gl), dev(gb), **rcnn_output, **cfg["rcnn"]["sampling"], seed=6)
    S = cfg["rcnn"]["sampling"]["num_samples"]
    assert tuple(smp["target_labels"].shape) == (B, S, nc + 1) and tuple(smp["pred_boxes"].shape) == (B, S, nc, 4)
    rl, rb = orc.generate_targets_batched(gl, gb, orc.to_absolute(rois, w.image_shape), w.image_shape, [0.5, 1.0], [0.0, 0.5])
    tl, ps, tb, pb = (host(smp[k]) for k in ("target_labels", "pred_scores", "target_boxes", "pred_boxes"))
    for b in range(B):
        for s in range(S):
            hit = np.nonzero((rl[b] == tl[b, s]).all(-1) & (ps7[b] == ps[b, s]).all(-1) & (dl7[b] == pb[b, s]).all((-1, -2)))[0]
            assert hit.size >= 1
            np.testing.assert_allclose(tb[b, s], rb[b, hit[0]], rtol=RTOL, atol=1e-6)
        n_fg = int((tl[b, :, 0] == 0).sum())
        assert 1 <= n_fg <= 16 and (tl[b, :n_fg, 0] == 0).all() and (tl[b, n_fg:, 0] == 1).all()
    pooled = rcnn.pool(dev(syn.feature_maps(rng, B, w.grid, 512)), dev(rois))
    assert tuple(pooled.shape) == (B, R, 49 * 512)


def test_sampling_quota_limits_are_rejected_not_hung(det, rng):
    """a foreground quota >= 4096 (the sort buffer) used to spin forever; the C ABI now refuses it, together with
    proportions outside [0, 1]; a quota just under the limit with more than 4096 foregrounds works (chunked selection)."""
    N = 9000
    tl = np.zeros((1, N, 2), F)
    tl[0, :6000, 1] = 1.0  # 6000 foregrounds
    tl[0, 6000:, 0] = 1.0
    with pytest.raises(ValueError):
        det.sample_and_gather_batched(dev(tl), 8192, 0.5, seed=1)      # quota 4096
    with pytest.raises(ValueError):
        det.sample_and_gather_batched(dev(tl), 64, 1.5, seed=1)
    out = det.sample_and_gather_batched(dev(tl), 8190, 0.5, seed=1)    # quota 4095 < 4096, 6000 candidates: two chunks
    idx = host(out["indices"])[0]
    assert len(set(idx[:4095].tolist())) == 4095 and (idx[:4095] < 6000).all() and (idx[4095:] >= 6000).all()
    # tf.math.round(num_samples * proportion) on the Python double, half to even: 10 * 0.35 = 3.5 -> 4 (float32 0.35 gave 3)
    tl2 = np.zeros((1, 40, 2), F)
    tl2[0, :20, 1] = 1.0
    tl2[0, 20:, 0] = 1.0
    out = det.sample_and_gather_batched(dev(tl2), 10, 0.35, seed=2)
    assert int((host(out["target_labels"])[0, :, 1] == 1).sum()) == 4


def test_train_step_engine_matches_composed_calls_and_is_graph_capturable(det, orc, syn, rng):
    """TrainStepEngine (pre-allocated, C-ABI calls only) == the same step composed from the public functions, and a CUDA
    graph of the step replays to the same bits (proposals, samples, losses, pooled features, arg-max codes, dFeature)."""
    cfg = syn.REFERENCE_CONFIG
    w = syn.WORKLOADS["kitti_b8"]
    B, C, nc = 2, 512, cfg["num_classes"]
    rpn_cfg = {"nms": dict(cfg["rpn"]["nms"]), "sampling": cfg["rpn"]["sampling"]}
    anchors = det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
    eng = det.engine.TrainStepEngine(w.image_shape, B, w.grid, C, anchors, nc, rpn_cfg, cfg["rcnn"])
    R = eng.R
    scores, deltas = syn.rpn_head_outputs(rng, B, w.num_anchors)
    gl, gb = syn.ground_truth(rng, B)
    f = syn.feature_maps(rng, B, w.grid, C)
    lg = rng.normal(0, 1, (B, R, nc + 1)).astype(F)
    ps7 = (np.exp(lg) / np.exp(lg).sum(-1, keepdims=True)).astype(F)
    dl7 = rng.normal(0, 0.1, (B, R, nc, 4)).astype(F)
    dy = rng.standard_normal((B, R, 7, 7, C)).astype(F)
    args = [dev(x) for x in (scores, deltas, f, gl, gb, ps7, dl7, dy)]
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        eng.step(*args, seed=11)
    torch.cuda.synchronize()
    # the same step from the public functions
    vidx = det.valid_anchor_indices(anchors, w.image_shape)
    va, sc_v, dl_v = det.remove_invalid_anchors_and_predictions(anchors, w.image_shape, args[0], args[1], vidx)
    out = det.postprocess_output(w.image_shape, va, sc_v, dl_v, **rpn_cfg["nms"])
    assert torch.equal(out["pred_boxes"], eng.rois) and torch.equal(out["num_valid_detections"], eng.num_valid)
    smp = det.get_training_samples(args[3], args[4], va, sc_v, dl_v, w.image_shape, **rpn_cfg["sampling"], objectness=True,
                                   seed=22, check_background=False)
    for k in ("target_labels", "pred_scores", "target_boxes", "pred_boxes"):
        assert torch.equal(smp[k], eng.rpn_samples[k]), k
    l_cls, l_reg = det.sampled_losses(smp["target_labels"], smp["pred_scores"], smp["target_boxes"], smp["pred_boxes"])
    assert l_cls.item() == eng.rpn_samples["losses"][0].item() and l_reg.item() == eng.rpn_samples["losses"][1].item()
    pooled, codes = det.roi_pool_forward_argmax(args[2], eng.rois, 7, 2)
    assert torch.equal(pooled, eng.pooled) and torch.equal(codes, eng.codes)
    smp7 = det.get_training_samples(args[3], args[4], det.to_absolute(eng.rois, w.image_shape), args[5], args[6], w.image_shape,
                                    **cfg["rcnn"]["sampling"], seed=23, check_background=False)
    for k in ("target_labels", "pred_scores", "target_boxes", "pred_boxes"):
        assert torch.equal(smp7[k], eng.rcnn_samples[k]), k
    g_ref = det.roi_pool_backward_argmax(tuple(args[2].shape), eng.rois, args[7], codes, 7, 2)
    # (TMA reduce-adds of different RoIs reach a pixel in a different order on every launch: equal to rounding only)
    assert float((g_ref - eng.d_feature).abs().max()) <= 2e-5 * float(g_ref.abs().max())
    ref_g = orc.roi_pool_backward(f, host(eng.rois), dy, 7, 2)
    np.testing.assert_allclose(host(eng.d_feature), ref_g, rtol=RTOL, atol=RTOL * np.abs(ref_g).max())
    eager = {k: v.clone() for k, v in (("rois", eng.rois), ("pooled", eng.pooled), ("codes", eng.codes), ("dfeat", eng.d_feature),
                                       ("rpn_losses", eng.rpn_samples["losses"]), ("rcnn_losses", eng.rcnn_samples["losses"]))}
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.step(*args, seed=11)
    eng.d_feature.zero_()
    eng.pooled.zero_()
    g.replay()
    torch.cuda.synchronize()
    for k, v in (("rois", eng.rois), ("pooled", eng.pooled), ("codes", eng.codes), ("rpn_losses", eng.rpn_samples["losses"]),
                 ("rcnn_losses", eng.rcnn_samples["losses"])):
        assert torch.equal(v, eager[k]), k
    # dFeature: TMA reduce-adds arrive in a different order on every run -> equal to rounding only
    assert float((eng.d_feature - eager["dfeat"]).abs().max()) <= 2e-5 * float(eager["dfeat"].abs().max())


# ------------------------------------------------------------------ round-2 API additions
def test_host_pipeline_submit_wait_two_batches_in_flight(det, orc, syn):
    """frcnn_pipeline_submit / _wait: two different batches in flight at once (double-buffered device and host buffers,
    C = 512: the persistent row-streaming kernels with one workspace per image) give exactly what separate runs give."""
    w, anchors, scores, deltas, _ = _decoded(orc, syn, "kitti_b8", 4, 12)
    f = syn.feature_maps(np.random.default_rng(8), 4, w.grid, 512)
    pipe = det.ProposalRoIPipeline(w.image_shape, 2, w.grid, 512, w.num_anchors, 0.0, 0.7, 300, pre_nms_top_n=2000)
    ref = []
    for h in range(2):
        out = pipe(anchors, scores[2 * h:2 * h + 2], deltas[2 * h:2 * h + 2], f[2 * h:2 * h + 2])
        ref.append([t.clone() for t in out])
    oref = orc.roi_pool_forward(f[:2], ref[0][0].numpy(), 7, 2)
    np.testing.assert_array_equal(ref[0][3].numpy(), oref)
    for rep in range(3):  # submit both, then wait in order; repeated so both device slots are reused
        t0 = pipe.submit(anchors, scores[:2], deltas[:2], f[:2])
        t1 = pipe.submit(anchors, scores[2:], deltas[2:], f[2:])
        o0 = [t.clone() for t in pipe.wait(t0)]
        o1 = [t.clone() for t in pipe.wait(t1)]
        for got, want in ((o0, ref[0]), (o1, ref[1])):
            for a, b in zip(got, want):
                assert torch.equal(a, b), rep
    pipe.close()


def test_roi_pool_sm_budget_and_workspace_contract(det, orc, syn, rng):
    """frcnn_roi_pool_set_max_sms: the persistent kernels give the same bits on 1, 7 or all SMs (dynamic item queue);
    the row-streaming shapes refuse a missing / short workspace instead of running without their records."""
    L = det._lib.load()
    f = syn.feature_maps(rng, 2, (24, 78), 512)
    rois = syn.random_rois(rng, 2, 40, num_valid=[37, 40])
    ref = orc.roi_pool_forward(f, rois, 7, 2)
    dy = rng.standard_normal(ref.shape).astype(F)
    gref = orc.roi_pool_backward(f, rois, dy, 7, 2)
    try:
        for sms in (1, 7, 0):
            assert L.frcnn_roi_pool_set_max_sms(sms) == 0
            out, codes = det.roi_pool_forward_argmax(dev(f), dev(rois), 7, 2)
            np.testing.assert_array_equal(host(out), ref)
            np.testing.assert_array_equal(host(det.roi_pool_forward(dev(f), dev(rois), 7, 2)), ref)
            g = host(det.roi_pool_backward_argmax(f.shape, dev(rois), dev(dy), codes, 7, 2))
            np.testing.assert_allclose(g, gref, rtol=RTOL, atol=RTOL * np.abs(gref).max())
    finally:
        L.frcnn_roi_pool_set_max_sms(0)
    need = L.frcnn_roi_pool_workspace_bytes(2, 40, 512, 7, 2)
    assert need >= 256 + 512 * 80 and L.frcnn_roi_pool_workspace_bytes(2, 40, 64, 7, 2) == 0  # C = 64: bin-row walk, no records
    ft, rt = dev(f), dev(rois)
    out = torch.empty((2, 40, 7, 7, 512), device="cuda")
    ws = torch.empty(need - 512, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    rc = L.frcnn_roi_pool_forward(ft.data_ptr(), 2, 24, 78, 512, rt.data_ptr(), 40, 7, 2, out.data_ptr(), ws.data_ptr(), need - 512, stream)
    assert rc == -2 and "workspace" in det._lib.last_error()  # FRCNN_EWORKSPACE
    rc = L.frcnn_roi_pool_forward(ft.data_ptr(), 2, 24, 78, 512, rt.data_ptr(), 40, 7, 2, out.data_ptr(), None, 0, stream)
    assert rc == -2


def test_roi_pool_prepared_calls_match_the_one_call_forms(det, orc, syn, rng):
    """frcnn_roi_pool_prepare + _forward_prepared / _backward_prepared (records built once, off the pooling stream; one
    forward and one backward per prepare) give the bits of frcnn_roi_pool_forward(_argmax) / the gradient of
    _backward_argmax; shapes without records are refused."""
    L = det._lib.load()
    B, R, C = 2, 60, 512
    f = syn.feature_maps(rng, B, (24, 78), C)
    rois = syn.random_rois(rng, B, R, num_valid=[51, 60])
    ref = orc.roi_pool_forward(f, rois, 7, 2)
    dy = rng.standard_normal(ref.shape).astype(F)
    gref = orc.roi_pool_backward(f, rois, dy, 7, 2)
    ft, rt, dyt = dev(f), dev(rois), dev(dy)
    need = L.frcnn_roi_pool_workspace_bytes(B, R, C, 7, 2)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    out = torch.empty((B, R, 7, 7, C), device="cuda")
    codes = torch.empty((B, R, 7, 7, C // 4), dtype=torch.uint8, device="cuda")
    dfeat = torch.full((B, 24, 78, C), 7.0, device="cuda")  # (overwritten: zero-initialised by the call)
    s = torch.cuda.current_stream().cuda_stream
    side = torch.cuda.Stream()
    for with_codes in (False, True):
        with torch.cuda.stream(side):  # the records come from another stream, as in the pipelined engine
            assert L.frcnn_roi_pool_prepare(rt.data_ptr(), B, R, 24, 78, C, 7, 2, ws.data_ptr(), need, side.cuda_stream) == 0
        torch.cuda.current_stream().wait_stream(side)
        out.zero_()
        rc = L.frcnn_roi_pool_forward_prepared(ft.data_ptr(), B, 24, 78, C, R, 7, 2, out.data_ptr(),
                                               codes.data_ptr() if with_codes else None, ws.data_ptr(), need, s)
        assert rc == 0, det._lib.last_error()
        np.testing.assert_array_equal(host(out), ref)
    rc = L.frcnn_roi_pool_backward_prepared(B, 24, 78, C, R, 7, 2, dyt.data_ptr(), codes.data_ptr(), dfeat.data_ptr(),
                                            ws.data_ptr(), need, s)
    assert rc == 0, det._lib.last_error()
    np.testing.assert_allclose(host(dfeat), gref, rtol=RTOL, atol=RTOL * np.abs(gref).max())
    dfeat.fill_(3.0)  # a second backward on the same records (the call re-arms its work counter)
    assert L.frcnn_roi_pool_backward_prepared(B, 24, 78, C, R, 7, 2, dyt.data_ptr(), codes.data_ptr(), dfeat.data_ptr(),
                                              ws.data_ptr(), need, s) == 0
    np.testing.assert_allclose(host(dfeat), gref, rtol=RTOL, atol=RTOL * np.abs(gref).max())
    g2 = host(det.roi_pool_backward_argmax(f.shape, rt, dyt, codes, 7, 2))
    np.testing.assert_allclose(host(dfeat), g2, rtol=RTOL, atol=RTOL * np.abs(gref).max())
    assert L.frcnn_roi_pool_prepare(rt.data_ptr(), B, R, 24, 78, 64, 7, 2, ws.data_ptr(), need, s) == -1  # C = 64: no records
    assert L.frcnn_roi_pool_prepare(rt.data_ptr(), B, R, 24, 78, C, 7, 2, ws.data_ptr(), need - 512, s) == -2


@pytest.mark.parametrize("pool_streams", [1, 2])
def test_pipelined_engine_matches_the_serial_engine(det, orc, syn, pool_streams):
    """ProposalRoIEngine with overlap (proposals of step i + 1 beside the pool of step i, records built with the proposals,
    two pooling streams and output buffers so that consecutive pooling kernels may overlap at their ends): the buffers of the
    last two steps hold the bits of the serial engine on the same inputs -- eagerly and replayed from one CUDA graph."""
    w, anchors, _, _, _ = _decoded(orc, syn, "kitti_b8", 2, 31)
    a = dev(anchors)
    sets = []
    for k in range(3):
        rng = np.random.default_rng(100 + k)
        sc, dl = syn.rpn_head_outputs(rng, 2, w.num_anchors)
        sets.append((dev(sc), dev(dl), dev(syn.feature_maps(rng, 2, w.grid, 1024))))
    mk = lambda overlap: det.engine.ProposalRoIEngine(w.image_shape, 2, w.grid, 1024, w.num_anchors, 0.0, 0.7, 300,  # noqa: E731
                                                      pre_nms_top_n=2000, overlap=overlap, pool_streams=pool_streams)
    serial = mk(False)
    want = []
    for k in range(3):
        rois, _, nv, pooled = serial.step(a, *sets[k])
        torch.cuda.synchronize()
        want.append((rois.clone(), nv.clone(), pooled.clone()))
    eng = mk(True)
    assert eng.n_pool_streams == pool_streams
    n = 7

    def run():
        outs = [eng.step(a, *sets[i % 3]) for i in range(n)]
        eng.drain()
        return outs

    eng.attach()
    outs = run()
    torch.cuda.synchronize()
    for i in (n - 2, n - 1):  # the two buffers still hold steps n - 2 and n - 1
        rois, _, nv, pooled = outs[i]
        assert torch.equal(rois, want[i % 3][0]) and torch.equal(nv, want[i % 3][1])
        if pool_streams == 2 or i == n - 1:  # (one pooling stream: one output buffer, it holds the last step)
            assert torch.equal(pooled, want[i % 3][2])
    assert eng.pooled.data_ptr() == outs[n - 1][3].data_ptr() and eng.rois.data_ptr() == outs[n - 1][0].data_ptr()
    eng.reset_pipeline()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g):
            eng.attach()
            outs = run()
    for t in outs[n - 1][2:] + outs[n - 2][2:]:
        t.zero_()
    g.replay()
    torch.cuda.synchronize()
    for i in (n - 2, n - 1):
        rois, _, nv, pooled = outs[i]
        assert torch.equal(rois, want[i % 3][0]) and torch.equal(nv, want[i % 3][1])
        if pool_streams == 2 or i == n - 1:  # (one pooling stream: one output buffer, it holds the last step)
            assert torch.equal(pooled, want[i % 3][2])
