"""Pins for the CPU oracle: hand-computed known answers, every parity trap of
SURVEY 8.1, agreement between the C oracle and the independent NumPy restatement, and
cross-checks against torchvision.ops.nms / torch grid_sample / torch autograd.

The reference itself ships no tests or golden vectors and TensorFlow is not installable here,
so these are the pins ("parity unpinned" by the reference)."""
import numpy as np
import pytest
import torch

from oracle import reference_np as rnp

F = np.float32
CFG_ANCHORS = dict(scales=[0.25, 0.5, 1.0, 2.0], aspect_ratios=[0.5, 1.0, 2.0], base_anchor_shape=[256, 256])


# ------------------------------------------------------------------ anchors (row 1, T15)
def test_anchor_known_values(orc):
    a = orc.generate_anchors((38, 125), **CFG_ANCHORS)
    assert a.shape == (57000, 4)
    np.testing.assert_allclose(a[0], [-22.627417, -45.254833, 22.627417, 45.254833], rtol=1e-6)
    w, h = a[:12, 2] - a[:12, 0], a[:12, 3] - a[:12, 1]
    np.testing.assert_allclose(w[:9], [45.2548, 90.5097, 181.019, 362.039, 64, 128, 256, 512, 90.5097], rtol=1e-5)
    np.testing.assert_allclose(h[:5], [90.5097, 181.019, 362.039, 724.077, 64], rtol=1e-5)
    # location outer / shape inner, centres at multiples of the stride starting at 0 (no half-stride offset)
    ctr = (a[:, :2] + a[:, 2:]) / 2
    np.testing.assert_allclose(ctr[12], [16, 0], atol=1e-4)
    np.testing.assert_allclose(ctr[125 * 12], [0, 16], atol=1e-4)
    np.testing.assert_allclose(ctr[(3 * 125 + 7) * 12 + 5], [7 * 16, 3 * 16], atol=1e-4)


def test_anchor_counts_match_survey(orc, syn):
    for name, n, n_valid in [("default_b2", 57000, 30833), ("kitti_b8_ref12", 22464, 8768), ("kitti_b8", 16848, 8736)]:
        w = syn.WORKLOADS[name]
        a = orc.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
        assert a.shape[0] == n
        assert len(orc.valid_anchor_indices(a, w.image_shape)) == n_valid


def test_anchor_c_vs_numpy(orc):
    a = orc.generate_anchors((5, 7), **CFG_ANCHORS)
    b = rnp.generate_anchors((5, 7), **CFG_ANCHORS)
    np.testing.assert_array_equal(a, b)


# ------------------------------------------------------------------ boxes.py (rows 2,4,5,10; T10)
def test_clip_to_window_is_x_first(orc):
    boxes = np.array([[-5, -7, 2000, 700], [10, 20, 30, 40]], F)
    out = orc.clip_to_window(boxes, [0, 0, 1987, 600])
    np.testing.assert_array_equal(out, [[0, 0, 1987, 600], [10, 20, 30, 40]])
    np.testing.assert_array_equal(out, rnp.clip_to_window(boxes, [0, 0, 1987, 600]))


def test_decode_zero_delta_is_identity_and_roundtrip(orc, rng):
    ref = np.array([[10, 20, 110, 220], [0, 0, 64, 64]], F)
    np.testing.assert_array_equal(orc.decode(np.zeros((2, 4), F), ref), ref)
    # hand value: tx=.1 on w=100 -> cx=60+10; tw=log 2 -> w=200
    d = orc.decode(np.array([[0.1, 0.0, np.log(2.0), 0.0]], F), ref[:1])
    np.testing.assert_allclose(d, [[70 - 100, 20, 70 + 100, 220]], rtol=1e-6)
    b = np.sort(rng.uniform(0, 500, (100, 2, 2)).astype(F), axis=1).reshape(100, 4) + F([0, 0, 1, 1])
    r = np.sort(rng.uniform(0, 500, (100, 2, 2)).astype(F), axis=1).reshape(100, 4) + F([0, 0, 1, 1])
    np.testing.assert_allclose(orc.decode(orc.encode(b, r), r), b, rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(orc.encode(b, r), rnp.encode(b, r), rtol=1e-6, atol=1e-6)
    # glibc expf vs numpy exp differ by an ulp; c -/+ 0.5*s cancels, so compare with an absolute floor
    np.testing.assert_allclose(orc.decode(b / 500, r), rnp.decode(b / 500, r), rtol=1e-5, atol=1e-3)


def test_decode_broadcast(orc, rng):
    ref = rng.uniform(0, 100, (5, 1, 4)).astype(F)
    t = rng.normal(0, 0.1, (3, 5, 2, 4)).astype(F)
    np.testing.assert_allclose(orc.decode(t, ref), rnp.decode(t, ref), rtol=1e-5, atol=1e-4)


def test_relative_absolute(orc):
    b = np.array([[1987, 600, 993.5, 300]], F)
    np.testing.assert_array_equal(orc.to_relative(b, (600, 1987, 3)), [[1, 1, 0.5, 0.5]])
    np.testing.assert_array_equal(orc.to_absolute(orc.to_relative(b, (600, 1987, 3)), (600, 1987, 3)), b)
    x = np.array([[7, 7, 7, 7]], F)  # true division, not multiply-by-reciprocal
    assert orc.to_relative(x, (600, 1987, 3))[0, 0] == F(7) / F(1987)


# ------------------------------------------------------------------ metrics.iou (row 9)
def test_iou_known_and_zero_area_guard(orc):
    a = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [0, 0, 0, 0]], F)
    b = np.array([[1, 1, 11, 11], [20, 20, 30, 30], [0, 0, 0, 0]], F)
    got = orc.iou(a, b, pairwise=True)
    assert got[0, 0] == F(81) / F(119)
    assert got[0, 1] == 0 and got[2, 2] == 0 and got[0, 2] == 0  # 0/0 guarded (metrics.py:208)
    np.testing.assert_array_equal(got, rnp.iou(a, b, pairwise=True))
    np.testing.assert_array_equal(orc.iou(a, b), np.diagonal(got))
    np.testing.assert_array_equal(orc.area(a), [100, 100, 0])
    assert orc.intersection(a, b, pairwise=True)[0, 0] == 81


# ------------------------------------------------------------------ CombinedNMS (row 6, T1-T4, T17)
def test_nms_three_box_kat(orc):
    boxes = np.array([[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30]], F)
    scores = np.array([0.9, 0.8, 0.7], F)
    np.testing.assert_array_equal(orc.nms(boxes, scores, 0.0, 0.5, 10), [0, 2])  # IoU 81/119 = .6807 > .5
    np.testing.assert_array_equal(orc.nms(boxes, scores, 0.0, 0.7, 10), [0, 1, 2])
    np.testing.assert_array_equal(orc.nms(boxes, scores, 0.0, 0.5, 1), [0])
    np.testing.assert_array_equal(orc.nms(boxes, scores, 0.75, 0.7, 10), [0, 1])  # 0.7 below the score threshold
    np.testing.assert_array_equal(orc.nms(boxes, scores, 0.7, 0.7, 10), [0, 1])  # thresholds are fp32: .7f > .7f false
    np.testing.assert_array_equal(orc.nms(boxes, np.array([0.9, 0.8, 0.0], F), 0.0, 0.5, 10), [0])  # 0 > 0 false (T3)


def test_nms_iou_exactly_at_threshold_is_kept(orc):
    # IoU = 50/150 = 1/3 exactly representable? use 0.5: boxes [0,0,2,1] & [1,0,3,1]: inter 1, union 3 -> 1/3.
    # use boxes with IoU exactly 0.5: [0,0,2,2] & [0,0,2,1]... inter 2, union 4 -> .5
    boxes = np.array([[0, 0, 2, 2], [0, 0, 2, 1]], F)
    assert orc.tf_nms_iou(boxes[0], boxes[1]) == 0.5
    np.testing.assert_array_equal(orc.nms(boxes, np.array([0.9, 0.8], F), 0.0, 0.5, 10), [0, 1])  # strict >
    np.testing.assert_array_equal(orc.nms(boxes, np.array([0.9, 0.8], F), 0.0, 0.49, 10), [0])


def test_nms_tie_rule_lower_index_first(orc):
    boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [50, 50, 60, 60]], F)
    scores = np.array([0.5, 0.5, 0.5], F)
    np.testing.assert_array_equal(orc.nms(boxes, scores, 0.0, 0.5, 10), [0, 2])
    np.testing.assert_array_equal(rnp.nms(boxes, scores, 0.0, 0.5, 10), [0, 2])


def test_nms_iou_normalises_corners_and_degenerate_boxes(orc):
    a, a_flipped = np.array([0, 0, 10, 10], F), np.array([10, 10, 0, 0], F)
    b = np.array([1, 1, 11, 11], F)
    assert orc.tf_nms_iou(a, b) == orc.tf_nms_iou(a_flipped, b) == rnp.tf_nms_iou(a, b) == F(81) / F(119)
    assert orc.tf_nms_iou(np.array([5, 5, 5, 9], F), b) == 0.0  # zero area -> 0


def test_nms_vs_torchvision(orc, rng):
    import torchvision
    for seed in range(5):
        r = np.random.default_rng(seed)
        xy = r.uniform(0, 100, (400, 2))
        wh = r.uniform(5, 40, (400, 2))
        boxes = np.concatenate([xy, xy + wh], -1).astype(F)
        scores = r.permutation(400).astype(F) / 400 + F(0.001)
        keep_tv = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.5).numpy()
        np.testing.assert_array_equal(orc.nms(boxes, scores, 0.0, 0.5, 400), keep_tv)
        np.testing.assert_array_equal(orc.nms(boxes, scores, 0.0, 0.5, 37), keep_tv[:37])
        np.testing.assert_array_equal(rnp.nms(boxes[:150], scores[:150], 0.0, 0.5, 150),
                                      orc.nms(boxes[:150], scores[:150], 0.0, 0.5, 150))


def test_nms_kept_set_properties(orc, rng):
    xy = rng.uniform(0, 1, (600, 2))
    wh = rng.uniform(0.05, 0.3, (600, 2))
    boxes = np.concatenate([xy, xy + wh], -1).astype(F)
    scores = (rng.permutation(600).astype(F) + 1) / 601
    kept = orc.nms(boxes, scores, 0.0, 0.6, 50)
    assert len(kept) == 50 and len(set(kept.tolist())) == 50
    assert np.all(np.diff(scores[kept]) < 0)
    for a in range(len(kept)):
        for b in range(a):
            assert not orc.tf_nms_iou(boxes[kept[a]], boxes[kept[b]]) > F(0.6)
    # every candidate ranked before the last kept one, and dropped, overlaps a higher-scored kept box
    order = np.argsort(-scores, kind="stable")
    last = np.nonzero(order == kept[-1])[0][0]
    ks = set(kept.tolist())
    for i in order[:last]:
        if int(i) not in ks:
            assert any(orc.tf_nms_iou(boxes[i], boxes[j]) > F(0.6) for j in kept if scores[j] > scores[i])


def test_pre_nms_top_n_is_extension_identical_when_not_binding(orc, rng):
    xy = rng.uniform(0, 1, (500, 2))
    boxes = np.concatenate([xy, xy + 0.1], -1).astype(F)
    scores = (rng.permutation(500).astype(F) + 1) / 501
    full = orc.nms(boxes, scores, 0.0, 0.5, 20)
    np.testing.assert_array_equal(orc.nms(boxes, scores, 0.0, 0.5, 20, pre_nms_top_n=500), full)
    capped = orc.nms(boxes, scores, 0.0, 0.5, 400, pre_nms_top_n=30)
    assert set(capped.tolist()) <= set(np.argsort(-scores)[:30].tolist())


def test_postprocess_uses_unclipped_boxes_for_iou_and_clips_output(orc):
    """T2: two decoded boxes that differ only outside the image overlap < thr unclipped but coincide once clipped."""
    image_shape = (100, 100, 3)
    regions = np.array([[0, 0, 100, 100], [-300, 0, 100, 100]], F)  # second extends far left of the image
    scores = np.array([[[0.1, 0.9], [0.2, 0.8]]], F)
    deltas = np.zeros((1, 2, 1, 4), F)
    out = orc.postprocess_output(image_shape, regions, scores, deltas, 0.0, 0.7, 300, 300, return_extra=True)
    assert out["num_valid_detections"][0] == 2  # IoU unclipped = .25 -> both kept
    np.testing.assert_array_equal(out["pred_boxes"][0, :2], [[0, 0, 1, 1], [0, 0, 1, 1]])  # clipped to [0,1] on output
    np.testing.assert_array_equal(out["decoded_boxes"][0, 1, 0], [-3, 0, 1, 1])
    np.testing.assert_array_equal(out["pred_boxes"][0, 2:], 0)
    np.testing.assert_array_equal(out["pred_scores"][0, :3], np.array([0.9, 0.8, 0.0], F))
    assert out["pred_classes"].dtype == np.int32 and out["num_valid_detections"].dtype == np.int32  # T14
    np.testing.assert_array_equal(out["kept_indices"][0, :3], [0, 1, -1])


def test_postprocess_c_vs_numpy_rpn_and_rcnn_forms(orc, syn, rng):
    # RPN form: regions [N,4], nc = 1
    anchors = orc.generate_anchors((6, 9), scales=[0.25, 0.5], aspect_ratios=[0.5, 1.0, 2.0], base_anchor_shape=[256, 256])
    N = anchors.shape[0]
    scores, deltas = syn.rpn_head_outputs(rng, 2, N)
    kw = dict(score_threshold=0.0, iou_threshold=0.7, max_output_size_per_class=20, max_total_size=20)
    a = orc.postprocess_output((96, 144, 3), anchors, scores, deltas, **kw, return_extra=True)
    b = rnp.postprocess_output((96, 144, 3), anchors, scores, deltas, **kw)
    for key in ("pred_classes", "num_valid_detections", "kept_indices"):
        np.testing.assert_array_equal(a[key], b[key])
    np.testing.assert_allclose(a["pred_boxes"], b["pred_boxes"], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(a["pred_scores"], b["pred_scores"])
    # RCNN form: regions [B,R,4], nc = 3, per-class cap < total cap, cross-class merge (T17)
    B, R, nc = 2, 40, 3
    rois = syn.random_rois(rng, B, R) * F([144, 96, 144, 96])
    logits = rng.normal(0, 1, (B, R, nc + 1)).astype(F)
    sc = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
    dl = rng.normal(0, 0.1, (B, R, nc, 4)).astype(F)
    kw = dict(score_threshold=0.05, iou_threshold=0.6, max_output_size_per_class=5, max_total_size=12)
    a = orc.postprocess_output((96, 144, 3), rois, sc.astype(F), dl, **kw, return_extra=True)
    b = rnp.postprocess_output((96, 144, 3), rois, sc.astype(F), dl, **kw)
    for key in ("pred_classes", "num_valid_detections", "kept_indices", "pred_scores"):
        np.testing.assert_array_equal(a[key], b[key])
    np.testing.assert_allclose(a["pred_boxes"], b["pred_boxes"], rtol=1e-6, atol=1e-7)
    assert (a["num_valid_detections"] <= 12).all()
    for bi in range(B):
        nv = a["num_valid_detections"][bi]
        assert np.all(np.diff(a["pred_scores"][bi, :nv]) <= 0)
        assert np.bincount(a["pred_classes"][bi, :nv], minlength=nc).max() <= 5


# ------------------------------------------------------------------ crop_and_resize + max-pool (rows 7/8, T6-T8)
def test_roi_pool_ramp_full_image_box(orc):
    """4x4 ramp, full-image box, S=4 (P=4,k=1) reproduces the ramp exactly ((H-1) scaling, T6)."""
    ramp = np.arange(16, dtype=F).reshape(1, 4, 4, 1)
    out = orc.roi_pool_forward(ramp, np.array([[[0, 0, 1, 1]]], F), 4, 1)
    np.testing.assert_array_equal(out.reshape(4, 4), ramp.reshape(4, 4))
    # P=2,k=2: max of each 2x2 block
    out = orc.roi_pool_forward(ramp, np.array([[[0, 0, 1, 1]]], F), 2, 2)
    np.testing.assert_array_equal(out.reshape(2, 2), [[5, 7], [13, 15]])


def test_roi_pool_padded_roi_broadcasts_pixel_00(orc, rng):
    f = rng.standard_normal((2, 5, 6, 8)).astype(F)
    rois = np.zeros((2, 3, 4), F)
    out = orc.roi_pool_forward(f, rois, 7, 2)
    for b in range(2):
        np.testing.assert_array_equal(out[b], np.broadcast_to(f[b, 0, 0], (3, 7, 7, 8)))
    # backward: all of dY lands on pixel (0,0) (T7), argmax = first window element
    dy = rng.standard_normal(out.shape).astype(F)
    df = orc.roi_pool_backward(f, rois, dy, 7, 2)
    np.testing.assert_allclose(df[:, 0, 0], dy.sum(axis=(1, 2, 3)), rtol=1e-5, atol=1e-5)
    assert np.count_nonzero(df[:, 1:]) == 0 and np.count_nonzero(df[:, 0, 1:]) == 0


def test_roi_pool_out_of_range_rows_are_zero_filled(orc):
    f = np.ones((1, 4, 4, 1), F)
    # y2 = 1.5 -> last sample rows fall beyond H-1 -> extrapolation 0; with k=1 visible directly
    out = orc.roi_pool_forward(f, np.array([[[0, 0, 1, 1.5]]], F), 4, 1).reshape(4, 4)
    np.testing.assert_array_equal(out[:3], 1)  # in_y = 0,1.5,3.0
    np.testing.assert_array_equal(out[3], 0)  # in_y = 4.5 > 3
    # x1 < 0: first column outside
    out = orc.roi_pool_forward(f, np.array([[[-0.5, 0, 1, 1]]], F), 4, 1).reshape(4, 4)
    np.testing.assert_array_equal(out[:, 0], 0)
    np.testing.assert_array_equal(out[:, 1:], 1)


def test_roi_pool_y2_equal_one_last_row_rounding(orc):
    """T6: clipped proposals have y2 == 1.0 exactly, so the last sample sits at H-1 +/- rounding; TF zero-fills the row
    when fp32 rounding pushes it above H-1.  Pin the op order: in_y = y1*(H-1) + 13*((y2-y1)*(H-1)/13), each op rounded."""
    zero_filled = []
    for H in (24, 38, 50, 78, 84, 125):
        f = np.arange(H, dtype=F).reshape(1, H, 1, 1) + 1
        f = np.broadcast_to(f, (1, H, 3, 1)).copy()
        for y1 in (0.0, 0.1, 0.25, 1 / 3, 0.7):
            out = orc.roi_pool_forward(f, np.array([[[0, y1, 1, 1.0]]], F), 14, 1).reshape(14, 14)
            scale = F(F(F(1.0) - F(y1)) * F(H - 1)) / F(13)
            in_y = F(F(F(y1) * F(H - 1)) + F(F(13) * scale))
            if in_y > F(H - 1):
                assert out[13, 0] == 0, (H, y1)
                zero_filled.append((H, y1))
            else:
                lo = int(np.floor(in_y))
                want = F(f[0, lo, 0, 0]) + F(F(f[0, int(np.ceil(in_y)), 0, 0] - f[0, lo, 0, 0]) * F(in_y - F(lo)))
                assert out[13, 0] == want != 0, (H, y1)
            np.testing.assert_array_equal(out, rnp.roi_pool_forward(f, np.array([[[0, y1, 1, 1.0]]], F), 14, 1).reshape(14, 14))
    assert (125, 0.1) in zero_filled  # the trap is real: a naive (H-1)*y2 formulation would not zero-fill here


def test_roi_pool_constant_map_gives_constant(orc, rng):
    f = np.full((1, 9, 11, 4), 3.25, F)
    rois = np.clip(rng.uniform(0, 1, (1, 16, 4)).astype(F), 0, 1)
    rois = np.concatenate([np.minimum(rois[..., :2], rois[..., 2:]), np.maximum(rois[..., :2], rois[..., 2:])], -1)
    np.testing.assert_array_equal(orc.roi_pool_forward(f, rois, 7, 2), 3.25)


def test_roi_pool_c_vs_numpy_and_argmax_first_max(orc, syn, rng):
    f = syn.feature_maps(rng, 2, (6, 9), 16, relu=True)  # exact zeros -> ties (T8)
    rois = syn.random_rois(rng, 2, 5, num_valid=[4, 3])
    out, am = orc.roi_pool_forward(f, rois, 7, 2, return_argmax=True)
    np.testing.assert_array_equal(out, rnp.roi_pool_forward(f, rois, 7, 2))
    dy = rng.standard_normal(out.shape).astype(F)
    np.testing.assert_allclose(orc.roi_pool_backward(f, rois, dy, 7, 2), rnp.roi_pool_backward(f, rois, dy, 7, 2),
                               rtol=1e-5, atol=1e-5)
    # ties on exact zeros happen and resolve to the first window element
    crops = rnp.crop_and_resize(f, rois.reshape(-1, 4)[:, [1, 0, 3, 2]], np.repeat(np.arange(2), 5), (14, 14))
    win = crops.reshape(10, 7, 2, 7, 2, 16).transpose(0, 1, 3, 2, 4, 5).reshape(2, 5, 7, 7, 4, 16)
    np.testing.assert_array_equal(am, win.argmax(4))
    assert ((win == win.max(4, keepdims=True)).sum(4) > 1).any()


def test_roi_pool_vs_torch_grid_sample_and_autograd(orc, rng):
    """Independent check: bilinear align_corners=True sampling == crop_and_resize for in-range samples; torch
    autograd through grid_sample + max_pool2d == the backward oracle (tie-free inputs)."""
    B, H, W, Cc, R, P, k = 2, 7, 10, 5, 6, 7, 2
    S = P * k
    f = rng.standard_normal((B, H, W, Cc)).astype(F)
    lo = rng.uniform(0.0, 0.5, (B, R, 2))
    hi = lo + rng.uniform(0.1, 0.5, (B, R, 2))
    rois = np.concatenate([lo, hi], -1).astype(F)
    ft = torch.tensor(f, dtype=torch.float64, requires_grad=True)
    t = torch.linspace(0, 1, S, dtype=torch.float64)
    outs = []
    for b in range(B):
        for r in range(R):
            x1, y1, x2, y2 = [float(v) for v in rois[b, r]]
            gx = (x1 + (x2 - x1) * t) * 2 - 1
            gy = (y1 + (y2 - y1) * t) * 2 - 1
            grid = torch.stack(torch.meshgrid(gy, gx, indexing="ij")[::-1], -1)[None]
            crop = torch.nn.functional.grid_sample(ft[b].permute(2, 0, 1)[None], grid, mode="bilinear",
                                                   align_corners=True)
            outs.append(torch.nn.functional.max_pool2d(crop, k)[0].permute(1, 2, 0))
    ref = torch.stack(outs).reshape(B, R, P, P, Cc)
    got = orc.roi_pool_forward(f, rois, P, k)
    np.testing.assert_allclose(got, ref.detach().numpy(), rtol=1e-4, atol=1e-5)
    dy = rng.standard_normal(got.shape).astype(F)
    ref.backward(torch.tensor(dy, dtype=torch.float64))
    np.testing.assert_allclose(orc.roi_pool_backward(f, rois, dy, P, k), ft.grad.numpy(), rtol=1e-4, atol=1e-4)
    # conservation: in-range RoIs spread exactly dY's mass
    np.testing.assert_allclose(orc.roi_pool_backward(f, rois, dy, P, k).sum(), dy.sum(), rtol=1e-3)


# ------------------------------------------------------------------ generate_targets (row 11, T11-T13)
def _targets_fixture():
    image_shape = (100, 200, 3)
    regions = np.array([[0, 0, 50, 50],      # IoU 1.0 with gt0 -> NOT in [0.7, 1.0) but forced fg (first global max)
                        [0, 0, 50, 40],      # IoU .8 -> fg
                        [0, 0, 50, 20],      # IoU .4 -> ignored
                        [100, 50, 150, 100], # IoU 0 -> bg
                        [0, 0, 50, 50]], F)  # IoU 1.0 again -> not forced (only the first), not in interval -> ignored
    gt_boxes = np.zeros((4, 4), F)
    gt_boxes[0] = [0, 0, 0.25, 0.5]
    gt_labels = np.zeros((4, 3), F)
    gt_labels[0, 2] = 1
    return image_shape, regions, gt_labels, gt_boxes


def test_generate_targets_intervals_and_forced_foreground(orc):
    image_shape, regions, gl, gb = _targets_fixture()
    tl, tb, mi, am = orc.generate_targets(gl, gb, regions, image_shape, [0.7, 1.0], [0.0, 0.3], return_extra=True)
    np.testing.assert_allclose(mi, [1.0, 0.8, 0.4, 0.0, 1.0], rtol=1e-6)
    np.testing.assert_array_equal(tl, [[0, 0, 1], [0, 0, 1], [0, 0, 0], [1, 0, 0], [0, 0, 0]])  # T12
    np.testing.assert_array_equal(tb[0], [[0, 0, 0, 0], [0, 0, 0, 0]])  # encode(gt, identical region) = 0, class 1 slot
    np.testing.assert_allclose(tb[1, 1], [0, 5 / 40, 0, np.log(50 / 40)], rtol=1e-6)
    np.testing.assert_array_equal(tb[1, 0], 0)
    np.testing.assert_array_equal(tb[2:], 0)
    tl2, tb2 = rnp.generate_targets(gl, gb, regions, image_shape, [0.7, 1.0], [0.0, 0.3])
    np.testing.assert_array_equal(tl, tl2)
    np.testing.assert_allclose(tb, tb2, rtol=1e-6, atol=1e-7)


def test_generate_targets_rpn_padding_not_removed(orc):
    """T11: RPN objectness labels turn padded rows into [1,0] (sum 1) so they survive the filter as zero-area gts."""
    image_shape, regions, gl, gb = _targets_fixture()
    obj = orc.rpn_objectness_labels(gl)
    np.testing.assert_array_equal(obj, [[0, 1], [1, 0], [1, 0], [1, 0]])
    tl, tb, mi, am = orc.generate_targets(obj, gb, regions, image_shape, [0.7, 1.0], [0.0, 0.3], return_extra=True)
    np.testing.assert_array_equal(tl, [[0, 1], [0, 1], [0, 0], [1, 0], [0, 0]])
    np.testing.assert_array_equal(am, [0, 0, 0, 0, 0])
    tl2, tb2 = rnp.generate_targets(obj, gb, regions, image_shape, [0.7, 1.0], [0.0, 0.3])
    np.testing.assert_array_equal(tl, tl2)
    np.testing.assert_allclose(tb, tb2, rtol=1e-6, atol=1e-7)


def test_generate_targets_random_c_vs_numpy(orc, syn, rng):
    w = syn.WORKLOADS["kitti_b8"]
    anchors = orc.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
    regions = anchors[orc.valid_anchor_indices(anchors, w.image_shape)][::7]
    gl, gb = syn.ground_truth(rng, 3)
    for b in range(3):
        for labels, fg, bg in ((orc.rpn_objectness_labels(gl[b]), [0.7, 1.0], [0.0, 0.3]), (gl[b], [0.5, 1.0], [0.0, 0.5])):
            a = orc.generate_targets(labels, gb[b], regions, w.image_shape, fg, bg)
            c = rnp.generate_targets(labels, gb[b], regions, w.image_shape, fg, bg)
            np.testing.assert_array_equal(a[0], c[0])
            np.testing.assert_allclose(a[1], c[1], rtol=2e-6, atol=1e-6)
            assert (a[0].sum(-1) <= 1).all() and a[0][:, 1:].sum() >= 1  # at least one foreground (forced)


def test_generate_targets_no_gt_defined_as_all_background(orc):
    """T13 (undefined in the reference; documented choice): no gt -> all background, zero targets."""
    regions = np.array([[0, 0, 10, 10], [5, 5, 20, 20]], F)
    tl, tb = orc.generate_targets(np.zeros((3, 4), F), np.zeros((3, 4), F), regions, (50, 50, 3), [0.5, 1.0], [0.0, 0.5])
    np.testing.assert_array_equal(tl, [[1, 0, 0, 0], [1, 0, 0, 0]])
    np.testing.assert_array_equal(tb, 0)
    tl2, tb2 = rnp.generate_targets(np.zeros((3, 4), F), np.zeros((3, 4), F), regions, (50, 50, 3), [0.5, 1.0], [0.0, 0.5])
    np.testing.assert_array_equal(tl, tl2)


def test_synthetic_scores_unique(syn, rng):
    s, d = syn.rpn_head_outputs(rng, 2, 16848)
    assert s.shape == (2, 16848, 2) and d.shape == (2, 16848, 1, 4)
    for b in range(2):
        assert np.unique(s[b, :, 1]).size == 16848
    assert (s[..., 1] > 0).all() and (s[..., 1] < 1).all()


def test_losses_known_answers():
    """utils/losses.py:4-43, hand-computed: CCE of a one-hot target against [0.5, 0.5] is ln 2, Keras
    renormalises rows and clips at 1e-7; Huber(delta 1) is quadratic up to |e| = 1, averaged over the 4 coordinates and
    SUMMED over the rows with a non-zero target."""
    t = np.array([[[0, 1], [1, 0]]], np.float32)
    assert abs(rnp.classification_loss(t, np.array([[[0.5, 0.5], [0.5, 0.5]]], np.float32)) - np.log(2.0)) < 1e-6
    assert abs(rnp.classification_loss(t, np.array([[[1.0, 1.0], [3.0, 3.0]]], np.float32)) - np.log(2.0)) < 1e-6  # renormalised
    clipped = rnp.classification_loss(np.array([[[0, 1]]], np.float32), np.array([[[1.0, 0.0]]], np.float32))
    assert abs(clipped - (-np.log(np.float32(1e-7)))) < 1e-4                                          # -log(eps) = 16.118
    # rows 2 and 3 are filtered out: the reference keeps a row iff the SUM of its target is non-zero (utils/losses.py:36),
    # so a non-zero target that sums to zero is dropped too
    tb = np.array([[[[1, 1, 1, 1]], [[0, 0, 0, 0]], [[2, -2, 0, 0]], [[0, 0, 0, 3]]]], np.float32)
    pb = np.array([[[[1.5, 1, 3, 1]], [[9, 9, 9, 9]], [[2, -2, 0, 1]], [[0, 0, 0, 3.5]]]], np.float32)
    assert abs(rnp.regression_loss(tb, pb) - ((0.125 + 0 + 1.5 + 0) / 4 + 0.125 / 4)) < 1e-6         # 0.40625 + 0.03125
    assert rnp.regression_loss(np.zeros((1, 2, 1, 4), np.float32), np.ones((1, 2, 1, 4), np.float32)) == 0.0


def test_average_precision_known_answers():
    """utils/metrics.py:4-84, hand-computed: two gt boxes, three predictions.  gt 0 is best matched by prediction 1
    (IoU 0.81 > 0.5), gt 1 by prediction 2 (IoU 1/7 <= 0.5): TP = [0, 1, 0].  Sorted by score (0.9, 0.8, 0.1):
    TP order [0, 1, 0] -> precision 0, 1/2, 1/3; recall 0, 1/2, 1/2.  11-point interpolation: recall levels 0..0.5 see
    max precision 1/2, levels 0.6..1.0 only the sentinel 0 -> AP = 6 * 0.5 / 11."""
    gt = np.array([[0, 0, 10, 10], [20, 20, 30, 30], [0, 0, 0, 0]], np.float32)      # last row: padding
    pred = np.array([[50, 50, 60, 60], [1, 1, 10, 10], [25, 25, 35, 35]], np.float32)
    tp, n = rnp.ap_true_positives(gt, pred, 0.5)
    assert tp.tolist() == [0, 1, 0] and n == 2
    ap = rnp.average_precision(tp, np.array([0.9, 0.8, 0.1], np.float32), n)
    assert abs(ap - 6 * 0.5 / 11) < 1e-6
    # a gt box that overlaps nothing selects prediction 0 (argmax of zeros) but IoU 0 is not > threshold
    tp2, _ = rnp.ap_true_positives(np.array([[100, 100, 110, 110]], np.float32), pred, 0.0)
    assert tp2.tolist() == [0, 0, 0]


def test_softmax_known_answers():
    """rpn_detector.py:84 Softmax over the (background, object) pair; stable for large logits."""
    from oracle import reference_np as ref
    np.testing.assert_array_equal(ref.softmax(np.zeros((3, 2), np.float32)), np.full((3, 2), 0.5, np.float32))
    np.testing.assert_allclose(ref.softmax(np.float32([[0.0, np.log(3.0)]])), [[0.25, 0.75]], rtol=1e-6)
    np.testing.assert_array_equal(ref.softmax(np.float32([[1000.0, 0.0], [-1000.0, 0.0]])), np.float32([[1, 0], [0, 1]]))
    x = np.random.default_rng(3).normal(0, 3, (50, 8)).astype(np.float32)
    p = ref.softmax(x)
    assert p.dtype == np.float32
    np.testing.assert_allclose(p.sum(-1), 1.0, atol=1e-6)
    e = np.exp(x.astype(np.float64))
    np.testing.assert_allclose(p, e / e.sum(-1, keepdims=True), rtol=1e-5)
    np.testing.assert_array_equal(p.argmax(-1), x.argmax(-1))
