"""Golden fixtures (tests/golden/hotpath_small.npz, produced by the independent NumPy restatement -- see
tests/golden/make_golden.py): the C oracle must reproduce them on the CPU, the CUDA path on the GPU."""
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_small.npz"))
F = np.float32


def _check_all(impl, to_np, exact_pool_bwd=False):
    image_shape = tuple(int(v) for v in G["rpn_image_shape"])
    anchors = to_np(impl.generate_anchors(tuple(int(v) for v in G["rpn_grid"]), [0.25, 0.5], [0.5, 1.0, 2.0], [128, 128]))
    np.testing.assert_array_equal(anchors, G["rpn_anchors"])
    np.testing.assert_array_equal(to_np(impl.clip_to_window(G["rpn_anchors"], [0, 0, image_shape[1], image_shape[0]])), G["rpn_regions"])
    out = impl.postprocess_output(image_shape, G["rpn_regions"], G["rpn_scores"], G["rpn_deltas"], 0.0, 0.7, 25, 25, return_extra=True)
    np.testing.assert_array_equal(to_np(out["kept_indices"]), G["rpn_out_idx"])
    np.testing.assert_array_equal(to_np(out["num_valid_detections"]), G["rpn_out_valid"])
    np.testing.assert_array_equal(to_np(out["pred_scores"]), G["rpn_out_scores"])
    np.testing.assert_allclose(to_np(out["pred_boxes"]), G["rpn_out_boxes"], rtol=1e-5, atol=1e-6)
    out = impl.postprocess_output(image_shape, G["rcnn_regions"], G["rcnn_scores"], G["rcnn_deltas"], 0.05, 0.6, 6, 14, return_extra=True)
    np.testing.assert_array_equal(to_np(out["kept_indices"]), G["rcnn_out_idx"])
    np.testing.assert_array_equal(to_np(out["pred_classes"]), G["rcnn_out_classes"])
    np.testing.assert_array_equal(to_np(out["num_valid_detections"]), G["rcnn_out_valid"])
    np.testing.assert_array_equal(to_np(out["pred_scores"]), G["rcnn_out_scores"])
    np.testing.assert_allclose(to_np(out["pred_boxes"]), G["rcnn_out_boxes"], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(to_np(impl.roi_pool_forward(G["pool_feat"], G["pool_rois"], 7, 2)), G["pool_out"])
    np.testing.assert_allclose(to_np(impl.roi_pool_backward(G["pool_feat"], G["pool_rois"], G["pool_dy"], 7, 2)), G["pool_dfeat"],
                               rtol=1e-5, atol=1e-5 * np.abs(G["pool_dfeat"]).max())
    tl, tb = impl.generate_targets(G["tgt_gt_labels"], G["tgt_gt_boxes"], G["tgt_regions"], image_shape, [0.5, 1.0], [0.0, 0.5])
    np.testing.assert_array_equal(to_np(tl), G["tgt_labels"])
    np.testing.assert_allclose(to_np(tb), G["tgt_boxes"], rtol=1e-5, atol=1e-6)
    tl, tb = impl.generate_targets(G["tgt_obj"], G["tgt_gt_boxes"], G["tgt_regions"], image_shape, [0.7, 1.0], [0.0, 0.3])
    np.testing.assert_array_equal(to_np(tl), G["tgt_rpn_labels"])
    np.testing.assert_allclose(to_np(tb), G["tgt_rpn_boxes"], rtol=1e-5, atol=1e-6)


def test_c_oracle_reproduces_golden(orc):
    _check_all(orc, np.asarray)


@pytest.mark.gpu
def test_cuda_path_reproduces_golden(det):
    _check_all(det, lambda t: t.detach().cpu().numpy())
