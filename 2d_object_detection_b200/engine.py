"""Pre-allocated device-resident forward engine: postprocess_output(rpn) -> ROIPooling with no per-step allocation,
so one step is exactly two C-ABI calls (four kernel launches) on the current torch stream.  Used by bench.py for
the device-resident throughput number and by anyone who wants the path CUDA-graph capturable."""
from __future__ import annotations

import torch

from ._tensor import as_device, check, lib, ptr, require_cuda, stream


class ProposalRoIEngine:
    KERNELS_PER_STEP = 4  # decode_filter, nms_select, finalize, roi_pool_fwd

    def __init__(self, image_shape, batch, grid_shape, channels, num_anchors, score_threshold, iou_threshold,
                 max_total_size, pooled_size=7, kernel_size=2, pre_nms_top_n=None):
        dev = require_cuda()
        self.image_shape = tuple(image_shape)
        self.B, (self.Hf, self.Wf), self.C, self.N = int(batch), grid_shape, int(channels), int(num_anchors)
        self.R, self.P, self.k = int(max_total_size), int(pooled_size), int(kernel_size)
        self.score_threshold, self.iou_threshold = float(score_threshold), float(iou_threshold)
        self.pre_nms_top_n = int(pre_nms_top_n or 0)
        self._lib = lib()
        self.ws_bytes = self._lib.frcnn_postprocess_workspace_bytes(self.B, self.N, 1, self.R, self.R)
        self.ws = torch.empty(max(self.ws_bytes, 16), dtype=torch.uint8, device=dev)
        self.rois = torch.empty((self.B, self.R, 4), dtype=torch.float32, device=dev)
        self.roi_scores = torch.empty((self.B, self.R), dtype=torch.float32, device=dev)
        self.classes = torch.empty((self.B, self.R), dtype=torch.int32, device=dev)
        self.num_valid = torch.empty((self.B,), dtype=torch.int32, device=dev)
        self.pooled = torch.empty((self.B, self.R, self.P, self.P, self.C), dtype=torch.float32, device=dev)

    # algorithmic bytes per step of each stage (BASELINE.md section 3 / SURVEY 8(d)), for roofline reporting
    def algorithmic_bytes(self, pre_nms=None):
        N, R, B = self.N, self.R, self.B
        K = pre_nms or self.pre_nms_top_n or N
        return {
            "decode_filter": 44 * N * B,
            "topk_sort": (4 * N + 40 * K) * B,
            "nms_bitmask": (16 * K + 16 * K * ((K + 63) // 64) + 24 * R) * B,
            "roi_pool_fwd": 4 * self.C * (self.Hf * self.Wf + self.P * self.P * R) * B,
        }

    def proposals(self, regions, pred_scores, pred_boxes):
        check(self._lib.frcnn_postprocess(float(self.image_shape[0]), float(self.image_shape[1]), ptr(regions), 0,
                                          ptr(pred_scores), ptr(pred_boxes), self.B, self.N, 1, self.score_threshold,
                                          self.iou_threshold, self.R, self.R, self.pre_nms_top_n, ptr(self.rois),
                                          ptr(self.roi_scores), ptr(self.classes), ptr(self.num_valid), None, None,
                                          ptr(self.ws), self.ws_bytes, stream()), "postprocess_output")

    def pool(self, feature_maps):
        check(self._lib.frcnn_roi_pool_forward(ptr(feature_maps), self.B, self.Hf, self.Wf, self.C, ptr(self.rois),
                                               self.R, self.P, self.k, ptr(self.pooled), stream()), "roi_pool_forward")

    def step(self, regions, pred_scores, pred_boxes, feature_maps):
        """All arguments: fp32 contiguous CUDA tensors of the constructor's shapes.  Results land in
        self.rois / self.roi_scores / self.num_valid / self.pooled."""
        self.proposals(regions, pred_scores, pred_boxes)
        self.pool(feature_maps)
        return self.rois, self.roi_scores, self.num_valid, self.pooled
