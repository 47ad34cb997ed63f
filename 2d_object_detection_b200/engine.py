"""Pre-allocated device-resident forward engine: postprocess_output(rpn) -> ROIPooling with no per-step allocation,
so one step is exactly two C-ABI calls (four kernel launches) on the current torch stream.  Used by bench.py for
the device-resident throughput number and by anyone who wants the path CUDA-graph capturable."""
from __future__ import annotations

import torch

from ._tensor import as_device, check, lib, ptr, require_cuda, stream


class ProposalRoIEngine:
    """Two-stream software pipeline over steps: proposals (decode + NMS + finalize; latency-bound, a handful of
    SMs) of step i+1 run on one stream while RoI pooling (bandwidth/ALU-bound, the whole chip) of step i runs on
    another.  The proposal buffers are double-buffered and guarded by events, so results are identical to the
    serial order.  ``overlap=False`` issues everything on the current stream (latency measurement)."""

    KERNELS_PER_STEP = 4  # decode_filter, nms_select, finalize, roi_pool_fwd

    def __init__(self, image_shape, batch, grid_shape, channels, num_anchors, score_threshold, iou_threshold,
                 max_total_size, pooled_size=7, kernel_size=2, pre_nms_top_n=None, overlap=True):
        dev = require_cuda()
        self.image_shape = tuple(image_shape)
        self.B, (self.Hf, self.Wf), self.C, self.N = int(batch), grid_shape, int(channels), int(num_anchors)
        self.R, self.P, self.k = int(max_total_size), int(pooled_size), int(kernel_size)
        self.score_threshold, self.iou_threshold = float(score_threshold), float(iou_threshold)
        self.pre_nms_top_n = int(pre_nms_top_n or 0)
        self.overlap = bool(overlap)
        self._lib = lib()
        self.ws_bytes = self._lib.frcnn_postprocess_workspace_bytes(self.B, self.N, 1, self.R, self.R)
        self.ws = torch.empty(max(self.ws_bytes, 16), dtype=torch.uint8, device=dev)
        mk = lambda *s, dtype=torch.float32: [torch.empty(s, dtype=dtype, device=dev) for _ in range(2)]  # noqa: E731
        self._rois = mk(self.B, self.R, 4)
        self._roi_scores = mk(self.B, self.R)
        self._classes = mk(self.B, self.R, dtype=torch.int32)
        self._num_valid = mk(self.B, dtype=torch.int32)
        self.pooled = torch.empty((self.B, self.R, self.P, self.P, self.C), dtype=torch.float32, device=dev)
        self._i = 0
        self.s_prop = torch.cuda.Stream(device=dev, priority=-1)  # few large CTAs: let them in ahead of the pool grid
        self.s_pool = torch.cuda.Stream(device=dev)
        self._ev_prop = [torch.cuda.Event(), torch.cuda.Event()]
        self._ev_pool = [torch.cuda.Event(), torch.cuda.Event()]
        self._pool_recorded = [False, False]

    # latest completed-in-order results (buffer of the last issued step)
    @property
    def rois(self):
        return self._rois[(self._i - 1) & 1]

    @property
    def roi_scores(self):
        return self._roi_scores[(self._i - 1) & 1]

    @property
    def num_valid(self):
        return self._num_valid[(self._i - 1) & 1]

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

    def _proposals(self, j, regions, pred_scores, pred_boxes):
        check(self._lib.frcnn_postprocess(float(self.image_shape[0]), float(self.image_shape[1]), ptr(regions), 0,
                                          ptr(pred_scores), ptr(pred_boxes), self.B, self.N, 1, self.score_threshold,
                                          self.iou_threshold, self.R, self.R, self.pre_nms_top_n, ptr(self._rois[j]),
                                          ptr(self._roi_scores[j]), ptr(self._classes[j]), ptr(self._num_valid[j]),
                                          None, None, ptr(self.ws), self.ws_bytes, stream()), "postprocess_output")

    def _pool(self, j, feature_maps):
        check(self._lib.frcnn_roi_pool_forward(ptr(feature_maps), self.B, self.Hf, self.Wf, self.C, ptr(self._rois[j]),
                                               self.R, self.P, self.k, ptr(self.pooled), stream()), "roi_pool_forward")

    def attach(self):
        """Make both pipeline streams wait for work already queued on the caller's current stream."""
        ev = torch.cuda.Event()
        ev.record()
        self.s_prop.wait_event(ev)
        self.s_pool.wait_event(ev)

    def drain(self):
        """Make the caller's current stream wait for everything issued so far."""
        cur = torch.cuda.current_stream()
        for s in (self.s_prop, self.s_pool):
            ev = torch.cuda.Event()
            ev.record(s)
            cur.wait_event(ev)

    def step(self, regions, pred_scores, pred_boxes, feature_maps, pool_events=None):
        """All arguments: fp32 contiguous CUDA tensors of the constructor's shapes.  Returns the buffers the results
        land in (rois, roi_scores, num_valid, pooled); with overlap they are complete once ``drain()``-ed.
        ``pool_events`` = (start, end) torch events recorded around the RoI-pool launch on its stream."""
        j = self._i & 1
        self._i += 1
        if not self.overlap:
            self._proposals(j, regions, pred_scores, pred_boxes)
            if pool_events:
                pool_events[0].record()
            self._pool(j, feature_maps)
            if pool_events:
                pool_events[1].record()
            return self._rois[j], self._roi_scores[j], self._num_valid[j], self.pooled
        if self._pool_recorded[j]:  # proposals(i) may overwrite buffer j only after pool(i-2) has read it
            self.s_prop.wait_event(self._ev_pool[j])
        with torch.cuda.stream(self.s_prop):
            self._proposals(j, regions, pred_scores, pred_boxes)
            self._ev_prop[j].record(self.s_prop)
        self.s_pool.wait_event(self._ev_prop[j])
        with torch.cuda.stream(self.s_pool):
            if pool_events:
                pool_events[0].record(self.s_pool)
            self._pool(j, feature_maps)
            if pool_events:
                pool_events[1].record(self.s_pool)
            self._ev_pool[j].record(self.s_pool)
        self._pool_recorded[j] = True
        return self._rois[j], self._roi_scores[j], self._num_valid[j], self.pooled
