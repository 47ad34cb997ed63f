"""Host-buffer front end (frcnn_pipeline_*): the inference slice of FasterRCNN.call
(models/faster_rcnn.py:52-55) taking and returning HOST arrays, copies included."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._tensor import check, lib, require_cuda


def _host_ptr(x):
    if isinstance(x, torch.Tensor):
        if x.is_cuda or not x.is_contiguous():
            raise ValueError("host pipeline takes contiguous CPU tensors (pinned recommended) or NumPy arrays")
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("host arrays must be C-contiguous")
        return x.ctypes.data
    raise TypeError(f"unsupported host buffer type {type(x)}")


class ProposalRoIPipeline:
    """postprocess_output(rpn_output) -> ROIPooling on host buffers.  Shapes are fixed at construction."""

    def __init__(self, image_shape, batch, grid_shape, channels, num_anchors, score_threshold, iou_threshold,
                 max_total_size, pooled_size=7, kernel_size=2, pre_nms_top_n=None, pin=True):
        require_cuda()
        self.image_shape = tuple(image_shape)
        self.B, (self.Hf, self.Wf), self.C, self.N = int(batch), grid_shape, int(channels), int(num_anchors)
        self.R, self.P, self.k = int(max_total_size), int(pooled_size), int(kernel_size)
        self.score_threshold, self.iou_threshold = float(score_threshold), float(iou_threshold)
        self.pre_nms_top_n = int(pre_nms_top_n or 0)
        self._h = C.c_void_p()
        self._seq = 0
        check(lib().frcnn_pipeline_create(C.byref(self._h), self.B, self.Hf, self.Wf, self.C, self.N, self.R, self.P,
                                          self.k), "pipeline_create")
        mk = (lambda *s, dtype=torch.float32: torch.empty(s, dtype=dtype, pin_memory=pin))
        # two sets of host output buffers: batch i + 1 may be submitted while batch i is still landing
        self._out = [(mk(self.B, self.R, 4), mk(self.B, self.R), mk(self.B, dtype=torch.int32),
                      mk(self.B, self.R, self.P, self.P, self.C)) for _ in range(2)]
        self._last = 0
        self._tickets = {}

    @property
    def h2d_bytes(self):
        return 4 * (4 * self.N + self.B * self.N * 6 + self.B * self.Hf * self.Wf * self.C)

    @property
    def d2h_bytes(self):
        return 4 * (self.B * self.R * 5 + self.B + self.B * self.R * self.P * self.P * self.C)

    # results of the most recently completed batch
    rois = property(lambda self: self._out[self._last][0])
    roi_scores = property(lambda self: self._out[self._last][1])
    num_valid = property(lambda self: self._out[self._last][2])
    pooled = property(lambda self: self._out[self._last][3])

    def submit(self, regions, pred_scores, pred_boxes, feature_maps, copy_pooled=True):
        """Enqueue one batch (copies + kernels) and return a ticket at once; at most two batches in flight.  The input
        arrays must stay untouched until ``wait(ticket)`` returns."""
        t = C.c_int32(-1)
        o = self._out[self._seq & 1]  # the library's device slot for this batch is ticket % 2 as well
        check(lib().frcnn_pipeline_submit(self._h, float(self.image_shape[0]), float(self.image_shape[1]),
                                          _host_ptr(regions), _host_ptr(pred_scores), _host_ptr(pred_boxes),
                                          _host_ptr(feature_maps), self.score_threshold, self.iou_threshold,
                                          self.pre_nms_top_n, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                          o[3].data_ptr() if copy_pooled else None, C.byref(t)), "pipeline_submit")
        self._tickets[t.value] = o
        self._seq = t.value + 1
        return t.value

    def wait(self, ticket):
        """Block until the batch behind ``ticket`` has landed; returns its host tensors
        (rois [B,R,4], roi_scores [B,R], num_valid [B], pooled [B,R,P,P,C])."""
        check(lib().frcnn_pipeline_wait(self._h, int(ticket)), "pipeline_wait")
        o = self._tickets.pop(int(ticket))
        self._last = int(ticket) & 1
        return o

    def __call__(self, regions, pred_scores, pred_boxes, feature_maps, copy_pooled=True):
        """regions [N,4], pred_scores [B,N,2], pred_boxes [B,N,1,4], feature_maps [B,Hf,Wf,C]: host fp32.
        Returns host tensors (rois [B,R,4], roi_scores [B,R], num_valid [B], pooled [B,R,P,P,C])."""
        return self.wait(self.submit(regions, pred_scores, pred_boxes, feature_maps, copy_pooled))

    def close(self):
        if self._h:
            lib().frcnn_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
