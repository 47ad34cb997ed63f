"""Pre-allocated device-resident forward engine: postprocess_output(rpn) -> ROIPooling with no per-step allocation,
so one step is exactly two C-ABI calls (four kernel launches) on the current torch stream.  Used by bench.py for
the device-resident throughput number and by anyone who wants the path CUDA-graph capturable."""
from __future__ import annotations

import torch

from ._tensor import as_device, check, lib, ptr, require_cuda, stream


class ProposalRoIEngine:
    """Software pipeline over steps: proposals (decode + NMS with its outputs + the RoI records; latency-bound, a handful
    of SMs) of step i+1 run on one stream while RoI pooling (bandwidth/ALU-bound, the rest of the chip) of step i runs on
    another -- on two alternating ones from four images per batch, so that consecutive pooling kernels overlap at their
    ends.  Proposal buffers, records and pooled outputs are double-buffered and guarded by events, so results are
    identical to the serial order.  ``overlap=False`` issues everything on the current stream (latency measurement)."""

    KERNELS_PER_STEP = 4  # decode_filter, nms_select (writes the padded single-class outputs itself), roi_records, roi_pool_fwd

    def __init__(self, image_shape, batch, grid_shape, channels, num_anchors, score_threshold, iou_threshold,
                 max_total_size, pooled_size=7, kernel_size=2, pre_nms_top_n=None, overlap=True, reserve_sms=None,
                 pool_streams=None):
        dev = require_cuda()
        self.image_shape = tuple(image_shape)
        self.B, (self.Hf, self.Wf), self.C, self.N = int(batch), grid_shape, int(channels), int(num_anchors)
        self.R, self.P, self.k = int(max_total_size), int(pooled_size), int(kernel_size)
        self.score_threshold, self.iou_threshold = float(score_threshold), float(iou_threshold)
        self.pre_nms_top_n = int(pre_nms_top_n or 0)
        self.overlap = bool(overlap)
        self._lib = lib()
        # The persistent RoI kernel takes one CTA per SM; the proposal path of the next step (one 1024-thread NMS CTA per
        # image) needs whole SMs to run beside it, so the pipelined engine keeps `reserve_sms` SMs out of the pool's grid.
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        # Half an SM per image: the proposal path of a batch then runs as two waves of NMS CTAs, ~80 us for the KITTI grid
        # against ~170 us of pooling for the same 8 images, and the pool keeps four more SMs than with one SM per image
        # (same-box sweep, profiles/r2_reserve_sms_sweep.txt: 46.3 k images/s with 4 reserved SMs, 45.7 k with 8).
        if reserve_sms is None:
            reserve_sms = min((self.B + 1) // 2, sms // 8) if self.overlap else 0
        self.pool_sms = max(1, sms - int(reserve_sms)) if reserve_sms else 0  # 0 = every SM
        self.ws_bytes = self._lib.frcnn_postprocess_workspace_bytes(self.B, self.N, 1, self.R, self.R)
        self.ws = torch.empty(max(self.ws_bytes, 16), dtype=torch.uint8, device=dev)
        mk = lambda *s, dtype=torch.float32: [torch.empty(s, dtype=dtype, device=dev) for _ in range(2)]  # noqa: E731
        self._rois = mk(self.B, self.R, 4)
        self._roi_scores = mk(self.B, self.R)
        self._classes = mk(self.B, self.R, dtype=torch.int32)
        self._num_valid = mk(self.B, dtype=torch.int32)
        # Two pooling streams (and two output buffers), alternating with the proposal buffers: the pooling kernel of step
        # i + 1 is not ordered behind the one of step i, so its CTAs start on the SMs that the CTAs of step i leave one by
        # one -- the tail of a persistent kernel (its warps finish their last items up to one long item apart) and the ramp
        # of the next one overlap instead of adding up (KITTI batch 8: 46.6 k -> 47.9 k images/s on one box, dense batch 16
        # + 2 %).  Nothing else may start earlier: proposals of step i + 2 wait for pool(i) anyway.  Not for small batches:
        # there the proposal path is the critical one and two overlapping pooling grids take its reserved SM for a while
        # (batch 1: 22.8 k -> 20.9 k images/s), so one pooling stream below four images.
        if pool_streams is None:
            pool_streams = 2 if self.B >= 4 else 1
        self.n_pool_streams = 2 if (overlap and int(pool_streams) >= 2) else 1
        self._pooled = [torch.empty((self.B, self.R, self.P, self.P, self.C), dtype=torch.float32, device=dev)
                        for _ in range(self.n_pool_streams)]
        self.pool_ws_bytes = self._lib.frcnn_roi_pool_workspace_bytes(self.B, self.R, self.C, self.P, self.k)
        # one workspace per proposal buffer: the records of step i + 1 are built (with its proposals) while the pool of step i
        # still reads its own
        self._pool_ws = [torch.empty(max(self.pool_ws_bytes, 16), dtype=torch.uint8, device=dev) for _ in range(2)]
        self._split = self.pool_ws_bytes > 0  # row-streaming configuration: records pre-kernel separate from the pooling kernel
        self._i = 0
        self.s_prop = torch.cuda.Stream(device=dev, priority=-1)  # few large CTAs: let them in ahead of the pool grid
        self._s_pool = [torch.cuda.Stream(device=dev) for _ in range(self.n_pool_streams)]
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

    @property
    def pool_ws(self):
        return self._pool_ws[(self._i - 1) & 1]

    @property
    def pooled(self):
        return self._pooled[((self._i - 1) & 1) % self.n_pool_streams]

    @property
    def s_pool(self):
        return self._s_pool[0]

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
        if self._split:
            # the per-RoI records of the pooling kernel belong to the proposal path (they only need the rois): built here,
            # on the proposal stream, they never sit between two pooling kernels
            check(self._lib.frcnn_roi_pool_prepare(ptr(self._rois[j]), self.B, self.R, self.Hf, self.Wf, self.C, self.P, self.k,
                                                   ptr(self._pool_ws[j]), self.pool_ws_bytes, stream()), "roi_pool_prepare")

    def _pool(self, j, feature_maps):
        self._lib.frcnn_roi_pool_set_max_sms(self.pool_sms)
        if self._split:
            check(self._lib.frcnn_roi_pool_forward_prepared(ptr(feature_maps), self.B, self.Hf, self.Wf, self.C, self.R, self.P,
                                                            self.k, ptr(self._pooled[j % self.n_pool_streams]), None,
                                                            ptr(self._pool_ws[j]), self.pool_ws_bytes, stream()),
                  "roi_pool_forward_prepared")
        else:
            check(self._lib.frcnn_roi_pool_forward(ptr(feature_maps), self.B, self.Hf, self.Wf, self.C, ptr(self._rois[j]),
                                                   self.R, self.P, self.k, ptr(self._pooled[j % self.n_pool_streams]),
                                                   ptr(self._pool_ws[j]), self.pool_ws_bytes, stream()), "roi_pool_forward")

    def reset_pipeline(self):
        """Forget the events of earlier steps (call after ``drain()`` + a synchronisation, e.g. before capturing a block of
        steps in a CUDA graph: a captured stream must not wait on an event recorded outside the capture)."""
        self._i = 0
        self._pool_recorded = [False, False]

    def attach(self):
        """Make both pipeline streams wait for work already queued on the caller's current stream."""
        ev = torch.cuda.Event()
        ev.record()
        self.s_prop.wait_event(ev)
        for sp in self._s_pool:
            sp.wait_event(ev)

    def drain(self):
        """Make the caller's current stream wait for everything issued so far."""
        cur = torch.cuda.current_stream()
        for s in (self.s_prop, *self._s_pool):
            ev = torch.cuda.Event()
            ev.record(s)
            cur.wait_event(ev)

    def step(self, regions, pred_scores, pred_boxes, feature_maps, pool_events=None):
        """All arguments: fp32 contiguous CUDA tensors of the constructor's shapes.  Returns the buffers the results
        land in (rois, roi_scores, num_valid, pooled); with overlap they are complete once ``drain()``-ed.
        ``pool_events`` = (start, end) torch events recorded around the RoI-pool launch on its stream (either may be None)."""
        j = self._i & 1
        self._i += 1
        if not self.overlap:
            self._proposals(j, regions, pred_scores, pred_boxes)
            if pool_events and pool_events[0] is not None:
                pool_events[0].record()
            self._pool(j, feature_maps)
            if pool_events and pool_events[1] is not None:
                pool_events[1].record()
            return self._rois[j], self._roi_scores[j], self._num_valid[j], self._pooled[0]
        if self._pool_recorded[j]:  # proposals(i) may overwrite buffer j only after pool(i-2) has read it
            self.s_prop.wait_event(self._ev_pool[j])
        with torch.cuda.stream(self.s_prop):
            self._proposals(j, regions, pred_scores, pred_boxes)
            self._ev_prop[j].record(self.s_prop)
        sp = self._s_pool[j % self.n_pool_streams]
        sp.wait_event(self._ev_prop[j])
        with torch.cuda.stream(sp):
            if pool_events and pool_events[0] is not None:
                pool_events[0].record(sp)
            self._pool(j, feature_maps)
            if pool_events and pool_events[1] is not None:
                pool_events[1].record(sp)
            self._ev_pool[j].record(sp)
        self._pool_recorded[j] = True
        return self._rois[j], self._roi_scores[j], self._num_valid[j], self._pooled[j % self.n_pool_streams]


class TrainStepEngine:
    """Pre-allocated training-path step (BASELINE configs[4]; models/faster_rcnn.py:59-117 minus the conv / Dense layers):

        RPN training filter (rpn_detector.py:201-230)      frcnn_gather_rows x 2
        proposals (post_processing.py:6-63)                 frcnn_postprocess               (3 kernels)
        RPN samples (rpn_detector.py:98-160)                frcnn_generate_targets_objectness + frcnn_sample_and_gather
        RPN losses (faster_rcnn.py:87-93)                   frcnn_sampled_losses
        ROIPooling forward, arg-max codes saved (:56)       frcnn_roi_pool_forward_argmax
        regions = to_absolute(rois) (fast_rcnn_detector.py:66)   frcnn_scale_boxes
        RCNN samples (fast_rcnn_detector.py:71-130)         frcnn_generate_targets + frcnn_sample_and_gather
        RCNN losses (faster_rcnn.py:94-100)                 frcnn_sampled_losses
        ROIPooling backward (:103 through :56)              frcnn_roi_pool_backward_argmax

    Every output and workspace is allocated once here; ``step`` only enqueues C-ABI calls on the current stream (no torch
    op, no allocation, no synchronisation), so it can be captured in a CUDA graph.  The Dense heads between the pooled
    features and the RCNN predictions are the caller's: ``rcnn_scores`` / ``rcnn_boxes`` / ``d_pooled`` are its tensors
    (bench.py passes synthetic ones)."""

    KERNELS_PER_STEP = 17  # 2 gathers, 2 postprocess (decode_filter + nms_select; one class: no finalize launch), 2+1 RPN targets/samples, 1 loss, 2 pool fwd (records + pool), 1 scale, 2+1 RCNN, 1 loss, 1 pool bwd (the forward's records) (+ 3 memsets)

    def __init__(self, image_shape, batch, grid_shape, channels, anchors, num_classes, rpn_config, rcnn_config,
                 max_num_objects=100, pre_nms_top_n=None):
        from .models.detectors.rpn_detector import _gather_rows, valid_anchor_indices
        dev = require_cuda()
        self._lib = lib()
        self.image_shape = tuple(image_shape)
        self.B, (self.Hf, self.Wf), self.C = int(batch), grid_shape, int(channels)
        self.nc = int(num_classes)
        self.G = int(max_num_objects)
        self.rpn, self.rcnn = rpn_config, rcnn_config
        nms = rpn_config["nms"]
        self.R = int(nms["max_total_size"])
        self.per_class = int(nms["max_output_size_per_class"])
        self.P, self.k = int(rcnn_config["roi_pooling"]["pooled_size"]), int(rcnn_config["roi_pooling"]["kernel_size"])
        self.pre_nms_top_n = int(pre_nms_top_n or 0)
        a = as_device(anchors)
        self.N = a.shape[0]
        self.valid_idx = valid_anchor_indices(a, image_shape)      # static per configuration (one-off synchronisation)
        self.valid_anchors = _gather_rows(a[None], self.valid_idx)[0]
        self.Nv = int(self.valid_idx.numel())
        B, Nv, R, nc, C = self.B, self.Nv, self.R, self.nc, self.C
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)  # noqa: E731
        i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)    # noqa: E731
        self.S_rpn, self.S_rcnn = int(rpn_config["sampling"]["num_samples"]), int(rcnn_config["sampling"]["num_samples"])
        self.scores_v, self.deltas_v = f32(B, Nv, 2), f32(B, Nv, 1, 4)
        self.rois, self.roi_scores, self.roi_classes, self.num_valid = f32(B, R, 4), f32(B, R), i32(B, R), i32(B)
        self.rois_abs = f32(B, R, 4)
        self.rpn_tl, self.rpn_tb = f32(B, Nv, 2), f32(B, Nv, 1, 4)
        self.rcnn_tl, self.rcnn_tb = f32(B, R, nc + 1), f32(B, R, nc, 4)
        mk_s = lambda S, L: {"target_labels": f32(B, S, L), "pred_scores": f32(B, S, L), "target_boxes": f32(B, S, L - 1, 4),  # noqa: E731
                             "pred_boxes": f32(B, S, L - 1, 4), "indices": torch.empty((B, S), dtype=torch.int64, device=dev),
                             "counts": i32(B, 2), "losses": f32(3), "d_pred_scores": f32(B, S, L), "d_pred_boxes": f32(B, S, L - 1, 4)}
        self.rpn_samples, self.rcnn_samples = mk_s(self.S_rpn, 2), mk_s(self.S_rcnn, nc + 1)
        self.pooled = f32(B, R, self.P, self.P, C)
        self.codes = torch.empty((B, R, self.P, self.P, C // 4), dtype=torch.uint8, device=dev)
        self.d_feature = f32(B, self.Hf, self.Wf, C)
        L_ = self._lib
        ws = lambda n: (torch.empty(max(int(n), 16), dtype=torch.uint8, device=dev), int(n))  # noqa: E731
        self.ws_post = ws(L_.frcnn_postprocess_workspace_bytes(B, Nv, 1, self.per_class, self.R))
        self.ws_tgt_rpn = ws(L_.frcnn_generate_targets_workspace_bytes(B, Nv))
        self.ws_tgt_rcnn = ws(L_.frcnn_generate_targets_workspace_bytes(B, R))
        self.ws_smp_rpn = ws(L_.frcnn_sample_indices_workspace_bytes(B, Nv))
        self.ws_smp_rcnn = ws(L_.frcnn_sample_indices_workspace_bytes(B, R))
        self.ws_pool = ws(L_.frcnn_roi_pool_workspace_bytes(B, R, C, self.P, self.k))
        self._step = 0

    def algorithmic_bytes(self):
        B, R, C, HW = self.B, self.R, self.C, self.Hf * self.Wf
        bins = self.P * self.P * R
        return {
            "roi_pool_fwd_codes": B * (4 * C * (HW + bins) + bins * C // 4),
            "roi_pool_bwd": B * (4 * C * (bins + 2 * HW) + bins * C // 4),
            "match_rpn": 40 * self.Nv * B,
        }

    def _samples(self, tl, tb, N, L, cfg, pred_scores, pred_boxes, out, ws, seed):
        check(self._lib.frcnn_sample_and_gather(ptr(tl), self.B, N, L, int(cfg["num_samples"]), float(cfg["foreground_proportion"]),
                                                seed, ptr(out["indices"]), ptr(out["counts"]), ptr(pred_scores), ptr(tb),
                                                ptr(pred_boxes), ptr(out["target_labels"]), ptr(out["pred_scores"]),
                                                ptr(out["target_boxes"]), ptr(out["pred_boxes"]), ptr(ws[0]), ws[1], stream()),
              "sample_and_gather")
        rows = self.B * int(cfg["num_samples"])
        check(self._lib.frcnn_sampled_losses(ptr(out["target_labels"]), ptr(out["pred_scores"]), rows, L, ptr(out["target_boxes"]),
                                             ptr(out["pred_boxes"]), rows * (L - 1), 1.0, 1.0, ptr(out["losses"]),
                                             ptr(out["d_pred_scores"]), ptr(out["d_pred_boxes"]), stream()), "sampled_losses")

    def forward(self, pred_scores, pred_boxes, feature_maps, gt_labels, gt_boxes, rcnn_scores, rcnn_boxes, seed=None,
                between=None):
        """Everything up to and including the four losses.  pred_scores [B,N,2] / pred_boxes [B,N,1,4]: the RPN heads'
        outputs over ALL anchors; rcnn_scores [B,R,nc+1] / rcnn_boxes [B,R,nc,4]: the Dense heads' outputs."""
        L_, B, Nv, R, H, W = self._lib, self.B, self.Nv, self.R, float(self.image_shape[0]), float(self.image_shape[1])
        seed = self._step if seed is None else int(seed)
        self._step += 1
        n_idx = self.valid_idx.numel()
        check(L_.frcnn_gather_rows(ptr(pred_scores), B, self.N, 2, ptr(self.valid_idx), n_idx, ptr(self.scores_v), stream()), "gather_rows")
        check(L_.frcnn_gather_rows(ptr(pred_boxes), B, self.N, 4, ptr(self.valid_idx), n_idx, ptr(self.deltas_v), stream()), "gather_rows")
        nms = self.rpn["nms"]
        check(L_.frcnn_postprocess(H, W, ptr(self.valid_anchors), 0, ptr(self.scores_v), ptr(self.deltas_v), B, Nv, 1,
                                   float(nms["score_threshold"]), float(nms["iou_threshold"]), self.per_class, R,
                                   self.pre_nms_top_n, ptr(self.rois), ptr(self.roi_scores), ptr(self.roi_classes),
                                   ptr(self.num_valid), None, None, ptr(self.ws_post[0]), self.ws_post[1], stream()),
              "postprocess_output")
        sm = self.rpn["sampling"]
        (fl, fh), (bl, bh) = sm["foreground_iou_interval"], sm["background_iou_interval"]
        check(L_.frcnn_generate_targets_objectness(ptr(gt_labels), ptr(gt_boxes), B, self.G, self.nc + 1, ptr(self.valid_anchors),
                                                   0, Nv, H, W, float(fl), float(fh), float(bl), float(bh), ptr(self.rpn_tl),
                                                   ptr(self.rpn_tb), None, None, ptr(self.ws_tgt_rpn[0]), self.ws_tgt_rpn[1],
                                                   stream()), "generate_targets")
        self._samples(self.rpn_tl, self.rpn_tb, Nv, 2, sm, self.scores_v, self.deltas_v, self.rpn_samples, self.ws_smp_rpn,
                      2 * seed)
        if self.ws_pool[1] > 0:  # row-streaming configuration: one set of records serves the forward and the backward
            check(L_.frcnn_roi_pool_prepare(ptr(self.rois), B, R, self.Hf, self.Wf, self.C, self.P, self.k, ptr(self.ws_pool[0]),
                                            self.ws_pool[1], stream()), "roi_pool_prepare")
            check(L_.frcnn_roi_pool_forward_prepared(ptr(feature_maps), B, self.Hf, self.Wf, self.C, R, self.P, self.k,
                                                     ptr(self.pooled), ptr(self.codes), ptr(self.ws_pool[0]), self.ws_pool[1],
                                                     stream()), "roi_pool_forward_prepared")
        else:
            check(L_.frcnn_roi_pool_forward_argmax(ptr(feature_maps), B, self.Hf, self.Wf, self.C, ptr(self.rois), R, self.P,
                                                   self.k, ptr(self.pooled), ptr(self.codes), ptr(self.ws_pool[0]),
                                                   self.ws_pool[1], stream()), "roi_pool_forward_argmax")
        if between is not None:
            between()  # e.g. the caller's Dense heads on self.pooled producing rcnn_scores / rcnn_boxes
        check(L_.frcnn_scale_boxes(ptr(self.rois), B * R, H, W, 0, ptr(self.rois_abs), stream()), "to_absolute")
        sm = self.rcnn["sampling"]
        (fl, fh), (bl, bh) = sm["foreground_iou_interval"], sm["background_iou_interval"]
        check(L_.frcnn_generate_targets(ptr(gt_labels), ptr(gt_boxes), B, self.G, self.nc, ptr(self.rois_abs), 1, R, H, W,
                                        float(fl), float(fh), float(bl), float(bh), ptr(self.rcnn_tl), ptr(self.rcnn_tb), None,
                                        None, ptr(self.ws_tgt_rcnn[0]), self.ws_tgt_rcnn[1], stream()), "generate_targets")
        self._samples(self.rcnn_tl, self.rcnn_tb, R, self.nc + 1, sm, rcnn_scores, rcnn_boxes, self.rcnn_samples,
                      self.ws_smp_rcnn, 2 * seed + 1)
        return self.rpn_samples["losses"], self.rcnn_samples["losses"]

    def backward(self, d_pooled):
        """dL/d(pooled) [B,R,P,P,C] -> dL/d(feature_maps) from the saved arg-max codes."""
        if self.ws_pool[1] > 0:  # (the records and the backward's work counter were prepared by forward())
            check(self._lib.frcnn_roi_pool_backward_prepared(self.B, self.Hf, self.Wf, self.C, self.R, self.P, self.k, ptr(d_pooled),
                                                             ptr(self.codes), ptr(self.d_feature), ptr(self.ws_pool[0]),
                                                             self.ws_pool[1], stream()), "roi_pool_backward_prepared")
        else:
            check(self._lib.frcnn_roi_pool_backward_argmax(self.B, self.Hf, self.Wf, self.C, ptr(self.rois), self.R, self.P,
                                                           self.k, ptr(d_pooled), ptr(self.codes), ptr(self.d_feature),
                                                           ptr(self.ws_pool[0]), self.ws_pool[1], stream()),
                  "roi_pool_backward_argmax")
        return self.d_feature

    def step(self, pred_scores, pred_boxes, feature_maps, gt_labels, gt_boxes, rcnn_scores, rcnn_boxes, d_pooled, seed=None,
             between=None):
        self.forward(pred_scores, pred_boxes, feature_maps, gt_labels, gt_boxes, rcnn_scores, rcnn_boxes, seed, between)
        return self.backward(d_pooled)
