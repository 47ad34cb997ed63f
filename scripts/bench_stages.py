#!/usr/bin/env python
"""Per-stage device timings (CUDA events, device-resident inputs) for the configurations of BASELINE.json that are
not the headline bench line: dense-anchor stress forward, and the training path (target matching + RoI-pool
backward).  Prints one JSON object per stage; algorithmic bytes follow BASELINE.md section 3."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
det = importlib.import_module("2d_object_detection_b200")
syn = det.synthetic
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()  # write a buffer larger than L2 between timed iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def report(name, workload, ms, alg_bytes, extra=None):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    d = {"stage": name, "workload": workload, "ms": ms, "algorithmic_bytes": alg_bytes, "GB/s": gbs, "frac_of_measured_hbm": gbs / PEAK}
    d.update(extra or {})
    print(json.dumps(d), flush=True)


def main():
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rng = np.random.default_rng(5)
    for wname in ("kitti_b8", "dense_b16"):
        w = syn.WORKLOADS[wname]
        B, N, R, C = w.batch, w.num_anchors, w.max_total_size, w.channels
        Hf, Wf = w.grid
        anchors = det.clip_to_window(det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape), [0, 0, w.image_shape[1], w.image_shape[0]])
        scores, deltas = syn.rpn_head_outputs(rng, B, N)
        scores, deltas = torch.as_tensor(scores).to(dev), torch.as_tensor(deltas).to(dev)
        feat = torch.as_tensor(syn.feature_maps(rng, B, w.grid, C)).to(dev)
        kw = w.nms_kwargs()
        out = det.postprocess_output(w.image_shape, anchors, scores, deltas, **kw, pre_nms_top_n=w.pre_nms_top_n, return_extra=True)
        K = w.pre_nms_top_n or N
        ms = timeit(lambda: det.postprocess_output(w.image_shape, anchors, scores, deltas, **kw, pre_nms_top_n=w.pre_nms_top_n), flush=flush)
        report("postprocess_output (decode+NMS+finalize)", wname, ms, B * (44 * N + 4 * N + 40 * K + 16 * K + 16 * K * ((K + 63) // 64) + 24 * R),
               {"exhausted_images": int(out["exhausted"].sum()), "num_valid_min": int(out["num_valid_detections"].min())})
        ms = timeit(lambda: det.postprocess_output(w.image_shape, anchors, scores, deltas, **kw, pre_nms_top_n=None), flush=flush)
        report("postprocess_output, no pre-NMS cap (reference behaviour)", wname, ms, B * (44 * N + 4 * N + 40 * N + 24 * R))
        # head epilogue (rpn_detector.py:81-84): softmax fused into the decode kernel vs a separate softmax launch
        logits = torch.randn((B, N, 2), device=dev) * 1.5
        ms = timeit(lambda: det.postprocess_output(w.image_shape, anchors, logits, deltas, **kw, pre_nms_top_n=w.pre_nms_top_n,
                                                   from_logits=True, return_softmax=True), flush=flush)
        report("postprocess_output from logits (softmax fused, probabilities written)", wname, ms, B * (44 * N + 8 * N + 4 * N))
        ms = timeit(lambda: det.postprocess_output(w.image_shape, anchors, torch.softmax(logits, -1), deltas, **kw,
                                                   pre_nms_top_n=w.pre_nms_top_n), flush=flush)
        report("torch.softmax + postprocess_output (unfused)", wname, ms, B * (44 * N + 16 * N + 4 * N))
        del logits
        rois = out["pred_boxes"]
        ms = timeit(lambda: det.roi_pool_forward(feat, rois, 7, 2), flush=flush)
        report("roi_pool_forward", wname, ms, 4 * C * (Hf * Wf + 49 * R) * B)
        dy = torch.randn((B, R, 7, 7, C), device=dev)
        ms = timeit(lambda: det.roi_pool_backward(feat, rois, dy, 7, 2), iters=10, flush=flush)
        report("roi_pool_backward", wname, ms, 4 * C * (49 * R + 3 * Hf * Wf) * B)
        out_a, codes = det.roi_pool_forward_argmax(feat, rois, 7, 2)
        ms = timeit(lambda: det.roi_pool_forward_argmax(feat, rois, 7, 2), flush=flush)
        report("roi_pool_forward + argmax codes (training forward)", wname, ms, (4 * C * (Hf * Wf + 49 * R) + 49 * R * C // 4) * B)
        ms = timeit(lambda: det.roi_pool_backward_argmax(feat.shape, rois, dy, codes, 7, 2), iters=10, flush=flush)
        report("roi_pool_backward from saved argmax codes", wname, ms, (4 * C * (49 * R + 2 * Hf * Wf) + 49 * R * C // 4) * B)
        del dy, out_a, codes
        # training path: target matching on the valid anchors (RPN) and on the proposals (RCNN)
        a_all = det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
        valid = a_all[det.valid_anchor_indices(a_all, w.image_shape)].contiguous()
        gl, gb = syn.ground_truth(rng, B)
        obj = det.rpn_objectness_labels(torch.as_tensor(gl).to(dev))
        glt, gbt = torch.as_tensor(gl).to(dev), torch.as_tensor(gb).to(dev)
        ms = timeit(lambda: det.generate_targets_batched(obj, gbt, valid, w.image_shape, [0.7, 1.0], [0.0, 0.3]), flush=flush)
        report("generate_targets RPN (N_valid x 100)", wname, ms, 40 * valid.shape[0] * B, {"N_valid": int(valid.shape[0])})
        tl_rpn, tb_rpn = det.generate_targets_batched(obj, gbt, valid, w.image_shape, [0.7, 1.0], [0.0, 0.3])
        ms = timeit(lambda: det.sample_and_gather_batched(tl_rpn, 256, 0.5, target_boxes=tb_rpn, seed=1, check_background=False), flush=flush)
        report("sample_and_gather RPN (256 of N_valid, labels + box targets)", wname, ms, B * (8 * valid.shape[0] + 256 * (8 + 8 + 16 + 16)))
        # RCNN-stage postprocess (7 classes, per-class cap 100, cross-class merge to 300): SURVEY 8(f) rank 1
        lg = torch.randn((B, R, 8), device=dev) * 1.5
        sc7 = torch.softmax(lg, -1).contiguous()
        dl7 = (torch.randn((B, R, 7, 4), device=dev) * 0.1).contiguous()
        rois_abs = det.to_absolute(rois, w.image_shape)
        kw7 = dict(score_threshold=0.0, iou_threshold=0.6, max_output_size_per_class=100, max_total_size=300)
        ms = timeit(lambda: det.postprocess_output(w.image_shape, rois_abs, sc7, dl7, **kw7), flush=flush)
        report("postprocess_output RCNN form (7 classes)", wname, ms, B * R * 7 * 44)
        ms = timeit(lambda: det.generate_targets_batched(glt, gbt, rois_abs, w.image_shape, [0.5, 1.0], [0.0, 0.5]), flush=flush)
        report("generate_targets RCNN (R x G', 7 classes)", wname, ms, (16 + 4 * 8 + 16 * 7) * R * B)


if __name__ == "__main__":
    main()
