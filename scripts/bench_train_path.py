#!/usr/bin/env python
"""Training-path step of BASELINE.json configs[4], one process per GPU:

    postprocess_output (RPN proposals)  ->  generate_targets RPN (N_valid x 100) + RCNN (300 x G')
    ->  sample_and_gather (256 / 64 training samples per image, fused with the gathers) -> sampled losses (RPN)
    ->  ROIPooling forward with saved arg-max codes  ->  ROIPooling backward (dY -> dFeature)
    ->  ncclAllReduce(sum, fp32) of a synthetic weight-gradient buffer (34.2 MB backbone, 51.0 MB with the heads)

The data path shards per image with no collective (every rank runs its own batch of 16 images: weak scaling); the
all-reduce is the one exchange of the training configuration and is issued on a second stream so that it overlaps the
RoI-pool backward of the same step, as the backbone gradients of a real step would.  Prints one JSON line (rank 0).

    python scripts/bench_train_path.py [--steps K] [--warmup W] [--grad-mb 34.2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_train_path.py
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
det = importlib.import_module("2d_object_detection_b200")
syn = det.synthetic
sh = importlib.import_module("2d_object_detection_b200.sharding")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--grad-mb", type=float, default=34.2)
    args = ap.parse_args()
    rank, world, local_rank = sh.rank_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = syn.WORKLOADS["kitti_b8"]
    B, R, C = args.batch, w.max_total_size, w.channels
    rng = np.random.default_rng(4040 + rank)
    a_all = det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
    vidx = det.valid_anchor_indices(a_all, w.image_shape)          # training mode: anchors fully inside the image (T16)
    valid = a_all[vidx].contiguous()
    n_sets = 3                                                        # rotate inputs: > 126 MB L2 together with dY
    sets = []
    for _ in range(n_sets):
        scores, deltas = syn.rpn_head_outputs(rng, B, w.num_anchors)
        feat = syn.feature_maps(rng, B, w.grid, C)
        gl, gb = syn.ground_truth(rng, B)
        sc = torch.as_tensor(scores).to(dev)[:, vidx].contiguous()
        dl = torch.as_tensor(deltas).to(dev)[:, vidx].contiguous()
        sets.append((sc, dl, torch.as_tensor(feat).to(dev), torch.as_tensor(gl).to(dev), torch.as_tensor(gb).to(dev)))
    dy = torch.randn((B, R, 7, 7, C), device=dev)
    grad = torch.randn(int(args.grad_mb * 1e6 / 4), device=dev)
    comm = torch.cuda.Stream()
    kw = w.nms_kwargs()

    def step(i):
        sc, dl, feat, gl, gb = sets[i % n_sets]
        out = det.postprocess_output(w.image_shape, valid, sc, dl, **kw)                     # no pre-NMS cap (reference)
        obj = det.rpn_objectness_labels(gl)
        tl, tb = det.generate_targets_batched(obj, gb, valid, w.image_shape, [0.7, 1.0], [0.0, 0.3])  # RPN targets
        smp = det.sample_and_gather_batched(tl, 256, 0.5, sc, tb, dl, seed=i, check_background=False)  # RPN minibatch (no sync)
        det.sampled_losses(smp["target_labels"], smp["pred_scores"], smp["target_boxes"], smp["pred_boxes"])  # rpn_cls, rpn_reg
        rois = out["pred_boxes"]
        tl2, tb2 = det.generate_targets_batched(gl, gb, det.to_absolute(rois, w.image_shape), w.image_shape, [0.5, 1.0],
                                                [0.0, 0.5])
        det.sample_and_gather_batched(tl2, 64, 0.25, target_boxes=tb2, seed=i, check_background=False)  # RCNN minibatch
        pooled, codes = det.roi_pool_forward_argmax(feat, rois, 7, 2)
        if world > 1:                                                                        # weight gradients of the step
            comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm):
                dist.all_reduce(grad)
        det.roi_pool_backward_argmax(tuple(feat.shape), rois, dy, codes, 7, 2)
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        step(i)
    t1.record()
    barrier()
    ms = sh.max_over_ranks(t0.elapsed_time(t1), dev) / args.steps
    # the all-reduce alone, for reference
    ar_ms = None
    if world > 1:
        for _ in range(3):
            dist.all_reduce(grad)
        barrier()
        t0.record()
        for _ in range(10):
            dist.all_reduce(grad)
        t1.record()
        barrier()
        ar_ms = sh.max_over_ranks(t0.elapsed_time(t1), dev) / 10
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({
            "metric": "training-path images/s (proposals + target matching + sampling + RoI pool fwd/bwd + weight-gradient all-reduce)",
            "value": B * world / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "ms_per_step": ms, "per_gpu_batch": B,
            "scaling": "weak", "N_valid": int(valid.shape[0]), "allreduce_mb": args.grad_mb, "allreduce_alone_ms": ar_ms,
            "config": "kitti 375x1242 training mode (anchors inside the image, no pre-NMS cap), 100 padded gt boxes, "
                      "dY [B,300,7,7,1024]"}), flush=True)


if __name__ == "__main__":
    main()
