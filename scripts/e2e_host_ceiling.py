#!/usr/bin/env python
"""Where does the end-to-end (host-buffer) number at N GPUs go?  Run under torchrun with one rank per GPU of ONE box:
for k = 1, 2, 4, ... active ranks (the others idle at a barrier) it times, per rank and in aggregate,
  d2h_1buf      plain cudaMemcpyAsync of the pooled-feature bytes (482 MB at KITTI batch 8) into ONE pinned buffer, back to back
  d2h_chunks    the same bytes as per-image chunks on two alternating streams into two alternating pinned buffers
  bidir         d2h_chunks with the step's input bytes (65 MB) going up on a third stream at the same time
  pipeline      frcnn_pipeline_submit / _wait, two batches in flight (what bench.py's e2e measures)
so the host-side ceiling and the pipeline can be compared at every k in one run.  Prints one JSON line per (k, mode) on rank 0."""
import importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
det = importlib.import_module("2d_object_detection_b200"); syn = det.synthetic
sh = importlib.import_module("2d_object_detection_b200.sharding")
import torch.distributed as dist
rank, world, local = sh.rank_world()
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
def barrier():
    if world > 1: dist.barrier(device_ids=[local])
    torch.cuda.synchronize()
SEC = float(sys.argv[1]) if len(sys.argv) > 1 else 0.8
w = syn.WORKLOADS["kitti_b8"]; B = w.batch
rng = np.random.default_rng(7 + rank)
anchors = det.clip_to_window(det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape), [0, 0, w.image_shape[1], w.image_shape[0]])
pin = []
for _ in range(2):
    sc, dl = syn.rpn_head_outputs(rng, B, w.num_anchors)
    ft = syn.feature_maps(rng, B, w.grid, w.channels)
    pin.append(tuple(torch.as_tensor(a).pin_memory() for a in (sc, dl, ft)))
anchors_h = anchors.cpu().pin_memory()
pipe = det.ProposalRoIPipeline(w.image_shape, B, w.grid, w.channels, w.num_anchors, w.score_threshold, w.iou_threshold,
                               w.max_total_size, w.pooled_size, w.kernel_size, w.pre_nms_top_n)
for i in range(3): pipe(anchors_h, *pin[i % 2])
n_out = pipe.d2h_bytes // 4
hb = [torch.empty(n_out, dtype=torch.float32).pin_memory() for _ in range(2)]
db = torch.zeros(n_out, dtype=torch.float32, device=dev)
din = [torch.empty_like(t, device=dev) for t in pin[0]]
streams = [torch.cuda.Stream() for _ in range(3)]
chunk = n_out // B

def d2h_1buf(n):
    for _ in range(n): hb[0].copy_(db, non_blocking=True)
def d2h_chunks(n, up=False):
    for i in range(n):
        if up:
            with torch.cuda.stream(streams[2]):
                for t, d in zip(pin[i % 2], din): d.copy_(t, non_blocking=True)
        for b in range(B):
            with torch.cuda.stream(streams[b & 1]):
                hb[i % 2][b * chunk:(b + 1) * chunk].copy_(db[b * chunk:(b + 1) * chunk], non_blocking=True)
def pipeline(n):
    pending = None
    for i in range(n):
        tk = pipe.submit(anchors_h, *pin[i % 2])
        if pending is not None: pipe.wait(pending)
        pending = tk
    pipe.wait(pending)
MODES = [("d2h_1buf", d2h_1buf), ("d2h_chunks", d2h_chunks), ("bidir", lambda n: d2h_chunks(n, True)), ("pipeline", pipeline)]
ks = [k for k in (1, 2, 4, 8) if k <= world]
for k in ks:
    for name, fn in MODES:
        active = rank < k
        barrier()
        n = 2
        if active:  # size the loop from a short probe (other active ranks probe at the same time)
            t0 = time.perf_counter(); fn(2); torch.cuda.synchronize(); one = (time.perf_counter() - t0) / 2
        else:
            one = 0.0
        one = sh.max_over_ranks(one, dev)
        n = int(max(3, min(200, np.ceil(SEC / max(one, 1e-4)))))
        barrier()
        t0 = time.perf_counter()
        if active:
            fn(n); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gbs = pipe.d2h_bytes * n / dt / 1e9 if active else 0.0
        slow = -sh.max_over_ranks(-gbs if active else -1e9, dev)   # slowest active rank
        agg = sh.sum_over_ranks(gbs, dev)
        if rank == 0:
            print(json.dumps({"active_ranks": k, "mode": name, "steps": n, "d2h_gbs_slowest_rank": round(slow, 2), "d2h_gbs_aggregate": round(agg, 2),
                              "ms_per_step_slowest": round(1e3 * pipe.d2h_bytes / (slow * 1e9), 2), "images_per_s_aggregate": round(agg * 1e9 / pipe.d2h_bytes * B, 1)}), flush=True)
barrier()
pipe.close()
if world > 1: dist.destroy_process_group()
