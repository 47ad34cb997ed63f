#!/usr/bin/env python
"""Benchmark of the hot path: images/sec through RPN decode + NMS + RoI pooling (BASELINE.json:metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload kitti_b8] [--impl b200|reference]

One "step" = one pass of the path over one per-GPU batch of synthetic KITTI-shaped input
(proposals for the batch, then RoI pooling of the 300 proposals of every image).  At N GPUs the batch shards
per image (weak scaling: every rank runs the per-GPU batch; no collective on the data path).  Rank 0 prints
ONE JSON line (see README / DESIGN.md for the fields).  `--impl reference` times the CPU restatement of the
reference's TF2 path (the oracle port: TensorFlow is not installable here) on the host cores instead.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "images/sec through RPN decode+NMS+RoIAlign"
UNIT = "images/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ inputs
def make_inputs(syn, w, batch, seed, channels=None):
    rng = np.random.default_rng(seed)
    scores, deltas = syn.rpn_head_outputs(rng, batch, w.num_anchors)
    feat = syn.feature_maps(rng, batch, w.grid, channels or w.channels, relu=True)
    return scores, deltas, feat


def config_dict(w, args, extra=None):
    cfg = {
        "workload": f"{w.name}: image {w.image_shape[0]}x{w.image_shape[1]}, grid {w.grid[0]}x{w.grid[1]}, "
                    f"{w.anchors_per_location} anchors/loc (N={w.num_anchors}), C={w.channels}, NMS iou>{w.iou_threshold} "
                    f"pre {w.pre_nms_top_n} / post {w.max_total_size}, RoI pool {w.pooled_size}x{w.pooled_size} "
                    f"(crop {w.pooled_size * w.kernel_size} + maxpool {w.kernel_size})",
        "per_gpu_batch": w.batch,
        "global_batch": w.batch * args.gpus,
        "parallelism": f"per-image shard x{args.gpus}, no collective",
    }
    cfg.update(extra or {})
    return cfg


# ------------------------------------------------------------------------------------------ CPU baseline
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def cpu_path_step(oracle, w, anchors, scores, deltas, feat):
    """The reference's CPU path for one batch: postprocess_output -> ROIPooling (oracle port, OpenMP)."""
    out = oracle.postprocess_output(w.image_shape, anchors, scores, deltas, **w.nms_kwargs(), pre_nms_top_n=w.pre_nms_top_n)
    pooled = oracle.roi_pool_forward(feat, out["pred_boxes"], w.pooled_size, w.kernel_size)
    return out, pooled


def time_cpu_baseline(w, syn, images, budget_s=20.0, min_reps=2, max_reps=50):
    import oracle
    oracle.build()
    oracle.set_num_threads(host_threads())
    anchors = oracle.clip_to_window(oracle.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape),
                                    [0, 0, w.image_shape[1], w.image_shape[0]])
    scores, deltas, feat = make_inputs(syn, w, images, 777)
    cpu_path_step(oracle, w, anchors, scores, deltas, feat)  # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < max_reps and (len(times) < min_reps or time.perf_counter() - t_start < budget_s):
        t0 = time.perf_counter()
        cpu_path_step(oracle, w, anchors, scores, deltas, feat)
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return {"value": images / med, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": f"{images} images of the same workload x {len(times)} repetitions (median), CPU restatement of the "
                      f"reference's TF2 path (TensorFlow not installable here), OpenMP over images/RoIs"}, med


def run_reference_arm(args, w, syn):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    oracle.build()
    oracle.set_num_threads(host_threads())  # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread
    anchors = oracle.clip_to_window(oracle.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape),
                                    [0, 0, w.image_shape[1], w.image_shape[0]])
    # bounded sample: a whole per-GPU batch unless that would push the run past a few minutes
    images = w.batch
    scores, deltas, feat = make_inputs(syn, w, images, 777)
    t0 = time.perf_counter()
    cpu_path_step(oracle, w, anchors, scores, deltas, feat)
    est = time.perf_counter() - t0
    while images > 1 and est * (args.steps + args.warmup) > 150.0:
        images = max(1, images // 2)
        est /= 2
        scores, deltas, feat = scores[:images], deltas[:images], feat[:images]
    for _ in range(args.warmup):
        cpu_path_step(oracle, w, anchors, scores, deltas, feat)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_path_step(oracle, w, anchors, scores, deltas, feat)
    dt = time.perf_counter() - t0
    value = images * args.steps / dt
    sample = (f"{images} images/step of the same workload, CPU restatement of the reference's TF2 path "
              f"(TensorFlow not installable here), OpenMP on all host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(w, args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cores": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs.
    NVML is initialised in the constructor (outside the timed region); ``sample_once`` lets the main thread take a
    guaranteed sample while the GPU is still busy with the queued steps (the timed region lasts only milliseconds)."""

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.period = period
        self.samples, self.reasons, self.max_mhz, self.err = [], set(), None, None
        self._stop_evt = threading.Event()
        self._lock = threading.Lock()
        self._h = None
        try:
            import pynvml as nv
            self._nv = nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = index
            if visible:
                try:
                    idx = int(visible.split(",")[index])
                except (ValueError, IndexError):
                    pass
            self._h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
            self._names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def sample_once(self):
        if self._h is None:
            return
        nv = self._nv
        try:
            mhz = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            with self._lock:
                self.samples.append(mhz)
                for bit, name in self._names.items():
                    if r & bit:
                        self.reasons.add(name)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        while not self._stop_evt.is_set():
            self.sample_once()
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.err:
            out["error"] = self.err
        return out


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200_arm(args, w, syn):
    import torch
    import torch.distributed as dist
    det = importlib.import_module("2d_object_detection_b200")
    sh = importlib.import_module("2d_object_detection_b200.sharding")
    rank, world, local_rank = sh.rank_world()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    peak_gbs, peak_src = load_peaks()
    B = w.batch
    anchors = det.clip_to_window(det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape),
                                 [0, 0, w.image_shape[1], w.image_shape[0]])
    eng = det.engine.ProposalRoIEngine(w.image_shape, B, w.grid, w.channels, w.num_anchors, w.score_threshold,
                                       w.iou_threshold, w.max_total_size, w.pooled_size, w.kernel_size, w.pre_nms_top_n)
    # Inputs larger than L2: rotate over n_sets input sets (each rank its own seeds = its own images)
    set_bytes = 4 * B * (w.grid[0] * w.grid[1] * w.channels + 6 * w.num_anchors)
    n_sets = max(3, int(np.ceil(1.6 * 126e6 / set_bytes)) + 1)
    host_sets = [make_inputs(syn, w, B, 20240 + 1000 * rank + s) for s in range(n_sets)]
    dev_sets = [tuple(torch.as_tensor(a).to(device) for a in hs) for hs in host_sets]

    def step(i, engine):
        sc, dl, ft = dev_sets[i % n_sets]
        evs = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        engine.step(anchors, sc, dl, ft, pool_events=evs)
        return evs

    # serial single-stream latency of one step (no overlap between consecutive steps)
    eng_serial = det.engine.ProposalRoIEngine(w.image_shape, B, w.grid, w.channels, w.num_anchors, w.score_threshold,
                                              w.iou_threshold, w.max_total_size, w.pooled_size, w.kernel_size,
                                              w.pre_nms_top_n, overlap=False)
    for i in range(3):
        step(i, eng_serial)
    torch.cuda.synchronize()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_lat = min(args.steps, 10)
    l0.record()
    for i in range(n_lat):
        step(i, eng_serial)
    l1.record()
    torch.cuda.synchronize()
    step_latency_ms = l0.elapsed_time(l1) / n_lat
    del eng_serial

    for i in range(max(args.warmup, 3)):
        step(i, eng)
    eng.drain()
    barrier()
    sampler = ClockSampler(local_rank)  # NVML initialised here, outside the timed region; no background thread, so
    t_begin = torch.cuda.Event(enable_timing=True)  # nothing competes with the launch loop for the GIL
    t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record()
    eng.attach()
    pool_events = [step(i, eng) for i in range(args.steps)]
    eng.drain()
    t_end.record()
    # the host runs ahead of the device: poll the clocks while the GPU is still inside the timed region
    sampler.sample_once()
    while not t_end.query():
        sampler.sample_once()
    barrier()
    clocks = sampler.stop()
    ms_total = sh.max_over_ranks(t_begin.elapsed_time(t_end), device)
    images_total = sh.sum_over_ranks(float(B * args.steps), device)
    value = images_total / (ms_total * 1e-3)
    pool_ms = float(np.mean([a.elapsed_time(b) for a, b in pool_events]))
    alg = eng.algorithmic_bytes()
    achieved = alg["roi_pool_fwd"] / (pool_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roi_pool_fwd_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")

    # ---- end to end through the host-buffer C-ABI call (pinned host buffers, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        pipe = det.ProposalRoIPipeline(w.image_shape, B, w.grid, w.channels, w.num_anchors, w.score_threshold,
                                       w.iou_threshold, w.max_total_size, w.pooled_size, w.kernel_size, w.pre_nms_top_n)
        pin = [tuple(torch.as_tensor(a).pin_memory() for a in hs) for hs in host_sets[:2]]
        anchors_h = anchors.cpu().pin_memory()
        e2e_steps = max(3, min(args.steps, args.e2e_steps))
        for i in range(3):
            pipe(anchors_h, *pin[i % 2])
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            pipe(anchors_h, *pin[i % 2])
        torch.cuda.synchronize()
        dt = sh.max_over_ranks(time.perf_counter() - t0, device)
        e2e = {"value": sh.sum_over_ranks(float(B * e2e_steps), device) / dt, "unit": UNIT,
               "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes, "steps": e2e_steps,
               "how": "frcnn_pipeline_run on pinned host buffers, wall clock, max over ranks"}
        pipe.close()

    cpu_baseline = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = time_cpu_baseline(w, syn, min(B, 8))

    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "step_latency_ms": step_latency_ms,
        "overlap": "proposals of step i+1 (stream A) overlap RoI pooling of step i (stream B); double-buffered rois",
        "config": config_dict(w, args, {"l2": f"inputs rotate over {n_sets} sets ({n_sets * set_bytes / 1e6:.0f} MB > 126 MB L2); "
                                              f"{4 * B * w.max_total_size * 49 * w.channels / 1e6:.0f} MB pooled output per step"}),
        "roofline": {"bound": "hbm", "kernel": "roi_pool_fwd_stream_kernel<1024>", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg["roi_pool_fwd"], "kernel_ms": pool_ms,
                     "frac_of_nominal_8TBps": achieved / 8000.0, "share_of_step": pool_ms / (ms_total / args.steps)},
        "cpu_baseline": cpu_baseline,
        "e2e": e2e,
        "gpu_launches": eng.KERNELS_PER_STEP * args.steps,
        "clocks": clocks,
        "host_cores": os.cpu_count(),
        "stage_algorithmic_bytes_per_step": alg,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kitti_b8")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # the synthetic-input module is NumPy only; load it without importing the CUDA package on the reference arm
    import importlib.util
    spec = importlib.util.spec_from_file_location("_frcnn_synthetic", os.path.join(ROOT, "2d_object_detection_b200", "synthetic.py"))
    syn = importlib.util.module_from_spec(spec)
    sys.modules["_frcnn_synthetic"] = syn
    spec.loader.exec_module(syn)
    w = syn.WORKLOADS[args.workload]
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched plainly with --gpus N: start one rank per GPU ourselves (the driver normally does this with torchrun)
        import socket
        import subprocess
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.impl == "reference":
        return run_reference_arm(args, w, syn)
    return run_b200_arm(args, w, syn)


if __name__ == "__main__":
    sys.exit(main())
