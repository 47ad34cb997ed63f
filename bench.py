#!/usr/bin/env python
"""Benchmark of the hot path: images/sec through RPN decode + NMS + RoI pooling (BASELINE.json:metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload kitti_b8|kitti_b1|dense_b16|train_b16]
                    [--impl b200|reference] [--min-seconds 1.0] [--no-others]

One "step" = one pass of the path over one per-GPU batch of synthetic KITTI-shaped input (proposals for the batch, then
RoI pooling of the proposals of every image; `train_b16`: the training-path step of BASELINE configs[4] incl. the
weight-gradient all-reduce).  At N GPUs the batch shards per image (weak scaling: every rank runs the per-GPU batch, no
collective on the inference data path).  Rank 0 prints ONE JSON line.

Timing method (B200 arm)
  * `value`: a block of K steps is captured ONCE in a CUDA graph (two streams inside: proposals of step i+1 overlap the
    RoI pooling of step i) and the block is replayed back to back until at least --min-seconds (default 1 s) have passed;
    `ms_per_step` = region / (K * replays).  Device timing with CUDA events, barrier + synchronize on both sides, max over
    ranks.  SM clocks / throttle reasons are sampled through NVML during the whole region.
  * `roofline`: the dominant kernel (RoI-pool forward) is timed with CUDA events around its launches on its own stream in an
    eager run of the same steps right after the graph region (events cannot be read back from inside a graph), same
    overlap with the proposal kernels; `achieved` = algorithmic bytes per launch / that mean time.
  * `e2e`: the host-buffer C-ABI call (frcnn_pipeline_submit / _wait, two batches in flight) on pinned host buffers,
    wall clock, copies inside the region; `pcie_gbs` next to a plain pinned copy of the same size measured in the run.
  * `other_workloads`: the remaining BASELINE configs (kitti_b1, dense_b16, train_b16) measured the same way in shorter
    regions, so that every config appears in the driver-run line.
`--impl reference` times the CPU restatement of the reference's TF2 path (the oracle port: TensorFlow is not installable
here) on the host cores instead.
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
L2_BYTES = 126e6
GRAD_MB = 34.2  # fp32 weight gradients of the reference's trainable backbone part + heads' share used by BASELINE configs[4]


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(workload, kernel):
    """DRAM bytes per launch of `kernel` on `workload` from an ncu --set full capture of this build (profiles/traffic.json,
    keyed by workload and kernel); None when that pair has not been captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(workload, {}).get(kernel)


# ------------------------------------------------------------------------------------------ inputs
def make_inputs(syn, w, batch, seed, channels=None):
    rng = np.random.default_rng(seed)
    scores, deltas = syn.rpn_head_outputs(rng, batch, w.num_anchors)
    feat = syn.feature_maps(rng, batch, w.grid, channels or w.channels, relu=True)
    return scores, deltas, feat


def input_sets(w):
    set_bytes = 4 * w.batch * (w.grid[0] * w.grid[1] * w.channels + 6 * w.num_anchors)
    return max(3, int(np.ceil(1.6 * L2_BYTES / set_bytes)) + 1), set_bytes


def config_dict(w, gpus, train=False):
    n_sets, set_bytes = input_sets(w)
    cfg = {
        "workload": f"{w.name}: image {w.image_shape[0]}x{w.image_shape[1]}, grid {w.grid[0]}x{w.grid[1]}, "
                    f"{w.anchors_per_location} anchors/loc (N={w.num_anchors}), C={w.channels}, NMS iou>{w.iou_threshold} "
                    f"pre {w.pre_nms_top_n} / post {w.max_total_size}, RoI pool {w.pooled_size}x{w.pooled_size} "
                    f"(crop {w.pooled_size * w.kernel_size} + maxpool {w.kernel_size})",
        "per_gpu_batch": w.batch,
        "global_batch": w.batch * gpus,
        "parallelism": f"per-image shard x{gpus}, " + ("weight-gradient all-reduce only" if train else "no collective"),
        "l2": f"inputs rotate over {n_sets} sets ({n_sets * set_bytes / 1e6:.0f} MB > 126 MB L2); "
              f"{4 * w.batch * w.max_total_size * 49 * w.channels / 1e6:.0f} MB pooled output per step",
    }
    return cfg


# ------------------------------------------------------------------------------------------ CPU baseline / reference arm
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


def cpu_train_step(oracle, w, anchors_valid, vidx, scores, deltas, feat, gl, gb, dy):
    """The training-path step on the CPU (oracle port): proposals on the valid anchors, both target matchings, RoI pool
    forward + backward (sampling and the tiny loss reductions are not ported: they are < 1 % of the CPU time)."""
    sc, dl = scores[:, vidx], deltas[:, vidx]
    out = oracle.postprocess_output(w.image_shape, anchors_valid, sc, dl, **w.nms_kwargs())
    oracle.generate_targets_batched(oracle.rpn_objectness_labels(gl), gb, anchors_valid, w.image_shape, [0.7, 1.0], [0.0, 0.3])
    rois = out["pred_boxes"]
    oracle.generate_targets_batched(gl, gb, oracle.to_absolute(rois, w.image_shape), w.image_shape, [0.5, 1.0], [0.0, 0.5])
    oracle.roi_pool_forward(feat, rois, w.pooled_size, w.kernel_size)
    oracle.roi_pool_backward(feat, rois, dy, w.pooled_size, w.kernel_size)


def _oracle_anchors(oracle, w):
    return oracle.clip_to_window(oracle.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape),
                                 [0, 0, w.image_shape[1], w.image_shape[0]])


def time_cpu_baseline(w, syn, images, budget_s=20.0, min_reps=2, max_reps=50):
    import oracle
    oracle.build()
    oracle.set_num_threads(host_threads())
    anchors = _oracle_anchors(oracle, w)
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
                      f"reference's TF2 path (TensorFlow not installable here), OpenMP over images/RoIs"}


def run_reference_arm(args, w, syn):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    oracle.build()
    oracle.set_num_threads(host_threads())  # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread
    train = args.workload == "train_b16"
    images = w.batch
    scores, deltas, feat = make_inputs(syn, w, images, 777)
    if train:
        full = oracle.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
        vidx = oracle.valid_anchor_indices(full, w.image_shape)
        anchors = full[vidx]
        rng = np.random.default_rng(778)
        gl, gb = syn.ground_truth(rng, images)
        dy = rng.standard_normal((images, w.max_total_size, 7, 7, w.channels)).astype(np.float32)
        step = lambda n: cpu_train_step(oracle, w, anchors, vidx, scores[:n], deltas[:n], feat[:n], gl[:n], gb[:n], dy[:n])  # noqa: E731
    else:
        anchors = _oracle_anchors(oracle, w)
        step = lambda n: cpu_path_step(oracle, w, anchors, scores[:n], deltas[:n], feat[:n])  # noqa: E731
    # bounded sample: a whole per-GPU batch unless that would push the run past a few minutes
    t0 = time.perf_counter()
    step(images)
    est = time.perf_counter() - t0
    while images > 1 and est * (args.steps + args.warmup) > 150.0:
        images = max(1, images // 2)
        est /= 2
    for _ in range(args.warmup):
        step(images)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(images)
    dt = time.perf_counter() - t0
    value = images * args.steps / dt
    sample = (f"{images} images/step of the same workload, CPU restatement of the reference's TF2 path "
              f"(TensorFlow not installable here), OpenMP on all host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(w, args.gpus, train),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cores": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while a timed region runs (NVML initialised in the
    constructor, outside the region).  The regions are graph replays or device-bound loops, so the sampling thread does not
    compete with a launch loop."""

    def __init__(self, index, period=0.01):
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
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
               "sm_mhz_min": float(min(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.err:
            out["error"] = self.err
        return out


def poll_until(event, sampler, period=0.01):
    """The host has queued everything: sample the clocks while the device works through the region."""
    sampler.sample_once()
    while not event.query():
        sampler.sample_once()
        time.sleep(period)


# ------------------------------------------------------------------------------------------ B200 arm
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.det = importlib.import_module("2d_object_detection_b200")
        self.sh = importlib.import_module("2d_object_detection_b200.sharding")
        self.rank, self.world, self.local_rank = self.sh.rank_world()
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.orig_affinity = os.sched_getaffinity(0)
        self.numa = bind_to_gpu_numa_node(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)
        self.peak_gbs, self.peak_src = load_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local_rank])
        self.torch.cuda.synchronize()

    def vmax(self, v):
        return self.sh.max_over_ranks(v, self.device)

    def vsum(self, v):
        return self.sh.sum_over_ranks(v, self.device)

    def close(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local_rank])
            self.dist.destroy_process_group()


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads (and so its first-touch pinned buffers) to the NUMA node of its GPU."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(visible.split(",")[local_rank]) if visible else local_rank
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"node": node, "bound": False}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"node": node, "bound": bool(allowed), "cpus": len(allowed)}
    except Exception as e:  # pragma: no cover
        return {"node": None, "bound": False, "error": repr(e)[:80]}


def measure_inference(ctx, w, syn, steps, warmup, min_seconds, with_e2e, e2e_seconds, reserve_sms=None, pool_streams=None):
    torch, det = ctx.torch, ctx.det
    device, rank = ctx.device, ctx.rank
    B = w.batch
    anchors = det.clip_to_window(det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape),
                                 [0, 0, w.image_shape[1], w.image_shape[0]])
    mk_engine = lambda overlap: det.engine.ProposalRoIEngine(  # noqa: E731
        w.image_shape, B, w.grid, w.channels, w.num_anchors, w.score_threshold, w.iou_threshold, w.max_total_size,
        w.pooled_size, w.kernel_size, w.pre_nms_top_n, overlap=overlap, reserve_sms=reserve_sms, pool_streams=pool_streams)
    eng = mk_engine(True)
    # Inputs larger than L2: rotate over n_sets input sets (each rank its own seeds = its own images)
    n_sets, set_bytes = input_sets(w)
    host_sets = [make_inputs(syn, w, B, 20240 + 1000 * rank + s) for s in range(n_sets)]
    dev_sets = [tuple(torch.as_tensor(a).to(device) for a in hs) for hs in host_sets]

    def step(i, engine, events=False):
        sc, dl, ft = dev_sets[i % n_sets]
        evs = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if events else None
        engine.step(anchors, sc, dl, ft, pool_events=evs)
        return evs

    def engine_step_events(engine, i, pool_events):
        sc, dl, ft = dev_sets[i % n_sets]
        engine.step(anchors, sc, dl, ft, pool_events=pool_events)

    # serial single-stream latency of one step (no overlap between consecutive steps)
    eng_serial = mk_engine(False)
    for i in range(3):
        step(i, eng_serial)
    torch.cuda.synchronize()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_lat = max(3, min(steps, 10))
    l0.record()
    for i in range(n_lat):
        step(i, eng_serial)
    l1.record()
    torch.cuda.synchronize()
    step_latency_ms = l0.elapsed_time(l1) / n_lat
    del eng_serial

    # ---- warm-up (eager), then capture ONE block of `steps` steps (both pipeline streams fork from / join the capture)
    warm = max(warmup, 3)
    for i in range(warm):
        step(i, eng)
    eng.drain()
    torch.cuda.synchronize()
    eng.reset_pipeline()
    # Two EXTERNAL events ride in the graph (event-record nodes on the pooling streams): the start of the pooling launch of
    # step `g_first` and the end of the one of the last step.  After the timed replays their distance is the span of the
    # pooling launches of the last replay -- the kernels' device time inside the timed region itself.
    # The block that is captured holds a multiple of `steps` steps (>= 200 of them): every replay of a block starts with an
    # empty pipeline (the pooling stream idles while the first proposals run, ~90 us), which a 20-step block meets every
    # 3.3 ms.  A stream of batches has no block boundaries; the region still consists of whole `steps`-step units.  (Measured
    # effect on the value: none -- under the power cap the idle time a short block leaves comes back as SM clock, 1.725 GHz
    # with 20-step blocks against 1.71 GHz with 200-step blocks on similar boxes; the long block is simply the steadier state.)
    block = steps * max(1, int(np.ceil(200.0 / max(steps, 1))))
    g_first = min(block // 2, block - 1)  # (mid-block: the pipeline is in its steady state)
    try:
        ev_g0 = torch.cuda.Event(enable_timing=True, external=True)
        ev_g1 = torch.cuda.Event(enable_timing=True, external=True)
    except TypeError:  # (a torch without external events: the eager run below is the only source)
        ev_g0 = ev_g1 = None
    def capture():
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            eng.attach()
            for i in range(block):
                pe = None
                if ev_g0 is not None and (i == g_first or i == block - 1):
                    pe = (ev_g0 if i == g_first else None, ev_g1 if i == block - 1 else None)
                engine_step_events(eng, i, pe)
            eng.drain()
        return g

    try:
        graph = capture()
    except RuntimeError:  # (event-record nodes refused: capture the plain block)
        torch.cuda.synchronize()
        eng.reset_pipeline()
        ev_g0 = ev_g1 = None
        graph = capture()
    graph.replay()  # untimed: instantiation / upload
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    graph.replay()
    t1.record()
    torch.cuda.synchronize()
    block_ms = t0.elapsed_time(t1)
    reps = max(1, int(np.ceil(min_seconds * 1e3 / max(block_ms, 1e-3))))
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for _ in range(reps):
        graph.replay()
    t_end.record()
    poll_until(t_end, sampler)
    ctx.barrier()
    clocks = sampler.stop()
    graph_span_ms = None
    if ev_g0 is not None and block - 1 > g_first:
        try:
            torch.cuda.synchronize()
            graph_span_ms = ev_g0.elapsed_time(ev_g1) / (block - g_first)
        except RuntimeError:
            graph_span_ms = None
    ms_total = ctx.vmax(t_begin.elapsed_time(t_end))
    images_total = ctx.vsum(float(B * block * reps))
    value = images_total / (ms_total * 1e-3)
    ms_per_step = ms_total / (block * reps)

    # ---- the dominant kernel, CUDA events around its launches on its stream (eager run of the same steps, same overlap)
    eng.reset_pipeline()
    n_ev = int(min(2000, max(steps, np.ceil(0.25e3 / max(ms_per_step, 1e-3)))))
    eng.attach()
    pool_events = [step(i, eng, events=True) for i in range(n_ev)]
    eng.drain()
    torch.cuda.synchronize()
    first = min(3, n_ev - 1)
    bracket_ms = float(np.mean([a.elapsed_time(b) for a, b in pool_events[first:]]))
    # With two pooling streams consecutive launches overlap at their ends (the CTAs of launch i + 1 start on the SMs the CTAs
    # of launch i leave), so the bracket of one launch contains time in which its CTAs wait for SMs that still belong to the
    # launch before: the launches' device time is then the span from the first bracket's start to the last bracket's end
    # divided by the launches in it (nothing else runs on those SMs).  With one stream the two agree whenever pooling bounds the
    # step; where it does not (batch 1) the span is the step time and the bracket is the kernel.
    span_ms = pool_events[first][0].elapsed_time(pool_events[-1][1]) / max(1, n_ev - first)
    # (the eager run pays Python launch bubbles: 1-3 % slower than the graph replay; the span inside the replayed graph, taken
    # from the two external events, is the same quantity without them)
    pool_ms = min(bracket_ms, span_ms, graph_span_ms) if graph_span_ms else min(bracket_ms, span_ms)
    alg = eng.algorithmic_bytes()
    achieved = alg["roi_pool_fwd"] / (pool_ms * 1e-3) / 1e9
    kernel = "roi_pool_fwd_persist_kernel<1024>" if w.channels == 1024 else "roi_pool_fwd_persist_kernel"
    # the same kernel timed ALONE in a short burst after a pause (all SMs, nothing beside it, clocks recovered): the
    # condition under which the peak in MEASURED_PEAKS.json was measured (a copy, best of 10); reported next to the number
    # above, which is the kernel inside the sustained, power-capped, overlapped pipeline
    time.sleep(0.25)
    torch.cuda.synchronize()
    det._lib.load().frcnn_roi_pool_set_max_sms(0)
    burst = []
    for i in range(12):
        sc, dl, ft = dev_sets[i % n_sets]
        a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_ev.record()
        det._tensor.check(det._lib.load().frcnn_roi_pool_forward(ft.data_ptr(), B, w.grid[0], w.grid[1], w.channels, eng.rois.data_ptr(),
                                                                   w.max_total_size, w.pooled_size, w.kernel_size, eng.pooled.data_ptr(),
                                                                   eng.pool_ws.data_ptr(), eng.pool_ws_bytes,
                                                                   torch.cuda.current_stream().cuda_stream), "roi_pool_forward")
        b_ev.record()
        burst.append((a_ev, b_ev))
    torch.cuda.synchronize()
    burst_ms = float(np.median([x.elapsed_time(y) for x, y in burst[2:]]))
    res = {
        "value": value, "ms_per_step": ms_per_step, "timed_region_s": ms_total * 1e-3, "block_reps": reps,
        "step_latency_ms": step_latency_ms, "clocks": clocks, "pool_sms": eng.pool_sms,
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": ctx.peak_gbs, "unit": "GB/s",
                     "frac": achieved / ctx.peak_gbs, "traffic": load_traffic(w.name, "roi_pool_fwd"),
                     "peak_source": ctx.peak_src, "algorithmic_bytes_per_launch": alg["roi_pool_fwd"], "kernel_ms": pool_ms,
                     "kernel_ms_how": ("smallest of: span of the pooling launches of the last replayed graph (two external CUDA events "
                                       "in the graph, start of launch %d to end of launch %d) / launches; span of %d launches of "
                                       "the eager run / launches; mean of the eager run's per-launch brackets (with two pooling "
                                       "streams a bracket includes waiting for SMs of the launch before)"
                                       % (g_first, block - 1, n_ev - first)),
                     "kernel_ms_bracketed": bracket_ms, "kernel_ms_eager_span": span_ms, "kernel_ms_graph_span": graph_span_ms,
                     "pool_streams": eng.n_pool_streams,
                     "kernel_launches_timed": n_ev, "frac_of_nominal_8TBps": achieved / 8000.0,
                     "share_of_step": pool_ms / ms_per_step,
                     "share_of_step_note": "kernel_ms comes from the eager run with events (%.4f ms per step), ms_per_step from the "
                                           "graph replay: pooling bounds the step, a share slightly above 1 is the eager run's "
                                           "launch bubbles" % span_ms,
                     "alone_burst": {"kernel_ms": burst_ms, "achieved": alg["roi_pool_fwd"] / (burst_ms * 1e-3) / 1e9,
                                     "frac": alg["roi_pool_fwd"] / (burst_ms * 1e-3) / 1e9 / ctx.peak_gbs,
                                     "how": "records + pool kernel alone on all SMs, 10 launches after a 0.25 s pause, median"},
                     "other_kernels": "decode_filter / nms_select (+ fused single-class finalize): latency-bound by size (a few MB per launch, "
                                      "one CTA per image for NMS), overlapped with the RoI pool of the previous step"},
        "gpu_launches": eng.KERNELS_PER_STEP * block * reps, "steps_per_graph": block,
        "stage_algorithmic_bytes_per_step": alg,
    }
    del graph

    # ---- end to end through the host-buffer C-ABI call (pinned host buffers, copies inside the timed region)
    if with_e2e:
        pipe = det.ProposalRoIPipeline(w.image_shape, B, w.grid, w.channels, w.num_anchors, w.score_threshold,
                                       w.iou_threshold, w.max_total_size, w.pooled_size, w.kernel_size, w.pre_nms_top_n)
        pin = [tuple(torch.as_tensor(a).pin_memory() for a in hs) for hs in host_sets[:2]]
        anchors_h = anchors.cpu().pin_memory()
        for i in range(3):
            pipe(anchors_h, *pin[i % 2])
        ctx.barrier()  # probe with every rank copying at once: the ranks share the host side, and all must time the same steps
        t0 = time.perf_counter()
        pipe(anchors_h, *pin[0])
        pipe(anchors_h, *pin[1])
        one = ctx.vmax((time.perf_counter() - t0) / 2)
        e2e_steps = int(max(4, min(400, np.ceil(e2e_seconds / max(one, 1e-4)))))
        ctx.barrier()
        t0 = time.perf_counter()
        pending = None
        for i in range(e2e_steps):  # two batches in flight: submit i, then wait for i - 1
            tk = pipe.submit(anchors_h, *pin[i % 2])
            if pending is not None:
                pipe.wait(pending)
            pending = tk
        pipe.wait(pending)
        torch.cuda.synchronize()
        dt = ctx.vmax(time.perf_counter() - t0)
        # plain pinned copies of the same sizes, all ranks at once: the ceiling the host side allows
        h_buf = pipe.pooled.reshape(-1)
        d_buf = torch.empty(h_buf.numel(), dtype=torch.float32, device=device)
        h_buf.copy_(d_buf, non_blocking=True)
        ctx.barrier()
        n_cp = int(max(6, min(100, np.ceil(0.4 * e2e_steps / max(e2e_seconds, 1e-3)))))  # ~0.4 s, the same count on every rank
        t0 = time.perf_counter()
        for _ in range(n_cp):
            h_buf.copy_(d_buf, non_blocking=True)
        torch.cuda.synchronize()
        d2h_s = ctx.vmax(time.perf_counter() - t0) / n_cp
        d2h_peak = h_buf.numel() * 4 / d2h_s / 1e9
        total_b = pipe.h2d_bytes + pipe.d2h_bytes
        res["e2e"] = {"value": ctx.vsum(float(B * e2e_steps)) / dt, "unit": UNIT,
                      "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes, "steps": e2e_steps,
                      "timed_region_s": dt, "ms_per_step": 1e3 * dt / e2e_steps,
                      "pcie_gbs": total_b * e2e_steps / dt / 1e9, "d2h_gbs": pipe.d2h_bytes * e2e_steps / dt / 1e9,
                      "pinned_d2h_copy_gbs": d2h_peak, "frac_of_pinned_d2h_copy": pipe.d2h_bytes * e2e_steps / dt / 1e9 / d2h_peak,
                      "limiter": "PCIe / host: the pooled features leaving the device; pinned_d2h_copy_gbs is a plain "
                                 "cudaMemcpyAsync of the same bytes issued by all ranks at once (per-rank, slowest rank). "
                                 "The host side of one box moves 56 / 71 / 73 / 120 GB/s in aggregate with 1 / 2 / 4 / 8 "
                                 "GPUs copying (profiles/r2_host_ceiling_n8.jsonl), so e2e cannot scale past ~2.1x from 1 to 8",
                      "numa": ctx.numa,
                      "how": "frcnn_pipeline_submit/_wait on pinned host buffers, two batches in flight, wall clock, max over ranks"}
        pipe.close()
        del d_buf
    del dev_sets, eng
    torch.cuda.empty_cache()
    return res


def measure_train(ctx, w, syn, steps, warmup, min_seconds):
    """BASELINE configs[4]: the training-path step on a 16-image per-GPU batch (TrainStepEngine: RPN filter, proposals, both
    get_training_samples, losses, RoI pool forward with codes + backward) and, at N > 1, the NCCL all-reduce of a synthetic
    fp32 weight-gradient buffer issued on a second stream between forward and backward, as the gradients of a real step
    would be.  The data path shards per image with no collective."""
    torch, det, dist = ctx.torch, ctx.det, ctx.dist
    device, rank, world = ctx.device, ctx.rank, ctx.world
    cfg = syn.REFERENCE_CONFIG
    B, R, C, nc = w.batch, w.max_total_size, w.channels, cfg["num_classes"]
    rpn_cfg = {"nms": dict(cfg["rpn"]["nms"], max_output_size_per_class=R, max_total_size=R), "sampling": cfg["rpn"]["sampling"]}
    anchors = det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape)
    eng = det.engine.TrainStepEngine(w.image_shape, B, w.grid, C, anchors, nc, rpn_cfg, cfg["rcnn"])
    rng = np.random.default_rng(4040 + rank)
    n_sets = 3
    sets = []
    for _ in range(n_sets):
        scores, deltas = syn.rpn_head_outputs(rng, B, w.num_anchors)
        feat = syn.feature_maps(rng, B, w.grid, C)
        gl, gb = syn.ground_truth(rng, B)
        lg = rng.normal(0, 1, (B, R, nc + 1)).astype(np.float32)
        ps7 = (np.exp(lg) / np.exp(lg).sum(-1, keepdims=True)).astype(np.float32)
        dl7 = rng.normal(0, 0.1, (B, R, nc, 4)).astype(np.float32)
        sets.append(tuple(torch.as_tensor(a).to(device) for a in (scores, deltas, feat, gl, gb, ps7, dl7)))
    dy = torch.randn((B, R, 7, 7, C), device=device)
    grad = torch.randn(int(GRAD_MB * 1e6 / 4), device=device)
    comm = torch.cuda.Stream()
    bwd_events = []

    def step(i, allreduce=True, events=False):
        sc, dl, feat, gl, gb, ps7, dl7 = sets[i % n_sets]
        eng.forward(sc, dl, feat, gl, gb, ps7, dl7, seed=i)
        if world > 1 and allreduce:  # weight gradients of the step: overlap the RoI-pool backward
            comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm):
                dist.all_reduce(grad)
        if events:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        eng.backward(dy)
        if events:
            e1.record()
            bwd_events.append((e0, e1))
        if world > 1 and allreduce:
            torch.cuda.current_stream().wait_stream(comm)

    def timed(n, **kw):
        ctx.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            step(i, **kw)
        b.record()
        return a, b

    for i in range(max(warmup, 3)):
        step(i)
    a, b = timed(max(3, min(steps, 10)))
    torch.cuda.synchronize()
    est = a.elapsed_time(b) / max(3, min(steps, 10))
    # every rank must run the SAME number of steps (each step holds an all-reduce): agree on the count
    n = int(ctx.vmax(float(max(steps, np.ceil(min_seconds * 1e3 / max(est, 1e-3))))))
    sampler = ClockSampler(ctx.local_rank)
    a, b = timed(n)
    poll_until(b, sampler)
    ctx.barrier()
    clocks = sampler.stop()
    ms_total = ctx.vmax(a.elapsed_time(b))
    ms = ms_total / n
    value = ctx.vsum(float(B * n)) / (ms_total * 1e-3)
    # dominant kernel: RoI-pool backward (memset + kernel), events on its stream
    n_ev = max(10, min(200, n))
    for i in range(n_ev):
        step(i, events=True)
    torch.cuda.synchronize()
    bwd_ms = float(np.mean([x.elapsed_time(y) for x, y in bwd_events[2:]]))
    alg = eng.algorithmic_bytes()
    achieved = alg["roi_pool_bwd"] / (bwd_ms * 1e-3) / 1e9
    res = {
        "value": value, "ms_per_step": ms, "timed_region_s": ms_total * 1e-3, "steps_timed": n, "clocks": clocks,
        "N_valid": eng.Nv,
        "roofline": {"bound": "hbm", "kernel": "roi_pool_bwd (memset + kernel)", "achieved": achieved, "peak": ctx.peak_gbs,
                     "unit": "GB/s", "frac": achieved / ctx.peak_gbs, "traffic": load_traffic("train_b16", "roi_pool_bwd"),
                     "peak_source": ctx.peak_src, "algorithmic_bytes_per_launch": alg["roi_pool_bwd"], "kernel_ms": bwd_ms,
                     "share_of_step": bwd_ms / ms},
        "gpu_launches": eng.KERNELS_PER_STEP * n,
        "allreduce": {"mb": GRAD_MB, "ranks": world, "backend": "nccl" if world > 1 else None},
    }
    if world > 1:
        a, b = timed(max(10, n // 4), allreduce=False)
        torch.cuda.synchronize()
        res["allreduce"]["ms_per_step_without"] = ctx.vmax(a.elapsed_time(b)) / max(10, n // 4)
        for _ in range(3):
            dist.all_reduce(grad)
        ctx.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            dist.all_reduce(grad)
        b.record()
        torch.cuda.synchronize()
        res["allreduce"]["alone_ms"] = ctx.vmax(a.elapsed_time(b)) / 20
        res["allreduce"]["exposed_ms"] = ms - res["allreduce"]["ms_per_step_without"]
    del sets, eng, dy, grad
    torch.cuda.empty_cache()
    return res


def run_b200_arm(args, syn):
    ctx = Ctx(args)
    w = syn.WORKLOADS[args.workload]
    train = args.workload == "train_b16"
    if train:
        res = measure_train(ctx, w, syn, args.steps, args.warmup, args.min_seconds)
    else:
        res = measure_inference(ctx, w, syn, args.steps, args.warmup, args.min_seconds, not args.no_e2e, args.e2e_seconds,
                                args.reserve_sms, args.pool_streams)
    others = []
    if not args.no_others:
        for name in ("kitti_b1", "kitti_b8", "dense_b16", "train_b16"):
            if name == args.workload:
                continue
            ow = syn.WORKLOADS[name]
            try:
                if name == "train_b16":
                    r = measure_train(ctx, ow, syn, args.steps, 3, args.other_seconds)
                else:
                    r = measure_inference(ctx, ow, syn, args.steps, 3, args.other_seconds, False, 0.0)
                others.append({"workload": name, "per_gpu_batch": ow.batch, "value": r["value"], "unit": UNIT,
                               "ms_per_step": r["ms_per_step"], "timed_region_s": r["timed_region_s"],
                               "dominant_kernel": r["roofline"]["kernel"], "kernel_ms": r["roofline"]["kernel_ms"],
                               "roofline_frac": r["roofline"]["frac"], "achieved_gbs": r["roofline"]["achieved"],
                               "clocks_sm_mhz": r["clocks"]["sm_mhz"],
                               **({"step_latency_ms": r["step_latency_ms"]} if "step_latency_ms" in r else {}),
                               **({"allreduce": r["allreduce"], "N_valid": r["N_valid"]} if "allreduce" in r else {})})
            except Exception as e:  # a side workload must never take the headline down
                others.append({"workload": name, "error": repr(e)[:300]})
    cpu_baseline = None
    if ctx.rank == 0 and args.gpus == 1 and not args.no_cpu_baseline and not train:
        os.sched_setaffinity(0, ctx.orig_affinity)  # the CPU baseline uses every host core, not only the GPU's NUMA node
        cpu_baseline = time_cpu_baseline(w, syn, min(w.batch, 8))
    ctx.close()
    if ctx.rank != 0:
        return 0
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "timed_region_s": res["timed_region_s"],
        "timing": (f"a CUDA graph of {res.get('steps_per_graph')} steps (= {res.get('steps_per_graph', 0) // max(args.steps, 1)} x --steps, so "
                   f"that the empty pipeline at the start of a replay is amortised) replayed {res.get('block_reps')} times back to back"
                   if not train else
                   f"{res.get('steps_timed')} eager steps back to back") + "; CUDA events, barrier + synchronize on both sides, max over ranks",
        "overlap": ("weight-gradient all-reduce (second stream) overlaps the RoI-pool backward" if train else
                    "proposals + RoI records of step i+1 (stream A, reserved SMs) overlap RoI pooling of step i; pooling kernels "
                    "alternate between two streams from 4 images per batch (their ends overlap); double-buffered rois / outputs"),
        "config": config_dict(w, args.gpus, train),
        "roofline": res["roofline"], "cpu_baseline": cpu_baseline, "e2e": res.get("e2e"),
        "gpu_launches": res["gpu_launches"], "clocks": res["clocks"], "host_cores": os.cpu_count(),
        "other_workloads": others,
    }
    for k in ("step_latency_ms", "block_reps", "steps_per_graph", "pool_sms", "stage_algorithmic_bytes_per_step", "allreduce", "N_valid",
              "steps_timed"):
        if k in res:
            line[k] = res[k]
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kitti_b8")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="length of the headline timed region")
    ap.add_argument("--other-seconds", type=float, default=0.3, help="timed region of each entry of other_workloads")
    ap.add_argument("--e2e-seconds", type=float, default=1.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true")
    ap.add_argument("--reserve-sms", type=int, default=None,
                    help="SMs the pipelined engine keeps free of the persistent RoI kernel for the proposal path (default: per-GPU batch)")
    ap.add_argument("--pool-streams", type=int, default=None,
                    help="pooling streams of the pipelined engine (2: consecutive pooling kernels may overlap at their ends; "
                         "default: 2 from four images per batch, else 1)")
    args = ap.parse_args()
    # the synthetic-input module is NumPy only; load it without importing the CUDA package on the reference arm
    import importlib.util
    spec = importlib.util.spec_from_file_location("_frcnn_synthetic", os.path.join(ROOT, "2d_object_detection_b200", "synthetic.py"))
    syn = importlib.util.module_from_spec(spec)
    sys.modules["_frcnn_synthetic"] = syn
    spec.loader.exec_module(syn)
    if args.workload not in syn.WORKLOADS:
        raise SystemExit(f"unknown workload {args.workload}; choose from {sorted(syn.WORKLOADS)}")
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
        return run_reference_arm(args, syn.WORKLOADS[args.workload], syn)
    return run_b200_arm(args, syn)


if __name__ == "__main__":
    sys.exit(main())
