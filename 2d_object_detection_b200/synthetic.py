"""Synthetic KITTI-shaped inputs for the hot path; NumPy only, no GPU needed.

Shared by tests/ and bench.py so the parity tests and the benchmark see the same
distributions: jittered-anchor deltas (heavy overlap, so NMS really suppresses),
softmax scores made unique per image (trap T4), post-ReLU-like feature maps
(exact zeros exercise the max-pool tie rule, trap T8) and zero-padded ground truth in
the layout of the reference's data/input_pipeline.py:119-128.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# config.json of the reference (values only; read as a spec).
REFERENCE_CONFIG = {
    "num_classes": 7,
    "image_shape": [600, 1987, 3],
    "rpn": {
        "anchors": {"scales": [0.25, 0.5, 1.0, 2.0], "aspect_ratios": [0.5, 1.0, 2.0], "base_anchor_shape": [256, 256]},
        "sampling": {"foreground_iou_interval": [0.7, 1.0], "background_iou_interval": [0.0, 0.3],
                     "num_samples": 256, "foreground_proportion": 0.5},
        "nms": {"score_threshold": 0.0, "iou_threshold": 0.7, "max_output_size_per_class": 300, "max_total_size": 300},
    },
    "rcnn": {
        "roi_pooling": {"pooled_size": 7, "kernel_size": 2},
        "sampling": {"foreground_iou_interval": [0.5, 1.0], "background_iou_interval": [0.0, 0.5],
                     "num_samples": 64, "foreground_proportion": 0.25},
        "nms": {"score_threshold": 0.0, "iou_threshold": 0.6, "max_output_size_per_class": 100, "max_total_size": 300},
    },
}


def feature_grid(image_hw):
    """Stride-16 grid of the Keras ResNet50 conv4 output: n -> floor((n-1)/2)+1, four times."""
    out = []
    for n in image_hw[:2]:
        for _ in range(4):
            n = (n - 1) // 2 + 1
        out.append(int(n))
    return tuple(out)


@dataclass
class Workload:
    """One named benchmark / parity configuration (BASELINE.json:configs)."""
    name: str
    image_shape: tuple
    scales: tuple
    aspect_ratios: tuple
    batch: int
    pre_nms_top_n: int | None
    max_total_size: int
    channels: int = 1024
    iou_threshold: float = 0.7
    score_threshold: float = 0.0
    pooled_size: int = 7
    kernel_size: int = 2
    base_anchor_shape: tuple = (256, 256)
    grid: tuple = field(init=False)

    def __post_init__(self):
        self.grid = feature_grid(self.image_shape)

    @property
    def anchors_per_location(self):
        return len(self.scales) * len(self.aspect_ratios)

    @property
    def num_anchors(self):
        return self.grid[0] * self.grid[1] * self.anchors_per_location

    def nms_kwargs(self):
        return dict(score_threshold=self.score_threshold, iou_threshold=self.iou_threshold,
                    max_output_size_per_class=self.max_total_size, max_total_size=self.max_total_size)


_R3 = (0.5, 1.0, 2.0)
WORKLOADS = {
    # BASELINE.json configs[0]/[1]/[2]: KITTI 375x1242, 24x78 grid, 9 anchors/location, pre 2000 / post 300.
    "kitti_b1": Workload("kitti_b1", (375, 1242, 3), (0.25, 0.5, 1.0), _R3, 1, 2000, 300),
    "kitti_b8": Workload("kitti_b8", (375, 1242, 3), (0.25, 0.5, 1.0), _R3, 8, 2000, 300),
    # the reference's own 12 anchors, no pre-NMS cap (exact reference behaviour, trap T1).
    "kitti_b8_ref12": Workload("kitti_b8_ref12", (375, 1242, 3), (0.25, 0.5, 1.0, 2.0), _R3, 8, None, 300),
    # configs[3]: dense-anchor stress 800x1333, 50x84 grid, 28 anchors/location, pre 12000 / post 1000.
    "dense_b16": Workload("dense_b16", (800, 1333, 3), (0.125, 0.177, 0.25, 0.354, 0.5, 0.707, 1.0),
                          (0.5, 1.0, 2.0, 3.0), 16, 12000, 1000),
    # configs[4]: training path, KITTI shape, batch 16 per GPU, training-mode RPN (anchors inside the image, no pre-NMS cap).
    "train_b16": Workload("train_b16", (375, 1242, 3), (0.25, 0.5, 1.0), _R3, 16, None, 300),
    # config.json default of the reference (600x1987, 12 anchors, no cap).
    "default_b2": Workload("default_b2", (600, 1987, 3), (0.25, 0.5, 1.0, 2.0), _R3, 2, None, 300),
}


def make_unique_per_row(scores: np.ndarray) -> np.ndarray:
    """Nudge duplicated fp32 values upward by ulps until every row is duplicate-free (trap T4)."""
    s = np.array(scores, dtype=np.float32, copy=True)
    for row in s.reshape(-1, s.shape[-1]):
        for _ in range(64):
            order = np.argsort(row, kind="stable")
            v = row[order]
            dup = np.nonzero(v[1:] == v[:-1])[0] + 1
            if dup.size == 0:
                break
            row[order[dup]] = np.nextafter(v[dup], np.float32(2.0))
        else:  # pragma: no cover
            raise RuntimeError("could not make scores unique")
    return s


def rpn_head_outputs(rng: np.random.Generator, batch: int, num_anchors: int, unique: bool = True):
    """pred_scores [B,N,2] (softmax, fg made unique) and pred_boxes [B,N,1,4] (jittered-anchor deltas)."""
    logits = rng.normal(0.0, 1.5, size=(batch, num_anchors, 2)).astype(np.float32)
    e = np.exp(logits - logits.max(-1, keepdims=True))
    p = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    fg = make_unique_per_row(p[..., 1]) if unique else p[..., 1]
    scores = np.stack([np.float32(1.0) - fg, fg], axis=-1).astype(np.float32)
    deltas = np.concatenate([rng.normal(0.0, 0.1, size=(batch, num_anchors, 1, 2)),
                             rng.normal(0.0, 0.2, size=(batch, num_anchors, 1, 2))], axis=-1).astype(np.float32)
    return scores, deltas


def feature_maps(rng: np.random.Generator, batch: int, grid, channels: int, relu: bool = True) -> np.ndarray:
    """[B,Hf,Wf,C] fp32 NHWC: set A N(0,1) (tie-free) or set B max(N(0,1),0) (post-ReLU-like)."""
    f = rng.standard_normal(size=(batch, grid[0], grid[1], channels), dtype=np.float32)
    return np.maximum(f, 0.0, out=f) if relu else f


def ground_truth(rng: np.random.Generator, batch: int, num_classes: int = 7, max_num_objects: int = 100,
                 max_objects: int = 20, allow_empty: bool = False):
    """gt_labels [B,G,nc+1] one-hot with bg column 0, gt_boxes [B,G,4] relative, zero padded
    (data/input_pipeline.py:119-128)."""
    labels = np.zeros((batch, max_num_objects, num_classes + 1), np.float32)
    boxes = np.zeros((batch, max_num_objects, 4), np.float32)
    for b in range(batch):
        n = int(rng.integers(0 if allow_empty else 1, max_objects + 1))
        w = rng.uniform(0.03, 0.3, size=n)
        h = rng.uniform(0.05, 0.5, size=n)
        x0 = rng.uniform(0.0, 1.0 - w)
        y0 = rng.uniform(0.0, 1.0 - h)
        boxes[b, :n] = np.stack([x0, y0, x0 + w, y0 + h], -1).astype(np.float32)
        labels[b, np.arange(n), 1 + rng.integers(0, num_classes, size=n)] = 1.0
    return labels, boxes


def random_rois(rng: np.random.Generator, batch: int, num_rois: int, num_valid=None) -> np.ndarray:
    """Relative [x1,y1,x2,y2] RoIs clipped to [0,1]; rows past num_valid are [0,0,0,0] (trap T7)."""
    cx, cy = rng.uniform(0, 1, (batch, num_rois)), rng.uniform(0, 1, (batch, num_rois))
    w, h = rng.uniform(0.02, 0.6, (batch, num_rois)), rng.uniform(0.05, 0.9, (batch, num_rois))
    r = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], -1).astype(np.float32)
    r = np.clip(r, 0.0, 1.0)
    if num_valid is not None:
        for b in range(batch):
            r[b, int(np.atleast_1d(num_valid)[b % np.size(num_valid)]):] = 0.0
    return r
