"""Drop-in for the reference's utils/metrics.py: the IoU primitives (:136-208) and the AP / mAP metrics (:4-133)."""
from __future__ import annotations

import torch

from .._tensor import as_device, check, lib, ptr, stream


def _boxes2d(x, name):
    t = as_device(x)
    if t.dim() != 2 or t.shape[1] != 4:
        raise ValueError(f"{name} must have shape [N, 4], got {tuple(t.shape)}")
    return t


def area(boxes):
    """utils/metrics.py:136-147 -> [N]."""
    b = _boxes2d(boxes, "boxes")
    out = torch.empty(b.shape[0], dtype=torch.float32, device=b.device)
    check(lib().frcnn_area(ptr(b), b.shape[0], ptr(out), stream()), "area")
    return out


def intersection(boxes_1, boxes_2, pairwise=False):
    """utils/metrics.py:150-180 -> [N, M] if pairwise else [N]."""
    b1, b2 = _boxes2d(boxes_1, "boxes_1"), _boxes2d(boxes_2, "boxes_2")
    if not pairwise:
        if b1.shape != b2.shape:
            raise ValueError("elementwise intersection needs equally shaped inputs")
        out = torch.empty(b1.shape[0], dtype=torch.float32, device=b1.device)
        check(lib().frcnn_intersection_elementwise(ptr(b1), ptr(b2), b1.shape[0], ptr(out), stream()), "intersection")
        return out
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    check(lib().frcnn_intersection_pairwise(ptr(b1), b1.shape[0], ptr(b2), b2.shape[0], ptr(out), stream()),
          "intersection")
    return out


def iou(boxes_1, boxes_2, pairwise=False):
    """utils/metrics.py:183-208: inter / (area_1 + area_2 - inter), 0 where inter == 0."""
    b1, b2 = _boxes2d(boxes_1, "boxes_1"), _boxes2d(boxes_2, "boxes_2")
    if pairwise:
        out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
        check(lib().frcnn_iou_pairwise(ptr(b1), b1.shape[0], ptr(b2), b2.shape[0], ptr(out), stream()), "iou")
        return out
    if b1.shape != b2.shape:
        raise ValueError("elementwise iou needs equally shaped inputs")
    out = torch.empty(b1.shape[0], dtype=torch.float32, device=b1.device)
    check(lib().frcnn_iou_elementwise(ptr(b1), ptr(b2), b1.shape[0], ptr(out), stream()), "iou")
    return out


class AveragePrecision:
    """Drop-in for utils/metrics.py:4-84 (a tf.keras.metrics.Metric in the reference): interpolated AP over
    ``num_points`` recall levels.  ``update_state`` matches a whole batch with one kernel launch
    (``frcnn_ap_match``: every non-padding ground-truth box marks its highest-IoU prediction as a true positive when
    that IoU exceeds the threshold); the accumulated state stays on the device.  ``result()`` is evaluation code off
    the hot path and uses torch's sort / cumsum exactly as the reference uses tf.argsort / tf.cumsum."""

    def __init__(self, iou_threshold, num_points=11, **kwargs):
        self.iou_threshold = float(iou_threshold)
        self.num_points = int(num_points)
        self.reset_states()

    def reset_states(self):
        self.true_count = 0      # ground-truth boxes seen (tensor once updated: no host synchronisation)
        self.pos_count = 0       # predictions seen, zero-padded ones included, as in the reference (:74)
        self._true_pos, self._scores = [], []

    def update_state(self, gt_boxes, pred_boxes, pred_scores):
        """gt_boxes [B,G,4] zero padded, pred_boxes [B,M,4], pred_scores [B,M]."""
        gt, pb, ps = as_device(gt_boxes), as_device(pred_boxes), as_device(pred_scores)
        if gt.dim() != 3 or pb.dim() != 3 or gt.shape[-1] != 4 or pb.shape[-1] != 4 or gt.shape[0] != pb.shape[0] \
                or tuple(ps.shape) != tuple(pb.shape[:2]):
            raise ValueError(f"expected gt [B,G,4], pred [B,M,4], scores [B,M]; got {tuple(gt.shape)}, {tuple(pb.shape)}, "
                             f"{tuple(ps.shape)}")
        B, G, _ = gt.shape
        M = pb.shape[1]
        tp = torch.empty((B, M), dtype=torch.int32, device=pb.device)
        ngt = torch.empty((B,), dtype=torch.int32, device=pb.device)
        check(lib().frcnn_ap_match(ptr(gt), B, G, ptr(pb), M, self.iou_threshold, ptr(tp), ptr(ngt), stream()), "ap_match")
        self.true_count = self.true_count + ngt.sum()
        self.pos_count += B * M
        self._true_pos.append(tp.reshape(-1))
        self._scores.append(ps.reshape(-1))

    def update_state_single_image(self, gt_boxes, pred_boxes, pred_scores):
        self.update_state(as_device(gt_boxes)[None], as_device(pred_boxes)[None], as_device(pred_scores)[None])

    def result(self):
        if not self._scores:
            return torch.zeros((), dtype=torch.float32)
        scores, true_pos = torch.cat(self._scores), torch.cat(self._true_pos)
        order = torch.argsort(scores, descending=True, stable=True)
        cum_tp = torch.cumsum(true_pos[order], 0).to(torch.float32)
        cum_pos = torch.arange(1, self.pos_count + 1, dtype=torch.float32, device=scores.device)
        zero, two = cum_tp.new_zeros(1), cum_tp.new_full((1,), 2.0)
        precisions = torch.cat([cum_tp / cum_pos, zero])
        recalls = torch.cat([cum_tp / torch.as_tensor(self.true_count, device=scores.device).to(torch.float32), two])
        levels = torch.arange(self.num_points, dtype=torch.float32, device=scores.device) / float(self.num_points - 1)
        interp = torch.where(recalls[None, :] >= levels[:, None], precisions[None, :], precisions.new_full((), -float("inf"))).amax(1)
        return interp.mean()


class MeanAveragePrecision:
    """utils/metrics.py:87-133: one AveragePrecision per class, fed with the boxes / predictions of that class (the others
    zeroed, exactly like the reference's tf.where masks), averaged over the classes."""

    def __init__(self, num_classes, iou_threshold, num_points=11, **kwargs):
        self.num_classes = int(num_classes)
        self.average_precisions = [AveragePrecision(iou_threshold, num_points) for _ in range(self.num_classes)]

    def reset_states(self):
        for ap in self.average_precisions:
            ap.reset_states()

    def result(self):
        return sum(ap.result() for ap in self.average_precisions) / float(self.num_classes)

    def update_state(self, gt_boxes, gt_class_labels, pred_boxes, pred_scores, pred_classes):
        gt, gl = as_device(gt_boxes), as_device(gt_class_labels)[..., 1:]
        pb, ps, pc = as_device(pred_boxes), as_device(pred_scores), as_device(pred_classes, dtype=torch.int32)
        for i, ap in enumerate(self.average_precisions):
            keep = pc == i
            ap.update_state(torch.where((gl[:, :, i] == 1.0)[..., None], gt, torch.zeros_like(gt)),
                            torch.where(keep[..., None], pb, torch.zeros_like(pb)),
                            torch.where(keep, ps, torch.zeros_like(ps)))
