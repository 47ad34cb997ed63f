"""Mirror of the reference's ``utils`` package for the hot path (boxes, metrics.iou, post_processing, training)."""
