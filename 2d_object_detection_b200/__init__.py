"""B200-native (sm_100a) region-proposal / RoI hot path of antoineBarbez/2D_object_detection.

The sub-packages mirror the reference's module paths so the path drops into ``faster_rcnn.py`` /
``train_faster_rcnn.py``:

    utils.boxes            clip_to_window, decode, encode, to_absolute, to_relative
    utils.metrics          area, intersection, iou, AveragePrecision, MeanAveragePrecision
    utils.post_processing  postprocess_output  (+ stage functions decode_filter, non_max_suppression)
    utils.training         generate_targets, generate_targets_batched, get_sample_indices, sample_and_gather_batched
    utils.losses           ClassificationLoss, RegressionLoss, sampled_losses
    models.detectors.rpn_detector        generate_anchors, valid_anchor_indices, RPNAnchors (= RPNDetector: anchors,
                                         training filter, get_training_samples)
    models.detectors.fast_rcnn_detector  ROIPooling (differentiable), FastRCNNDetector (pool, regions, get_training_samples)
    engine                 ProposalRoIEngine (inference), TrainStepEngine (training path), pre-allocated + graph-capturable
    pipeline               ProposalRoIPipeline (host buffers in, host buffers out)

Python only marshals arguments; all arithmetic runs in hand-written CUDA kernels behind the C ABI of
``include/frcnn_b200.h`` (``libfrcnn_b200.so``).  There is no CPU, PyTorch-eager or Triton fallback: importing
this package without the built library raises.

The directory name starts with a digit, so import it with
``importlib.import_module("2d_object_detection_b200")`` or through the ``b200det`` alias module.
"""
from . import _lib

_lib.load()  # fail loudly at import if the CUDA library has not been built

from . import engine, sharding, synthetic  # noqa: E402
from .models.detectors.fast_rcnn_detector import (FastRCNNDetector, ROIPooling, roi_pool_backward, roi_pool_backward_argmax,  # noqa: E402
                                                  roi_pool_forward, roi_pool_forward_argmax)
from .models.detectors.rpn_detector import (RPNAnchors, RPNDetector, generate_anchors,  # noqa: E402
                                            remove_invalid_anchors_and_predictions, valid_anchor_indices)
from .pipeline import ProposalRoIPipeline  # noqa: E402
from .utils.boxes import clip_to_window, decode, encode, to_absolute, to_relative  # noqa: E402
from .utils.metrics import AveragePrecision, MeanAveragePrecision, area, intersection, iou  # noqa: E402
from .utils.post_processing import decode_filter, non_max_suppression, postprocess_output  # noqa: E402
from .utils.losses import ClassificationLoss, RegressionLoss, sampled_losses  # noqa: E402
from .utils.training import (generate_targets, generate_targets_batched, get_sample_indices, get_training_samples,  # noqa: E402
                             rpn_objectness_labels, sample_and_gather_batched)

__version__ = "0.1.0"
