"""ctypes binding of libfrcnn_b200.so (the C ABI declared in include/frcnn_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and every op raises if
no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libfrcnn_b200.so")

EINVAL, EWORKSPACE, ELIMIT = -1, -2, -3

_f = C.c_float
_i = C.c_int
_i64 = C.c_int64
_p = C.c_void_p
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/frcnn_b200.h declaration by declaration
SIGNATURES = {
    "frcnn_version": (_i, []),
    "frcnn_last_error": (C.c_char_p, []),
    "frcnn_clip_to_window": (_i, [_p, _i64, _f, _f, _f, _f, _p, _p]),
    "frcnn_decode": (_i, [_p, _p, _i64, _i64, _i64, _p, _p]),
    "frcnn_encode": (_i, [_p, _p, _i64, _i64, _i64, _p, _p]),
    "frcnn_scale_boxes": (_i, [_p, _i64, _f, _f, _i, _p, _p]),
    "frcnn_generate_anchors": (_i, [_i, _i, C.POINTER(_f), _i, C.POINTER(_f), _i, _f, _f, _f, _f, _p, _p]),
    "frcnn_valid_anchor_indices": (_i, [_p, _i64, _f, _f, _p, _p, _p]),
    "frcnn_gather_rows": (_i, [_p, _i64, _i64, _i64, _p, _i64, _p, _p]),
    "frcnn_area": (_i, [_p, _i64, _p, _p]),
    "frcnn_intersection_pairwise": (_i, [_p, _i64, _p, _i64, _p, _p]),
    "frcnn_iou_pairwise": (_i, [_p, _i64, _p, _i64, _p, _p]),
    "frcnn_iou_elementwise": (_i, [_p, _p, _i64, _p, _p]),
    "frcnn_intersection_elementwise": (_i, [_p, _p, _i64, _p, _p]),
    "frcnn_postprocess_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "frcnn_postprocess": (_i, [_f, _f, _p, _i, _p, _p, _i, _i, _i, _f, _f, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p,
                               _sz, _p]),
    "frcnn_postprocess_logits": (_i, [_f, _f, _p, _i, _p, _p, _p, _i, _i, _i, _f, _f, _i, _i, _i, _p, _p, _p, _p, _p, _p,
                                      _p, _sz, _p]),
    "frcnn_decode_filter": (_i, [_f, _f, _p, _i, _p, _p, _i, _i, _i, _f, _p, _p, _p]),
    "frcnn_nms_workspace_bytes": (_sz, [_i, _i]),
    "frcnn_nms": (_i, [_p, _p, _i, _i, _f, _f, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "frcnn_roi_pool_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "frcnn_roi_pool_set_max_sms": (_i, [_i]),
    "frcnn_roi_pool_forward": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _i, _p, _p, _sz, _p]),
    "frcnn_roi_pool_backward": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _i, _p, _p, _p]),
    "frcnn_roi_pool_prepare": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "frcnn_roi_pool_forward_prepared": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "frcnn_roi_pool_backward_prepared": (_i, [_i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "frcnn_roi_pool_argmax_supported": (_i, [_i, _i, _i]),
    "frcnn_roi_pool_forward_argmax": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "frcnn_roi_pool_backward_argmax": (_i, [_i, _i, _i, _i, _p, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "frcnn_generate_targets_workspace_bytes": (_sz, [_i, _i]),
    "frcnn_generate_targets": (_i, [_p, _p, _i, _i, _i, _p, _i, _i, _f, _f, _f, _f, _f, _f, _p, _p, _p, _p, _p, _sz,
                                    _p]),
    "frcnn_generate_targets_objectness": (_i, [_p, _p, _i, _i, _i, _p, _i, _i, _f, _f, _f, _f, _f, _f, _p, _p, _p, _p, _p,
                                               _sz, _p]),
    "frcnn_objectness_labels": (_i, [_p, _i64, _i, _p, _p]),
    "frcnn_ap_match": (_i, [_p, _i, _i, _p, _i, _f, _p, _p, _p]),
    "frcnn_sampled_losses": (_i, [_p, _p, _i64, _i, _p, _p, _i64, _f, _f, _p, _p, _p, _p]),
    "frcnn_sample_indices_workspace_bytes": (_sz, [_i, _i]),
    "frcnn_sample_and_gather": (_i, [_p, _i, _i, _i, _i, C.c_double, C.c_uint64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "frcnn_pipeline_create": (_i, [C.POINTER(_p), _i, _i, _i, _i, _i, _i, _i, _i]),
    "frcnn_pipeline_destroy": (None, [_p]),
    "frcnn_pipeline_run": (_i, [_p, _f, _f, _p, _p, _p, _p, _f, _f, _i, _p, _p, _p, _p]),
    "frcnn_pipeline_submit": (_i, [_p, _f, _f, _p, _p, _p, _p, _f, _f, _i, _p, _p, _p, _p, C.POINTER(C.c_int32)]),
    "frcnn_pipeline_wait": (_i, [_p, C.c_int32]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().frcnn_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    """Map the C ABI's return code to the reference-style Python exceptions."""
    if rc == 0:
        return
    msg = f"{what}: {last_error()} (code {rc})"
    if rc in (EINVAL, ELIMIT):
        raise ValueError(msg)
    raise RuntimeError(msg)
