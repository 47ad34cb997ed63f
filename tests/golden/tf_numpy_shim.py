"""A NumPy stand-in for the handful of ``tensorflow`` symbols the reference's hot-path modules touch, so that the
reference's OWN Python source (utils/boxes.py, utils/metrics.py, utils/training.py, utils/post_processing.py and the
anchor / RoI-pooling methods of models/detectors/*.py) can be imported and executed in the build container, where
TensorFlow itself is not installable.  Test infrastructure only: used by tests/golden/make_reference_golden.py to
produce tests/golden/reference_exec.npz.

STATUS: written but NOT YET RUN against the reference.  Executing the reference's source needs a human's authorisation in
the build environment (the tree is untrusted public content); until the generator has been run and its output
committed, the oracle stays "parity unpinned" (DESIGN.md section 2).

What this pins and what it does not:
  * every line of the reference's Python on the path is executed as written (argument order, axis conventions,
    one-hot / where / gather logic, interval tests, broadcasting);
  * element-wise float32 ops (+ - * / min max exp log sqrt, comparisons) are NumPy's, which are IEEE-identical to
    TensorFlow's CPU kernels except that exp / log may differ in the last ulp;
  * the three TF *kernels* the path calls -- tf.image.combined_non_max_suppression, tf.image.crop_and_resize and
    MaxPool2D -- are NOT TensorFlow's code here: they delegate to the NumPy restatement of their published algorithms
    (oracle/reference_np.py).  Their semantics stay "restated", pinned only by the known-answer tests.
Python scalars and lists are converted to the dtype of the tensor operand (TensorFlow's rule), never to float64.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import reference_np as _rnp  # noqa: E402

float32, int32, int64 = np.float32, np.int32, np.int64
bool = np.bool_  # noqa: A001  (tf.bool)


def _t(x, like=None):
    """convert_to_tensor: arrays pass through; Python numbers / lists take the dtype of the other operand, else
    float32 for floats and int32 for ints."""
    if isinstance(x, np.ndarray) or isinstance(x, np.generic):
        return x
    a = np.asarray(x)
    if like is not None and a.dtype != np.bool_:
        return a.astype(np.asarray(like).dtype)
    if a.dtype == np.float64:
        return a.astype(np.float32)
    if a.dtype == np.int64:
        return a.astype(np.int32)
    return a


def _bin(a, b):
    if isinstance(a, (np.ndarray, np.generic)):
        return a, _t(b, a)
    if isinstance(b, (np.ndarray, np.generic)):
        return _t(a, b), b
    return _t(a), _t(b)


def cast(x, dtype): return np.asarray(x).astype(dtype)
def constant(x, dtype=None): return _t(x) if dtype is None else np.asarray(x, dtype)
def stop_gradient(x): return x
def unstack(x, axis=0): return [np.take(x, i, axis) for i in np.arange(np.shape(x)[axis])]
def stack(xs, axis=0): return np.stack([_t(v) for v in xs], axis)
def concat(xs, axis): return np.concatenate([_t(v) for v in xs], axis)
def maximum(a, b): return np.maximum(*_bin(a, b))
def minimum(a, b): return np.minimum(*_bin(a, b))
def multiply(a, b): return np.multiply(*_bin(a, b))
def divide(a, b): return np.divide(*_bin(a, b))
def truediv(a, b): return np.divide(*_bin(a, b))
def equal(a, b): return np.equal(*_bin(a, b))
def sqrt(x): return np.sqrt(_t(x))
def shape(x): return np.asarray(np.shape(x), np.int32)
def rank(x): return np.int32(np.ndim(x))
def size(x): return np.int32(np.size(x))
def expand_dims(x, axis): return np.expand_dims(_t(x), axis)
def squeeze(x, axis=None): return np.squeeze(x, tuple(axis) if axis is not None else None)
def reshape(x, shp): return np.reshape(x, [int(v) for v in np.atleast_1d(shp)])
def tile(x, multiples): return np.tile(x, [int(v) for v in np.atleast_1d(multiples)])
def transpose(x): return np.transpose(x)
def zeros(shp, dtype=np.float32): return np.zeros([int(v) for v in np.atleast_1d(shp)], dtype)
def zeros_like(x): return np.zeros_like(x)
def repeat(x, repeats): return np.repeat(x, int(repeats))
def gather(params, indices, axis=0): return np.take(params, np.asarray(indices), axis=axis)
def gather_nd(params, indices): return np.asarray(params)[tuple(np.asarray(indices).T)]
def reduce_sum(x, axis=None): return np.sum(x, axis=axis, dtype=np.asarray(x).dtype)
def reduce_max(x, axis=None): return np.max(x, axis=axis)
def reduce_mean(x, axis=None): return np.mean(x, axis=axis, dtype=np.asarray(x).dtype)


def range(*args, dtype=None):  # noqa: A001
    r = np.arange(*[int(a) for a in args])
    return r.astype(dtype if dtype is not None else np.int32)


def meshgrid(a, b): return [np.ascontiguousarray(m) for m in np.meshgrid(_t(a), _t(b))]  # default indexing="xy", as TF


def split(x, num_or_size_splits, axis=0):
    if np.ndim(num_or_size_splits) == 0:
        return np.split(x, int(num_or_size_splits), axis)
    return np.split(x, np.cumsum(np.asarray(num_or_size_splits))[:-1], axis)


def where(cond, x=None, y=None):
    if x is None:
        return np.argwhere(cond).astype(np.int64)  # [n, rank], row-major order, as tf.where
    return np.where(cond, x, y)


def one_hot(indices, depth, on_value=None, off_value=None, dtype=None):
    idx = np.asarray(indices)
    if dtype is None:
        if on_value is None:
            dtype = np.float32
        elif isinstance(on_value, (type(True), np.bool_)):
            dtype = np.bool_
        elif isinstance(on_value, int):
            dtype = np.int32
        elif isinstance(on_value, float):
            dtype = np.float32
        else:
            dtype = np.asarray(on_value).dtype
    on = np.asarray(1 if on_value is None else on_value).astype(dtype)
    off = np.asarray(0 if off_value is None else off_value).astype(dtype)
    hit = idx[..., None] == np.arange(int(depth))
    return np.where(hit, on, off).astype(dtype)


math = types.SimpleNamespace(
    exp=lambda x: np.exp(_t(x)), log=lambda x: np.log(_t(x)), argmax=lambda x, axis=None: np.argmax(x, axis=axis).astype(np.int64),
    round=lambda x: np.round(x))


def _combined_nms(boxes, scores, max_output_size_per_class, max_total_size, iou_threshold=0.5, score_threshold=float("-inf"),
                  pad_per_class=False, clip_boxes=True):
    assert not pad_per_class and clip_boxes
    ob, os_, oc, ov, _ = _rnp.combined_non_max_suppression(boxes, scores, int(max_output_size_per_class), int(max_total_size),
                                                           iou_threshold, score_threshold)
    return ob, os_, oc.astype(np.float32), ov  # TF returns the classes as float32; the reference casts them (:61)


def _crop_and_resize(image, boxes, box_indices, crop_size, method="bilinear", extrapolation_value=0.0, name=None):
    assert method == "bilinear" and extrapolation_value == 0.0
    return _rnp.crop_and_resize(image, boxes, np.asarray(box_indices), tuple(int(v) for v in crop_size))


image = types.SimpleNamespace(combined_non_max_suppression=_combined_nms, crop_and_resize=_crop_and_resize)


class _Layer:
    """tf.keras.layers.Layer / tf.keras.Model / Metric / Loss base: stores nothing, calls ``call``."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


class _MaxPool2D(_Layer):
    def __init__(self, pool_size, name=None):
        self.k = int(pool_size)

    def call(self, x):  # strides = pool_size, padding "valid" (the Keras defaults)
        n, h, w, c = x.shape
        k = self.k
        return x[:, :h // k * k, :w // k * k].reshape(n, h // k, k, w // k, k, c).max(axis=(2, 4))


class _Flatten(_Layer):
    def __init__(self, name=None):
        pass

    def call(self, x):
        return x.reshape(x.shape[0], -1)


class _Anything:
    """Constructors the detectors call in __init__ (Conv2D, Dense, initializers, ...): never executed on this path."""

    def __init__(self, *args, **kwargs):
        pass

    def __getattr__(self, name):
        return _Anything


_layers = types.SimpleNamespace(Layer=_Layer, MaxPool2D=_MaxPool2D, Flatten=_Flatten, Conv2D=_Anything, Dense=_Anything,
                                Reshape=_Anything, Softmax=_Anything)
keras = types.SimpleNamespace(Model=_Layer, layers=_layers, metrics=types.SimpleNamespace(Metric=_Layer),
                              losses=types.SimpleNamespace(Loss=_Layer, Huber=_Anything, CategoricalCrossentropy=_Anything,
                                                           Reduction=types.SimpleNamespace(NONE="none")),
                              initializers=_Anything(), regularizers=_Anything())
VariableSynchronization = types.SimpleNamespace(ON_READ="on_read")
Variable = _Anything
TensorShape = _Anything
__version__ = "2.numpy-shim"


def install():
    """Register this module as ``tensorflow`` (only if the real one is absent)."""
    sys.modules.setdefault("tensorflow", sys.modules[__name__])
