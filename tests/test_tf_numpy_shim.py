"""The NumPy stand-in for TensorFlow (tests/golden/tf_numpy_shim.py) follows TensorFlow's conversion and indexing rules
for the symbols the reference's hot path uses.  No reference code is touched here."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import tf_numpy_shim as tf  # noqa: E402

F = np.float32


def test_python_operands_take_the_tensor_dtype():
    a = F([[1, 2, 3, 4], [0, 0, 2, 2]])
    for out in (tf.multiply(a, [2, 2, 2, 2]), tf.divide(a, [128, 80, 128, 80]), tf.maximum(0.0, a), tf.minimum(a, 3),
                tf.truediv(a, a + 1), tf.math.exp(a), tf.math.log(a + 1), tf.sqrt([0.5, 1.0, 2.0]), tf.reduce_sum(a, -1),
                tf.range(3, dtype=tf.float32) * 16, tf.zeros([2, 3]), tf.one_hot(tf.zeros(4, dtype=tf.int32), 3)):
        assert np.asarray(out).dtype == np.float32
    s, r = tf.meshgrid([0.5, 1.0], [1.0, 2.0, 3.0])  # indexing "xy": [len(b), len(a)]
    assert s.shape == (3, 2) and s.dtype == np.float32 and (s[:, 1] == 1.0).all() and (r[2] == 3.0).all()


def test_structure_ops():
    a = F([[1, 2, 3, 4], [0, 0, 0, 0], [5, 6, 7, 8]])
    lo, hi = tf.split(a, num_or_size_splits=2, axis=-1)
    assert lo.shape == hi.shape == (3, 2)
    parts = tf.split(np.arange(6), [2, 2, 2])
    assert [p.tolist() for p in parts] == [[0, 1], [2, 3], [4, 5]]
    idx = tf.where(tf.reduce_sum(a, -1) != 0.0)
    assert idx.dtype == np.int64 and idx.tolist() == [[0], [2]]
    np.testing.assert_array_equal(tf.gather_nd(a, idx), a[[0, 2]])
    np.testing.assert_array_equal(tf.gather(a, [1, 0, 3, 2], axis=-1)[0], [2, 1, 4, 3])
    assert tf.where(a[:, 0] == 5)[0][0] == 2
    np.testing.assert_array_equal(tf.where(a > 2, a, tf.zeros_like(a))[0], [0, 0, 3, 4])
    assert tf.shape(a).tolist() == [3, 4] and int(tf.rank(a)) == 2 and int(tf.size(a)) == 12
    assert tf.tile(tf.expand_dims(a[:, 0] > 0, 1), [1, 3]).shape == (3, 3)
    assert tf.squeeze(a[:, :1], [-1]).shape == (3,) and tf.transpose(a[:, :1]).shape == (1, 3)
    assert [v.tolist() for v in tf.unstack(tf.cast([1, 2], dtype=tf.float32))] == [1.0, 2.0]
    assert tf.repeat(tf.range(2), 3).tolist() == [0, 0, 0, 1, 1, 1]
    assert tf.math.argmax(F([[1, 3, 3], [2, 1, 0]]), axis=-1).tolist() == [1, 0]  # first maximum


def test_one_hot_variants():
    assert tf.one_hot(np.int64(1), 3, on_value=True, off_value=False).tolist() == [False, True, False]
    m = tf.one_hot(tf.rank(np.zeros((2, 5, 1, 4))) - 2, tf.rank(np.zeros((2, 5, 1, 4))), on_value=np.int32(7), off_value=1,
                   dtype=tf.int32)
    assert m.tolist() == [1, 1, 7, 1] and m.dtype == np.int32
    assert tf.one_hot(np.int32([0, 5]), 2).tolist() == [[1.0, 0.0], [0.0, 0.0]]  # out-of-range index -> all-zero row


def test_layers():
    x = np.arange(2 * 4 * 4 * 1, dtype=F).reshape(2, 4, 4, 1)
    y = tf.keras.layers.MaxPool2D(2, name="p")(x)
    assert y.shape == (2, 2, 2, 1) and y[0, 0, 0, 0] == 5 and y[0, 1, 1, 0] == 15
    assert tf.keras.layers.Flatten(name="f")(y).shape == (2, 4)

    class L(tf.keras.layers.Layer):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.d = tf.keras.layers.Dense(3, kernel_initializer=tf.keras.initializers.VarianceScaling(scale=1.0),
                                           kernel_regularizer=tf.keras.regularizers.l2(1e-4))

        def call(self, v):
            return v + 1

    assert L(name="l")(1) == 2
