"""
TEST INFRASTRUCTURE ONLY -- an eager, NumPy-float32 stand-in for the handful of
TensorFlow 1.x symbols that the reference's per-anchor hot path touches.

Why it exists: TensorFlow ("tensorflow 1.12", reference README.md:22) is not
installed in the build container and cannot be installed (no network).  With this
directory put FIRST on sys.path, the reference's own, unmodified source files
(/root/reference/detector/{anchor_generator,training_target_creation,losses,ssd}.py,
detector/utils/{box_utils,nms}.py) import and execute line by line, so the op
ORDER, temporaries, thresholds and quirks of the reference are exercised exactly
as written.  What it does NOT pin is the arithmetic inside TensorFlow's own C++
kernels: every op here is the IEEE float32 NumPy equivalent of the TF op's
documented semantics (first-max argmax, ascending tf.where, order preserving
boolean_mask, ...), transcendental functions are NumPy's (last-ulp differences
from Eigen are possible), and tf.image.non_max_suppression is a scalar
restatement of TF 1.12's NonMaxSuppressionV3 CPU kernel (see image.py).

It is used only by tests/golden/make_golden.py (run in the build container, where
/root/reference exists) to produce the committed fixtures under tests/golden/.
Nothing in the product imports it.
"""
import builtins as _b
import contextlib

import numpy as np

from . import image, nn, summary  # noqa: F401  (tf.image / tf.nn / tf.summary)
from ._tensor import Tensor, _arr, _wrap

float32 = np.float32
int32 = np.int32
int64 = np.int64
bool = np.bool_  # noqa: A001


@contextlib.contextmanager
def name_scope(name):
    yield


variable_scope = name_scope


def constant(value, dtype=None):
    return _wrap(np.asarray(_arr(value), dtype=dtype))


def convert_to_tensor(value, dtype=None):
    return constant(value, dtype)


def to_float(x):
    return _wrap(np.asarray(_arr(x)).astype(np.float32))


def to_int32(x):
    return _wrap(np.asarray(_arr(x)).astype(np.int32))


def cast(x, dtype):
    return _wrap(np.asarray(_arr(x)).astype(dtype))


def ceil(x):
    return _wrap(np.ceil(_arr(x)))


def sqrt(x):
    return _wrap(np.sqrt(_arr(x)))


def exp(x):
    return _wrap(np.exp(_arr(x)))


def log(x):
    return _wrap(np.log(_arr(x)))


def log1p(x):
    return _wrap(np.log1p(_arr(x)))


def abs(x):  # noqa: A001
    return _wrap(np.abs(_arr(x)))


def square(x):
    a = _arr(x)
    return _wrap(a * a)


def pow(x, y):  # noqa: A001
    a = _arr(x)
    return _wrap(np.power(a, np.asarray(_arr(y), dtype=a.dtype)))


def sigmoid(x):
    # Generic Eigen scalar_logistic_op: 1 / (1 + exp(-x)), evaluated in float32.
    a = _arr(x)
    one = np.float32(1.0)
    return _wrap(one / (one + np.exp(-a)))


def _binary(fn):
    def op(x, y):
        xa, ya = _arr(x), _arr(y)
        if not isinstance(xa, np.ndarray) and isinstance(ya, np.ndarray):
            xa = np.asarray(xa, dtype=ya.dtype)
        elif isinstance(xa, np.ndarray) and not isinstance(ya, np.ndarray):
            ya = np.asarray(ya, dtype=xa.dtype)
        return _wrap(fn(xa, ya))
    return op


minimum = _binary(np.minimum)
maximum = _binary(np.maximum)
divide = _binary(np.true_divide)
greater = _binary(np.greater)
greater_equal = _binary(np.greater_equal)
less = _binary(np.less)
equal = _binary(np.equal)
less_equal = _binary(np.less_equal)
add = _binary(np.add)


def clip_by_value(x, lo, hi):
    a = _arr(x)
    # TF: minimum(maximum(x, lo), hi)  (clip_ops.py) -- same result as np.clip for lo<=hi.
    return _wrap(np.minimum(np.maximum(a, a.dtype.type(lo)), a.dtype.type(hi)))


def range(n):  # noqa: A001
    return _wrap(np.arange(int(_arr(n)), dtype=np.int32))


def meshgrid(x, y):
    xx, yy = np.meshgrid(_arr(x), _arr(y))  # 'xy' indexing in both libraries
    return _wrap(xx), _wrap(yy)


def stack(values, axis=0):
    return _wrap(np.stack([np.asarray(_arr(v)) for v in values], axis=axis))


def unstack(x, axis=0):
    a = _arr(x)
    return [_wrap(np.take(a, i, axis=axis)) for i in _b.range(a.shape[axis])]


def split(x, num_or_size_splits, axis=0):
    return [_wrap(p) for p in np.split(_arr(x), num_or_size_splits, axis=axis)]


def concat(values, axis):
    return _wrap(np.concatenate([_arr(v) for v in values], axis=axis))


def expand_dims(x, axis):
    return _wrap(np.expand_dims(_arr(x), axis))


def squeeze(x, axis=None):
    return _wrap(np.squeeze(_arr(x), axis=axis))


def _ints(seq):
    return [int(_arr(s)) for s in seq]


def tile(x, multiples):
    return _wrap(np.tile(_arr(x), _ints(multiples)))


def reshape(x, shape):
    return _wrap(np.reshape(_arr(x), _ints(shape)))


def transpose(x, perm=None):
    return _wrap(np.transpose(_arr(x), perm))


def shape(x):
    return _wrap(np.asarray(np.shape(_arr(x)), dtype=np.int32))


def size(x):
    return _wrap(np.asarray(np.size(_arr(x)), dtype=np.int32))


def fill(dims, value):
    dt = np.int32 if isinstance(value, int) else np.float32
    return _wrap(np.full(_ints(dims), value, dtype=dt))


def zeros(shape, dtype=np.float32):  # noqa: A002
    return _wrap(np.zeros(_ints(shape), dtype=dtype))


def ones_like(x):
    return _wrap(np.ones_like(_arr(x)))


def zeros_like(x):
    return _wrap(np.zeros_like(_arr(x)))


def identity(x, name=None):
    return x


def cond(pred, true_fn, false_fn):
    return true_fn() if _b.bool(_arr(pred)) else false_fn()


def argmax(x, axis, output_type=np.int64):
    # np.argmax, like tf.argmax, returns the FIRST index of the maximum.
    return _wrap(np.argmax(_arr(x), axis=axis).astype(output_type))


def _axes(axis):
    if isinstance(axis, (list, tuple)):
        return tuple(axis)
    return axis


def reduce_max(x, axis=None):
    return _wrap(np.max(_arr(x), axis=_axes(axis)))


def reduce_sum(x, axis=None):
    a = _arr(x)
    return _wrap(np.sum(a, axis=_axes(axis), dtype=a.dtype))


def reduce_mean(x, axis=None):
    a = _arr(x)
    return _wrap(np.mean(a, axis=_axes(axis), dtype=a.dtype))


def one_hot(indices, depth, dtype=np.float32, axis=-1):
    idx = np.asarray(_arr(indices))
    depth = int(_arr(depth))
    out = (idx[..., None] == np.arange(depth, dtype=idx.dtype)).astype(dtype)
    if axis not in (-1, out.ndim - 1):
        out = np.moveaxis(out, -1, axis)
    return _wrap(out)


def logical_not(x):
    return _wrap(np.logical_not(np.asarray(_arr(x))))


def reduce_any(x, axis=None):
    return _wrap(np.any(np.asarray(_arr(x)), axis=axis))


def where(condition, x=None, y=None):
    c = np.asarray(_arr(condition))
    if x is None:
        return _wrap(np.argwhere(c).astype(np.int64))  # ascending, shape [n, rank]
    return _wrap(np.where(c, _arr(x), _arr(y)))


def gather(params, indices):
    return _wrap(np.asarray(_arr(params))[np.asarray(_arr(indices))])


def boolean_mask(x, mask):
    return _wrap(np.asarray(_arr(x))[np.asarray(_arr(mask))])


def dynamic_stitch(indices, data):
    idx = [np.asarray(_arr(i)) for i in indices]
    dat = [np.asarray(_arr(d)) for d in data]
    n = max([int(i.max()) + 1 for i in idx if i.size] + [0])
    out = np.zeros((n,) + dat[0].shape[1:], dtype=dat[0].dtype)
    for i, d in zip(idx, dat):
        out[i] = d
    return _wrap(out)


def pad(x, paddings):
    pw = [tuple(int(_arr(v)) for v in p) for p in paddings]
    return _wrap(np.pad(_arr(x), pw))


def map_fn(fn, elems, dtype=None, parallel_iterations=None, back_prop=None,
           swap_memory=None, infer_shape=None):
    arrs = [np.asarray(_arr(e)) for e in elems]
    outs = [fn([_wrap(a[i]) for a in arrs]) for i in _b.range(arrs[0].shape[0])]
    return tuple(_wrap(np.stack([np.asarray(_arr(o[k])) for o in outs], axis=0))
                 for k in _b.range(len(outs[0])))
