"""Eager tensor wrapper for the TF stand-in (test infrastructure only; see __init__.py)."""
import numpy as np


class Dim(int):
    """tf.Dimension: `tensor.shape[i].value` is used at reference nms.py:65."""

    @property
    def value(self):
        return int(self)


def _arr(x):
    """Unwrap to ndarray / python scalar."""
    if isinstance(x, Tensor):
        return x.a
    if isinstance(x, (list, tuple)) and any(isinstance(v, Tensor) for v in x):
        return np.stack([np.asarray(_arr(v)) for v in x])
    return x


def _wrap(a):
    return Tensor(np.asarray(a))


def _coerce(other, like):
    """Python scalars take the tensor's dtype, as tf.convert_to_tensor does in binary ops."""
    o = _arr(other)
    if isinstance(o, np.ndarray):
        return o
    return np.asarray(o, dtype=like.dtype)


class Tensor:
    __array_priority__ = 1000

    def __init__(self, a):
        self.a = a

    # -- introspection ------------------------------------------------------
    @property
    def shape(self):
        return tuple(Dim(d) for d in self.a.shape)

    @property
    def dtype(self):
        return self.a.dtype

    def set_shape(self, shape):
        assert tuple(int(s) for s in shape) == tuple(self.a.shape), (shape, self.a.shape)

    def numpy(self):
        return self.a

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    def __index__(self):
        return int(self.a)

    def __int__(self):
        return int(self.a)

    def __float__(self):
        return float(self.a)

    def __bool__(self):
        return bool(self.a)

    def __len__(self):
        return len(self.a)

    def __repr__(self):
        return 'ShimTensor(%r)' % (self.a,)

    def __getitem__(self, key):
        def fix(k):
            if isinstance(k, Tensor):
                return int(k.a)
            if isinstance(k, slice):
                return slice(*[None if v is None else int(_arr(v)) for v in (k.start, k.stop, k.step)])
            return k
        key = tuple(fix(k) for k in key) if isinstance(key, tuple) else fix(key)
        return _wrap(self.a[key])

    # -- arithmetic (result dtype = tensor dtype; no silent float64 promotion) --
    def _bin(self, other, fn, swap=False):
        o = _coerce(other, self.a)
        x, y = (o, self.a) if swap else (self.a, o)
        r = fn(x, y)
        if self.a.dtype == np.float32 and isinstance(r, np.ndarray) and r.dtype == np.float64:
            raise TypeError('float64 promotion in shim op')
        return _wrap(r)

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.true_divide)
    def __rtruediv__(self, o): return self._bin(o, np.true_divide, True)
    def __neg__(self): return _wrap(-self.a)
    def __ge__(self, o): return self._bin(o, np.greater_equal)
    def __gt__(self, o): return self._bin(o, np.greater)
    def __le__(self, o): return self._bin(o, np.less_equal)
    def __lt__(self, o): return self._bin(o, np.less)
