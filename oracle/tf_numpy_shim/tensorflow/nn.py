"""tf.nn stand-ins (test infrastructure only; see __init__.py)."""
import numpy as np

from ._tensor import _arr, _wrap


def sigmoid_cross_entropy_with_logits(labels=None, logits=None):
    # tensorflow/python/ops/nn_impl.py (1.12):
    #   cond = logits >= 0; relu_logits = where(cond, logits, 0)
    #   neg_abs_logits = where(cond, -logits, logits)
    #   add(relu_logits - logits * labels, log1p(exp(neg_abs_logits)))
    x = _arr(logits)
    z = np.asarray(_arr(labels), dtype=x.dtype)
    cond = x >= 0
    zeros = np.zeros_like(x)
    relu_logits = np.where(cond, x, zeros)
    neg_abs_logits = np.where(cond, -x, x)
    return _wrap((relu_logits - x * z) + np.log1p(np.exp(neg_abs_logits)))


def relu(x):
    a = _arr(x)
    return _wrap(np.maximum(a, a.dtype.type(0)))


def top_k(x, k, sorted=True):  # noqa: A002  (summaries only, reference ssd.py:146)
    a = np.asarray(_arr(x))
    k = int(_arr(k))
    idx = np.argsort(-a, axis=-1, kind='stable')[..., :k]
    return _wrap(np.take_along_axis(a, idx, axis=-1)), _wrap(idx.astype(np.int32))
