"""Oracle gradients of the SSD loss (test infrastructure; see oracle/__init__.py).

The reference never writes a backward pass: model.py:115-118 asks TensorFlow to differentiate
    total_loss = localization_loss_weight * localization_loss + classification_loss_weight * classification_loss
(model.py:86-91) through detector/ssd.py:89-133 and detector/losses.py:4-50, with targets, matches and weights
constant (ssd.py:197 back_prop=False; losses.py:30,33,46 stop_gradient).  TensorFlow cannot run here, so this file
states the derivative of exactly those forward formulas in float64 (closed form), and
tests/test_oracle_golden.py::test_gradient_oracle_* pin it two ways: `forward64` below must agree with the float32
forward oracle (which is pinned to the reference's own code through the golden fixtures), and the closed-form
gradient must agree with central finite differences of `forward64`.  PARITY UNPINNED only in the sense that TF's
autodiff rounding (float32 op by op) is not reproduced; the GPU tolerance is 1e-5 relative.

tf.where routes the gradient to the selected branch, d|x|/dx = sign(x) (0 at 0), tf.pow(x, 2) -> 2x:
    smooth-L1 (losses.py:16-19):  d/dp = d if |d| < 1 else sign(d),  d = p - t
    focal (losses.py:34-50), p = sigmoid(x), sp = softplus(x):
        z = 0:  (1-alpha) p^gamma     [ gamma (1-p) sp + p ]
        z = 1:   alpha   (1-p)^gamma  [ gamma p log p - (1-p) ]
"""
import numpy as np

from .constants import NEGATIVES_THRESHOLD, POSITIVES_THRESHOLD
from .ssd import create_targets_batch


def _parts(anchors, groundtruth, num_classes, positives_threshold, negatives_threshold):
    reg_t, cls_t, matches = create_targets_batch(anchors, groundtruth, positives_threshold, negatives_threshold)
    onehot = (cls_t[:, :, None] == np.arange(1, num_classes + 1, dtype=np.int32))            # ssd.py:96-100
    matched = (matches >= 0).astype(np.float64)                                              # ssd.py:89
    not_ignore = (matches >= -1).astype(np.float64)                                          # ssd.py:103
    return reg_t.astype(np.float64), onehot, matched, not_ignore


def forward64(anchors, encoded_boxes, class_predictions, groundtruth, params, num_classes, num_matches=None,
              positives_threshold=POSITIVES_THRESHOLD, negatives_threshold=NEGATIVES_THRESHOLD, parts=None):
    """(localization_loss, classification_loss) in float64; `num_matches` overrides the local count (sharded batches)."""
    reg_t, onehot, matched, not_ignore = parts or _parts(anchors, groundtruth, num_classes, positives_threshold, negatives_threshold)
    gamma, alpha = float(params['gamma']), float(params['alpha'])
    x = np.asarray(class_predictions, np.float64)
    d = np.abs(np.asarray(encoded_boxes, np.float64) - reg_t)
    loc = (np.where(d < 1.0, 0.5 * d * d, d - 0.5).sum(axis=2) * matched).sum()
    p = 1.0 / (1.0 + np.exp(-x))
    nlpt = np.where(onehot, np.logaddexp(0.0, -x), np.logaddexp(0.0, x))
    p_t = np.where(onehot, p, 1.0 - p)
    a_t = np.where(onehot, alpha, 1.0 - alpha)
    cls = ((np.power(1.0 - p_t, gamma) * a_t * nlpt).sum(axis=2) * not_ignore).sum()
    n = max(float(matched.sum()) if num_matches is None else float(num_matches), 1.0)        # ssd.py:121-123
    return loc / n, cls / n


def ssd_loss_grad(anchors, encoded_boxes, class_predictions, groundtruth, params, num_classes, upstream=(1.0, 1.0),
                  num_matches=None, positives_threshold=POSITIVES_THRESHOLD, negatives_threshold=NEGATIVES_THRESHOLD, parts=None):
    """d(upstream[0] * localization_loss + upstream[1] * classification_loss) / d(class_predictions, encoded_boxes)."""
    reg_t, onehot, matched, not_ignore = parts or _parts(anchors, groundtruth, num_classes, positives_threshold, negatives_threshold)
    gamma, alpha = float(params['gamma']), float(params['alpha'])
    n = max(float(matched.sum()) if num_matches is None else float(num_matches), 1.0)
    x = np.asarray(class_predictions, np.float64)
    p = 1.0 / (1.0 + np.exp(-x))
    q = 1.0 / (1.0 + np.exp(x))                                   # 1 - p
    sp = np.logaddexp(0.0, x)
    logp = -np.logaddexp(0.0, -x)
    with np.errstate(invalid='ignore', divide='ignore'):
        g_neg = (1.0 - alpha) * np.power(p, gamma) * (gamma * q * sp + p)
        g_pos = alpha * np.power(q, gamma) * (gamma * p * logp - q)
    g_logits = np.where(onehot, g_pos, g_neg) * not_ignore[:, :, None] * (float(upstream[1]) / n)
    d = np.asarray(encoded_boxes, np.float64) - reg_t
    g_codes = np.where(np.abs(d) < 1.0, d, np.sign(d)) * matched[:, :, None] * (float(upstream[0]) / n)
    return {'class_predictions': g_logits, 'encoded_boxes': g_codes, 'num_matches': matched.sum()}
