"""Oracle losses (test infrastructure; see oracle/__init__.py).

Follows reference detector/losses.py: localization_loss :4-19, focal_loss :22-50;
sigmoid_cross_entropy_with_logits as in tensorflow/python/ops/nn_impl.py (1.12)."""
import numpy as np

f32 = np.float32


def localization_loss(predictions, targets, weights):
    d = np.abs(predictions - targets)                               # :16
    lt1 = d < f32(1.0)                                              # :17
    loss = np.where(lt1, f32(0.5) * (d * d), d - f32(0.5))          # :18
    return weights * np.sum(loss, axis=2, dtype=np.float32)         # :19


def sigmoid_cross_entropy_with_logits(labels, logits):
    cond = logits >= 0
    relu_logits = np.where(cond, logits, f32(0))
    neg_abs = np.where(cond, -logits, logits)
    return (relu_logits - logits * labels) + np.log1p(np.exp(neg_abs))


def sigmoid(x):
    return f32(1.0) / (f32(1.0) + np.exp(-x))


def focal_loss(predictions, targets, weights, gamma=2.0, alpha=0.25):
    pos = targets == f32(1.0)                                       # :34
    nlpt = sigmoid_cross_entropy_with_logits(targets, predictions)  # :36
    p = sigmoid(predictions)                                        # :37
    p_t = np.where(pos, p, f32(1.0) - p)                            # :38
    mod = np.power(f32(1.0) - p_t, f32(gamma))                      # :41
    wl = np.where(pos, f32(alpha) * nlpt, f32(1.0 - alpha) * nlpt)  # :42-46
    fl = mod * wl                                                   # :47
    return weights * np.sum(fl, axis=2, dtype=np.float32)           # :50
