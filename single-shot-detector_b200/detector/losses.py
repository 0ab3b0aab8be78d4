"""Per-anchor losses with the reference's interface (detector/losses.py), computed by csrc/box_ops.cu.
SSD.loss does not go through these two helpers: it uses the fused streaming kernel of csrc/loss.cu."""
import torch

from .. import _lib
from .._tensors import Call, ptr


def localization_loss(predictions, targets, weights):
    """reference :4-19.  [B,A,4], [B,A,4], [B,A] -> [B,A]."""
    call = Call()
    p = call.tensor(predictions, torch.float32)
    B, A = p.shape[0], p.shape[1]
    t = call.tensor(targets, torch.float32, (B, A, 4))
    w = call.tensor(weights, torch.float32, (B, A))
    out = call.empty([B, A], torch.float32)
    _lib.check(_lib.load().ssdk_localization_loss(call.ctx(), ptr(p), ptr(t), ptr(w), B, A, ptr(out)))
    return call.result(out)


def focal_loss(predictions, targets, weights, gamma=2.0, alpha=0.25):
    """reference :22-50.  logits [B,A,C], one-hot float targets [B,A,C], weights [B,A] -> [B,A]."""
    call = Call()
    x = call.tensor(predictions, torch.float32)
    B, A, C = x.shape
    z = call.tensor(targets, torch.float32, (B, A, C))
    w = call.tensor(weights, torch.float32, (B, A))
    out = call.empty([B, A], torch.float32)
    _lib.check(_lib.load().ssdk_focal_loss(call.ctx(), ptr(x), ptr(z), ptr(w), B, A, C, float(gamma), float(alpha), ptr(out)))
    return call.result(out)
