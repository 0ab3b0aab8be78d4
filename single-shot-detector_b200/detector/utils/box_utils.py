"""Box helpers with the reference's interface (detector/utils/box_utils.py), computed by csrc/box_ops.cu.
All boxes are [ymin, xmin, ymax, xmax], normalised to [0, 1]."""
import torch

from ... import _lib
from ..._tensors import Call, ptr


def area(boxes):
    """reference :53-61.  [N,4] -> [N]."""
    call = Call()
    b = call.tensor(boxes, torch.float32, (-1, 4))
    out = call.empty([b.shape[0]], torch.float32)
    _lib.check(_lib.load().ssdk_area(call.ctx(), ptr(b), b.shape[0], ptr(out)))
    return call.result(out)


def _pairwise(fn_name, boxes1, boxes2):
    call = Call()
    b1 = call.tensor(boxes1, torch.float32, (-1, 4))
    b2 = call.tensor(boxes2, torch.float32, (-1, 4))
    out = call.empty([b1.shape[0], b2.shape[0]], torch.float32)
    _lib.check(getattr(_lib.load(), fn_name)(call.ctx(), ptr(b1), b1.shape[0], ptr(b2), b2.shape[0], ptr(out)))
    return call.result(out)


def intersection(boxes1, boxes2):
    """reference :30-50.  [N,4], [M,4] -> [N,M]."""
    return _pairwise('ssdk_intersection', boxes1, boxes2)


def iou(boxes1, boxes2):
    """reference :14-27.  [N,4], [M,4] -> [N,M]."""
    return _pairwise('ssdk_iou', boxes1, boxes2)


def _coder(fn_name, x, anchors):
    call = Call()
    a = call.tensor(x, torch.float32, (-1, 4))
    b = call.tensor(anchors, torch.float32, (-1, 4))
    assert a.shape[0] == b.shape[0]
    out = call.empty([a.shape[0], 4], torch.float32)
    _lib.check(getattr(_lib.load(), fn_name)(call.ctx(), ptr(a), ptr(b), a.shape[0], ptr(out)))
    return call.result(out)


def encode(boxes, anchors):
    """reference :80-111.  [N,4], [N,4] -> codes [N,4] = [ty, tx, th, tw]."""
    return _coder('ssdk_encode', boxes, anchors)


def decode(codes, anchors):
    """reference :114-142.  [N,4], [N,4] -> boxes [N,4]."""
    return _coder('ssdk_decode', codes, anchors)


def batch_decode(box_encodings, anchors):
    """reference :145-173.  [B,A,4], [A,4] -> clipped boxes [B,A,4]."""
    call = Call()
    e = call.tensor(box_encodings, torch.float32)
    B, A = e.shape[0], e.shape[1]
    a = call.tensor(anchors, torch.float32, (A, 4))
    out = call.empty([B, A, 4], torch.float32)
    _lib.check(_lib.load().ssdk_batch_decode(call.ctx(), ptr(e), ptr(a), B, A, ptr(out)))
    return call.result(out)
