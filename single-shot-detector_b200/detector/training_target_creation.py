"""Target assignment with the reference's interface (detector/training_target_creation.py), computed by
csrc/matcher.cu.  The single-image functions mirror the reference one to one; batch_training_targets is the
batched form that SSD._create_targets uses (one launch for all images instead of tf.map_fn)."""
import torch

from .. import _lib
from .._tensors import Call, ptr


def get_training_targets(anchors, groundtruth_boxes, groundtruth_labels,
                         positives_threshold=0.5, negatives_threshold=0.4):
    """reference :5-45.  Returns reg_targets [A,4] f32, cls_targets [A] i32, matches [A] i32."""
    call = Call()
    a = call.tensor(anchors, torch.float32, (-1, 4))
    gt = call.tensor(groundtruth_boxes, torch.float32, (-1, 4))
    lab = call.tensor(groundtruth_labels, torch.int32, (-1,))
    A, N = a.shape[0], gt.shape[0]
    reg = call.empty([A, 4], torch.float32)
    cls = call.empty([A], torch.int32)
    matches = call.empty([A], torch.int32)
    _lib.check(_lib.load().ssdk_training_targets(
        call.ctx(), ptr(a), A, ptr(gt), ptr(lab), None, 1, N, float(positives_threshold),
        float(negatives_threshold), ptr(reg), ptr(cls), ptr(matches)))
    return call.result(reg, cls, matches)


def match_boxes(anchors, groundtruth_boxes, positives_threshold=0.5, negatives_threshold=0.4,
                force_match_groundtruth=True):
    """reference :48-130.  Returns matches [A] i32 in {-2, -1, 0..N-1}."""
    assert positives_threshold >= negatives_threshold                                       # :86
    call = Call()
    a = call.tensor(anchors, torch.float32, (-1, 4))
    gt = call.tensor(groundtruth_boxes, torch.float32, (-1, 4))
    A, N = a.shape[0], gt.shape[0]
    matches = call.empty([A], torch.int32)
    _lib.check(_lib.load().ssdk_match_boxes(
        call.ctx(), ptr(a), A, ptr(gt), None, 1, N, float(positives_threshold), float(negatives_threshold),
        1 if force_match_groundtruth else 0, ptr(matches)))
    return call.result(matches)


def create_targets(anchors, groundtruth_boxes, groundtruth_labels, matches):
    """reference :133-176.  Returns reg_targets [A,4], cls_targets [A]."""
    call = Call()
    a = call.tensor(anchors, torch.float32, (-1, 4))
    gt = call.tensor(groundtruth_boxes, torch.float32, (-1, 4))
    lab = call.tensor(groundtruth_labels, torch.int32, (-1,))
    m = call.tensor(matches, torch.int32, (-1,))
    A, N = a.shape[0], gt.shape[0]
    reg = call.empty([A, 4], torch.float32)
    cls = call.empty([A], torch.int32)
    _lib.check(_lib.load().ssdk_create_targets(call.ctx(), ptr(a), A, ptr(gt), ptr(lab), 1, N, ptr(m), ptr(reg), ptr(cls)))
    return call.result(reg, cls)


def batch_training_targets(anchors, boxes, labels, num_boxes, positives_threshold=0.5, negatives_threshold=0.4):
    """All images of a batch at once: boxes [B,Gmax,4], labels [B,Gmax], num_boxes [B] (padded format of the
    input pipeline, reference pipeline.py:61-62; slicing [:num_boxes] as ssd.py:183)."""
    call = Call()
    a = call.tensor(anchors, torch.float32, (-1, 4))
    gt = call.tensor(boxes, torch.float32)
    B, G = gt.shape[0], gt.shape[1]
    lab = call.tensor(labels, torch.int32, (B, G))
    num = call.tensor(num_boxes, torch.int32, (B,))
    A = a.shape[0]
    reg = call.empty([B, A, 4], torch.float32)
    cls = call.empty([B, A], torch.int32)
    matches = call.empty([B, A], torch.int32)
    _lib.check(_lib.load().ssdk_training_targets(
        call.ctx(), ptr(a), A, ptr(gt), ptr(lab), ptr(num), B, G, float(positives_threshold),
        float(negatives_threshold), ptr(reg), ptr(cls), ptr(matches)))
    return call.result(reg, cls, matches)
