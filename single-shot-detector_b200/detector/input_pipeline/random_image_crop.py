"""Ground-truth side of the random-crop augmentation with the reference's interface
(detector/input_pipeline/random_image_crop.py: ioa :190, prune_completely_outside_window :102,
prune_non_overlapping_boxes :134, change_coordinate_frame :162), computed by csrc/crop_ops.cu, plus `crop_boxes`, the
batched form of the box part of randomly_crop_image (:86-99).  The crop window is an input: drawing it
(tf.image.sample_distorted_bounding_box) and decoding the JPEG belong to the input pipeline, which is out of scope."""
import torch

from ... import _lib
from ..._tensors import Call, ptr


def ioa(boxes1, boxes2):
    """reference :190-209.  [N,4], [M,4] -> [N,M] intersection over the area of boxes2 (not symmetric)."""
    call = Call()
    b1 = call.tensor(boxes1, torch.float32, (-1, 4))
    b2 = call.tensor(boxes2, torch.float32, (-1, 4))
    n, m = b1.shape[0], b2.shape[0]
    out = call.empty([n, m], torch.float32)
    _lib.check(_lib.load().ssdk_ioa(call.ctx(), ptr(b1), n, ptr(b2), m, ptr(out)))
    return call.result(out)


def change_coordinate_frame(boxes, window):
    """reference :162-187.  [N,4], [4] -> [N,4] relative to the window, clipped to [0,1]."""
    call = Call()
    b = call.tensor(boxes, torch.float32, (-1, 4))
    w = call.tensor(window, torch.float32, (4,))
    out = call.empty([b.shape[0], 4], torch.float32)
    _lib.check(_lib.load().ssdk_change_coordinate_frame(call.ctx(), ptr(b), b.shape[0], ptr(w), ptr(out)))
    return call.result(out)


def _pruned(call, out_boxes, out_idx, out_num):
    n = int(out_num[0].item())          # M_out is data dependent: one host read, as tf.where's output shape would need
    return call.result(out_boxes[:n], out_idx[:n].to(torch.int64))


def prune_completely_outside_window(boxes, window):
    """reference :102-131.  -> (boxes [M_out,4] not clipped, valid_indices [M_out] int64 ascending)."""
    call = Call()
    b = call.tensor(boxes, torch.float32, (-1, 4))
    w = call.tensor(window, torch.float32, (4,))
    n = b.shape[0]
    ob, oi, on = call.empty([n, 4], torch.float32), call.empty([n], torch.int32), call.empty([1], torch.int32)
    _lib.check(_lib.load().ssdk_prune_completely_outside_window(call.ctx(), ptr(b), n, ptr(w), ptr(ob), ptr(oi), ptr(on)))
    return _pruned(call, ob, oi, on)


def prune_non_overlapping_boxes(boxes1, boxes2, min_overlap):
    """reference :134-159.  Keeps the boxes of boxes1 whose IOA with at least one box of boxes2 is >= min_overlap."""
    call = Call()
    b1 = call.tensor(boxes1, torch.float32, (-1, 4))
    b2 = call.tensor(boxes2, torch.float32, (-1, 4))
    n = b1.shape[0]
    ob, oi, on = call.empty([n, 4], torch.float32), call.empty([n], torch.int32), call.empty([1], torch.int32)
    _lib.check(_lib.load().ssdk_prune_non_overlapping_boxes(call.ctx(), ptr(b1), n, ptr(b2), b2.shape[0], float(min_overlap),
                                                            ptr(ob), ptr(oi), ptr(on)))
    return _pruned(call, ob, oi, on)


def crop_boxes(boxes, num_boxes, windows, overlap_thresh=0.3):
    """The box part of randomly_crop_image (reference :86-99) for a batch in the pipeline's padded format
    (pipeline.py:61-62): boxes [B,Gmax,4], num_boxes [B] (or None), windows [B,4] -> (boxes in each window's frame
    [B,Gmax,4] zero padded, keep_indices [B,Gmax] int32 with -1 padding -- gather the labels with it, :36 --, new num_boxes [B]).
    One launch, no host synchronisation."""
    call = Call()
    b = call.tensor(boxes, torch.float32)
    B, G = b.shape[0], b.shape[1]
    w = call.tensor(windows, torch.float32, (B, 4))
    nb = None if num_boxes is None else call.tensor(num_boxes, torch.int32, (B,))
    ob, oi, on = call.empty([B, G, 4], torch.float32), call.empty([B, G], torch.int32), call.empty([B], torch.int32)
    _lib.check(_lib.load().ssdk_crop_boxes(call.ctx(), ptr(b), ptr(nb), ptr(w), B, G, float(overlap_thresh), ptr(ob), ptr(oi), ptr(on)))
    return call.result(ob, oi, on)
