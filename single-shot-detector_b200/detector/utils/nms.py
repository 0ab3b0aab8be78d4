"""Post-processing with the reference's interface (detector/utils/nms.py), computed by csrc/postprocess.cu."""
import torch

from ... import _lib
from ..._tensors import Call, ptr


def _postprocess(call, codes, anchors, scores, flags, score_threshold, iou_threshold, max_boxes_per_class,
                 return_anchor_indices=False, box_scaler=None, final_score_threshold=None):
    B, A, C = scores.shape
    K = int(max_boxes_per_class)
    M = C * K
    boxes = call.empty([B, M, 4], torch.float32)
    out_scores = call.empty([B, M], torch.float32)
    classes = call.empty([B, M], torch.int32)
    num = call.empty([B], torch.int32)
    aidx = call.empty([B, M], torch.int32) if return_anchor_indices else None
    if box_scaler is not None or final_score_threshold is not None:
        # the consumers that follow get_predictions in the reference (model.py:67-68, inference/detector.py:54-58)
        assert not return_anchor_indices
        sc = None if box_scaler is None else call.tensor(box_scaler, torch.float32, (B, 4))
        thr2 = float('-inf') if final_score_threshold is None else float(final_score_threshold)
        _lib.check(_lib.load().ssdk_detect(
            call.ctx(), ptr(codes), ptr(anchors), ptr(scores), flags, B, A, C, float(score_threshold),
            float(iou_threshold), K, ptr(sc), thr2, ptr(boxes), ptr(out_scores), ptr(classes), ptr(num)))
        return boxes, out_scores, classes, num
    _lib.check(_lib.load().ssdk_postprocess(
        call.ctx(), ptr(codes), ptr(anchors), ptr(scores), flags, B, A, C, float(score_threshold),
        float(iou_threshold), K, ptr(boxes), ptr(out_scores), ptr(classes), ptr(num), ptr(aidx)))
    if return_anchor_indices:
        return boxes, out_scores, classes, num, aidx
    return boxes, out_scores, classes, num


def multiclass_non_max_suppression(boxes, scores, score_threshold, iou_threshold, max_boxes_per_class, return_indices=False):
    """reference :6-45.  boxes [N,4] (already decoded), scores [N,C] ->
    selected_boxes [N',4], selected_scores [N'], selected_classes [N'] (class-major, score-descending).
    return_indices=True appends the selected row indices [N'] (what tf.image.non_max_suppression returns per class, :33)."""
    call = Call()
    s = call.tensor(scores, torch.float32)
    N, C = s.shape
    b = call.tensor(boxes, torch.float32, (1, N, 4))
    out = _postprocess(call, b, None, s.reshape(1, N, C), _lib.SSDK_INPUT_SCORES | _lib.SSDK_BOXES_DECODED,
                       score_threshold, iou_threshold, max_boxes_per_class, return_anchor_indices=return_indices)
    ob, os_, oc, on = out[:4]
    n = int(on[0].item())          # N' is data dependent: one host read, as tf.shape() would need
    if return_indices:
        return call.result(ob[0, :n], os_[0, :n], oc[0, :n], out[4][0, :n])
    return call.result(ob[0, :n], os_[0, :n], oc[0, :n])


def batch_multiclass_non_max_suppression(encoded_boxes, anchors, scores, score_threshold, iou_threshold,
                                         max_boxes_per_class, scores_are_logits=False, return_anchor_indices=False,
                                         box_scaler=None, final_score_threshold=None):
    """reference :48-102.  encoded_boxes [B,N,4], anchors [N,4], scores [B,N,C] ->
    boxes [B,N',4], scores [B,N'], classes [B,N'], num_detections [B], N' = C * max_boxes_per_class.
    scores_are_logits=True fuses the sigmoid of ssd.py:60 into the streaming pass."""
    call = Call()
    s = call.tensor(scores, torch.float32)
    B, A, C = s.shape
    e = call.tensor(encoded_boxes, torch.float32, (B, A, 4))
    a = call.tensor(anchors, torch.float32, (A, 4))
    flags = (_lib.SSDK_INPUT_LOGITS if scores_are_logits else _lib.SSDK_INPUT_SCORES) | _lib.SSDK_BOXES_ENCODED
    return call.result(*_postprocess(call, e, a, s, flags, score_threshold, iou_threshold, max_boxes_per_class,
                                     return_anchor_indices, box_scaler, final_score_threshold))
