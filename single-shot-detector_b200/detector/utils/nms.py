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
                                         box_scaler=None, final_score_threshold=None, phase=None):
    """reference :48-102.  encoded_boxes [B,N,4], anchors [N,4], scores [B,N,C] ->
    boxes [B,N',4], scores [B,N'], classes [B,N'], num_detections [B], N' = C * max_boxes_per_class.
    scores_are_logits=True fuses the sigmoid of ssd.py:60 into the streaming pass."""
    call = Call()
    s = call.tensor(scores, torch.float32)
    B, A, C = s.shape
    e = call.tensor(encoded_boxes, torch.float32, (B, A, 4))
    a = call.tensor(anchors, torch.float32, (A, 4))
    flags = (_lib.SSDK_INPUT_LOGITS if scores_are_logits else _lib.SSDK_INPUT_SCORES) | _lib.SSDK_BOXES_ENCODED
    # phase (an extension): 'scan' enqueues only the HBM-bound score scan, a following call with phase='finish' and the same
    # arguments the latency-bound rest (include/ssdk.h: SSDK_POST_SCAN_ONLY / SSDK_POST_FINISH_ONLY)
    assert phase in (None, 'scan', 'finish')
    flags |= {None: 0, 'scan': _lib.SSDK_POST_SCAN_ONLY, 'finish': _lib.SSDK_POST_FINISH_ONLY}[phase]
    out = _postprocess(call, e, a, s, flags, score_threshold, iou_threshold, max_boxes_per_class,
                       return_anchor_indices, box_scaler, final_score_threshold)
    return None if phase == 'scan' else call.result(*out)


def batch_detections_by_label(encoded_boxes, anchors, scores, score_threshold, iou_threshold, max_boxes_per_class,
                              scores_are_logits=False, box_scaler=None, final_score_threshold=None, image_ids=None):
    """The per-label detection lists of the reference's evaluator (metrics.py:113-123: add_detections over the images of a batch in
    order), straight from the NMS results (label-major pack kernel, ssdk_detect_by_label).
    Returns (boxes [C, B*K, 4], scores [C, B*K], image [C, B*K] int, counts [C] int): label c has counts[c] records --
    image 0's boxes of that class in descending score, then image 1's, ...; `image` holds image_ids[b] (default b)."""
    call = Call()
    s = call.tensor(scores, torch.float32)
    B, A, C = s.shape
    K = int(max_boxes_per_class)
    e = call.tensor(encoded_boxes, torch.float32, (B, A, 4))
    a = call.tensor(anchors, torch.float32, (A, 4))
    sc = None if box_scaler is None else call.tensor(box_scaler, torch.float32, (B, 4))
    ids = None if image_ids is None else call.tensor(image_ids, torch.int32, (B,))
    thr2 = float('-inf') if final_score_threshold is None else float(final_score_threshold)
    boxes, out_scores = call.empty([C, B * K, 4], torch.float32), call.empty([C, B * K], torch.float32)
    image, counts = call.empty([C, B * K], torch.int32), call.empty([C], torch.int32)
    flags = (_lib.SSDK_INPUT_LOGITS if scores_are_logits else _lib.SSDK_INPUT_SCORES) | _lib.SSDK_BOXES_ENCODED
    _lib.check(_lib.load().ssdk_detect_by_label(call.ctx(), ptr(e), ptr(a), ptr(s), flags, B, A, C, float(score_threshold),
                                                float(iou_threshold), K, ptr(sc), thr2, ptr(ids), ptr(boxes), ptr(out_scores),
                                                ptr(image), ptr(counts)))
    return call.result(boxes, out_scores, image, counts)


def batch_coco_detections(encoded_boxes, anchors, scores, image_sizes, score_threshold, iou_threshold, max_boxes_per_class,
                          scores_are_logits=False, box_scaler=None, final_score_threshold=None, image_ids=None, category_ids=None):
    """batch_multiclass_non_max_suppression plus, per detection, the row of the COCO results file that
    inference/evaluate_on_COCO.ipynb (cell 10) builds: bbox = [int(xmin), int(ymin), int(xmax - xmin), int(ymax - ymin)] of
    box * [height, width, height, width], category_id (through `category_ids`, the notebook's integer_to_coco_id), image_id.
    image_sizes: [B, 2] (height, width).  Returns (boxes, scores, classes, num_detections, bbox_xywh [B,N',4] int,
    category_id [B,N'] int, image_id [B,N'] int)."""
    call = Call()
    s = call.tensor(scores, torch.float32)
    B, A, C = s.shape
    K = int(max_boxes_per_class)
    M = C * K
    e = call.tensor(encoded_boxes, torch.float32, (B, A, 4))
    a = call.tensor(anchors, torch.float32, (A, 4))
    sizes = call.tensor(image_sizes, torch.float32, (B, 2))
    sc = None if box_scaler is None else call.tensor(box_scaler, torch.float32, (B, 4))
    ids = None if image_ids is None else call.tensor(image_ids, torch.int32, (B,))
    cats = None if category_ids is None else call.tensor(category_ids, torch.int32, (C,))
    thr2 = float('-inf') if final_score_threshold is None else float(final_score_threshold)
    boxes, out_scores = call.empty([B, M, 4], torch.float32), call.empty([B, M], torch.float32)
    classes, num = call.empty([B, M], torch.int32), call.empty([B], torch.int32)
    xywh, cat, img = call.empty([B, M, 4], torch.int32), call.empty([B, M], torch.int32), call.empty([B, M], torch.int32)
    flags = (_lib.SSDK_INPUT_LOGITS if scores_are_logits else _lib.SSDK_INPUT_SCORES) | _lib.SSDK_BOXES_ENCODED
    _lib.check(_lib.load().ssdk_detect_coco(call.ctx(), ptr(e), ptr(a), ptr(s), flags, B, A, C, float(score_threshold),
                                            float(iou_threshold), K, ptr(sc), thr2, ptr(sizes), ptr(ids), ptr(cats), ptr(boxes),
                                            ptr(out_scores), ptr(classes), ptr(num), ptr(xywh), ptr(cat), ptr(img)))
    return call.result(boxes, out_scores, classes, num, xywh, cat, img)
