"""
tf.image stand-ins (test infrastructure only; see __init__.py).

non_max_suppression restates TensorFlow 1.12's NonMaxSuppressionV3 CPU kernel
(tensorflow/core/kernels/non_max_suppression_op.cc; the source is NOT in
/root/reference, the reference only calls it at detector/utils/nms.py:33):

  * candidates: every i with scores[i] > score_threshold             (strict)
  * std::priority_queue ordered by score -> popped in descending score.
    Equal scores: the 1.12 heap order is unspecified; this restatement pops the
    LOWER index first (what TF >= 2 does explicitly).
  * each popped candidate is compared against the already selected boxes,
    newest first, and dropped iff IoU > iou_threshold                (strict)
  * IoU: corners are min/max-normalised; if either area <= 0 the pair never
    suppresses; iou = inter / (area_i + area_j - inter), no epsilon, float32.
  * stops once max_output_size boxes are selected.
"""
import numpy as np

from ._tensor import _arr, _wrap


class ResizeMethod:
    NEAREST_NEIGHBOR = 1
    BILINEAR = 0


def _iou_gt(boxes, i, j, thr):
    f = np.float32
    bi, bj = boxes[i], boxes[j]
    ymin_i, xmin_i = min(bi[0], bi[2]), min(bi[1], bi[3])
    ymax_i, xmax_i = max(bi[0], bi[2]), max(bi[1], bi[3])
    ymin_j, xmin_j = min(bj[0], bj[2]), min(bj[1], bj[3])
    ymax_j, xmax_j = max(bj[0], bj[2]), max(bj[1], bj[3])
    area_i = f(f(ymax_i - ymin_i) * f(xmax_i - xmin_i))
    area_j = f(f(ymax_j - ymin_j) * f(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return False
    iy0, ix0 = max(ymin_i, ymin_j), max(xmin_i, xmin_j)
    iy1, ix1 = min(ymax_i, ymax_j), min(xmax_i, xmax_j)
    inter = f(max(f(iy1 - iy0), f(0)) * max(f(ix1 - ix0), f(0)))
    iou = f(inter / f(f(area_i + area_j) - inter))
    return bool(iou > f(thr))


def non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5,
                        score_threshold=float('-inf')):
    boxes = np.asarray(_arr(boxes), dtype=np.float32)
    scores = np.asarray(_arr(scores), dtype=np.float32)
    k = int(_arr(max_output_size))
    cand = np.nonzero(scores > np.float32(score_threshold))[0]
    order = cand[np.argsort(-scores[cand], kind='stable')]  # score desc, index asc
    selected = []
    for i in order:
        if len(selected) >= k:
            break
        keep = True
        for j in reversed(selected):
            if _iou_gt(boxes, i, j, iou_threshold):
                keep = False
                break
        if keep:
            selected.append(int(i))
    return _wrap(np.asarray(selected, dtype=np.int32))
