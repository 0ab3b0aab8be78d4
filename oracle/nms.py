"""Oracle post-processing (test infrastructure; see oracle/__init__.py).

Follows reference detector/utils/nms.py: multiclass_non_max_suppression :6-45,
batch_multiclass_non_max_suppression :48-102.  tf.image.non_max_suppression
(TF 1.12 NonMaxSuppressionV3, external C++) is restated in
non_max_suppression_v3 below and, for speed, in oracle/csrc/oracle_nms.c.
That op is pinned to TensorFlow's own published unit-test vectors
(tests/golden/tf_nms_vectors.py, restating non_max_suppression_op_test.cc) and cross-checked
against torchvision.ops.nms on random boxes (tests/test_oracle_golden.py)."""
import ctypes
import os

import numpy as np

from .box_utils import decode

f32 = np.float32
_LIB = None


def _lib():
    """Load oracle/_build/liboracle.so, building it with gcc on first use."""
    global _LIB
    if _LIB is None:
        here = os.path.dirname(os.path.abspath(__file__))
        so = os.path.join(here, '_build', 'liboracle.so')
        if not os.path.exists(so):
            import subprocess
            subprocess.check_call(['make', '-s', '-C', here, '_build/liboracle.so'])
        lib = ctypes.CDLL(so)
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
        lib.oracle_nms_v3.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_float, ctypes.c_float, ip]
        lib.oracle_nms_v3.restype = ctypes.c_int
        lib.oracle_multiclass_nms.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_float, ctypes.c_float, ip, ip]
        lib.oracle_multiclass_nms.restype = None
        _LIB = lib
    return _LIB


def _iou_greater(bi, bj, thr):
    ymin_i, xmin_i = min(bi[0], bi[2]), min(bi[1], bi[3])
    ymax_i, xmax_i = max(bi[0], bi[2]), max(bi[1], bi[3])
    ymin_j, xmin_j = min(bj[0], bj[2]), min(bj[1], bj[3])
    ymax_j, xmax_j = max(bj[0], bj[2]), max(bj[1], bj[3])
    area_i = f32(f32(ymax_i - ymin_i) * f32(xmax_i - xmin_i))
    area_j = f32(f32(ymax_j - ymin_j) * f32(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return False
    ih = f32(min(ymax_i, ymax_j) - max(ymin_i, ymin_j))
    iw = f32(min(xmax_i, xmax_j) - max(xmin_i, xmin_j))
    inter = f32(max(ih, f32(0)) * max(iw, f32(0)))
    return bool(f32(inter / f32(f32(area_i + area_j) - inter)) > f32(thr))


def non_max_suppression_v3(boxes, scores, max_output_size, iou_threshold, score_threshold,
                           use_c=True):
    """Indices selected by TF 1.12 NonMaxSuppressionV3, descending score (ties: lower index)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    if not (0.0 <= float(iou_threshold) <= 1.0):        # OP_REQUIRES of the TF kernel
        raise ValueError('iou_threshold must be in [0, 1]')
    if use_c:
        out = np.zeros([max(int(max_output_size), 1)], dtype=np.int32)
        k = _lib().oracle_nms_v3(
            boxes.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
            scores.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 1, n,
            int(max_output_size), float(iou_threshold), float(score_threshold),
            out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        return out[:k].copy()
    cand = np.nonzero(scores > f32(score_threshold))[0]
    order = cand[np.argsort(-scores[cand], kind='stable')]
    sel = []
    for i in order:
        if len(sel) >= max_output_size:
            break
        if not any(_iou_greater(boxes[i], boxes[j], iou_threshold) for j in reversed(sel)):
            sel.append(int(i))
    return np.asarray(sel, dtype=np.int32)


def multiclass_non_max_suppression(boxes, scores, score_threshold, iou_threshold,
                                   max_boxes_per_class, use_c=True, return_indices=False):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n, C = scores.shape
    K = int(max_boxes_per_class)
    if use_c:
        idx = np.zeros([C, max(K, 1)], dtype=np.int32)
        cnt = np.zeros([C], dtype=np.int32)
        _lib().oracle_multiclass_nms(
            boxes.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
            scores.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n, C, K,
            float(iou_threshold), float(score_threshold),
            idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
            cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        per_class = [idx[c, :cnt[c]] for c in range(C)]
    else:
        per_class = [non_max_suppression_v3(boxes, scores[:, c], K, iou_threshold,
                                            score_threshold, use_c=False) for c in range(C)]  # :31-36
    sel_boxes = np.concatenate([boxes[i] for i in per_class], axis=0).reshape(-1, 4)  # :38,42
    sel_scores = np.concatenate([scores[i, c] for c, i in enumerate(per_class)])      # :39,43
    sel_classes = np.concatenate([np.full(i.shape, c, dtype=np.int32)                 # :40,44
                                  for c, i in enumerate(per_class)])
    if return_indices:
        return sel_boxes, sel_scores, sel_classes, np.concatenate(per_class).astype(np.int32)
    return sel_boxes, sel_scores, sel_classes


def batch_multiclass_non_max_suppression(encoded_boxes, anchors, scores, score_threshold,
                                         iou_threshold, max_boxes_per_class, use_c=True,
                                         return_anchor_indices=False):
    encoded_boxes = np.asarray(encoded_boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    anchors = np.asarray(anchors, dtype=np.float32)
    B, _, C = scores.shape
    K = int(max_boxes_per_class)
    M = K * C                                                       # :82
    out_b = np.zeros([B, M, 4], np.float32); out_s = np.zeros([B, M], np.float32)
    out_c = np.zeros([B, M], np.int32); out_n = np.zeros([B], np.int32)
    out_a = np.full([B, M], -1, np.int32)
    for b in range(B):                                              # tf.map_fn :96-101
        conf = np.max(scores[b], axis=1) >= f32(score_threshold)    # :71  (>=)
        keep = np.nonzero(conf)[0]
        boxes = decode(encoded_boxes[b][keep], anchors[keep])       # :72-76
        boxes = np.minimum(np.maximum(boxes, f32(0.0)), f32(1.0))   # :77
        sb, ss, sc, si = multiclass_non_max_suppression(
            boxes, scores[b][keep], score_threshold, iou_threshold, K, use_c, True)  # :79-82
        n = sb.shape[0]
        out_b[b, :n] = sb; out_s[b, :n] = ss; out_c[b, :n] = sc; out_n[b] = n  # :83-93 (zero pad)
        out_a[b, :n] = keep[si]
    if return_anchor_indices:
        return out_b, out_s, out_c, out_n, out_a
    return out_b, out_s, out_c, out_n
