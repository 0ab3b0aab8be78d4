"""Oracle target assignment (test infrastructure; see oracle/__init__.py).

Follows reference detector/training_target_creation.py: get_training_targets
:5-45, match_boxes :48-130, create_targets :133-176."""
import numpy as np

from .box_utils import encode, iou


def get_training_targets(anchors, groundtruth_boxes, groundtruth_labels,
                         positives_threshold=0.5, negatives_threshold=0.4):
    gt = np.asarray(groundtruth_boxes, dtype=np.float32).reshape(-1, 4)
    A = anchors.shape[0]
    if gt.shape[0] > 0:                                             # :28-36
        matches = match_boxes(anchors, gt, positives_threshold, negatives_threshold, True)
    else:
        matches = np.full([A], -1, dtype=np.int32)                  # :26,37
    reg_targets, cls_targets = create_targets(anchors, gt, groundtruth_labels, matches)  # :41-44
    return reg_targets, cls_targets, matches


def match_boxes(anchors, groundtruth_boxes, positives_threshold=0.5,
                negatives_threshold=0.4, force_match_groundtruth=True, return_similarity=False):
    assert positives_threshold >= negatives_threshold               # :86
    sim = iou(groundtruth_boxes, anchors)                           # :89  [N, A]
    matches = np.argmax(sim, axis=0).astype(np.int32)               # :90  first maximum
    vals = np.max(sim, axis=0)                                      # :91
    is_pos = (vals >= np.float32(positives_threshold)).astype(np.int32)  # :92
    if positives_threshold == negatives_threshold:                  # :94 (python float ==)
        is_neg = 1 - is_pos
        matches = matches * is_pos + (-1 * is_neg)                  # :96
    else:
        is_neg = (np.float32(negatives_threshold) > vals).astype(np.int32)  # :98
        ign = (1 - is_pos) * (1 - is_neg)                           # :99
        matches = matches * is_pos + (-1 * is_neg) + (-2 * ign)     # :100
    if force_match_groundtruth:                                     # :105
        A = anchors.shape[0]
        forced_ids = np.argmax(sim, axis=1).astype(np.int32)        # :112  [N], first max
        # one_hot [N, A] int32 (:116), then argmax over rows BEFORE masking (:117):
        # for an anchor picked by several GTs the lowest GT index wins regardless of is_okay.
        row_ids = np.zeros([A], dtype=np.int32)
        picked = np.zeros([A], dtype=bool)
        for g in range(sim.shape[0] - 1, -1, -1):                   # descending so lowest g is last
            row_ids[forced_ids[g]] = g
            picked[forced_ids[g]] = True
        forced_vals = np.max(sim, axis=1)                           # :120
        is_okay = forced_vals >= np.float32(0.1)                    # :121-122
        mask = np.zeros([A], dtype=bool)                            # :123,125: any okay GT picked a
        mask[forced_ids[is_okay]] = True
        matches = np.where(mask, row_ids, matches).astype(np.int32)  # :126
    if return_similarity:
        return matches, sim
    return matches


def create_targets(anchors, groundtruth_boxes, groundtruth_labels, matches):
    A = anchors.shape[0]
    gt = np.asarray(groundtruth_boxes, dtype=np.float32).reshape(-1, 4)
    labels = np.asarray(groundtruth_labels, dtype=np.int32).reshape(-1)
    matched = np.nonzero(matches >= 0)[0]                           # :148-149 (ascending)
    gt_idx = matches[matched]                                       # :154
    reg = np.zeros([A, 4], dtype=np.float32)                        # :163 + stitch :166-169
    cls = np.zeros([A], dtype=np.int32)                             # :164 + stitch :171-174
    if matched.size:
        reg[matched] = encode(gt[gt_idx], anchors[matched])         # :155-158
        cls[matched] = labels[gt_idx] + 1                           # :159-160
    return reg, cls
