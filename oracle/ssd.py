"""Oracle detector head (test infrastructure; see oracle/__init__.py).

Follows reference detector/ssd.py: get_predictions :42-69, loss :71-133,
_create_targets :165-199 (minus TensorBoard summaries :125-129,135-163)."""
import numpy as np

from .constants import NEGATIVES_THRESHOLD, POSITIVES_THRESHOLD
from .losses import focal_loss, localization_loss, sigmoid
from .nms import batch_multiclass_non_max_suppression
from .training_target_creation import get_training_targets

f32 = np.float32


def create_targets_batch(anchors, groundtruth, positives_threshold=POSITIVES_THRESHOLD,
                         negatives_threshold=NEGATIVES_THRESHOLD):
    boxes, labels, num = groundtruth['boxes'], groundtruth['labels'], groundtruth['num_boxes']
    reg, cls, mat = [], [], []
    for b in range(boxes.shape[0]):                                 # tf.map_fn :193-198
        n = int(num[b])
        r, c, m = get_training_targets(anchors, boxes[b, :n], labels[b, :n],  # :183-189
                                       positives_threshold, negatives_threshold)
        reg.append(r); cls.append(c); mat.append(m)
    return np.stack(reg), np.stack(cls), np.stack(mat)


def loss(anchors, encoded_boxes, class_predictions, groundtruth, params, num_classes,
         positives_threshold=POSITIVES_THRESHOLD, negatives_threshold=NEGATIVES_THRESHOLD,
         return_all=False):
    reg_t, cls_t, matches = create_targets_batch(anchors, groundtruth, positives_threshold,
                                                 negatives_threshold)  # :84
    weights = (matches >= 0).astype(np.float32)                     # :89
    onehot = (cls_t[:, :, None] == np.arange(num_classes + 1, dtype=np.int32)).astype(np.float32)  # :96
    onehot = onehot[:, :, 1:]                                       # :100
    not_ignore = (matches >= -1).astype(np.float32)                 # :103
    cls_losses = focal_loss(class_predictions, onehot, not_ignore,  # :106-109
                            gamma=params['gamma'], alpha=params['alpha'])
    loc_losses = localization_loss(encoded_boxes, reg_t, weights)   # :117
    per_image = np.sum(weights, axis=1, dtype=np.float32)           # :121
    num_matches = np.sum(per_image, axis=0, dtype=np.float32)       # :122
    normalizer = np.maximum(num_matches, f32(1.0))                  # :123
    loc = np.sum(loc_losses, dtype=np.float32)                      # :131
    cls = np.sum(cls_losses, dtype=np.float32)                      # :132
    out = {'localization_loss': f32(loc / normalizer), 'classification_loss': f32(cls / normalizer)}
    if return_all:
        out.update(reg_targets=reg_t, cls_targets=cls_t, matches=matches, cls_losses=cls_losses,
                   loc_losses=loc_losses, num_matches=num_matches,
                   loc_sum64=np.sum(loc_losses, dtype=np.float64),
                   cls_sum64=np.sum(cls_losses, dtype=np.float64))
    return out


def get_predictions(anchors, encoded_boxes, class_predictions, score_threshold=0.05,
                    iou_threshold=0.5, max_boxes_per_class=20):
    scores = sigmoid(np.asarray(class_predictions, dtype=np.float32))  # :60
    b, s, c, n = batch_multiclass_non_max_suppression(              # :64-68
        encoded_boxes, anchors, scores, score_threshold, iou_threshold, max_boxes_per_class)
    return {'boxes': b, 'labels': c, 'scores': s, 'num_boxes': n}   # :69
