"""TensorFlow's own published known-answer vectors for NonMaxSuppressionV3, the one op on the hot path whose algorithm
lives outside the reference tree (called at /root/reference/detector/utils/nms.py:33 as tf.image.non_max_suppression).

Restated (numbers only, no code) from the public TensorFlow unit tests
    tensorflow/core/kernels/non_max_suppression_op_test.cc   -- class NonMaxSuppressionV3OpTest (TF r1.12)
    tensorflow/python/ops/image_ops_test.py                  -- NonMaxSuppressionTest.testSelectFromThreeClusters
TensorFlow is not installable in this container, so the vectors are typed in from the published test file; each
case names the TEST_F it restates.  They pin: greedy order by descending score, suppression iff IoU > threshold,
corner min/max normalisation ("flipped coordinates"), the strict score threshold, the cut at max_output_size,
duplicates (ten identical boxes -> only the first survives) and the empty input.

Used by tests/test_oracle_golden.py (oracle, Python and C) and tests/test_gpu_parity.py (CUDA, through
multiclass_non_max_suppression with one class)."""
import numpy as np

_THREE_CLUSTERS = [[0, 0, 1, 1], [0, 0.1, 1, 1.1], [0, -0.1, 1, 0.9],
                   [0, 10, 1, 11], [0, 10.1, 1, 11.1], [0, 100, 1, 101]]
_THREE_CLUSTERS_FLIPPED = [[1, 1, 0, 0], [0, 0.1, 1, 1.1], [0, .9, 1, -0.1],
                           [0, 10, 1, 11], [1, 10.1, 0, 11.1], [1, 101, 0, 100]]
_SCORES = [.9, .75, .6, .95, .5, .3]

# (name of the TEST_F, boxes [n,4], scores [n], max_output_size, iou_threshold, score_threshold, expected indices)
CASES = [
    ('TestSelectFromThreeClusters', _THREE_CLUSTERS, _SCORES, 3, .5, 0.0, [3, 0, 5]),
    ('TestSelectFromThreeClustersWithScoreThreshold', _THREE_CLUSTERS, _SCORES, 3, .5, 0.4, [3, 0]),
    ('TestSelectFromThreeClustersFlippedCoordinates', _THREE_CLUSTERS_FLIPPED, _SCORES, 3, .5, 0.0, [3, 0, 5]),
    ('TestSelectAtMostTwoBoxesFromThreeClusters', _THREE_CLUSTERS, _SCORES, 2, .5, 0.0, [3, 0]),
    ('TestSelectAtMostThirtyBoxesFromThreeClusters', _THREE_CLUSTERS, _SCORES, 30, .5, 0.0, [3, 0, 5]),
    ('TestSelectSingleBox', [[0, 0, 1, 1]], [.9], 3, .5, 0.0, [0]),
    ('TestSelectFromTenIdenticalBoxes', [[0, 0, 1, 1]] * 10, [.9] * 10, 3, .5, 0.0, [0]),
    ('TestEmptyInput', np.zeros([0, 4]), np.zeros([0]), 30, .5, 0.0, []),
]

# TestInvalidIOUThreshold: iou_threshold = 1.2 -> "iou_threshold must be in [0, 1]"
INVALID_IOU_THRESHOLD = (_THREE_CLUSTERS[:1], [.9], 3, 1.2, 0.0)


def cases():
    for name, boxes, scores, k, iou, thr, want in CASES:
        yield (name, np.asarray(boxes, np.float32).reshape(-1, 4), np.asarray(scores, np.float32), int(k), float(iou),
               float(thr), np.asarray(want, np.int32))
