"""Oracle of the ground-truth side of the random-crop augmentation (test infrastructure; see oracle/__init__.py).

Follows reference detector/input_pipeline/random_image_crop.py: ioa :190-209, prune_completely_outside_window :102-131,
prune_non_overlapping_boxes :134-159, change_coordinate_frame :162-187, and the box part of randomly_crop_image :86-99.
The crop window is an INPUT here (the reference draws it with tf.image.sample_distorted_bounding_box, :68-78)."""
import numpy as np

from .box_utils import area, intersection
from .constants import EPSILON

f32 = np.float32


def ioa(boxes1, boxes2):
    intersections = intersection(boxes1, boxes2)                                          # :206
    areas = area(boxes2)[None, :]                                                         # :207
    return np.minimum(np.maximum(intersections / (areas + EPSILON), f32(0.0)), f32(1.0))  # :208


def prune_completely_outside_window(boxes, window):
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 4)
    w = np.asarray(window, np.float32)
    y_min, x_min, y_max, x_max = [boxes[:, i:i + 1] for i in range(4)]                    # :117
    violations = np.concatenate([y_min >= w[2], x_min >= w[3], y_max <= w[0], x_max <= w[1]], axis=1)  # :122-125
    valid = np.nonzero(~violations.any(axis=1))[0].astype(np.int64)                       # :126-129
    return boxes[valid], valid                                                            # :130-131


def prune_non_overlapping_boxes(boxes1, boxes2, min_overlap):
    boxes1 = np.ascontiguousarray(boxes1, np.float32).reshape(-1, 4)
    overlap = ioa(boxes2, boxes1)                                                         # :152  [M, N]
    overlap = overlap.max(axis=0) if overlap.shape[0] else np.zeros([boxes1.shape[0]], np.float32)  # :153
    keep = np.nonzero(overlap >= f32(min_overlap))[0].astype(np.int64)                    # :155-156
    return boxes1[keep], keep                                                             # :158-159


def change_coordinate_frame(boxes, window):
    b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 4)
    w = np.asarray(window, np.float32)
    ymin, xmin, ymax, xmax = b[:, 0] - w[0], b[:, 1] - w[1], b[:, 2] - w[0], b[:, 3] - w[1]   # :174-178
    win_height, win_width = w[2] - w[0], w[3] - w[1]                                          # :180-181
    out = np.stack([ymin / win_height, xmin / win_width, ymax / win_height, xmax / win_width], axis=1)  # :182-185
    return np.minimum(np.maximum(out, f32(0.0)), f32(1.0))                                    # :186


def crop_boxes(boxes, window, overlap_thresh=0.3):
    """randomly_crop_image :86-99 with the window given: (boxes in the window's frame, indices into the input boxes)."""
    boxes, inside_window_ids = prune_completely_outside_window(boxes, window)             # :87
    boxes, keep_indices = prune_non_overlapping_boxes(boxes, np.asarray(window, np.float32)[None], overlap_thresh)  # :90-93
    boxes = change_coordinate_frame(boxes, window)                                        # :96
    return boxes, inside_window_ids[keep_indices]                                         # :98
