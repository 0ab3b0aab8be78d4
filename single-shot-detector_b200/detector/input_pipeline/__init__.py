"""Only the ground-truth box arithmetic of the reference's input pipeline lives here (random_image_crop.py); image decoding,
tf.data and the augmentation samplers are out of scope."""
from .random_image_crop import (change_coordinate_frame, crop_boxes, ioa, prune_completely_outside_window,  # noqa: F401
                                prune_non_overlapping_boxes)
