"""B200-native per-anchor detection hot path with the call surface of TropComplique/single-shot-detector.

    AnchorGenerator, SSD                       detector/anchor_generator.py, detector/ssd.py
    get_training_targets, match_boxes, ...     detector/training_target_creation.py
    focal_loss, localization_loss              detector/losses.py
    reshape_and_concatenate                    detector/box_predictor.py:67-104 (a view: SSD consumes the per-level
                                               tower outputs without the transpose / concat copy)
    iou, encode, decode, batch_multiclass_non_max_suppression, ...   detector/utils

All arithmetic runs in hand-written sm_100a CUDA kernels (csrc/) behind the C ABI of include/ssdk.h.
The directory name contains a hyphen: import it with importlib.import_module('single-shot-detector_b200')
(or `import ssd_b200`, the alias module at the repository root).
"""
from . import _lib, config, graph, parallel  # noqa: F401
from .detector import SSD  # noqa: F401
from .detector.anchor_generator import AnchorGenerator  # noqa: F401
from .detector.box_predictor import HeadPredictions, reshape_and_concatenate  # noqa: F401
from .detector.input_pipeline import (change_coordinate_frame, crop_boxes, ioa, prune_completely_outside_window,  # noqa: F401
                                      prune_non_overlapping_boxes)
from .detector.losses import focal_loss, localization_loss  # noqa: F401
from .detector.training_target_creation import (batch_training_targets, create_targets,  # noqa: F401
                                                get_training_targets, match_boxes)
from .detector.utils import (area, batch_decode, batch_multiclass_non_max_suppression, decode, encode,  # noqa: F401
                             intersection, iou, multiclass_non_max_suppression)

__version__ = '0.1.0'
