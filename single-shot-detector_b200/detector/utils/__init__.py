# same exports as reference detector/utils/__init__.py (minus the conv-net layer helpers, which are out of scope)
from .box_utils import iou, area, intersection, encode, decode, batch_decode
from .nms import batch_multiclass_non_max_suppression, multiclass_non_max_suppression
