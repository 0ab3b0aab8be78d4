"""AnchorGenerator with the reference's interface (detector/anchor_generator.py:12-120), computed by
csrc/anchors.cu through ssdk_anchors."""
import ctypes
import itertools

import numpy as np
import torch

from .. import _lib
from .._tensors import Call, ptr


class AnchorGenerator:
    def __init__(self, strides=[8, 16, 32, 64, 128], scales=[32, 64, 128, 256, 512],
                 scale_multipliers=[1.0, 1.4142], aspect_ratios=[1.0, 2.0, 0.5]):
        """Same arguments as the reference (anchor_generator.py:13-38)."""
        assert len(strides) == len(scales)
        self.strides = list(strides)
        self.scales = list(scales)
        self.scale_multipliers = list(scale_multipliers)
        self.aspect_ratios = list(aspect_ratios)
        self.num_anchors_per_location = len(aspect_ratios) * len(scale_multipliers)

    def _host_params(self):
        pairs = list(itertools.product(self.scale_multipliers, self.aspect_ratios))      # :70
        ratios = np.array([a for _, a in pairs], dtype=np.float32)                       # :71
        # python-double product, then one cast to float32 (:75)
        scales = np.array([[m * s for m, _ in pairs] for s in self.scales], dtype=np.float32)
        strides = np.array(self.strides, dtype=np.int32)
        return strides, scales, ratios

    def count(self, image_height, image_width):
        """(total, per-level list) without touching the GPU."""
        strides, _, _ = self._host_params()
        total = ctypes.c_int64(0)
        per_level = np.zeros([len(self.strides)], np.int32)
        _lib.check(_lib.load().ssdk_num_anchors(
            int(image_height), int(image_width), strides.ctypes.data, len(self.strides),
            self.num_anchors_per_location, ctypes.byref(total), per_level.ctypes.data))
        return int(total.value), [int(v) for v in per_level]

    def __call__(self, image_height, image_width, device=None):
        """Returns a float32 CUDA tensor [num_anchors, 4], normalised, not clipped (:40-120)."""
        image_height, image_width = int(image_height), int(image_width)
        total, per_level = self.count(image_height, image_width)
        self.num_anchors_per_feature_map = per_level                                      # :65
        call = Call(device)
        strides, scales, ratios = self._host_params()
        out = call.empty([total, 4], torch.float32)
        raw = call.empty([total, 4], torch.float32)
        _lib.check(_lib.load().ssdk_anchors(
            call.ctx(), image_height, image_width, strides.ctypes.data, scales.ctypes.data, ratios.ctypes.data,
            len(self.strides), self.num_anchors_per_location, ptr(out), ptr(raw)))
        self.raw_anchors = list(torch.split(raw, per_level, dim=0))                        # :105
        return out
