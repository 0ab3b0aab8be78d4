"""Oracle anchor generation (test infrastructure; see oracle/__init__.py).

Follows reference detector/anchor_generator.py:13-120 (AnchorGenerator) and
:123-170 (tile_anchors)."""
import itertools

import numpy as np

f32 = np.float32


class AnchorGenerator:
    def __init__(self, strides=(8, 16, 32, 64, 128), scales=(32, 64, 128, 256, 512),
                 scale_multipliers=(1.0, 1.4142), aspect_ratios=(1.0, 2.0, 0.5)):
        assert len(strides) == len(scales)                          # :33
        self.strides = list(strides)
        self.scales = list(scales)
        self.scale_multipliers = list(scale_multipliers)
        self.aspect_ratios = list(aspect_ratios)
        self.num_anchors_per_location = len(aspect_ratios) * len(scale_multipliers)  # :38

    def __call__(self, image_height, image_width):
        H, W = f32(image_height), f32(image_width)                  # :53-54
        info, per_map = [], []
        for stride in self.strides:                                 # :58-62
            h = int(np.ceil(H / f32(stride)))
            w = int(np.ceil(W / f32(stride)))
            info.append((stride, h, w))
            per_map.append(h * w * self.num_anchors_per_location)
        self.num_anchors_per_feature_map = per_map                  # :65

        pairs = list(itertools.product(self.scale_multipliers, self.aspect_ratios))  # :70
        ratios = np.array([a for _, a in pairs], dtype=np.float32)  # :71
        levels = []
        for i, (stride, h, w) in enumerate(info):
            # python-double product, then cast (:75)
            scales = np.array([m * self.scales[i] for m, _ in pairs], dtype=np.float32)
            s = f32(stride)
            off_y = f32(0.5) * (H - (f32(h) - f32(1.0)) * s)        # :92
            off_x = f32(0.5) * (W - (f32(w) - f32(1.0)) * s)        # :93
            levels.append(tile_anchors(h, w, scales, ratios, (s, s), (off_y, off_x)))
        self.raw_anchors = levels                                   # :105
        anchors = np.concatenate(levels, axis=0)                    # :107
        scaler = np.array([H, W, H, W], dtype=np.float32)           # :110-113
        return anchors / scaler                                     # :114, unclipped (:116-118)


def tile_anchors(grid_height, grid_width, scales, aspect_ratios, anchor_stride, anchor_offset):
    """reference :123-170; returns [grid_height*grid_width*N, 4] absolute boxes."""
    n = scales.shape[0]
    ratio_sqrts = np.sqrt(aspect_ratios)                            # :144
    heights = scales / ratio_sqrts                                  # :145
    widths = scales * ratio_sqrts                                   # :146
    yc = np.arange(grid_height, dtype=np.int32).astype(np.float32) * anchor_stride[0] + anchor_offset[0]  # :151
    xc = np.arange(grid_width, dtype=np.int32).astype(np.float32) * anchor_stride[1] + anchor_offset[1]   # :152
    xc, yc = np.meshgrid(xc, yc)                                    # :153
    centers = np.stack([yc, xc], axis=2)[:, :, None, :]             # :156-157
    centers = np.tile(centers, [1, 1, n, 1])                        # :158
    sizes = np.stack([heights, widths], axis=1)[None, None]         # :161-162
    sizes = np.tile(sizes, [grid_height, grid_width, 1, 1])         # :163
    half = f32(0.5) * sizes
    boxes = np.concatenate([centers - half, centers + half], axis=3)  # :166
    return boxes.reshape(-1, 4).astype(np.float32)                  # :168
