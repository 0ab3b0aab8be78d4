"""Oracle box math (test infrastructure; see oracle/__init__.py).

Follows reference detector/utils/box_utils.py: iou :14-27, intersection :30-50,
area :53-61, to_center_coordinates :64-77, encode :80-111, decode :114-142,
batch_decode :145-173.  Every op is a separate float32 NumPy op (no fusion),
as TF executes one kernel per Python-level op."""
import numpy as np

from .constants import EPSILON, SCALE_FACTORS

f32 = np.float32


def _f(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def area(boxes):
    b = _f(boxes)
    return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])                # :60-61


def intersection(boxes1, boxes2):
    b1, b2 = _f(boxes1), _f(boxes2)
    ymin1, xmin1, ymax1, xmax1 = [b1[:, i:i + 1] for i in range(4)]  # :38
    ymin2, xmin2, ymax2, xmax2 = [b2[:, i:i + 1].T for i in range(4)]  # :39 + transposes
    ih = np.maximum(f32(0.0), np.minimum(ymax1, ymax2) - np.maximum(ymin1, ymin2))  # :42-44
    iw = np.maximum(f32(0.0), np.minimum(xmax1, xmax2) - np.maximum(xmin1, xmin2))  # :45-47
    return ih * iw                                                  # :50


def iou(boxes1, boxes2):
    inter = intersection(boxes1, boxes2)                            # :23
    a1, a2 = area(boxes1), area(boxes2)                             # :24-25
    unions = a1[:, None] + a2[None, :] - inter                      # :26  ((a1+a2) - inter)
    q = inter / (unions + EPSILON)                                  # :27
    return np.minimum(np.maximum(q, f32(0.0)), f32(1.0))            # clip_by_value


def to_center_coordinates(boxes):
    b = _f(boxes)
    ymin, xmin, ymax, xmax = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    h, w = ymax - ymin, xmax - xmin                                 # :75
    cy, cx = ymin + f32(0.5) * h, xmin + f32(0.5) * w               # :76
    return [cy, cx, h, w]


def encode(boxes, anchors):
    cya, cxa, ha, wa = to_center_coordinates(anchors)               # :92
    cy, cx, h, w = to_center_coordinates(boxes)                     # :93
    ha = ha + EPSILON; wa = wa + EPSILON; h = h + EPSILON; w = w + EPSILON  # :96-99
    ty = (cy - cya) / ha                                            # :101
    tx = (cx - cxa) / wa                                            # :102
    th = np.log(h / ha)                                             # :103
    tw = np.log(w / wa)                                             # :104
    ty = ty * f32(SCALE_FACTORS[0]); tx = tx * f32(SCALE_FACTORS[1])  # :106-107
    th = th * f32(SCALE_FACTORS[2]); tw = tw * f32(SCALE_FACTORS[3])  # :108-109
    return np.stack([ty, tx, th, tw], axis=1)                       # :111


def decode(codes, anchors):
    cya, cxa, ha, wa = to_center_coordinates(anchors)               # :127
    c = _f(codes)
    ty = c[:, 0] / f32(SCALE_FACTORS[0]); tx = c[:, 1] / f32(SCALE_FACTORS[1])  # :130-131
    th = c[:, 2] / f32(SCALE_FACTORS[2]); tw = c[:, 3] / f32(SCALE_FACTORS[3])  # :132-133
    h = np.exp(th) * ha                                             # :135
    w = np.exp(tw) * wa                                             # :136
    cy = ty * ha + cya                                              # :137 (mul, then add)
    cx = tx * wa + cxa                                              # :138
    hh, hw = f32(0.5) * h, f32(0.5) * w
    return np.stack([cy - hh, cx - hw, cy + hh, cx + hw], axis=1)   # :140-142


def batch_decode(box_encodings, anchors):
    e = _f(box_encodings)
    B, A = e.shape[0], e.shape[1]
    tiled = np.tile(_f(anchors)[None], [B, 1, 1])                   # :161-164
    d = decode(e.reshape(-1, 4), tiled.reshape(-1, 4)).reshape(B, A, 4)  # :165-170
    return np.minimum(np.maximum(d, f32(0.0)), f32(1.0))            # :171
