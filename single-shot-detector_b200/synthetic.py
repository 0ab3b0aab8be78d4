"""Synthetic workloads for the parity tests and bench.py.

Pure NumPy, deterministic per (config, image index): seed = 1000*config + image.
These are INPUT generators only -- no reference arithmetic is restated here apart
from the pairwise IoU needed to plant object-like logits on anchors near a GT box.
"""
import numpy as np

# name -> (H, W, scale_multipliers, num_classes, batch, gt_per_image)
CONFIGS = {
    1: dict(H=640, W=640, scale_multipliers=[1.0, 1.4142], C=80, B=1, G=20),
    2: dict(H=640, W=896, scale_multipliers=[1.0, 2 ** (1 / 3), 2 ** (2 / 3)], C=90, B=16, G=20),
    3: dict(H=640, W=896, scale_multipliers=[1.0, 2 ** (1 / 3), 2 ** (2 / 3)], C=90, B=32, G=20),
    4: dict(H=640, W=896, scale_multipliers=[1.0, 2 ** (1 / 3), 2 ** (2 / 3)], C=90, B=256, G=20),
    5: dict(H=896, W=1344, scale_multipliers=[1.0, 1.4142], C=80, B=8, G=300),
}
STRIDES = [8, 16, 32, 64, 128]
SCALES = [32, 64, 128, 256, 512]
ASPECT_RATIOS = [1.0, 2.0, 0.5]


def make_gt_boxes(rng, n, H, W):
    """n valid boxes [ymin,xmin,ymax,xmax] float32 in [0,1] (degenerate ones re-drawn)."""
    out = np.zeros([n, 4], np.float32)
    k = 0
    while k < n:
        s = np.exp(rng.uniform(np.log(16.0), np.log(0.6 * min(H, W))))
        r = np.exp(rng.uniform(np.log(0.5), np.log(2.0)))
        h, w = s / np.sqrt(r) / H, s * np.sqrt(r) / W
        cy, cx = rng.uniform(0, 1, 2)
        b = np.clip(np.array([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2]), 0, 1).astype(np.float32)
        if b[0] < b[2] and b[1] < b[3]:
            out[k] = b
            k += 1
    return out


def make_groundtruth(config, B, G, H, W, C, first_image=0, vary_count=False):
    """Padded GT in the input pipeline's format (reference pipeline.py:61-62)."""
    boxes = np.zeros([B, G, 4], np.float32)
    labels = np.zeros([B, G], np.int32)
    num = np.zeros([B], np.int32)
    for b in range(B):
        rng = np.random.default_rng(1000 * config + first_image + b)
        n = int(rng.integers(0, G + 1)) if vary_count else G
        boxes[b, :n] = make_gt_boxes(rng, n, H, W)
        labels[b, :n] = rng.integers(0, C, n)
        num[b] = n
    return {'boxes': boxes, 'labels': labels, 'num_boxes': num}


def _pair_iou(gt, anchors):
    ih = np.maximum(0, np.minimum(gt[:, None, 2], anchors[None, :, 2]) - np.maximum(gt[:, None, 0], anchors[None, :, 0]))
    iw = np.maximum(0, np.minimum(gt[:, None, 3], anchors[None, :, 3]) - np.maximum(gt[:, None, 1], anchors[None, :, 1]))
    inter = ih * iw
    a1 = (gt[:, 2] - gt[:, 0]) * (gt[:, 3] - gt[:, 1])
    a2 = (anchors[:, 2] - anchors[:, 0]) * (anchors[:, 3] - anchors[:, 1])
    return inter / (a1[:, None] + a2[None, :] - inter + 1e-8)


def make_logits(kind, config, B, A, C, anchors=None, groundtruth=None, first_image=0, out=None):
    """kind: 'train' N(-4.595,1) | 'dense' N(-2,1.5) | 'realistic' N(-7,1) + N(1.5,1.5) planted."""
    if out is None:
        out = np.empty([B, A, C], np.float32)
    for b in range(B):
        rng = np.random.default_rng(1000 * config + first_image + b + 500_000)
        x = rng.standard_normal([A, C], dtype=np.float32)
        if kind == 'train':
            x += np.float32(-4.595)
        elif kind == 'dense':
            x *= np.float32(1.5); x += np.float32(-2.0)
        elif kind == 'realistic':
            x += np.float32(-7.0)
            n = int(groundtruth['num_boxes'][b])
            if n:
                sim = _pair_iou(groundtruth['boxes'][b, :n], anchors)
                for g in range(n):
                    idx = np.nonzero(sim[g] >= 0.4)[0]
                    x[idx, groundtruth['labels'][b, g]] = (
                        1.5 + 1.5 * rng.standard_normal(idx.size)).astype(np.float32)
        else:
            raise ValueError(kind)
        out[b] = x
    return out


def make_codes(config, B, A, first_image=0, out=None):
    if out is None:
        out = np.empty([B, A, 4], np.float32)
    for b in range(B):
        rng = np.random.default_rng(1000 * config + first_image + b + 900_000)
        out[b] = rng.standard_normal([A, 4], dtype=np.float32)
    return out
