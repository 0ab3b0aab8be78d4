"""Workload for ncu captures of the inference side: cfg3 geometry, `realistic` logits, anchor-major and head layouts.
    ncu --set full --import-source on -k regex:'filter_kernel|nms' ... python scripts/profile_infer.py [batch] [reps]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('single-shot-detector_b200')
syn = importlib.import_module('single-shot-detector_b200.synthetic')

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = syn.CONFIGS[3]
H, W, C, G = cfg['H'], cfg['W'], cfg['C'], cfg['G']
gen = pkg.AnchorGenerator(scale_multipliers=cfg['scale_multipliers'])
anchors = gen(H, W)
A, n = anchors.shape[0], gen.num_anchors_per_location
gt = syn.make_groundtruth(3, B, G, H, W, C)
logits = torch.from_numpy(syn.make_logits('realistic', 3, B, A, C, anchors.cpu().numpy(), gt)).cuda()
codes = torch.from_numpy(syn.make_codes(3, B, A)).cuda()
shapes = [(-(-H // s), -(-W // s)) for s in gen.strides]


def to_levels(t, D):
    out, off = [], 0
    for h, w in shapes:
        cnt = h * w * n
        out.append(t[:, off:off + cnt].reshape(B, h, w, n * D).permute(0, 3, 1, 2).contiguous())
        off += cnt
    return out


flat = pkg.SSD.from_predictions(H, W, {'encoded_boxes': codes, 'class_predictions': logits}, gen, C)
head = pkg.SSD.from_head_outputs(H, W, to_levels(codes, 4), to_levels(logits, C), gen, C)
for _ in range(reps):
    a = flat.get_predictions(0.05, 0.5, 100)
    b = head.get_predictions(0.05, 0.5, 100)
torch.cuda.synchronize()
print('ok', int(a['num_boxes'].sum()), all(torch.equal(a[k], b[k]) for k in a))
