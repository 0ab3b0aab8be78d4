"""Workload for ncu captures of the head-layout kernels and the matcher: cfg2 training batch (16 images, 107,415 anchors,
90 classes) as channels_first tower outputs; a few forward, forward+backward and detect calls.
    ncu --set full --import-source on -k regex:'head_|match_kernel' ... python scripts/profile_head.py [reps]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('single-shot-detector_b200')
syn = importlib.import_module('single-shot-detector_b200.synthetic')

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = syn.CONFIGS[2]
H, W, C, B, G = cfg['H'], cfg['W'], cfg['C'], cfg['B'], cfg['G']
gen = pkg.AnchorGenerator(scale_multipliers=cfg['scale_multipliers'])
anchors = gen(H, W)
A, n = anchors.shape[0], gen.num_anchors_per_location
g = torch.Generator(device='cuda').manual_seed(0)
shapes = [(-(-H // s), -(-W // s)) for s in gen.strides]
lv_cls = [torch.randn([B, n * C, h, w], device='cuda', generator=g) - 4.595 for h, w in shapes]
lv_box = [torch.randn([B, n * 4, h, w], device='cuda', generator=g) for h, w in shapes]
gt = {k: torch.from_numpy(v).cuda() for k, v in syn.make_groundtruth(2, B, G, H, W, C).items()}
ssd = pkg.SSD.from_head_outputs(H, W, lv_box, lv_cls, gen, C)
params = {'gamma': 2.0, 'alpha': 0.25}
up = torch.ones(2, device='cuda')
for _ in range(reps):
    ssd.loss(gt, params)
    ssd.loss_with_gradients(gt, params, upstream=up)
    ssd.get_predictions(0.05, 0.5, 100)
torch.cuda.synchronize()
print('ok', float(ssd.num_matches))
