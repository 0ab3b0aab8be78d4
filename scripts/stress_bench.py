"""BASELINE.json configs[4] ("stress"): 1344x896 images, 300 GT boxes per image, the shipped 6-anchors-per-location set
(150,402 anchors, 80 classes), dense scores N(-2, 1.5) (~73 % of all (anchor, class) pairs above the 0.05 threshold),
batch 8.  SURVEY.md section 8(d): at this size the matcher is FP32/ALU-bound (G*A = 45.1 M IoU pairs per image), not
HBM-bound, so it is reported as IoU pairs per second; the dense post-processing is sort / NMS bound.

    python scripts/stress_bench.py [--steps K] > profiles/<tag>_stress.json        (on the GPU box)"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('single-shot-detector_b200')
syn = importlib.import_module('single-shot-detector_b200.synthetic')

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--batch', type=int, default=0)
args = ap.parse_args()

cfg = syn.CONFIGS[5]
H, W, C, G = cfg['H'], cfg['W'], cfg['C'], cfg['G']
B = args.batch or cfg['B']
gen = pkg.AnchorGenerator(scale_multipliers=cfg['scale_multipliers'])
anchors = gen(H, W)
A = anchors.shape[0]
gt = {k: torch.from_numpy(v).cuda() for k, v in syn.make_groundtruth(5, B, G, H, W, C).items()}
g = torch.Generator(device='cuda').manual_seed(5)
logits = torch.randn([B, A, C], device='cuda', generator=g) * 1.5 - 2.0          # 'dense'
codes = torch.randn([B, A, 4], device='cuda', generator=g)
ssd = pkg.SSD.from_predictions(H, W, {'encoded_boxes': codes, 'class_predictions': logits}, gen, C)
params = {'gamma': 2.0, 'alpha': 0.25}


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


ms_targets, tg = timed(lambda: ssd._create_targets(gt), args.steps)
ms_loss, ls = timed(lambda: ssd.loss(gt, params), args.steps)
ms_fb, _ = timed(lambda: ssd.loss_with_gradients(gt, params), args.steps)
ms_pp, pred = timed(lambda: ssd.get_predictions(0.05, 0.5, 100), max(2, args.steps // 2))
pkg._lib.set_profiling(True)
pkg._lib.profile_read()
n = max(2, args.steps // 2)
for _ in range(n):
    ssd.loss(gt, params)
    ssd.get_predictions(0.05, 0.5, 100)
prof = pkg._lib.profile_read()
pkg._lib.set_profiling(False)
matches = tg[2]
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
except Exception:
    pass
peak = float(peaks.get('hbm_gbs', 6650.0))
cand = float((torch.sigmoid(logits[0]) > 0.05).float().mean())
out = {
    'config': 'stress (BASELINE.json configs[4]): %dx%d, %d anchors (6 per location), %d classes, batch %d, %d GT boxes/image, dense logits N(-2,1.5)'
              % (H, W, A, C, B, G),
    'matcher': {'ms': ms_targets, 'iou_pairs_per_s': B * G * A / (ms_targets * 1e-3), 'images_per_s': B / (ms_targets * 1e-3),
                'positives': int((matches >= 0).sum()), 'bound': 'FP32/ALU issue (SURVEY.md 8d)'},
    'train_forward': {'ms': ms_loss, 'images_per_s': B / (ms_loss * 1e-3),
                      'frac_of_hbm_roofline': ((4 * A * C + 56 * A + 20 * G) * B / (ms_loss * 1e-3) / 1e9) / peak,
                      'localization_loss': float(ls['localization_loss']), 'classification_loss': float(ls['classification_loss'])},
    'train_forward_backward': {'ms': ms_fb, 'images_per_s': B / (ms_fb * 1e-3),
                               'frac_of_hbm_roofline': ((8 * A * C + 72 * A + 20 * G) * B / (ms_fb * 1e-3) / 1e9) / peak},
    'postprocess_dense': {'ms': ms_pp, 'images_per_s': B / (ms_pp * 1e-3), 'candidate_fraction_image0': cand,
                          'candidates_per_image': cand * A * C, 'detections_image0': int(pred['num_boxes'][0]),
                          'frac_of_hbm_roofline': ((4 * A * C + 32 * A + 24 * C * 100 + 4) * B / (ms_pp * 1e-3) / 1e9) / peak},
    'kernel_ms_per_call': {k: v[0] / n for k, v in prof.items() if v[1]},
    'peak_hbm_gbs': peak,
}
print(json.dumps(out))
