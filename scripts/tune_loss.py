"""Sweep the loss kernel's tuning knobs (SSDK_LOSS_RPW / STAGES / CTAS) on the cfg2 batch; prints kernel ms per setting."""
import importlib, os, sys, itertools
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module('single-shot-detector_b200'); syn = importlib.import_module('single-shot-detector_b200.synthetic')
cfg = syn.CONFIGS[2]; H, W, C, G, B = cfg['H'], cfg['W'], cfg['C'], cfg['G'], cfg['B']
gen = pkg.AnchorGenerator(scale_multipliers=cfg['scale_multipliers']); anchors = gen(H, W); A = anchors.shape[0]
gt = {k: torch.from_numpy(v).cuda() for k, v in syn.make_groundtruth(2, B, G, H, W, C).items()}
lg = torch.from_numpy(syn.make_logits('train', 2, B, A, C)).cuda(); cd = torch.from_numpy(syn.make_codes(2, B, A)).cuda()
ssd = pkg.SSD.from_predictions(H, W, {'encoded_boxes': cd, 'class_predictions': lg}, gen, C)
P = {'gamma': 2.0, 'alpha': 0.25}
def measure(env):
    for k in ('SSDK_LOSS_RPW', 'SSDK_LOSS_STAGES', 'SSDK_LOSS_CTAS'):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        for _ in range(3): ssd.loss(gt, P)
        pkg._lib.set_profiling(True); 
        for _ in range(10): r = ssd.loss(gt, P)
        prof = pkg._lib.profile_read(); pkg._lib.set_profiling(False)
        return prof['ssd_loss'][0] / prof['ssd_loss'][1], float(r['classification_loss'])
    except Exception as e:
        return None, str(e)[:80]
print('default', measure({}))
for rpw, st, ct in [(8,2,4),(8,3,3),(8,4,2),(4,4,4),(4,3,5),(4,2,5),(12,2,3),(12,3,2),(16,2,2),(16,3,1),(8,2,3),(8,3,2),(20,2,1),(8,2,5)]:
    print(dict(rpw=rpw, stages=st, ctas=ct), measure({'SSDK_LOSS_RPW': rpw, 'SSDK_LOSS_STAGES': st, 'SSDK_LOSS_CTAS': ct}))
