"""Graph-replay time of the stand-alone target assignment (cfg2, and the stress configuration) and of the forward+backward
training step for the library build named by SSDK_LIB; one JSON line."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('single-shot-detector_b200')
syn = importlib.import_module('single-shot-detector_b200.synthetic')
params = {'gamma': 2.0, 'alpha': 0.25}


def timeit(fn, reps=40):
    cap = pkg.graph.capture(fn, warmup=2)
    best = 1e9
    for _ in range(4):
        for _ in range(3):
            cap.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            cap.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    cap.release()
    return round(best, 5)


out = {'lib': os.path.basename(os.path.dirname(os.environ.get('SSDK_LIB', 'default/x')))}
for name, cid, B in (('cfg2', 2, 16), ('cfg5', 5, 8), ('cfg2_B4', 2, 4)):
    cfg = syn.CONFIGS[cid]
    H, W, C, G = cfg['H'], cfg['W'], cfg['C'], cfg['G']
    gen = pkg.AnchorGenerator(scale_multipliers=cfg['scale_multipliers'])
    anchors = gen(H, W)
    A = anchors.shape[0]
    gt = {k: torch.from_numpy(v).cuda() for k, v in syn.make_groundtruth(cid, B, G, H, W, C).items()}
    out[name + '_matcher_ms'] = timeit(lambda: pkg.SSD.assign_targets(anchors, gt))
    if cid == 2 and B == 16:
        g = torch.Generator(device='cuda').manual_seed(2)
        logits = torch.randn([B, A, C], device='cuda', generator=g) - 4.595
        codes = torch.randn([B, A, 4], device='cuda', generator=g)
        ssd = pkg.SSD.from_predictions(H, W, {'encoded_boxes': codes, 'class_predictions': logits}, gen, C)
        out['cfg2_fwd_bwd_ms'] = timeit(lambda: ssd.loss_with_gradients(gt, params, upstream=(1.0, 1.0)))
        tg = pkg.SSD.assign_targets(anchors, gt)
        out['cfg2_matches_sum'] = int(tg['matches'].sum())
        del ssd, logits, codes
print(json.dumps(out))
