"""Sweep of the fused training step's role split (matcher CTAs per SM x the matcher CTAs' share of the streaming) over several
workloads, graph-replay timings; one JSON object.   python scripts/sweep_train_split.py"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('single-shot-detector_b200')
syn = importlib.import_module('single-shot-detector_b200.synthetic')
L = pkg._lib
params = {'gamma': 2.0, 'alpha': 0.25}


def timeit(fn, reps=30):
    cap = pkg.graph.capture(fn, warmup=2)
    for _ in range(3):
        cap.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            cap.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    cap.release()
    return best


def workload(cfg_id, B, G=None, C=None):
    cfg = syn.CONFIGS[cfg_id]
    H, W = cfg['H'], cfg['W']
    C = C or cfg['C']
    G = G or cfg['G']
    gen = pkg.AnchorGenerator(scale_multipliers=cfg['scale_multipliers'])
    anchors = gen(H, W)
    A = anchors.shape[0]
    gt = {k: torch.from_numpy(v).cuda() for k, v in syn.make_groundtruth(cfg_id, B, G, H, W, C).items()}
    g = torch.Generator(device='cuda').manual_seed(cfg_id)
    logits = torch.randn([B, A, C], device='cuda', generator=g) - 4.595
    codes = torch.randn([B, A, 4], device='cuda', generator=g)
    ssd = pkg.SSD.from_predictions(H, W, {'encoded_boxes': codes, 'class_predictions': logits}, gen, C)
    return ssd, gt, A, C, G


out = {}
cases = [('cfg2_B16', 2, 16, None, None), ('cfg2_B32', 2, 32, None, None), ('cfg2_B4', 2, 4, None, None), ('cfg2_B16_G100', 2, 16, 100, None),
         ('cfg2_B16_C20', 2, 16, None, 20), ('cfg5_B8_G300', 5, 8, None, None), ('cfg1_B1', 1, 1, None, None)]
if os.environ.get('SWEEP_SMALL_BATCHES'):
    cases = [('cfg2_B%d' % b, 2, b, None, None) for b in (1, 2, 3, 4, 6, 8, 12)] + [('cfg1_B%d' % b, 1, b, None, None) for b in (2, 4, 8)]
for name, cid, B, G, C in cases:
    ssd, gt, A, C, G = workload(cid, B, G, C)
    res = {'A': A, 'C': C, 'G': G, 'B': B}
    L.set_option(L.SSDK_OPT_MATCH_CTAS_PER_SM, 0)
    L.set_option(L.SSDK_OPT_MATCH_FLAT_SHARE_PCT, -1)
    res['auto'] = round(timeit(lambda: ssd.loss(gt, params)), 5)
    sweep = {}
    for ctas in (2, 3, 4, 5):
        for share in (0, 10, 20, 30):
            L.set_option(L.SSDK_OPT_MATCH_CTAS_PER_SM, ctas)
            L.set_option(L.SSDK_OPT_MATCH_FLAT_SHARE_PCT, share)
            sweep['m%d_s%d' % (ctas, share)] = round(timeit(lambda: ssd.loss(gt, params)), 5)
    best = min(sweep, key=sweep.get)
    res['best'] = best
    res['best_ms'] = sweep[best]
    res['sweep'] = sweep
    res['roofline_ms'] = (4 * A * C + 56 * A + 20 * G) * B / 6448.4e9 * 1e3
    out[name] = res
    del ssd, gt
    torch.cuda.empty_cache()
L.set_option(L.SSDK_OPT_MATCH_CTAS_PER_SM, 0)
L.set_option(L.SSDK_OPT_MATCH_FLAT_SHARE_PCT, -1)
print(json.dumps(out))
