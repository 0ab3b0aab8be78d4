"""How should the two sub-paths of the bench step share the GPU?  Graph replays of the step with the sub-paths on two streams:
issue order, stream priorities, resident CTAs per SM of the (persistent) training-step kernel."""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('single-shot-detector_b200')
syn = importlib.import_module('single-shot-detector_b200.synthetic')
L = pkg._lib
cfg = syn.CONFIGS[2]
H, W, C, Bt, G = cfg['H'], cfg['W'], cfg['C'], cfg['B'], cfg['G']
Bi = syn.CONFIGS[3]['B']
gen = pkg.AnchorGenerator(scale_multipliers=cfg['scale_multipliers'])
anchors = gen(H, W)
A = anchors.shape[0]
g = torch.Generator(device='cuda').manual_seed(2)
gt = {k: torch.from_numpy(v).cuda() for k, v in syn.make_groundtruth(2, Bt, G, H, W, C).items()}
tlog = torch.randn([Bt, A, C], device='cuda', generator=g) - 4.595
tcod = torch.randn([Bt, A, 4], device='cuda', generator=g)
ilog = torch.randn([Bi, A, C], device='cuda', generator=g) - 7.0
igt = syn.make_groundtruth(3, Bi, G, H, W, C)
anc_np = anchors.cpu().numpy()
for b in range(Bi):
    sim = syn._pair_iou(igt['boxes'][b], anc_np)
    for gi in range(G):
        idx = torch.from_numpy(np.nonzero(sim[gi] >= 0.4)[0]).cuda()
        ilog[b, idx, int(igt['labels'][b, gi])] = 1.5 + 1.5 * torch.randn([idx.numel()], device='cuda', generator=g)
icod = torch.randn([Bi, A, 4], device='cuda', generator=g)
st = pkg.SSD.from_predictions(H, W, {'encoded_boxes': tcod, 'class_predictions': tlog}, gen, C)
si = pkg.SSD.from_predictions(H, W, {'encoded_boxes': icod, 'class_predictions': ilog}, gen, C)
params = {'gamma': 2.0, 'alpha': 0.25}
train = lambda: st.loss(gt, params)
infer = lambda: si.get_predictions(0.05, 0.5, 100)

def timeit(fn, reps=40):
    cap = pkg.graph.capture(fn, warmup=2)
    best = 1e9
    for _ in range(4):
        for _ in range(3): cap.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): cap.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    cap.release()
    return round(best, 5)

scan = lambda: si.get_predictions(0.05, 0.5, 100, phase='scan')
finish = lambda: si.get_predictions(0.05, 0.5, 100, phase='finish')
lref = train()
ref = infer()
scan()
two = finish()
out = {'split_phase_equals_whole': bool(all(torch.equal(ref[k], two[k]) for k in ref)),
       'train': timeit(train), 'infer': timeit(infer), 'infer_split': timeit(lambda: (scan(), finish())),
       'scan_only': timeit(scan), 'sequential': timeit(lambda: (train(), infer()))}
# the scan first (alone on the GPU), then the training step next to the latency-bound rest of the post-processing
for ctas in (0,):
    for name, pr in (('', None), ('_finish_hi', (0, -1))):
        conc = pkg.graph.concurrent(train, finish, priorities=pr, train_ctas_per_sm=ctas)
        out['scan_then_train_and_finish_ctas%d%s' % (ctas, name)] = timeit(lambda: (scan(), conc()))
    conc2 = pkg.graph.concurrent(finish, train, train_ctas_per_sm=ctas)
    out['scan_then_finish_and_train_ctas%d' % ctas] = timeit(lambda: (scan(), conc2()))
# the same with the training step's chunks handed out dynamically (SSDK_OPT_TRAIN_DYNAMIC_CHUNKS)
L.set_option(L.SSDK_OPT_TRAIN_DYNAMIC_CHUNKS, 1)
ld = train()
out['dyn_loss_equal_to_1e-7'] = bool(abs(float(ld['classification_loss']) - float(lref['classification_loss'])) <= 1e-7 * abs(float(lref['classification_loss'])))
out['dyn_loss_repeatable'] = bool(all(float(train()['classification_loss']) == float(ld['classification_loss']) for _ in range(5)))
out['dyn_train'] = timeit(train)
for mm in (1, 2, 3):
    L.set_option(L.SSDK_OPT_MATCH_CTAS_PER_SM, mm)
    out['dyn_train_m%d' % mm] = timeit(train)
L.set_option(L.SSDK_OPT_MATCH_CTAS_PER_SM, 0)
for name, fns, pr in (('dyn_scan_then_finish_and_train', (finish, train), None), ('dyn_scan_then_finish_hi_and_train', (finish, train), (-1, 0)),
                      ('dyn_scan_then_train_and_finish', (train, finish), None), ('dyn_scan_then_train_and_finish_hi', (train, finish), (0, -1))):
    conc = pkg.graph.concurrent(*fns, priorities=pr)
    out[name] = timeit(lambda: (scan(), conc()))
out['dyn_concurrent_infer_first_hi'] = timeit(pkg.graph.concurrent(infer, train, priorities=(-1, 0)))
out['dyn_concurrent_train_first'] = timeit(pkg.graph.concurrent(train, infer))
L.set_option(L.SSDK_OPT_TRAIN_DYNAMIC_CHUNKS, 0)
# the training step first, then the scan, then the rest
out['train_scan_finish'] = timeit(lambda: (train(), scan(), finish()))
L.set_option(L.SSDK_OPT_TRAIN_CTAS_PER_SM, 0)
print(json.dumps(out))
