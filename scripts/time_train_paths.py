"""Graph-replay timings of the training-side sub-paths (both layouts) on cfg2: forward, fused forward+backward."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
exec(open(os.path.join(ROOT, 'scripts', 'profile_head.py')).read().split("for _ in range(reps):")[0])
flat_logits = pkg.reshape_and_concatenate(lv_box, lv_cls, C, n, lazy=False)
fssd = pkg.SSD.from_predictions(H, W, flat_logits, gen, C)
def timeit(fn, reps_=30):
    cap = pkg.graph.capture(fn, warmup=2)
    for _ in range(3): cap.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps_): cap.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps_
print('head: fwd %.4f  fwd+bwd %.4f   anchor-major: fwd %.4f  fwd+bwd %.4f ms' % (
    timeit(lambda: ssd.loss(gt, params)), timeit(lambda: ssd.loss_with_gradients(gt, params, upstream=up)),
    timeit(lambda: fssd.loss(gt, params)), timeit(lambda: fssd.loss_with_gradients(gt, params, upstream=up))))
