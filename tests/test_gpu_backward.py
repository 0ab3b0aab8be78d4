"""GPU parity of the loss backward (ssdk_ssd_loss_backward) against the float64 gradient oracle (oracle/losses_grad.py,
itself pinned by finite differences of the pinned forward in tests/test_oracle_golden.py).  Tolerance: 1e-5 relative
per element (north star: losses to 1e-5), with an absolute floor for gradients that underflow float32 arithmetic."""
import importlib

import numpy as np
import pytest

from conftest import load_pkg

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

RTOL = 1e-5
# d = prediction - target is formed in float32 (as the reference's graph does, losses.py:16) from operands of magnitude
# up to ~10 whose own last bit (target = log/divide, 1 ulp) is not pinned: absolute slack of a few ulp(10) on d, i.e. on N * grad
CODES_ATOL = 4e-6


@pytest.fixture(scope='module')
def pkg():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return load_pkg()


def cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def grad_close(got, want, name, atol=1e-30):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    err = np.abs(got - want)
    tol = RTOL * np.abs(want) + atol
    bad = err > tol
    assert not bad.any(), '%s: %d elements off, worst rel err %.3g at %s (got %g want %g)' % (
        name, bad.sum(), (err / (np.abs(want) + 1e-300)).max(), np.unravel_index(np.argmax(err / (np.abs(want) + 1e-300)), want.shape),
        got.flat[np.argmax(err / (np.abs(want) + 1e-300))], want.flat[np.argmax(err / (np.abs(want) + 1e-300))])


def make_case(pkg, syn, H, W, C, B, G, seed, kind='realistic', sm=(1.0, 1.4142)):
    from oracle.anchor_generator import AnchorGenerator as OracleGen
    anchors = OracleGen(scale_multipliers=list(sm))(H, W)
    A = anchors.shape[0]
    gt = syn.make_groundtruth(seed, B, G, H, W, C, vary_count=True)
    logits = syn.make_logits(kind, seed, B, A, C, anchors, gt) if kind == 'realistic' else syn.make_logits(kind, seed, B, A, C)
    codes = (syn.make_codes(seed, B, A) * np.float32(0.8)).astype(np.float32)
    gen = pkg.AnchorGenerator(scale_multipliers=list(sm))
    return anchors, gt, logits, codes, gen


@pytest.mark.parametrize('thr', [(0.5, 0.5), (0.5, 0.4)])
@pytest.mark.parametrize('gamma,alpha', [(2.0, 0.25), (1.5, 0.4)])
def test_backward_matches_oracle(pkg, syn, thr, gamma, alpha):
    from oracle import losses_grad as og
    ssd_mod = importlib.import_module('single-shot-detector_b200.detector.ssd')
    H, W, C, B, G = 256, 320, 12, 3, 7           # A = 7,680: ragged last tile (B*A % 64 != 0 is covered by the next test)
    anchors, gt, logits, codes, gen = make_case(pkg, syn, H, W, C, B, G, seed=11)
    params = {'gamma': gamma, 'alpha': alpha}
    up = (0.7, 1.3)
    want = og.ssd_loss_grad(anchors, codes, logits, gt, params, C, upstream=up, positives_threshold=thr[0], negatives_threshold=thr[1])
    ssd = pkg.SSD.from_predictions(H, W, {'encoded_boxes': cuda(codes), 'class_predictions': cuda(logits)}, gen, C)
    old = ssd_mod.POSITIVES_THRESHOLD, ssd_mod.NEGATIVES_THRESHOLD
    ssd_mod.POSITIVES_THRESHOLD, ssd_mod.NEGATIVES_THRESHOLD = thr
    try:
        losses, grads = ssd.loss_with_gradients({k: cuda(v) for k, v in gt.items()}, params, upstream=up)
    finally:
        ssd_mod.POSITIVES_THRESHOLD, ssd_mod.NEGATIVES_THRESHOLD = old
    assert float(ssd.num_matches) == float(want['num_matches']) > 0
    grad_close(grads['class_predictions'].cpu().numpy(), want['class_predictions'], 'grad logits')
    grad_close(grads['encoded_boxes'].cpu().numpy(), want['encoded_boxes'], 'grad codes', atol=CODES_ATOL * max(up) / float(want['num_matches']))
    loc, cls = og.forward64(anchors, codes, logits, gt, params, C, positives_threshold=thr[0], negatives_threshold=thr[1])
    assert abs(losses['localization_loss'].item() - loc) <= RTOL * abs(loc)
    assert abs(losses['classification_loss'].item() - cls) <= RTOL * abs(cls)


@pytest.mark.parametrize('H,W,C,B', [(200, 333, 7, 1), (64, 96, 1, 5), (128, 160, 91, 2)])
def test_backward_ragged_shapes(pkg, syn, H, W, C, B):
    """Anchor counts that do not fill the last tile, one class, class counts that are not multiples of 4."""
    from oracle import losses_grad as og
    anchors, gt, logits, codes, gen = make_case(pkg, syn, H, W, C, B, 5, seed=23, kind='dense')
    params = {'gamma': 2.0, 'alpha': 0.25}
    want = og.ssd_loss_grad(anchors, codes, logits, gt, params, C)
    ssd = pkg.SSD.from_predictions(H, W, {'encoded_boxes': cuda(codes), 'class_predictions': cuda(logits)}, gen, C)
    _, grads = ssd.loss_with_gradients({k: cuda(v) for k, v in gt.items()}, params)
    grad_close(grads['class_predictions'].cpu().numpy(), want['class_predictions'], 'grad logits')
    grad_close(grads['encoded_boxes'].cpu().numpy(), want['encoded_boxes'], 'grad codes', atol=CODES_ATOL / max(float(want['num_matches']), 1.0))


def test_autograd_path_equals_explicit_backward(pkg, syn):
    """(w_loc * loc + w_cls * cls).backward() through SSD.loss == loss_with_gradients(upstream=(w_loc, w_cls)) (model.py:86-91,115-118)."""
    H, W, C, B, G = 256, 320, 12, 2, 6
    anchors, gt, logits, codes, gen = make_case(pkg, syn, H, W, C, B, G, seed=31)
    params = {'gamma': 2.0, 'alpha': 0.25, 'localization_loss_weight': 2.0, 'classification_loss_weight': 0.5}
    d_gt = {k: cuda(v) for k, v in gt.items()}
    lg, cd = cuda(logits).requires_grad_(True), cuda(codes).requires_grad_(True)
    ssd = pkg.SSD.from_predictions(H, W, {'encoded_boxes': cd, 'class_predictions': lg}, gen, C)
    total = pkg.config.total_loss(ssd.loss(d_gt, params), params)
    total.backward()
    ssd2 = pkg.SSD.from_predictions(H, W, {'encoded_boxes': cuda(codes), 'class_predictions': cuda(logits)}, gen, C)
    losses, grads = ssd2.loss_with_gradients(d_gt, params, upstream=(2.0, 0.5))
    assert torch.equal(lg.grad, grads['class_predictions']) and torch.equal(cd.grad, grads['encoded_boxes'])
    assert abs(total.item() - (2.0 * losses['localization_loss'].item() + 0.5 * losses['classification_loss'].item())) < 1e-6 * abs(total.item())
    with torch.no_grad():                                                  # no graph when grads are off
        assert not ssd.loss(d_gt, params)['classification_loss'].requires_grad


def test_backward_properties_full_size(pkg, syn):
    """cfg2 shapes (107,415 anchors, 90 classes), 2 images: structure of the gradient that holds at any size."""
    cfg = syn.CONFIGS[2]
    H, W, C, G, B = cfg['H'], cfg['W'], cfg['C'], cfg['G'], 2
    gen = pkg.AnchorGenerator(scale_multipliers=cfg['scale_multipliers'])
    anchors = gen(H, W)
    A = anchors.shape[0]
    gt = syn.make_groundtruth(2, B, G, H, W, C)
    logits = cuda(syn.make_logits('train', 2, B, A, C))
    codes = cuda(syn.make_codes(2, B, A))
    d_gt = {k: cuda(v) for k, v in gt.items()}
    params = {'gamma': 2.0, 'alpha': 0.25}
    ssd = pkg.SSD.from_predictions(H, W, {'encoded_boxes': codes, 'class_predictions': logits}, gen, C)
    _, g11 = ssd.loss_with_gradients(d_gt, params, upstream=(1.0, 1.0))
    _, g10 = ssd.loss_with_gradients(d_gt, params, upstream=(1.0, 0.0))
    _, g03 = ssd.loss_with_gradients(d_gt, params, upstream=(0.0, 3.0))
    _, _, matches = ssd._create_targets(d_gt)
    n = float((matches >= 0).sum())
    # linear in the upstream pair; the two losses touch disjoint tensors
    assert torch.count_nonzero(g10['class_predictions']) == 0 and torch.count_nonzero(g03['encoded_boxes']) == 0
    torch.testing.assert_close(g03['class_predictions'], 3.0 * g11['class_predictions'], rtol=2e-6, atol=0)
    assert torch.equal(g10['encoded_boxes'], g11['encoded_boxes'])
    # codes: zero off the matched anchors, |g| <= 1/N on them
    unmatched = (matches < 0)
    assert torch.count_nonzero(g11['encoded_boxes'][unmatched]) == 0
    assert g11['encoded_boxes'].abs().max().item() <= 1.0 / n * (1 + 1e-6)
    # logits: negative-class gradients are >= 0 (pushing logits down), exactly one negative entry per matched anchor
    neg_entries = (g11['class_predictions'] < 0).sum(dim=2)
    assert torch.equal(neg_entries, (matches >= 0).to(neg_entries.dtype))
    # against the float64 oracle on one image
    from oracle import losses_grad as og
    a_np = anchors.cpu().numpy()
    one = {k: v[:1] for k, v in gt.items()}
    ssd1 = pkg.SSD.from_predictions(H, W, {'encoded_boxes': codes[:1], 'class_predictions': logits[:1]}, gen, C)
    _, g = ssd1.loss_with_gradients({k: cuda(v) for k, v in one.items()}, params)
    want = og.ssd_loss_grad(a_np, codes[:1].cpu().numpy(), logits[:1].cpu().numpy(), one, params, C)
    grad_close(g['class_predictions'].cpu().numpy(), want['class_predictions'], 'grad logits (cfg2)')
    grad_close(g['encoded_boxes'].cpu().numpy(), want['encoded_boxes'], 'grad codes (cfg2)', atol=CODES_ATOL / float(want['num_matches']))


def test_fused_forward_backward_equals_two_pass(pkg, syn):
    """ssdk_ssd_loss_forward_backward (one read of the logits) == forward kernel + backward kernel."""
    for (H, W, C, B, G, kind, gamma) in [(256, 320, 12, 3, 7, 'realistic', 2.0), (200, 333, 7, 2, 5, 'dense', 2.0),
                                          (128, 160, 91, 2, 5, 'dense', 1.5)]:
        anchors, gt, logits, codes, gen = make_case(pkg, syn, H, W, C, B, G, seed=41, kind=kind)
        params = {'gamma': gamma, 'alpha': 0.3}
        d_gt = {k: cuda(v) for k, v in gt.items()}
        ssd = pkg.SSD.from_predictions(H, W, {'encoded_boxes': cuda(codes), 'class_predictions': cuda(logits)}, gen, C)
        l2, g2 = ssd.loss_with_gradients(d_gt, params, upstream=(0.5, 2.0), fused=False)
        n2 = float(ssd.num_matches)
        l1, g1 = ssd.loss_with_gradients(d_gt, params, upstream=(0.5, 2.0), fused=True)
        assert float(ssd.num_matches) == n2
        # identical per-element arithmetic for the gradients; the loss sums use the general (2-MUFU) negative form here and
        # the polynomial fast path in the forward kernel: 1e-5 relative, like every loss comparison
        assert torch.equal(g1['encoded_boxes'], g2['encoded_boxes'])
        torch.testing.assert_close(g1['class_predictions'], g2['class_predictions'], rtol=1e-6, atol=1e-30)
        for k in ('localization_loss', 'classification_loss'):
            assert abs(l1[k].item() - l2[k].item()) <= RTOL * abs(l2[k].item()), (k, l1[k].item(), l2[k].item())


def test_precomputed_targets_step_equals_the_full_step(pkg, syn):
    """SSD.assign_targets (anchors + ground truth only, e.g. issued on a side stream during the network's forward pass) followed by
    loss_with_gradients(targets=...) == the step that assigns targets itself, for both layouts."""
    from oracle import box_predictor as obp
    H, W, C, B, G = 256, 320, 9, 3, 7
    anchors, gt, logits, codes, gen = make_case(pkg, syn, H, W, C, B, G, seed=23)
    dgt = {k: cuda(v) for k, v in gt.items()}
    params = {'gamma': 2.0, 'alpha': 0.25}
    ssd = pkg.SSD.from_predictions(H, W, {'encoded_boxes': cuda(codes), 'class_predictions': cuda(logits)}, gen, C)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        targets = pkg.SSD.assign_targets(ssd.anchors, dgt)
    torch.cuda.current_stream().wait_stream(side)
    assert float(targets['count']) == float((targets['matches'] >= 0).sum())
    l0, g0 = ssd.loss_with_gradients(dgt, params, upstream=(0.7, 1.3))
    l1, g1 = ssd.loss_with_gradients(None, params, upstream=(0.7, 1.3), targets=targets)
    assert float(l0['localization_loss']) == float(l1['localization_loss']) and float(l0['classification_loss']) == float(l1['classification_loss'])
    assert torch.equal(g0['class_predictions'], g1['class_predictions']) and torch.equal(g0['encoded_boxes'], g1['encoded_boxes'])
    shapes = obp.level_shapes(H, W, gen.strides)
    n = gen.num_anchors_per_location
    head = pkg.SSD.from_head_outputs(H, W, [cuda(t) for t in obp.split_to_levels(codes, shapes, n)],
                                     [cuda(t) for t in obp.split_to_levels(logits, shapes, n)], gen, C)
    h0, hg0 = head.loss_with_gradients(dgt, params)
    h1, hg1 = head.loss_with_gradients(None, params, targets=targets)
    assert float(h0['classification_loss']) == float(h1['classification_loss'])
    assert all(torch.equal(a, b) for a, b in zip(hg0['class_predictions'], hg1['class_predictions']))
