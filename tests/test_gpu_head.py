"""GPU parity of the head-layout path (csrc/head.cu, ssdk_head_*): the per-level tower outputs of
detector/box_predictor.py are consumed WITHOUT reshape_and_concatenate (:67-104), and the results must equal what the
reference computes after it -- checked against the committed golden fixture, the NumPy oracle and the anchor-major CUDA
path.  Bit-exact for the concatenation itself, matches, labels and NMS kept sets; 1e-5 relative for losses, gradients,
boxes and scores.  Also the observability by-products (ssdk_level_summaries)."""
import numpy as np
import pytest

from conftest import load_pkg

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

RTOL = 1e-5
SM = [1.0, 1.4142]
STRIDES = [8, 16, 32, 64, 128]


@pytest.fixture(scope='module')
def pkg():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return load_pkg()


def cuda(x, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).cuda()


def close(a, b, rtol=RTOL, atol=0.0):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


def _head_inputs(seed, B, C, n, shapes):
    """Same generator as tests/golden/make_golden.py::head_inputs."""
    rng = np.random.default_rng(seed)
    boxes = [rng.standard_normal([B, n * 4, h, w]).astype(np.float32) for h, w in shapes]
    classes = [rng.standard_normal([B, n * C, h, w]).astype(np.float32) for h, w in shapes]
    return boxes, classes


def _case(pkg, syn, H, W, C, B, G, seed, kind='realistic', sm=SM, data_format='channels_first', vary_count=True):
    """Anchor-major synthetic tensors + the same data laid out as tower outputs."""
    from oracle import box_predictor as obp
    from oracle.anchor_generator import AnchorGenerator as OracleGen
    anchors = OracleGen(scale_multipliers=list(sm))(H, W)
    A = anchors.shape[0]
    n = 3 * len(sm)
    gt = syn.make_groundtruth(seed, B, G, H, W, C, vary_count=vary_count)
    logits = syn.make_logits(kind, seed, B, A, C, anchors, gt) if kind == 'realistic' else syn.make_logits(kind, seed, B, A, C)
    codes = (syn.make_codes(seed, B, A) * np.float32(0.6)).astype(np.float32)
    shapes = obp.level_shapes(H, W, STRIDES)
    lv_boxes = obp.split_to_levels(codes, shapes, n, data_format)
    lv_classes = obp.split_to_levels(logits, shapes, n, data_format)
    gen = pkg.AnchorGenerator(scale_multipliers=list(sm))
    return dict(anchors=anchors, A=A, n=n, gt=gt, logits=logits, codes=codes, shapes=shapes, lv_boxes=lv_boxes,
                lv_classes=lv_classes, gen=gen, H=H, W=W, C=C, B=B, data_format=data_format)


def _head_ssd(pkg, c):
    return pkg.SSD.from_head_outputs(c['H'], c['W'], [cuda(t) for t in c['lv_boxes']], [cuda(t) for t in c['lv_classes']],
                                     c['gen'], c['C'], data_format=c['data_format'])


def _flat_ssd(pkg, c):
    raw = {'encoded_boxes': cuda(c['codes']), 'class_predictions': cuda(c['logits'])}
    return pkg.SSD.from_predictions(c['H'], c['W'], raw, c['gen'], c['C'])


def _set_thresholds(pkg, pos, neg):
    mod = load_pkg('detector.ssd')
    mod.POSITIVES_THRESHOLD, mod.NEGATIVES_THRESHOLD = pos, neg


@pytest.fixture(autouse=True)
def _restore_thresholds():
    yield
    mod = load_pkg('detector.ssd')
    mod.POSITIVES_THRESHOLD, mod.NEGATIVES_THRESHOLD = 0.5, 0.5


# ------------------------------------------------------------------------------------------------ reshape_and_concatenate
def test_head_concat_golden(pkg, golden):
    """ssdk_head_concat == the reference's reshape_and_concatenate run on the TF shim (bit-exact: data movement)."""
    import hashlib
    g = golden('head')
    B, C, n = [int(v) for v in g['tiny/params']]
    shapes = [tuple(int(v) for v in s) for s in g['tiny/shapes']]
    boxes = [cuda(g['tiny/boxes%d' % i]) for i in range(len(shapes))]
    classes = [cuda(g['tiny/classes%d' % i]) for i in range(len(shapes))]
    r = pkg.reshape_and_concatenate(boxes, classes, C, n)
    assert isinstance(r, pkg.HeadPredictions) and sorted(r.keys()) == ['class_predictions', 'encoded_boxes']
    assert np.array_equal(r['encoded_boxes'].cpu().numpy(), g['tiny/encoded_boxes'])
    assert np.array_equal(r['class_predictions'].cpu().numpy(), g['tiny/class_predictions'])
    B, C, n = [int(v) for v in g['big/params']]
    shapes = [tuple(int(v) for v in s) for s in g['big/shapes']]
    boxes, classes = _head_inputs(22, B, C, n, shapes)
    r = pkg.reshape_and_concatenate([cuda(t) for t in boxes], [cuda(t) for t in classes], C, n, lazy=False)
    assert hashlib.sha256(r['encoded_boxes'].cpu().numpy().tobytes()).hexdigest() == str(g['big/encoded_boxes_sha256'])
    assert hashlib.sha256(r['class_predictions'].cpu().numpy().tobytes()).hexdigest() == str(g['big/class_predictions_sha256'])


@pytest.mark.parametrize('data_format', ['channels_first', 'channels_last'])
def test_head_concat_vs_oracle_ragged(pkg, data_format):
    from oracle import box_predictor as obp
    B, C, n, shapes = 3, 5, 4, [(25, 42), (13, 21), (7, 11), (4, 6), (2, 3), (1, 1)]
    boxes, classes = _head_inputs(5, B, C, n, shapes)
    if data_format == 'channels_last':
        boxes = [np.ascontiguousarray(t.transpose(0, 2, 3, 1)) for t in boxes]
        classes = [np.ascontiguousarray(t.transpose(0, 2, 3, 1)) for t in classes]
    want = obp.reshape_and_concatenate(boxes, classes, C, n, data_format)
    got = pkg.reshape_and_concatenate([cuda(t) for t in boxes], [cuda(t) for t in classes], C, n, data_format=data_format)
    assert np.array_equal(got['encoded_boxes'].cpu().numpy(), want['encoded_boxes'])
    assert np.array_equal(got['class_predictions'].cpu().numpy(), want['class_predictions'])
    with pytest.raises(ValueError):
        pkg.reshape_and_concatenate([cuda(t) for t in boxes], [cuda(t) for t in classes], C + 1, n, data_format=data_format)


# ------------------------------------------------------------------------------------------------ loss
@pytest.mark.parametrize('thr', [(0.5, 0.5), (0.5, 0.4)])
@pytest.mark.parametrize('gamma,alpha', [(2.0, 0.25), (1.5, 0.4)])
@pytest.mark.parametrize('data_format', ['channels_first', 'channels_last'])
def test_head_loss_matches_oracle(pkg, syn, thr, gamma, alpha, data_format):
    from oracle import ssd as ossd
    c = _case(pkg, syn, 256, 320, 7, 3, 9, 77, data_format=data_format)
    c['gt']['num_boxes'][1] = 0                                       # an image with no boxes
    c['logits'][0, :50] = np.random.default_rng(5).uniform(-30, 30, [50, 7]).astype(np.float32)
    from oracle import box_predictor as obp
    c['lv_classes'] = obp.split_to_levels(c['logits'], c['shapes'], c['n'], data_format)
    _set_thresholds(pkg, *thr)
    params = {'gamma': gamma, 'alpha': alpha}
    dgt = {k: cuda(v) for k, v in c['gt'].items()}
    ssd = _head_ssd(pkg, c)
    sums = ssd.loss_sums(dgt, params).cpu().numpy()
    res = ssd.loss(dgt, params)
    o = ossd.loss(c['anchors'], c['codes'], c['logits'], c['gt'], params, 7, positives_threshold=thr[0], negatives_threshold=thr[1],
                  return_all=True)
    assert sums[2] == float(o['num_matches'])
    if thr[1] < thr[0]:
        assert (o['matches'] == -2).sum() > 0                         # the ignore band is really exercised
    close(sums[0], o['loc_sum64'])
    close(sums[1], o['cls_sum64'])
    close(res['localization_loss'].item(), o['localization_loss'])
    close(res['classification_loss'].item(), o['classification_loss'])
    # and against the anchor-major CUDA path on the concatenated tensors
    flat = _flat_ssd(pkg, c).loss_sums(dgt, params).cpu().numpy()
    assert flat[2] == sums[2]
    close(sums[:2], flat[:2], rtol=2e-6)


@pytest.mark.parametrize('H,W,C,B', [(200, 333, 7, 2), (64, 96, 1, 5), (128, 160, 91, 2)])
def test_head_loss_ragged_shapes(pkg, syn, H, W, C, B):
    """Level sizes that are not multiples of 4 floats (scalar tails of the flat pass), C = 1, C > 64."""
    from oracle import ssd as ossd
    c = _case(pkg, syn, H, W, C, B, 6, 31)
    _set_thresholds(pkg, 0.5, 0.4)
    params = {'gamma': 2.0, 'alpha': 0.25}
    dgt = {k: cuda(v) for k, v in c['gt'].items()}
    sums = _head_ssd(pkg, c).loss_sums(dgt, params).cpu().numpy()
    o = ossd.loss(c['anchors'], c['codes'], c['logits'], c['gt'], params, C, positives_threshold=0.5, negatives_threshold=0.4,
                  return_all=True)
    assert sums[2] == float(o['num_matches'])
    close(sums[0], o['loc_sum64'])
    close(sums[1], o['cls_sum64'])


# ------------------------------------------------------------------------------------------------ gradients
def _grad_close(got, want, name, rtol=RTOL, atol=1e-30):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    err = np.abs(got - want)
    bad = err > rtol * np.abs(want) + atol
    assert not bad.any(), '%s: %d elements off, worst rel err %.3g' % (name, bad.sum(), (err / (np.abs(want) + 1e-300)).max())


@pytest.mark.parametrize('thr', [(0.5, 0.5), (0.5, 0.4)])
@pytest.mark.parametrize('gamma,alpha', [(2.0, 0.25), (1.5, 0.4)])
@pytest.mark.parametrize('data_format', ['channels_first', 'channels_last'])
def test_head_forward_backward_matches_oracle(pkg, syn, thr, gamma, alpha, data_format):
    from oracle import box_predictor as obp, losses_grad as og, ssd as ossd
    c = _case(pkg, syn, 200, 333, 6, 2, 7, 41, data_format=data_format)
    _set_thresholds(pkg, *thr)
    params = {'gamma': gamma, 'alpha': alpha}
    up = (0.7, 1.3)
    dgt = {k: cuda(v) for k, v in c['gt'].items()}
    ssd = _head_ssd(pkg, c)
    losses, grads = ssd.loss_with_gradients(dgt, params, upstream=up)
    o = ossd.loss(c['anchors'], c['codes'], c['logits'], c['gt'], params, 6, positives_threshold=thr[0], negatives_threshold=thr[1])
    close(losses['localization_loss'].item(), o['localization_loss'])
    close(losses['classification_loss'].item(), o['classification_loss'])
    want = og.ssd_loss_grad(c['anchors'], c['codes'], c['logits'], c['gt'], params, 6, upstream=up, positives_threshold=thr[0],
                            negatives_threshold=thr[1])
    want_cls = obp.split_to_levels(want['class_predictions'], c['shapes'], c['n'], data_format)
    want_box = obp.split_to_levels(want['encoded_boxes'], c['shapes'], c['n'], data_format)
    norm = max(float(want['num_matches']), 1.0)
    for l in range(len(c['shapes'])):
        assert tuple(grads['class_predictions'][l].shape) == c['lv_classes'][l].shape
        _grad_close(grads['class_predictions'][l].cpu().numpy(), want_cls[l], 'grad class level %d' % l)
        _grad_close(grads['encoded_boxes'][l].cpu().numpy(), want_box[l], 'grad boxes level %d' % l, atol=4e-6 * max(up) / norm)
    # the anchor-major fused pass gives the same numbers
    _, gflat = _flat_ssd(pkg, c).loss_with_gradients(dgt, params, upstream=up)
    got_cat = obp.reshape_and_concatenate([g.cpu().numpy() for g in grads['encoded_boxes']],
                                          [g.cpu().numpy() for g in grads['class_predictions']], 6, c['n'], data_format)
    np.testing.assert_allclose(got_cat['class_predictions'], gflat['class_predictions'].cpu().numpy(), rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(got_cat['encoded_boxes'], gflat['encoded_boxes'].cpu().numpy(), rtol=1e-6, atol=1e-12)
    # backward-only entry (out_sums == NULL) after a plain forward
    ssd2 = _head_ssd(pkg, c)
    ssd2._loss_forward(dgt, params, keep_targets=True)
    g2 = ssd2.loss_backward(up)
    for l in range(len(c['shapes'])):
        assert torch.equal(g2['class_predictions'][l], grads['class_predictions'][l])
        assert torch.equal(g2['encoded_boxes'][l], grads['encoded_boxes'][l])


def test_head_autograd(pkg, syn):
    """SSD.loss on tower outputs that require grad: torch.autograd reaches the per-level tensors."""
    c = _case(pkg, syn, 128, 160, 5, 2, 5, 43)
    params = {'gamma': 2.0, 'alpha': 0.25}
    dgt = {k: cuda(v) for k, v in c['gt'].items()}
    lv_b = [cuda(t).requires_grad_(True) for t in c['lv_boxes']]
    lv_c = [cuda(t).requires_grad_(True) for t in c['lv_classes']]
    ssd = pkg.SSD.from_head_outputs(c['H'], c['W'], lv_b, lv_c, c['gen'], c['C'])
    out = ssd.loss(dgt, params)
    (0.5 * out['localization_loss'] + 2.0 * out['classification_loss']).backward()
    _, want = _head_ssd(pkg, c).loss_with_gradients(dgt, params, upstream=(0.5, 2.0))
    for l in range(len(lv_b)):
        assert torch.equal(lv_c[l].grad, want['class_predictions'][l])
        assert torch.equal(lv_b[l].grad, want['encoded_boxes'][l])


# ------------------------------------------------------------------------------------------------ post-processing
@pytest.mark.parametrize('data_format', ['channels_first', 'channels_last'])
@pytest.mark.parametrize('kind,K', [('realistic', 10), ('dense', 7)])
def test_head_detect_equals_anchor_major_path_and_oracle(pkg, syn, data_format, kind, K):
    from oracle import losses as olosses, nms as onms
    c = _case(pkg, syn, 200, 333, 6, 3, 8, 88, kind=kind, data_format=data_format, vary_count=False)
    head = _head_ssd(pkg, c)
    flat = _flat_ssd(pkg, c)
    got = head._get_predictions_head(head._head(), 0.05, 0.5, K, return_anchor_indices=True)
    ref = flat.get_predictions(score_threshold=0.05, iou_threshold=0.5, max_boxes_per_class=K)
    for k in ('boxes', 'labels', 'scores', 'num_boxes'):              # same kernels after the scan: bit-identical
        assert torch.equal(got[k], ref[k]), k
    want = onms.batch_multiclass_non_max_suppression(c['codes'], c['anchors'], olosses.sigmoid(c['logits']), 0.05, 0.5, K)
    assert np.array_equal(got['num_boxes'].cpu().numpy(), want[3]) and want[3].sum() > 0
    assert np.array_equal(got['labels'].cpu().numpy(), want[2])
    close(got['boxes'].cpu().numpy(), want[0], atol=1e-7)
    close(got['scores'].cpu().numpy(), want[1])
    # public API incl. the folded consumers
    scaler = np.array([[1.0, 0.8, 1.0, 0.8]] * c['B'], np.float32)
    a = head.get_predictions(0.05, 0.5, K, box_scaler=cuda(scaler), final_score_threshold=0.3)
    b = flat.get_predictions(0.05, 0.5, K, box_scaler=cuda(scaler), final_score_threshold=0.3)
    for k in ('boxes', 'labels', 'scores', 'num_boxes'):
        assert torch.equal(a[k], b[k]), k


def test_head_full_size_equals_anchor_major(pkg, syn):
    """cfg2 / cfg3 geometry (640x896, 9 anchors per location, 90 classes), 2 images: head path == anchor-major path."""
    cfg = syn.CONFIGS[2]
    c = _case(pkg, syn, cfg['H'], cfg['W'], cfg['C'], 2, cfg['G'], 2, sm=cfg['scale_multipliers'], vary_count=False)
    assert c['A'] == 107415
    params = {'gamma': 2.0, 'alpha': 0.25}
    dgt = {k: cuda(v) for k, v in c['gt'].items()}
    head, flat = _head_ssd(pkg, c), _flat_ssd(pkg, c)
    _set_thresholds(pkg, 0.5, 0.4)
    s_head, s_flat = head.loss_sums(dgt, params).cpu().numpy(), flat.loss_sums(dgt, params).cpu().numpy()
    assert s_head[2] == s_flat[2] and s_head[2] > 0
    close(s_head[:2], s_flat[:2], rtol=2e-6)
    (lh, gh), (lf, gf) = head.loss_with_gradients(dgt, params), flat.loss_with_gradients(dgt, params)
    close(lh['classification_loss'].item(), lf['classification_loss'].item(), rtol=2e-6)
    cat = pkg.reshape_and_concatenate(gh['encoded_boxes'], gh['class_predictions'], cfg['C'], c['n'])
    np.testing.assert_allclose(cat['class_predictions'].cpu().numpy(), gf['class_predictions'].cpu().numpy(), rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(cat['encoded_boxes'].cpu().numpy(), gf['encoded_boxes'].cpu().numpy(), rtol=1e-6, atol=1e-12)
    a = head.get_predictions(0.05, 0.5, 100)
    b = flat.get_predictions(0.05, 0.5, 100)
    assert int(a['num_boxes'].sum()) > 0
    for k in ('boxes', 'labels', 'scores', 'num_boxes'):
        assert torch.equal(a[k], b[k]), k


def test_head_argument_checks(pkg, syn):
    c = _case(pkg, syn, 64, 96, 3, 1, 2, 3)
    gen9 = pkg.AnchorGenerator(scale_multipliers=[1.0, 1.26, 1.59])
    with pytest.raises(ValueError):                                   # 6 anchors per location in the tensors, generator says 9
        pkg.SSD.from_head_outputs(64, 96, [cuda(t) for t in c['lv_boxes']], [cuda(t) for t in c['lv_classes']], gen9, 3)
    ssd = pkg.SSD.from_head_outputs(128, 96, [cuda(t) for t in c['lv_boxes']], [cuda(t) for t in c['lv_classes']], c['gen'], 3)
    with pytest.raises(ValueError):                                   # anchors of a 128x96 image do not fit 64x96 towers
        ssd.get_predictions()


# ------------------------------------------------------------------------------------------------ observability
def test_level_summaries(pkg, syn):
    """ssd.py:125-129,135-163: per-level matched counts and top-20 % loss statistics vs the oracle restatement."""
    from oracle import box_predictor as obp, ssd as ossd
    c = _case(pkg, syn, 256, 320, 7, 3, 9, 77)
    params = {'gamma': 2.0, 'alpha': 0.25}
    dgt = {k: cuda(v) for k, v in c['gt'].items()}
    ssd = _flat_ssd(pkg, c)
    _, extra = ssd.loss_sums(dgt, params, per_anchor=True)
    out = ssd.level_summaries(cls_losses=extra['cls_losses'], loc_losses=extra['loc_losses'], matches=extra['matches'])
    per_level = ssd.num_anchors_per_feature_map
    o = ossd.loss(c['anchors'], c['codes'], c['logits'], c['gt'], params, 7, return_all=True)
    per, lvl, total = obp.matches_summaries(o['matches'], per_level)
    assert np.array_equal(out['matches'].cpu().numpy(), per)
    close(out['mean_matches_per_image_on_level'].cpu().numpy(), lvl, rtol=1e-6)
    close(out['total_mean_matches_per_image'].item(), total, rtol=1e-6)
    assert np.array_equal(out['matches'].cpu().numpy(), ssd.matches_per_level(extra['matches']).cpu().numpy())
    for name, key in (('classification_losses', 'cls_losses'), ('localization_losses', 'loc_losses')):
        v = extra[key].cpu().numpy()                                   # selection is exact on the values it is given
        mean, kth, hist = obp.top_fraction_summaries(v, per_level)
        assert np.array_equal(out[name]['topk_kth'].cpu().numpy(), kth), name
        close(out[name]['topk_mean'].cpu().numpy(), mean, rtol=1e-6, atol=1e-12)
        close(out[name]['histogram_mean'].cpu().numpy(), hist, rtol=1e-6, atol=1e-12)
    # ties and zeros: a constant vector and an all-zero one
    const = torch.full([2, sum(per_level)], 0.25, device='cuda')
    r = ssd.level_summaries(cls_losses=const, loc_losses=torch.zeros_like(const))
    assert (r['classification_losses']['topk_mean'] == 0.25).all() and (r['classification_losses']['topk_kth'] == 0.25).all()
    assert (r['localization_losses']['topk_mean'] == 0).all()


def test_head_c_abi_argument_errors(pkg):
    """The C entry points reject malformed head descriptors with a status and a message (no crash, no fallback)."""
    import ctypes
    lib_mod = pkg._lib
    lib = lib_mod.load()
    ctx = lib_mod.context(0)
    x = torch.zeros([1, 2 * 3, 2, 2], device='cuda')
    bx = torch.zeros([1, 2 * 4, 2, 2], device='cuda')
    sums = torch.zeros(3, dtype=torch.float64, device='cuda')
    reg = torch.zeros([1, 8, 4], device='cuda'); cls = torch.zeros([1, 8], dtype=torch.int32, device='cuda')
    mat = torch.full([1, 8], -1, dtype=torch.int32, device='cuda')

    def desc(levels=1, n=2, fmt=1, h=2, w=2, cp=None, bp=None):
        d = lib_mod.SsdkHead()
        d.num_levels, d.anchors_per_location, d.data_format = levels, n, fmt
        d.height[0], d.width[0] = h, w
        d.class_predictions[0] = x.data_ptr() if cp is None else cp
        d.encoded_boxes[0] = bx.data_ptr() if bp is None else bp
        return d

    def call(d, A=8, C=3):
        return lib.ssdk_head_ssd_loss(ctx, ctypes.byref(d), reg.data_ptr(), cls.data_ptr(), mat.data_ptr(), 1, A, C, 2.0, 0.25,
                                      sums.data_ptr())
    assert call(desc()) == 0
    torch.cuda.synchronize()
    assert sums[2].item() == 0 and sums[0].item() == 0 and abs(sums[1].item() - 0.75 * 24 * 0.25 * np.log(2.0)) < 1e-6
    for bad, code in ((desc(levels=0), -1), (desc(levels=9), -1), (desc(n=0), -1), (desc(fmt=7), -1), (desc(h=3), -2),
                      (desc(cp=x.data_ptr() + 4), -2), (desc(cp=0), -1)):
        assert call(bad) == code, lib.ssdk_last_error()
        assert lib.ssdk_last_error()
    assert call(desc(), A=9) == -2 and b'anchors' in lib.ssdk_last_error()
    assert lib.ssdk_head_ssd_loss(ctx, None, reg.data_ptr(), cls.data_ptr(), mat.data_ptr(), 1, 8, 3, 2.0, 0.25, sums.data_ptr()) == -1
    assert lib.ssdk_head_detect(ctx, ctypes.byref(desc()), None, 1, 1, 8, 3, 0.05, 0.5, 10, None, float('-inf'), None, None, None, None,
                                None) == -1


def test_raw_predictions_stay_differentiable_and_inplace_edits_are_caught(pkg, syn):
    """ADVICE (round 1): user code that indexes raw_predictions as in the reference (auxiliary losses, regularisers) must get
    gradients back to the towers -- the lazily materialised tensors come from differentiable torch ops when a level requires
    grad -- and an in-place edit of a head tensor between SSD.loss() and backward() must raise (save_for_backward)."""
    torch.manual_seed(0)
    B, C, n = 2, 5, 6
    lv_cls = [torch.randn([B, n * C, h, w], device='cuda', requires_grad=True) for h, w in ((4, 5), (2, 3))]
    lv_box = [torch.randn([B, n * 4, h, w], device='cuda', requires_grad=True) for h, w in ((4, 5), (2, 3))]
    head = pkg.reshape_and_concatenate(lv_box, lv_cls, C, n)
    cp, eb = head['class_predictions'], head['encoded_boxes']
    assert cp.requires_grad and eb.requires_grad and cp.shape == (B, (20 + 6) * n, C)
    (cp.square().sum() + eb.sum()).backward()
    for t in lv_cls:
        assert torch.allclose(t.grad, 2 * t.detach())
    for t in lv_box:
        assert torch.equal(t.grad, torch.ones_like(t))
    # same values as the kernel path (no grad): ssdk_head_concat
    with torch.no_grad():
        plain = pkg.reshape_and_concatenate([t.detach() for t in lv_box], [t.detach() for t in lv_cls], C, n, lazy=False)
    assert torch.equal(plain['class_predictions'], cp.detach()) and torch.equal(plain['encoded_boxes'], eb.detach())
    # in-place edit between forward and backward
    c = _case(pkg, syn, 64, 96, 4, 2, 3, seed=4)
    lvl_c = [cuda(t).requires_grad_(True) for t in c['lv_classes']]
    lvl_b = [cuda(t).requires_grad_(True) for t in c['lv_boxes']]
    ssd = pkg.SSD.from_head_outputs(c['H'], c['W'], lvl_b, lvl_c, c['gen'], c['C'])
    dgt = {k: cuda(v) for k, v in c['gt'].items()}
    res = ssd.loss(dgt, {'gamma': 2.0, 'alpha': 0.25})
    with torch.no_grad():
        lvl_c[0].add_(1.0)
    with pytest.raises(RuntimeError):
        (res['classification_loss'] + res['localization_loss']).backward()


def test_two_streams_share_one_context_safely(pkg, syn):
    """ADVICE (round 1): the library keeps one context per (thread, device) and re-points its stream per call; workspaces are
    shared.  A prefetching SSD.assign_targets on a side stream next to SSD.loss() on the main stream (the pattern the docstring
    recommends), and the same entry point issued from two streams back to back, must give the single-stream results."""
    c = _case(pkg, syn, 128, 160, 6, 3, 5, seed=6)
    params = {'gamma': 2.0, 'alpha': 0.25}
    ssd = _head_ssd(pkg, c)
    c['dgt'] = {k: cuda(v) for k, v in c['gt'].items()}
    want = ssd.loss(c['dgt'], params)
    want_t = pkg.SSD.assign_targets(ssd.anchors, c['dgt'])
    torch.cuda.synchronize()
    side, side2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(20):
        with torch.cuda.stream(side):
            t1 = pkg.SSD.assign_targets(ssd.anchors, c['dgt'])
        got = ssd.loss(c['dgt'], params)
        with torch.cuda.stream(side2):
            t2 = pkg.SSD.assign_targets(ssd.anchors, c['dgt'])
            got2 = ssd.loss(c['dgt'], params)
        torch.cuda.synchronize()
        for t in (t1, t2):
            assert torch.equal(t['matches'], want_t['matches']) and torch.equal(t['reg_targets'], want_t['reg_targets'])
            assert t['count'].item() == want_t['count'].item()
        for g in (got, got2):
            assert float(g['classification_loss']) == float(want['classification_loss'])
            assert float(g['localization_loss']) == float(want['localization_loss'])
