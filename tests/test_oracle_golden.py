"""The NumPy oracle vs. fixtures produced by the reference's own code (tests/golden/make_golden.py).

Transcendental-free outputs (anchors, IoU, matches, labels, kept sets) must be bit-exact;
outputs downstream of exp/log are compared at a few ulp."""
import hashlib

import numpy as np
import pytest

from oracle import box_utils, losses, nms, ssd
from oracle.anchor_generator import AnchorGenerator
from oracle.training_target_creation import create_targets, get_training_targets, match_boxes


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


ANCHOR_CASES = ['shipped_640x640', 'nine_640x896', 'shipped_896x1344', 'shipped_200x333']


@pytest.mark.parametrize('name', ANCHOR_CASES)
def test_anchors_bit_exact(golden, name):
    g = golden('anchors')
    H, W, _ = g[name + '/args']
    gen = AnchorGenerator(scale_multipliers=list(g[name + '/sm']))
    a = gen(int(H), int(W))
    assert a.dtype == np.float32 and tuple(a.shape) == tuple(g[name + '/shape'])
    assert list(gen.num_anchors_per_feature_map) == list(g[name + '/per_map'])
    assert np.array_equal(a[:64], g[name + '/head']) and np.array_equal(a[-64:], g[name + '/tail'])
    assert sha(a) == str(g[name + '/sha256'])
    if name == 'shipped_200x333':
        assert np.array_equal(a, g[name + '/full'])
        for i, r in enumerate(gen.raw_anchors):
            assert np.array_equal(r, g[name + '/raw%d' % i])


def test_known_answers_from_survey():
    a = AnchorGenerator()(640, 640)
    assert a.shape == (51150, 4)
    assert np.array_equal(a[0] * 640, np.array([-12, -12, 20, 20], np.float32))
    np.testing.assert_allclose(a[1] * 640, [-7.3137083, -18.627417, 15.313709, 26.627417], rtol=1e-7)
    b = np.array([[0.1, 0.2, 0.5, 0.9]], np.float32)
    assert np.array_equal(box_utils.encode(b, b), np.zeros([1, 4], np.float32))
    small = np.array([[0.5, 0.5, 0.5001, 0.5001]], np.float32)
    assert box_utils.iou(small, small)[0, 0] < 1.0     # area/(area+1e-8)


def test_box_utils(golden):
    g = golden('box_utils')
    assert np.array_equal(box_utils.iou(g['b1'], g['b2']), g['iou'])
    assert np.array_equal(box_utils.intersection(g['b1'], g['b2']), g['intersection'])
    assert np.array_equal(box_utils.area(g['b2']), g['area'])
    assert np.array_equal(box_utils.encode(g['pair'], g['b2']), g['encode'])
    assert np.array_equal(box_utils.decode(g['codes'], g['b2']), g['decode'])
    assert np.array_equal(box_utils.batch_decode(g['bcodes'], g['anchors']), g['batch_decode'])


MATCH_CASES = ['random12', 'random40', 'tiny_ties', 'quirk', 'single', 'empty']
THR = {'p5n5': (0.5, 0.5), 'p5n4': (0.5, 0.4), 'p7n3': (0.7, 0.3)}


@pytest.mark.parametrize('case', MATCH_CASES)
@pytest.mark.parametrize('tag', list(THR))
def test_matching(golden, case, tag):
    g = golden('matching')
    anc, gt, lab = g['anchors'], g[case + '/gt'], g[case + '/labels']
    pt, nt = THR[tag]
    reg, cls, m = get_training_targets(anc, gt, lab, pt, nt)
    assert m.dtype == np.int32 and cls.dtype == np.int32 and reg.dtype == np.float32
    assert np.array_equal(m, g['%s/%s/matches' % (case, tag)])
    assert np.array_equal(cls, g['%s/%s/cls' % (case, tag)])
    assert np.array_equal(reg, g['%s/%s/reg' % (case, tag)])


@pytest.mark.parametrize('case', [c for c in MATCH_CASES if c != 'empty'])
def test_matching_no_force(golden, case):
    g = golden('matching')
    anc, gt, lab = g['anchors'], g[case + '/gt'], g[case + '/labels']
    m = match_boxes(anc, gt, 0.5, 0.4, force_match_groundtruth=False)
    assert np.array_equal(m, g[case + '/noforce_p5n4/matches'])
    reg, cls = create_targets(anc, gt, lab, m)
    assert np.array_equal(cls, g[case + '/noforce_p5n4/cls'])
    assert np.array_equal(reg, g[case + '/noforce_p5n4/reg'])


def test_quirk_fixture_really_hits_the_quirk(golden):
    """GT0 and GT1 force the same anchor; GT0 is below 0.1 IoU, GT1 is not -> the anchor gets GT0."""
    g = golden('matching')
    ids, vals = g['quirk/forced_ids'], g['quirk/forced_vals']
    assert ids[0] == ids[1] and vals[0] < 0.1 <= vals[1]
    assert g['quirk/p5n5/matches'][ids[0]] == 0
    assert g['quirk/noforce_p5n4/matches'][ids[0]] != 0


@pytest.mark.parametrize('tag', ['p5n5', 'p5n4'])
def test_cfg1_full_size(golden, syn, tag):
    g = golden('cfg1_matching')
    cfg = syn.CONFIGS[1]
    anc = AnchorGenerator(scale_multipliers=cfg['scale_multipliers'])(cfg['H'], cfg['W'])
    gt = syn.make_groundtruth(1, 1, cfg['G'], cfg['H'], cfg['W'], cfg['C'])
    assert np.array_equal(gt['boxes'], g['gt_boxes']), 'synthetic generator drifted from the fixture'
    pt, nt = THR[tag]
    reg, cls, m = get_training_targets(anc, gt['boxes'][0], gt['labels'][0], pt, nt)
    assert sha(m) == str(g[tag + '/matches_sha256'])
    assert sha(cls) == str(g[tag + '/cls_sha256'])
    idx = g[tag + '/nonbg_idx']
    assert np.array_equal(m[idx], g[tag + '/nonbg_matches'])
    assert np.array_equal(reg[idx], g[tag + '/reg_rows'])


@pytest.mark.parametrize('tag', ['p5n5', 'p5n4'])
def test_losses(golden, tag):
    g = golden('losses')
    gt = {'boxes': g['gt_boxes'], 'labels': g['gt_labels'], 'num_boxes': g['num_boxes']}
    C = int(g['C'])
    pt, nt = THR[tag]
    for gamma, alpha in [(2.0, 0.25), (1.5, 0.4)]:
        r = ssd.loss(g['anchors'], g['codes'], g['logits'], gt, {'gamma': gamma, 'alpha': alpha}, C,
                     pt, nt, return_all=True)
        key = '%s/g%s_a%s/' % (tag, gamma, alpha)
        assert r['localization_loss'] == g[key + 'localization_loss']
        assert r['classification_loss'] == g[key + 'classification_loss']
        assert np.array_equal(r['matches'], g[tag + '/matches'])
        assert np.array_equal(r['cls_targets'], g[tag + '/cls_targets'])
        assert np.array_equal(r['reg_targets'], g[tag + '/reg_targets'])
        if gamma == 2.0:
            assert np.array_equal(r['cls_losses'], g[tag + '/cls_losses'])
            assert np.array_equal(r['loc_losses'], g[tag + '/loc_losses'])
    assert (g[tag + '/matches'] == -2).any() == (tag == 'p5n4')


@pytest.mark.parametrize('tag', ['s05_i5_k10', 's15_i6_k25', 's30_i3_k3'])
@pytest.mark.parametrize('use_c', [True, False])
def test_postprocess(golden, tag, use_c):
    g = golden('postprocess')
    st, it, K = g[tag + '/params']
    b, s, c, n = nms.batch_multiclass_non_max_suppression(g['codes'], g['anchors'], g['scores'], st, it,
                                                          int(K), use_c=use_c)
    assert np.array_equal(n, g[tag + '/num']) and n.sum() > 0
    assert np.array_equal(c, g[tag + '/classes'])
    assert np.array_equal(s, g[tag + '/scores'])
    assert np.array_equal(b, g[tag + '/boxes'])


def test_get_predictions(golden):
    g = golden('postprocess')
    p = ssd.get_predictions(g['anchors'], g['codes'], g['logits'])
    for k in ['boxes', 'labels', 'scores', 'num_boxes']:
        assert np.array_equal(p[k], g['get_predictions/' + k]), k
    assert np.array_equal(losses.sigmoid(g['logits']), g['scores'])


def _tf_vectors():
    import importlib.util, os
    spec = importlib.util.spec_from_file_location('tf_nms_vectors', os.path.join(os.path.dirname(__file__), 'golden', 'tf_nms_vectors.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize('use_c', [True, False])
def test_nms_v3_tensorflow_published_vectors(use_c):
    """NonMaxSuppressionV3 (called at reference detector/utils/nms.py:33) pinned to TensorFlow's own unit-test vectors
    (tests/golden/tf_nms_vectors.py restates non_max_suppression_op_test.cc)."""
    tfv = _tf_vectors()
    seen = 0
    for name, boxes, scores, k, iou, thr, want in tfv.cases():
        got = nms.non_max_suppression_v3(boxes, scores, k, iou, thr, use_c=use_c)
        assert np.array_equal(got, want), (name, got, want)
        # the same case through the reference's call site (nms.py:31-44): one class
        sb, ss, sc, si = nms.multiclass_non_max_suppression(boxes, scores.reshape(-1, 1), thr, iou, k, use_c=use_c,
                                                            return_indices=True)
        assert np.array_equal(si, want) and np.array_equal(sb, boxes[want]) and np.array_equal(ss, scores[want]), name
        seen += 1
    assert seen == 8
    b, s, k, iou, thr = tfv.INVALID_IOU_THRESHOLD
    with pytest.raises(ValueError):
        nms.non_max_suppression_v3(np.asarray(b, np.float32), np.asarray(s, np.float32), k, iou, thr, use_c=use_c)


def test_nms_restatement_vs_torchvision():
    """Independent cross-check of the NonMaxSuppressionV3 restatement (kept sets)."""
    torch = pytest.importorskip('torch')
    tv = pytest.importorskip('torchvision')
    rng = np.random.default_rng(3)
    for trial in range(5):
        n = 400
        ctr = rng.uniform(0, 1, (n, 2)); sz = rng.uniform(0.02, 0.3, (n, 2))
        b = np.concatenate([ctr - sz / 2, ctr + sz / 2], 1).astype(np.float32).clip(0, 1)
        s = rng.uniform(0, 1, n).astype(np.float32)
        mine = nms.non_max_suppression_v3(b, s, 50, 0.5, 0.05)
        keep = tv.ops.nms(torch.from_numpy(b[:, [1, 0, 3, 2]]), torch.from_numpy(s), 0.5).numpy()
        keep = np.array([k for k in keep if s[k] > 0.05][:50])
        assert np.array_equal(mine, keep)
        assert np.array_equal(mine, nms.non_max_suppression_v3(b, s, 50, 0.5, 0.05, use_c=False))


# ------------------------------------------------------------------------------------------------ gradient oracle
def _small_train_case(syn, seed=5, B=2, C=4, G=3, H=64, W=96):
    from oracle.anchor_generator import AnchorGenerator as OracleGen
    anchors = OracleGen()(H, W)
    A = anchors.shape[0]
    gt = syn.make_groundtruth(seed, B, G, H, W, C)
    logits = syn.make_logits('realistic', seed, B, A, C, anchors, gt)
    codes = (syn.make_codes(seed, B, A) * np.float32(0.7)).astype(np.float32)
    return anchors, gt, logits, codes, C


@pytest.mark.parametrize('thr', [(0.5, 0.5), (0.5, 0.4)])
def test_gradient_oracle_forward64_matches_pinned_forward(syn, thr):
    """forward64 (the function the gradient oracle differentiates) == the float32 forward oracle, which the golden
    fixtures pin to the reference's own code."""
    from oracle import losses_grad as og, ssd as ossd
    anchors, gt, logits, codes, C = _small_train_case(syn)
    params = {'gamma': 2.0, 'alpha': 0.25}
    want = ossd.loss(anchors, codes, logits, gt, params, C, positives_threshold=thr[0], negatives_threshold=thr[1])
    loc, cls = og.forward64(anchors, codes, logits, gt, params, C, positives_threshold=thr[0], negatives_threshold=thr[1])
    assert abs(loc - float(want['localization_loss'])) <= 2e-6 * abs(loc)
    assert abs(cls - float(want['classification_loss'])) <= 2e-6 * abs(cls)


@pytest.mark.parametrize('gamma,alpha', [(2.0, 0.25), (1.5, 0.4)])
def test_gradient_oracle_matches_finite_differences(syn, gamma, alpha):
    from oracle import losses_grad as og
    anchors, gt, logits, codes, C = _small_train_case(syn)
    params = {'gamma': gamma, 'alpha': alpha}
    up = (0.7, 1.3)
    parts = og._parts(anchors, gt, C, 0.5, 0.4)
    g = og.ssd_loss_grad(anchors, codes, logits, gt, params, C, upstream=up, parts=parts)
    assert g['num_matches'] > 0 and (parts[3] == 0).any(), 'case must contain matched and ignored anchors'

    def total(lg, cd):
        loc, cls = og.forward64(anchors, cd, lg, gt, params, C, parts=parts)
        return up[0] * loc + up[1] * cls
    rng = np.random.default_rng(0)
    x64, c64 = logits.astype(np.float64), codes.astype(np.float64)
    matched_rows = np.argwhere(parts[2] > 0)
    picks = [tuple(rng.integers(0, s) for s in x64.shape) for _ in range(40)]
    picks += [(b, a, int(np.argmax(parts[1][b, a]))) for b, a in matched_rows[:20]]            # positive-class elements
    h = 1e-5
    for idx in picks:
        xp, xm = x64.copy(), x64.copy()
        xp[idx] += h; xm[idx] -= h
        fd = (total(xp, c64) - total(xm, c64)) / (2 * h)
        assert abs(fd - g['class_predictions'][idx]) <= 1e-6 * abs(fd) + 3e-10, (idx, fd, g['class_predictions'][idx])   # 3e-10: float64 round-off of the O(1) total / h
    for b, a in list(matched_rows[:10]) + [(0, 0)]:
        for k in range(4):
            cp, cm = c64.copy(), c64.copy()
            cp[b, a, k] += h; cm[b, a, k] -= h
            fd = (total(x64, cp) - total(x64, cm)) / (2 * h)
            assert abs(fd - g['encoded_boxes'][b, a, k]) <= 1e-6 * abs(fd) + 3e-10


# ---------------------------------------------------------------- head layout (box_predictor.py:67-104)
def _head_inputs(seed, B, C, n, shapes):
    """Same generator as tests/golden/make_golden.py::head_inputs."""
    rng = np.random.default_rng(seed)
    boxes = [rng.standard_normal([B, n * 4, h, w]).astype(np.float32) for h, w in shapes]
    classes = [rng.standard_normal([B, n * C, h, w]).astype(np.float32) for h, w in shapes]
    return boxes, classes


def test_reshape_and_concatenate_matches_the_reference(golden):
    from oracle import box_predictor as obp
    g = golden('head')
    B, C, n = [int(v) for v in g['tiny/params']]
    shapes = [tuple(int(v) for v in s) for s in g['tiny/shapes']]
    boxes = [g['tiny/boxes%d' % i] for i in range(len(shapes))]
    classes = [g['tiny/classes%d' % i] for i in range(len(shapes))]
    r = obp.reshape_and_concatenate(boxes, classes, C, n)
    assert np.array_equal(r['encoded_boxes'], g['tiny/encoded_boxes'])
    assert np.array_equal(r['class_predictions'], g['tiny/class_predictions'])
    B, C, n = [int(v) for v in g['big/params']]
    shapes = [tuple(int(v) for v in s) for s in g['big/shapes']]
    boxes, classes = _head_inputs(22, B, C, n, shapes)
    r = obp.reshape_and_concatenate(boxes, classes, C, n)
    assert hashlib.sha256(r['encoded_boxes'].tobytes()).hexdigest() == str(g['big/encoded_boxes_sha256'])
    assert hashlib.sha256(r['class_predictions'].tobytes()).hexdigest() == str(g['big/class_predictions_sha256'])
    assert np.array_equal(r['class_predictions'][:, :40], g['big/class_predictions_head'])
    # the inverse used to manufacture tower-shaped inputs, both data formats
    for fmt in ('channels_first', 'channels_last'):
        lv_b = obp.split_to_levels(r['encoded_boxes'], shapes, n, fmt)
        lv_c = obp.split_to_levels(r['class_predictions'], shapes, n, fmt)
        back = obp.reshape_and_concatenate(lv_b, lv_c, C, n, fmt)
        assert np.array_equal(back['encoded_boxes'], r['encoded_boxes'])
        assert np.array_equal(back['class_predictions'], r['class_predictions'])
    assert all(np.array_equal(a, b) for a, b in zip(obp.split_to_levels(r['class_predictions'], shapes, n), classes))
    assert obp.level_shapes(896, 1344, [8, 16, 32, 64, 128])[-1] == (7, 11)          # 1344/128 = 10.5 -> ceil


def test_summary_oracle_known_answers():
    from oracle import box_predictor as obp
    v = np.array([[5, 1, 4, 2, 3, 9, 8, 7, 6, 0, 10, 11]], np.float32)
    mean, kth, hist = obp.top_fraction_summaries(v, [10, 2], top_fraction=0.20)      # k = ceil(2.0) = 2, ceil(0.4) = 1
    assert mean.tolist() == [[8.5, 11.0]] and kth.tolist() == [[8.0, 11.0]] and hist.tolist() == [8.5, 11.0]
    m = np.array([[0, -1, 3, -2, -1, 2], [-1, -1, -1, 1, -2, -1]], np.int32)
    per, lvl, total = obp.matches_summaries(m, [4, 2])
    assert per.tolist() == [[2.0, 1.0], [1.0, 0.0]] and lvl.tolist() == [1.5, 0.5] and float(total) == 2.0


# ---------------------------------------------------------------- random-crop box ops (input_pipeline/random_image_crop.py)
def test_crop_box_ops_match_the_reference(golden):
    from oracle import random_image_crop as oric
    g = golden('crop')
    for case in range(6):
        pre = 'c%d/' % case
        boxes, window, thr = g[pre + 'boxes'], g[pre + 'window'], float(g[pre + 'thr'])
        b1, i1 = oric.prune_completely_outside_window(boxes, window)
        assert np.array_equal(b1, g[pre + 'outside_boxes']) and np.array_equal(i1, g[pre + 'outside_idx'])
        b2, i2 = oric.prune_non_overlapping_boxes(b1, window[None], thr)
        assert np.array_equal(b2, g[pre + 'overlap_boxes']) and np.array_equal(i2, g[pre + 'overlap_idx'])
        assert np.array_equal(oric.change_coordinate_frame(b2, window), g[pre + 'changed'])
        cb, keep = oric.crop_boxes(boxes, window, thr)
        assert np.array_equal(cb, g[pre + 'changed']) and np.array_equal(keep, g[pre + 'keep'])
        assert np.array_equal(oric.ioa(g[pre + 'others'], boxes), g[pre + 'ioa'])
        b3, i3 = oric.prune_non_overlapping_boxes(boxes, g[pre + 'others'], 0.25)
        assert np.array_equal(b3, g[pre + 'multi_boxes']) and np.array_equal(i3, g[pre + 'multi_idx'])
    assert sum(len(g['c%d/keep' % c]) for c in range(6)) > 10 and len(g['c3/keep']) == 7      # full-image window keeps everything
