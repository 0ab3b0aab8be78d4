"""Property tests of the oracle (hypothesis; CPU only, small sizes): invariants that the reference's algorithms imply and
that the GPU parity tests rely on -- IoU range and symmetry, encode/decode round trip, matching invariants of both threshold
branches and of the forced matches (incl. the row-id quirk), greedy NMS invariants (score order, pairwise IoU, idempotence,
K-truncation == prefix), and the head-layout transform being a pure permutation."""
import numpy as np
import pytest

from oracle import box_predictor as obp
from oracle import box_utils, nms
from oracle.training_target_creation import create_targets, match_boxes

hyp = pytest.importorskip('hypothesis')
from hypothesis import given, settings, strategies as st  # noqa: E402

f32 = np.float32
FAST = settings(max_examples=60, deadline=None)


def boxes_from(seed, n, degenerate=False):
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.0, 1.0, [n, 2])
    hw = np.exp(rng.uniform(np.log(0.01), np.log(0.7), [n, 2]))
    if degenerate and n:
        hw[rng.integers(0, n)] = 0.0
    b = np.concatenate([c - hw / 2, c + hw / 2], axis=1)
    return np.clip(b, 0.0, 1.0).astype(np.float32)


@FAST
@given(st.integers(0, 2 ** 31), st.integers(1, 12), st.integers(1, 40))
def test_iou_range_symmetry_and_self(seed, n, m):
    a, b = boxes_from(seed, n), boxes_from(seed + 1, m)
    s = box_utils.iou(a, b)
    assert s.shape == (n, m) and s.dtype == np.float32
    assert (s >= 0).all() and (s <= 1).all()
    assert np.array_equal(s, box_utils.iou(b, a).T)                       # every op is symmetric in its operands
    inter = box_utils.intersection(a, b)
    assert (inter <= np.minimum(box_utils.area(a)[:, None], box_utils.area(b)[None, :]) * f32(1 + 1e-6) + f32(1e-12)).all()
    big = box_utils.area(a) > 1e-4
    np.testing.assert_allclose(np.diag(box_utils.iou(a, a))[big], 1.0, rtol=1e-3)   # inter / (area + 1e-8)


@FAST
@given(st.integers(0, 2 ** 31), st.integers(1, 50))
def test_encode_decode_round_trip(seed, n):
    boxes, anchors = boxes_from(seed, n), boxes_from(seed + 7, n)
    keep = (box_utils.area(boxes) > 1e-5) & (box_utils.area(anchors) > 1e-5)
    back = box_utils.decode(box_utils.encode(boxes, anchors), anchors)
    np.testing.assert_allclose(back[keep], boxes[keep], atol=2e-5)


@FAST
@given(st.integers(0, 2 ** 31), st.integers(0, 8), st.integers(1, 120), st.sampled_from([(0.5, 0.5), (0.5, 0.4), (0.7, 0.3)]))
def test_matching_invariants(seed, n_gt, n_anchors, thr):
    pos, neg = thr
    gt, anchors = boxes_from(seed, n_gt, degenerate=True), boxes_from(seed + 3, n_anchors)
    if n_gt == 0:
        return
    m_plain = match_boxes(anchors, gt, pos, neg, force_match_groundtruth=False)
    m = match_boxes(anchors, gt, pos, neg, force_match_groundtruth=True)
    sim = box_utils.iou(gt, anchors)
    assert m.dtype == np.int32 and m.shape == (n_anchors,)
    assert ((m >= -2) & (m < n_gt)).all()
    if pos == neg:
        assert (m_plain != -2).all()                                      # the '==' branch never ignores (:94-96)
    best = sim.max(axis=0)
    first = sim.argmax(axis=0)                                            # first maximum, as tf.argmax
    assert np.array_equal(m_plain >= 0, best >= f32(pos))
    assert np.array_equal(m_plain[m_plain >= 0], first[m_plain >= 0])
    assert np.array_equal(m_plain == -2, (best < f32(pos)) & ~(f32(neg) > best))
    # forced matches only ever ADD matches or re-point them; every GT whose best IoU >= 0.1 owns its first-argmax anchor unless
    # a lower-indexed GT picked the same anchor (the row-id quirk, training_target_creation.py:117)
    fid = sim.argmax(axis=1)
    ok = sim.max(axis=1) >= f32(0.1)
    changed = np.nonzero(m != m_plain)[0]
    assert set(changed) <= set(fid[ok])
    for g_ in np.nonzero(ok)[0]:
        owners = np.nonzero(fid == fid[g_])[0]
        assert m[fid[g_]] == owners.min()
    reg, cls = create_targets(anchors, gt, np.arange(n_gt, dtype=np.int32), m)
    assert np.array_equal(cls > 0, m >= 0) and np.array_equal(cls[m >= 0] - 1, m[m >= 0])
    assert (reg[m < 0] == 0).all()


@FAST
@given(st.integers(0, 2 ** 31), st.integers(1, 80), st.integers(1, 12), st.sampled_from([0.3, 0.5, 0.7]))
def test_greedy_nms_invariants(seed, n, K, iou_thr):
    rng = np.random.default_rng(seed)
    boxes = boxes_from(seed, n, degenerate=True)
    scores = rng.permutation(np.linspace(0.01, 0.99, n)).astype(np.float32)          # tie-free
    thr = f32(0.2)
    kept_all = nms.non_max_suppression_v3(boxes, scores, n, iou_thr, thr, use_c=False)
    kept = nms.non_max_suppression_v3(boxes, scores, K, iou_thr, thr, use_c=False)
    assert np.array_equal(kept, kept_all[:K])                             # truncation at K == prefix of the untruncated run
    assert (scores[kept_all] > thr).all()                                 # strict '>'
    assert (np.diff(scores[kept_all]) < 0).all()                          # descending score order
    for i, a in enumerate(kept_all):
        for b in kept_all[:i]:
            assert not nms._iou_greater(boxes[a], boxes[b], f32(iou_thr))
    # every candidate that was dropped is suppressed by some better kept box
    cand = [i for i in np.argsort(-scores) if scores[i] > thr]
    for i in cand:
        if i not in set(kept_all):
            assert any(scores[k] > scores[i] and nms._iou_greater(boxes[i], boxes[k], f32(iou_thr)) for k in kept_all)
    # idempotence
    again = nms.non_max_suppression_v3(boxes[kept_all], scores[kept_all], n, iou_thr, thr, use_c=False)
    assert np.array_equal(again, np.arange(len(kept_all)))
    # the C restatement agrees with the Python one
    assert np.array_equal(nms.non_max_suppression_v3(boxes, scores, K, iou_thr, thr, use_c=True), kept)


@FAST
@given(st.integers(0, 2 ** 31), st.integers(1, 3), st.integers(1, 5), st.integers(1, 4),
       st.lists(st.tuples(st.integers(1, 5), st.integers(1, 6)), min_size=1, max_size=4),
       st.sampled_from(['channels_first', 'channels_last']))
def test_head_layout_is_a_permutation(seed, B, C, n, shapes, data_format):
    rng = np.random.default_rng(seed)
    A = sum(h * w * n for h, w in shapes)
    logits = rng.standard_normal([B, A, C]).astype(np.float32)
    codes = rng.standard_normal([B, A, 4]).astype(np.float32)
    lv_c = obp.split_to_levels(logits, shapes, n, data_format)
    lv_b = obp.split_to_levels(codes, shapes, n, data_format)
    for (h, w), t in zip(shapes, lv_c):
        assert t.shape == ((B, n * C, h, w) if data_format == 'channels_first' else (B, h, w, n * C))
    back = obp.reshape_and_concatenate(lv_b, lv_c, C, n, data_format)
    assert np.array_equal(back['class_predictions'], logits) and np.array_equal(back['encoded_boxes'], codes)
    # element (b, level, y, x, k, c) sits at channel k*C + c -- the address arithmetic of csrc/common.cuh::head_elem
    b, lvl = int(rng.integers(0, B)), int(rng.integers(0, len(shapes)))
    h, w = shapes[lvl]
    y, x, k, c = int(rng.integers(0, h)), int(rng.integers(0, w)), int(rng.integers(0, n)), int(rng.integers(0, C))
    a = sum(hh * ww * n for hh, ww in shapes[:lvl]) + (y * w + x) * n + k
    got = lv_c[lvl][b, k * C + c, y, x] if data_format == 'channels_first' else lv_c[lvl][b, y, x, k * C + c]
    assert got == logits[b, a, c]


@FAST
@given(st.integers(0, 2 ** 31), st.integers(1, 4), st.lists(st.integers(1, 40), min_size=1, max_size=4))
def test_top_fraction_summaries_invariants(seed, B, per_level):
    rng = np.random.default_rng(seed)
    v = np.abs(rng.standard_normal([B, sum(per_level)])).astype(np.float32)
    mean, kth, hist = obp.top_fraction_summaries(v, per_level)
    off = 0
    for i, n in enumerate(per_level):
        k = int(np.ceil(f32(n) * f32(0.2)))
        sl = v[:, off:off + n]
        assert (kth[:, i] <= sl.max(axis=1)).all() and (mean[:, i] >= kth[:, i] - 1e-6).all()
        assert ((sl >= kth[:, i:i + 1]).sum(axis=1) >= k).all() and ((sl > kth[:, i:i + 1]).sum(axis=1) < k).all()
        off += n
    np.testing.assert_allclose(hist, mean.mean(axis=0), rtol=1e-6)
