"""
Generates tests/golden/*.npz by executing the REFERENCE'S OWN, UNMODIFIED Python
sources (/root/reference/detector/...) on top of oracle/tf_numpy_shim (an eager
NumPy stand-in for TensorFlow, which cannot be installed here).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The fixtures are committed; the GPU box never runs this script.
"""
import hashlib
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'tf_numpy_shim'))

import tensorflow as tf  # noqa: E402  (the shim)
assert 'tf_numpy_shim' in tf.__file__
from detector import SSD  # noqa: E402  (reference code)
from detector.anchor_generator import AnchorGenerator  # noqa: E402
from detector.box_predictor import reshape_and_concatenate  # noqa: E402
from detector.losses import focal_loss, localization_loss  # noqa: E402
from detector.training_target_creation import create_targets, get_training_targets, match_boxes  # noqa: E402
from detector.utils import area, batch_decode, batch_multiclass_non_max_suppression, encode, intersection, iou  # noqa: E402
from detector.utils.box_utils import decode  # noqa: E402

syn = importlib.import_module('single-shot-detector_b200.synthetic')


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def T(x):
    return tf.constant(np.asarray(x))


def N(x):
    return np.asarray(x.a if hasattr(x, 'a') else x)


def save(name, **kw):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **kw)
    print('%-28s %8.1f KB' % (name, os.path.getsize(path) / 1024))


def ref_anchors(H, W, sm, ar=(1.0, 2.0, 0.5), strides=(8, 16, 32, 64, 128), scales=(32, 64, 128, 256, 512)):
    g = AnchorGenerator(strides=list(strides), scales=list(scales), scale_multipliers=list(sm),
                        aspect_ratios=list(ar))
    a = N(g(H, W))
    return g, a


# ---------------------------------------------------------------- anchors
def golden_anchors():
    out = {}
    cases = {
        'shipped_640x640': (640, 640, [1.0, 1.4142]),
        'nine_640x896': (640, 896, [1.0, 2 ** (1 / 3), 2 ** (2 / 3)]),
        'shipped_896x1344': (896, 1344, [1.0, 1.4142]),      # 1344/128 = 10.5 -> ceil
        'shipped_200x333': (200, 333, [1.0, 1.4142]),        # nothing divisible
    }
    for name, (H, W, sm) in cases.items():
        g, a = ref_anchors(H, W, sm)
        out[name + '/sha256'] = np.array(sha(a))
        out[name + '/shape'] = np.array(a.shape)
        out[name + '/per_map'] = np.array([int(v) for v in g.num_anchors_per_feature_map])
        out[name + '/head'] = a[:64]
        out[name + '/tail'] = a[-64:]
        out[name + '/args'] = np.array([H, W, len(sm)])
        out[name + '/sm'] = np.array(sm, np.float64)
    g, a = ref_anchors(200, 333, [1.0, 1.4142])
    out['shipped_200x333/full'] = a
    for i, r in enumerate(g.raw_anchors):
        out['shipped_200x333/raw%d' % i] = N(r)
    save('anchors', **out)


# ---------------------------------------------------------------- box utils
def golden_box_utils():
    rng = np.random.default_rng(7)
    b1 = syn.make_gt_boxes(rng, 37, 480, 640)
    _, anc = ref_anchors(128, 160, [1.0, 1.4142])
    sel = rng.choice(anc.shape[0], 300, replace=False)
    b2 = anc[np.sort(sel)]
    codes = rng.standard_normal([300, 4]).astype(np.float32) * 2
    pair = syn.make_gt_boxes(rng, 300, 480, 640)
    bcodes = rng.standard_normal([3, anc.shape[0], 4]).astype(np.float32)
    save('box_utils', b1=b1, b2=b2, codes=codes, pair=pair, anchors=anc, bcodes=bcodes,
         iou=N(iou(T(b1), T(b2))), intersection=N(intersection(T(b1), T(b2))),
         area=N(area(T(b2))), encode=N(encode(T(pair), T(b2))), decode=N(decode(T(codes), T(b2))),
         batch_decode=N(batch_decode(T(bcodes), T(anc))))


# ---------------------------------------------------------------- matching
def golden_matching():
    out = {}
    H, W = 256, 320
    _, anc = ref_anchors(H, W, [1.0, 1.4142])
    out['anchors'] = anc
    out['HW'] = np.array([H, W])
    rng = np.random.default_rng(11)
    cases = {}
    cases['random12'] = (syn.make_gt_boxes(rng, 12, H, W), rng.integers(0, 80, 12).astype(np.int32))
    cases['random40'] = (syn.make_gt_boxes(rng, 40, H, W), rng.integers(0, 80, 40).astype(np.int32))
    # tiny boxes inside one cell -> tied IoU across same-size anchors, low IoU (< 0.1 and >= 0.1 mix)
    tiny = np.array([[0.50, 0.50, 0.52, 0.52], [0.20, 0.70, 0.26, 0.74], [0.501, 0.501, 0.519, 0.519],
                     [0.0, 0.0, 0.01, 0.01], [0.9, 0.9, 1.0, 1.0]], np.float32)
    cases['tiny_ties'] = (tiny, np.array([1, 2, 3, 4, 5], np.int32))
    # force-match quirk (training_target_creation.py:117): GT0 and GT1 pick the same anchor;
    # GT0's best IoU < 0.1 (not okay), GT1's >= 0.1 (okay) -> the anchor is assigned GT0.
    # GT0 is a 1 x 3.6 px sliver: its best anchor a* has IoU ~0.0035 (< 0.1, "not okay").
    # GT1 is that same anchor a* shrunk to 75 % -> unique max at a*, IoU 0.5625 ("okay").
    g0 = np.array([[(124 + 14.8) / H, (164 + 12.2) / W, (124 + 15.8) / H, (164 + 15.8) / W]], np.float32)
    astar = anc[int(np.argmax(N(iou(T(g0), T(anc)))[0]))]
    cy, cx = (astar[0] + astar[2]) / 2, (astar[1] + astar[3]) / 2
    hh, hw = 0.375 * (astar[2] - astar[0]), 0.375 * (astar[3] - astar[1])
    quirk = np.array([g0[0], [cy - hh, cx - hw, cy + hh, cx + hw], [0.1, 0.1, 0.6, 0.7]], np.float32)
    cases['quirk'] = (quirk, np.array([7, 8, 9], np.int32))
    # a GT that overlaps nothing (outside every anchor is impossible; use zero-area-free far corner sliver)
    cases['single'] = (np.array([[0.3, 0.3, 0.7, 0.8]], np.float32), np.array([0], np.int32))
    cases['empty'] = (np.zeros([0, 4], np.float32), np.zeros([0], np.int32))
    for name, (gt, lab) in cases.items():
        out[name + '/gt'] = gt
        out[name + '/labels'] = lab
        for tag, (pt, nt) in {'p5n5': (0.5, 0.5), 'p5n4': (0.5, 0.4), 'p7n3': (0.7, 0.3)}.items():
            r, c, m = get_training_targets(T(anc), T(gt), T(lab), positives_threshold=pt,
                                           negatives_threshold=nt)
            out['%s/%s/reg' % (name, tag)] = N(r)
            out['%s/%s/cls' % (name, tag)] = N(c)
            out['%s/%s/matches' % (name, tag)] = N(m)
        if gt.shape[0]:
            m = match_boxes(T(anc), T(gt), positives_threshold=0.5, negatives_threshold=0.4,
                            force_match_groundtruth=False)
            out[name + '/noforce_p5n4/matches'] = N(m)
            r, c = create_targets(T(anc), T(gt), T(lab), m)
            out[name + '/noforce_p5n4/reg'] = N(r)
            out[name + '/noforce_p5n4/cls'] = N(c)
    # did the quirk case really hit the quirk?  (documented in the fixture itself)
    sim = N(iou(T(quirk), T(anc)))
    out['quirk/forced_ids'] = np.argmax(sim, axis=1)
    out['quirk/forced_vals'] = np.max(sim, axis=1)
    save('matching', **out)


# ---------------------------------------------------------------- full-size config 1 (hashes)
def golden_cfg1():
    cfg = syn.CONFIGS[1]
    H, W, C, G = cfg['H'], cfg['W'], cfg['C'], cfg['G']
    _, anc = ref_anchors(H, W, cfg['scale_multipliers'])
    gt = syn.make_groundtruth(1, 1, G, H, W, C)
    out = {'gt_boxes': gt['boxes'], 'gt_labels': gt['labels'], 'num_boxes': gt['num_boxes']}
    for tag, (pt, nt) in {'p5n5': (0.5, 0.5), 'p5n4': (0.5, 0.4)}.items():
        r, c, m = get_training_targets(T(anc), T(gt['boxes'][0]), T(gt['labels'][0]),
                                       positives_threshold=pt, negatives_threshold=nt)
        m = N(m)
        idx = np.nonzero(m != -1)[0].astype(np.int32)
        out[tag + '/nonbg_idx'] = idx
        out[tag + '/nonbg_matches'] = m[idx]
        out[tag + '/matches_sha256'] = np.array(sha(m))
        out[tag + '/cls_sha256'] = np.array(sha(N(c)))
        out[tag + '/reg_rows'] = N(r)[idx]
    save('cfg1_matching', **out)


# ---------------------------------------------------------------- losses through SSD.loss
class _FakeSSD(SSD):
    """Reference SSD with the (out-of-scope) network replaced by given head outputs."""

    def __init__(self, H, W, anchor_generator, raw, num_classes):
        images = T(np.broadcast_to(np.zeros([1, 1, 1, 1], np.float32), [raw['encoded_boxes'].a.shape[0], H, W, 3]))
        super().__init__(images, lambda x: None, anchor_generator, lambda f: raw, num_classes)


def golden_losses():
    out = {}
    H, W, C, B, G = 256, 320, 7, 3, 9
    sm = [1.0, 1.4142]
    gen = AnchorGenerator(scale_multipliers=sm)
    anc = N(gen(H, W))
    A = anc.shape[0]
    gt = syn.make_groundtruth(77, B, G, H, W, C, vary_count=True)
    gt['num_boxes'][1] = 0                                          # an image with no boxes
    logits = syn.make_logits('realistic', 77, B, A, C, anc, gt)
    logits[0, :50] = np.random.default_rng(5).uniform(-30, 30, [50, C]).astype(np.float32)  # extremes
    codes = syn.make_codes(77, B, A)
    out.update(HW=np.array([H, W]), C=np.array(C), anchors=anc, gt_boxes=gt['boxes'], gt_labels=gt['labels'],
               num_boxes=gt['num_boxes'], logits=logits, codes=codes)
    import detector.ssd as ssd_mod
    for tag, (pt, nt) in {'p5n5': (0.5, 0.5), 'p5n4': (0.5, 0.4)}.items():
        ssd_mod.POSITIVES_THRESHOLD, ssd_mod.NEGATIVES_THRESHOLD = pt, nt   # module constants (ssd.py:187-188)
        raw = {'encoded_boxes': T(codes), 'class_predictions': T(logits)}
        ssd = _FakeSSD(H, W, AnchorGenerator(scale_multipliers=sm), raw, C)
        groundtruth = {k: T(v) for k, v in gt.items()}
        for g_, a_ in [(2.0, 0.25), (1.5, 0.4)]:
            res = ssd.loss(groundtruth, {'gamma': g_, 'alpha': a_})
            out['%s/g%s_a%s/localization_loss' % (tag, g_, a_)] = N(res['localization_loss'])
            out['%s/g%s_a%s/classification_loss' % (tag, g_, a_)] = N(res['classification_loss'])
        reg, cls, mat = ssd._create_targets(groundtruth)
        reg, cls, mat = N(reg), N(cls), N(mat)
        out[tag + '/matches'] = mat
        out[tag + '/cls_targets'] = cls
        out[tag + '/reg_targets'] = reg
        onehot = np.eye(C + 1, dtype=np.float32)[cls][:, :, 1:]
        out[tag + '/cls_losses'] = N(focal_loss(T(logits), T(onehot), T((mat >= -1).astype(np.float32)),
                                                gamma=2.0, alpha=0.25))
        out[tag + '/loc_losses'] = N(localization_loss(T(codes), T(reg), T((mat >= 0).astype(np.float32))))
    ssd_mod.POSITIVES_THRESHOLD, ssd_mod.NEGATIVES_THRESHOLD = 0.5, 0.5
    save('losses', **out)


# ---------------------------------------------------------------- post-processing
def golden_postprocess():
    out = {}
    H, W, C, B, G = 256, 320, 6, 3, 8
    sm = [1.0, 1.4142]
    gen = AnchorGenerator(scale_multipliers=sm)
    anc = N(gen(H, W))
    A = anc.shape[0]
    gt = syn.make_groundtruth(88, B, G, H, W, C)
    logits = syn.make_logits('realistic', 88, B, A, C, anc, gt)
    # image 2: moderately dense scores so that NMS has real work and hits the per-class cap
    rng = np.random.default_rng(9)
    logits[2] = (rng.standard_normal([A, C]) * 1.5 - 3.0).astype(np.float32)
    codes = (syn.make_codes(88, B, A) * np.float32(0.5)).astype(np.float32)
    scores = N(tf.sigmoid(T(logits)))
    out.update(HW=np.array([H, W]), C=np.array(C), anchors=anc, logits=logits, codes=codes, scores=scores)
    for tag, (st, it, K) in {'s05_i5_k10': (0.05, 0.5, 10), 's15_i6_k25': (0.15, 0.6, 25),
                             's30_i3_k3': (0.30, 0.3, 3)}.items():
        b, s, c, n = batch_multiclass_non_max_suppression(T(codes), T(anc), T(scores), score_threshold=st,
                                                          iou_threshold=it, max_boxes_per_class=K)
        out[tag + '/boxes'] = N(b); out[tag + '/scores'] = N(s)
        out[tag + '/classes'] = N(c); out[tag + '/num'] = N(n)
        out[tag + '/params'] = np.array([st, it, K], np.float64)
    raw = {'encoded_boxes': T(codes), 'class_predictions': T(logits)}
    ssd = _FakeSSD(H, W, AnchorGenerator(scale_multipliers=sm), raw, C)
    p = ssd.get_predictions()   # defaults 0.05 / 0.5 / 20  (ssd.py:42)
    for k, v in p.items():
        out['get_predictions/' + k] = N(v)
    save('postprocess', **out)


# ---------------------------------------------------------------- head layout (reshape_and_concatenate)
def head_inputs(seed, B, C, n, shapes):
    """Seeded channels_first tower outputs; regenerated identically by the tests (never stored for the big case)."""
    rng = np.random.default_rng(seed)
    boxes = [rng.standard_normal([B, n * 4, h, w]).astype(np.float32) for h, w in shapes]
    classes = [rng.standard_normal([B, n * C, h, w]).astype(np.float32) for h, w in shapes]
    return boxes, classes


def golden_head():
    out = {}
    # tiny case, stored in full
    B, C, n, shapes = 2, 3, 2, [(3, 4), (2, 2), (1, 1)]
    boxes, classes = head_inputs(21, B, C, n, shapes)
    r = reshape_and_concatenate([T(b) for b in boxes], [T(c) for c in classes], C, n)
    out['tiny/params'] = np.array([B, C, n])
    out['tiny/shapes'] = np.array(shapes)
    for i in range(len(shapes)):
        out['tiny/boxes%d' % i] = boxes[i]
        out['tiny/classes%d' % i] = classes[i]
    out['tiny/encoded_boxes'] = N(r['encoded_boxes'])
    out['tiny/class_predictions'] = N(r['class_predictions'])
    # the geometry of the 'losses' fixture (256x320, 6 anchors per location, 7 classes), hashes only
    B, C, n, shapes = 3, 7, 6, [(32, 40), (16, 20), (8, 10), (4, 5), (2, 3)]
    boxes, classes = head_inputs(22, B, C, n, shapes)
    r = reshape_and_concatenate([T(b) for b in boxes], [T(c) for c in classes], C, n)
    out['big/params'] = np.array([B, C, n])
    out['big/shapes'] = np.array(shapes)
    out['big/encoded_boxes_sha256'] = np.array(sha(N(r['encoded_boxes'])))
    out['big/class_predictions_sha256'] = np.array(sha(N(r['class_predictions'])))
    out['big/class_predictions_head'] = N(r['class_predictions'])[:, :40]
    save('head', **out)


# ---------------------------------------------------------------- random-crop box ops (input_pipeline/random_image_crop.py)
def golden_crop():
    import importlib.util
    # the file is loaded by path: detector/input_pipeline/__init__.py pulls in the tf.data pipeline, which the shim does not cover
    spec = importlib.util.spec_from_file_location('ref_random_image_crop', '/root/reference/detector/input_pipeline/random_image_crop.py')
    ric = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ric)
    rng = np.random.default_rng(31)
    out = {}
    for case in range(6):
        n = [12, 40, 1, 7, 25, 3][case]
        boxes = syn.make_gt_boxes(rng, n, 480, 640)
        c = rng.uniform(0.2, 0.8, 2)
        hw = rng.uniform(0.15, 0.6, 2)
        window = np.clip(np.array([c[0] - hw[0], c[1] - hw[1], c[0] + hw[0], c[1] + hw[1]]), 0, 1).astype(np.float32)
        if case == 3:
            window = np.array([0.0, 0.0, 1.0, 1.0], np.float32)
        thr = [0.3, 0.3, 0.3, 0.3, 0.6, 0.05][case]
        pre = 'c%d/' % case
        out[pre + 'boxes'], out[pre + 'window'], out[pre + 'thr'] = boxes, window, np.float32(thr)
        b1, i1 = ric.prune_completely_outside_window(T(boxes), T(window))
        out[pre + 'outside_boxes'], out[pre + 'outside_idx'] = N(b1).reshape(-1, 4), N(i1).reshape(-1)
        b2, i2 = ric.prune_non_overlapping_boxes(b1, tf.expand_dims(T(window), 0), min_overlap=thr)
        out[pre + 'overlap_boxes'], out[pre + 'overlap_idx'] = N(b2).reshape(-1, 4), N(i2).reshape(-1)
        out[pre + 'changed'] = N(ric.change_coordinate_frame(b2, T(window))).reshape(-1, 4)
        out[pre + 'keep'] = N(tf.gather(i1, i2)).reshape(-1)
        others = syn.make_gt_boxes(rng, 5, 480, 640)
        out[pre + 'others'] = others
        out[pre + 'ioa'] = N(ric.ioa(T(others), T(boxes)))
        b3, i3 = ric.prune_non_overlapping_boxes(T(boxes), T(others), min_overlap=0.25)
        out[pre + 'multi_boxes'], out[pre + 'multi_idx'] = N(b3).reshape(-1, 4), N(i3).reshape(-1)
    save('crop', **out)


if __name__ == '__main__':
    golden_crop()
    golden_head()
    golden_anchors()
    golden_box_utils()
    golden_matching()
    golden_cfg1()
    golden_losses()
    golden_postprocess()
