"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/ssdk.h declares (no compute
calls without a GPU), the Python mirror has the reference's call surface, config loading, image sharding, and the
N>1 path (world_size-2 gloo): shard -> per-shard sums -> all-reduce -> normalise == the full-batch losses."""
import ctypes
import inspect
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_pkg

HEADER = os.path.join(ROOT, 'include', 'ssdk.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'SSDK_API\s+[\w\s\*]+?\b(ssdk_\w+)\s*\(', text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ('ssdk_anchors', 'ssdk_match_boxes', 'ssdk_training_targets', 'ssdk_create_targets', 'ssdk_ssd_loss',
                 'ssdk_loss_finalize', 'ssdk_ssd_targets_and_loss', 'ssdk_ssd_targets_and_loss_host', 'ssdk_postprocess',
                 'ssdk_postprocess_host', 'ssdk_iou', 'ssdk_encode', 'ssdk_decode', 'ssdk_focal_loss',
                 'ssdk_localization_loss'):
        assert must in syms
    assert len(syms) >= 29


def test_library_exports_every_declared_symbol():
    lib_mod = load_pkg('_lib')
    assert os.path.exists(lib_mod.LIB_PATH), 'libssdk.so not built: run __graft_entry__.build()'
    lib = ctypes.CDLL(lib_mod.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes table mirrors the header declaration by declaration
    assert sorted(lib_mod._SIGNATURES) == declared_symbols()
    lib.ssdk_version.restype = ctypes.c_int
    assert lib.ssdk_version() >= 100


def test_header_is_plain_c():
    """The boundary must compile as C (no C++ / torch types in the signatures)."""
    src = '#include "ssdk.h"\nint main(void) { ssdk_status s = SSDK_OK; ssdk_ctx* c = 0; (void)c; return (int)s; }\n'
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-fsyntax-only', '-I', os.path.join(ROOT, 'include'), '-x', 'c', '-'],
                       input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr


def test_ctypes_structs_match_the_c_layout(tmp_path):
    """ssdk_head / ssdk_head_grads are passed by pointer: the ctypes mirrors must have the C compiler's layout."""
    lib_mod = load_pkg('_lib')
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "ssdk.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %d\\n", '
           'sizeof(ssdk_head), offsetof(ssdk_head, height), offsetof(ssdk_head, class_predictions), '
           'offsetof(ssdk_head, encoded_boxes), sizeof(ssdk_head_grads), SSDK_MAX_LEVELS); return 0; }\n')
    exe = str(tmp_path / 'layout')
    r = subprocess.run(['gcc', '-std=c99', '-I', os.path.join(ROOT, 'include'), '-x', 'c', '-', '-o', exe], input=src, text=True,
                       capture_output=True)
    assert r.returncode == 0, r.stderr
    got = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    H, G = lib_mod.SsdkHead, lib_mod.SsdkHeadGrads
    assert got == [ctypes.sizeof(H), H.height.offset, H.class_predictions.offset, H.encoded_boxes.offset, ctypes.sizeof(G),
                   lib_mod.SSDK_MAX_LEVELS]


def build_c_example(tmp_path):
    """examples/c_abi_example.c: a plain-C (gcc -std=c99) user of the library -- no C++, torch or Python on the boundary."""
    lib_dir = os.path.dirname(load_pkg('_lib').LIB_PATH)
    exe = str(tmp_path / 'ssdk_example')
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), '-I', '/usr/local/cuda/include',
                        os.path.join(ROOT, 'examples', 'c_abi_example.c'), '-L', lib_dir, '-lssdk', '-L', '/usr/local/cuda/lib64',
                        '-lcudart', '-lm', '-Wl,-rpath,' + lib_dir, '-o', exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_plain_c_program_links_and_runs_the_host_part(tmp_path):
    import torch
    exe = build_c_example(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'anchors: 107415 (per level 80640 20160 5040 1260 315)' in r.stdout
    if not torch.cuda.is_available():
        assert 'no CPU fallback' in r.stdout                            # refuses, loudly, to compute without a GPU


def test_host_only_entry_points_and_error_reporting():
    """ssdk_num_anchors is pure shape arithmetic (anchor_generator.py:59-62) and needs no GPU; a compute entry point
    without a device must fail loudly with a message, never fall back."""
    lib = load_pkg('_lib').load()
    strides = (ctypes.c_int * 5)(8, 16, 32, 64, 128)
    total = ctypes.c_int64(0)
    per = (ctypes.c_int32 * 5)()
    assert lib.ssdk_num_anchors(640, 896, ctypes.cast(strides, ctypes.c_void_p), 5, 9, ctypes.byref(total),
                                ctypes.cast(per, ctypes.c_void_p)) == 0
    assert total.value == 107415 and list(per) == [80640, 20160, 5040, 1260, 315]
    assert lib.ssdk_num_anchors(896, 1344, ctypes.cast(strides, ctypes.c_void_p), 5, 6, ctypes.byref(total), None) == 0
    assert total.value == 150402                                     # 1344/128 = 10.5 -> ceil
    assert lib.ssdk_num_anchors(0, 10, ctypes.cast(strides, ctypes.c_void_p), 5, 6, ctypes.byref(total), None) == -1
    assert b'bad arguments' in lib.ssdk_last_error()
    import torch
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        st = lib.ssdk_ctx_create(0, None, ctypes.byref(h))
        assert st == -3 and b'no CPU fallback' in lib.ssdk_last_error()
        pkg = load_pkg()
        with pytest.raises(pkg._lib.SsdkError):
            pkg.AnchorGenerator()(640, 640)
        with pytest.raises(pkg._lib.SsdkError):
            pkg.iou(np.zeros([2, 4], np.float32), np.zeros([3, 4], np.float32))


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, 'single-shot-detector_b200')
    for d, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(d, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), os.path.join(d, f)
                assert 'liboracle' not in text, os.path.join(d, f)


REFERENCE_SURFACE = {
    # name -> (positional parameter names, defaults) as in the reference
    'get_training_targets': (['anchors', 'groundtruth_boxes', 'groundtruth_labels', 'positives_threshold', 'negatives_threshold'],
                             {'positives_threshold': 0.5, 'negatives_threshold': 0.4}),
    'match_boxes': (['anchors', 'groundtruth_boxes', 'positives_threshold', 'negatives_threshold', 'force_match_groundtruth'],
                    {'positives_threshold': 0.5, 'negatives_threshold': 0.4, 'force_match_groundtruth': True}),
    'create_targets': (['anchors', 'groundtruth_boxes', 'groundtruth_labels', 'matches'], {}),
    'localization_loss': (['predictions', 'targets', 'weights'], {}),
    'focal_loss': (['predictions', 'targets', 'weights', 'gamma', 'alpha'], {'gamma': 2.0, 'alpha': 0.25}),
    'iou': (['boxes1', 'boxes2'], {}),
    'intersection': (['boxes1', 'boxes2'], {}),
    'area': (['boxes'], {}),
    'encode': (['boxes', 'anchors'], {}),
    'decode': (['codes', 'anchors'], {}),
    'batch_decode': (['box_encodings', 'anchors'], {}),
    'multiclass_non_max_suppression': (['boxes', 'scores', 'score_threshold', 'iou_threshold', 'max_boxes_per_class'], {}),
    'batch_multiclass_non_max_suppression': (['encoded_boxes', 'anchors', 'scores', 'score_threshold', 'iou_threshold',
                                              'max_boxes_per_class'], {}),
    'reshape_and_concatenate': (['encoded_boxes', 'class_predictions', 'num_classes', 'num_anchors_per_location'], {}),
    'ioa': (['boxes1', 'boxes2'], {}),
    'prune_completely_outside_window': (['boxes', 'window'], {}),
    'prune_non_overlapping_boxes': (['boxes1', 'boxes2', 'min_overlap'], {}),
    'change_coordinate_frame': (['boxes', 'window'], {}),
}


@pytest.mark.parametrize('name', sorted(REFERENCE_SURFACE))
def test_mirror_keeps_the_reference_signatures(name):
    pkg = load_pkg()
    names, defaults = REFERENCE_SURFACE[name]
    sig = inspect.signature(getattr(pkg, name))
    got = list(sig.parameters)
    assert got[:len(names)] == names, (name, got)
    for k, v in defaults.items():
        assert sig.parameters[k].default == v
    for extra in got[len(names):]:                                   # extensions must be optional
        assert sig.parameters[extra].default is not inspect.Parameter.empty


def test_ssd_and_anchor_generator_surface():
    pkg = load_pkg()
    assert list(inspect.signature(pkg.SSD.__init__).parameters)[1:] == \
        ['images', 'feature_extractor', 'anchor_generator', 'box_predictor', 'num_classes']          # ssd.py:10
    gp = inspect.signature(pkg.SSD.get_predictions).parameters
    assert (gp['score_threshold'].default, gp['iou_threshold'].default, gp['max_boxes_per_class'].default) == (0.05, 0.5, 20)
    assert list(inspect.signature(pkg.SSD.loss).parameters)[1:] == ['groundtruth', 'params']          # ssd.py:71
    g = pkg.AnchorGenerator()                                                                        # anchor_generator.py:13-16
    assert (g.strides, g.scales, g.scale_multipliers, g.aspect_ratios) == \
        ([8, 16, 32, 64, 128], [32, 64, 128, 256, 512], [1.0, 1.4142], [1.0, 2.0, 0.5])
    assert g.num_anchors_per_location == 6                                                           # :38
    assert g.count(640, 640) == (51150, [38400, 9600, 2400, 600, 150])
    with pytest.raises(AssertionError):                                                              # :33
        pkg.AnchorGenerator(strides=[8, 16], scales=[32])
    with pytest.raises(AssertionError):                                                              # training_target_creation.py:86
        pkg.match_boxes(np.zeros([4, 4], np.float32), np.zeros([1, 4], np.float32), 0.4, 0.5)
    c = load_pkg('detector.constants')
    assert (c.POSITIVES_THRESHOLD, c.NEGATIVES_THRESHOLD) == (0.5, 0.5)                              # constants.py:25-26


def test_config_loader_accepts_the_reference_configs(tmp_path):
    """config_mobilenet.json / config_shufflenet.json: the hot-path keys with the shipped values."""
    pkg = load_pkg()
    shipped = {'model_dir': 'models/run00', 'pretrained_checkpoint': 'pretrained/mobilenet_v1_1.0_224.ckpt', 'backbone': 'mobilenet',
               'depth_multiplier': 1.0, 'weight_decay': 1e-4, 'num_classes': 80, 'score_threshold': 0.15, 'iou_threshold': 0.6,
               'max_boxes_per_class': 25, 'localization_loss_weight': 1.0, 'classification_loss_weight': 1.0,
               'gamma': 2.0, 'alpha': 0.25, 'num_steps': 300000, 'initial_learning_rate': 1e-3,
               'min_dimension': 640, 'batch_size': 14, 'image_height': 640, 'image_width': 640}
    p = tmp_path / 'config_mobilenet.json'
    p.write_text(json.dumps(shipped))
    params = pkg.config.load_config(str(p))
    assert pkg.config.postprocess_kwargs(params) == {'score_threshold': 0.15, 'iou_threshold': 0.6, 'max_boxes_per_class': 25}
    assert pkg.config.total_loss({'localization_loss': 2.0, 'classification_loss': 3.0}, params) == 5.0
    (tmp_path / 'bad.json').write_text(json.dumps({'num_classes': 80}))
    with pytest.raises(ValueError):
        pkg.config.load_config(str(tmp_path / 'bad.json'))


def test_shard_range_partitions_the_batch():
    par = load_pkg('parallel')
    for B in (0, 1, 7, 16, 256, 257):
        for world in (1, 2, 3, 4, 8):
            spans = [par.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))        # contiguous, no overlap
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert [par.shard_range(256, r, 8) for r in (0, 7)] == [(0, 32), (224, 256)]          # cfg4: 32 images per GPU


WORKER = r'''
import importlib, json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
par = importlib.import_module('single-shot-detector_b200.parallel')
syn = importlib.import_module('single-shot-detector_b200.synthetic')
from oracle import ssd as ossd                                  # the checker stands in for the GPU kernels on CPU
from oracle.anchor_generator import AnchorGenerator as OracleGen
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
assert par.connect_peers() is False                # no GPUs here: the NVLink peer exchange is unavailable, NCCL/gloo stays in use
H, W, C, B, G = 128, 160, 5, 5, 4
anchors = OracleGen()(H, W); A = anchors.shape[0]
gt = syn.make_groundtruth(77, B, G, H, W, C, vary_count=True)
logits = syn.make_logits('realistic', 77, B, A, C, anchors, gt); codes = syn.make_codes(77, B, A)
lo, hi = par.shard_range(B, rank, world)
sgt = par.shard_groundtruth(gt, rank, world)
params = dict(gamma=2.0, alpha=0.25)
if hi > lo:
    r = ossd.loss(anchors, codes[lo:hi], logits[lo:hi], sgt, params, C, return_all=True)
    sums = torch.tensor([r['loc_sum64'], r['cls_sum64'], float(r['num_matches'])], dtype=torch.float64)
else:
    sums = torch.zeros(3, dtype=torch.float64)
local = sums.clone()
par.all_reduce_sums(sums)
loc, cls = par.finalize_losses(sums)
full = ossd.loss(anchors, codes, logits, gt, params, C, return_all=True)
out = dict(rank=rank, lo=lo, hi=hi, local=local.tolist(), sums=sums.tolist(), loc=float(loc), cls=float(cls),
           full_loc=float(full['localization_loss']), full_cls=float(full['classification_loss']), full_n=float(full['num_matches']))
print('RESULT ' + json.dumps(out), flush=True)
dist.destroy_process_group()
'''


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharded_loss_equals_full_batch(tmp_path):
    """cfg4's structure at toy size: images sharded over 2 ranks, ONE all-reduce of (sum loc, sum cls, num_matches),
    normaliser = max(global count, 1) (ssd.py:121-133).  Both ranks must return the full-batch losses."""
    script = tmp_path / 'worker.py'
    script.write_text(WORKER.format(root=ROOT))
    import socket
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                   OMP_NUM_THREADS='1', GLOO_SOCKET_IFNAME='lo')
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    res = []
    for p in procs:
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0, err[-2000:]
        res.append(json.loads([l for l in out.splitlines() if l.startswith('RESULT ')][0][7:]))
    r0, r1 = sorted(res, key=lambda r: r['rank'])
    assert (r0['lo'], r0['hi'], r1['lo'], r1['hi']) == (0, 3, 3, 5)
    assert r0['sums'] == r1['sums']                                           # every rank holds the global sums
    np.testing.assert_allclose(r0['sums'], np.add(r0['local'], r1['local']), rtol=1e-12)
    assert r0['sums'][2] == r0['full_n'] and r0['full_n'] > 0                 # the count is exact
    for r in (r0, r1):
        assert abs(r['loc'] - r['full_loc']) <= 1e-6 * abs(r['full_loc'])
        assert abs(r['cls'] - r['full_cls']) <= 1e-6 * abs(r['full_cls'])
    # a shard that normalised with its LOCAL count would be wrong -- that is what the all-reduce is for
    assert abs(r0['local'][0] / max(r0['local'][2], 1.0) - r0['full_loc']) > 1e-4 * abs(r0['full_loc'])


def test_finalize_losses_uses_max_count_one():
    import torch
    par = load_pkg('parallel')
    loc, cls = par.finalize_losses(torch.tensor([3.0, 6.0, 0.0], dtype=torch.float64))      # no matches: normaliser 1
    assert (float(loc), float(cls)) == (3.0, 6.0) and loc.dtype == torch.float32
    loc, cls = par.finalize_losses(torch.tensor([3.0, 6.0, 4.0], dtype=torch.float64))
    assert (float(loc), float(cls)) == (0.75, 1.5)
