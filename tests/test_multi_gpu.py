"""N>1 on real GPUs (skipped unless the box has >= 2): images sharded over 2 ranks, one all-reduce of
(sum loc, sum cls, num_matches) -- through NCCL and through the hand-written NVLink peer-memory kernel (csrc/comm.cu);
every rank must return the full-batch losses of the oracle, both ways, and the sharded `matches` / detections must
equal the single-GPU ones bit for bit."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

WORKER = r'''
import importlib, json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
pkg = importlib.import_module('single-shot-detector_b200')
syn = importlib.import_module('single-shot-detector_b200.synthetic')
H, W, C, B, G = 256, 320, 12, 6, 6
gen = pkg.AnchorGenerator()
anchors = gen(H, W)
A = anchors.shape[0]
a_np = anchors.cpu().numpy()
gt = syn.make_groundtruth(91, B, G, H, W, C, vary_count=True)
logits = syn.make_logits('realistic', 91, B, A, C, a_np, gt); codes = (syn.make_codes(91, B, A) * np.float32(0.5)).astype(np.float32)
lo, hi = pkg.parallel.shard_range(B, rank, world)
cuda = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
raw = dict(encoded_boxes=cuda(codes[lo:hi]), class_predictions=cuda(logits[lo:hi]))
ssd = pkg.SSD.from_predictions(H, W, raw, gen, C)
ssd.process_group = True
sgt = {{k: cuda(v) for k, v in pkg.parallel.shard_groundtruth(gt, rank, world).items()}}
res = ssd.loss(sgt, dict(gamma=2.0, alpha=0.25))
_, _, matches = ssd._create_targets(sgt)
pred = ssd.get_predictions(0.05, 0.5, 10)
# ---- the same exchange through the peer-memory kernel: plain all-reduce (many epochs, both parity slots), fused
#      finalisation, fused forward + backward, and CUDA-graph replay
peer = pkg.parallel.connect_peers()
peer_out = dict(connected=bool(peer))
if peer:
    vals = []
    for it in range(5):
        t = torch.tensor([rank + 1.0 + it, 10.0 * (rank + 1), 0.5, -2.0 * rank], dtype=torch.float64, device='cuda')
        pkg.parallel.peer_all_reduce_sum(t)
        vals.append(t.cpu().numpy().tolist())
    ssd.peer_all_reduce = True
    res_p = ssd.loss(sgt, dict(gamma=2.0, alpha=0.25))
    l_fb, g_fb = ssd.loss_with_gradients(sgt, dict(gamma=2.0, alpha=0.25))
    ssd.peer_all_reduce = False
    l_nc, g_nc = ssd.loss_with_gradients(sgt, dict(gamma=2.0, alpha=0.25))
    ssd.peer_all_reduce = True
    step = pkg.graph.capture(lambda: ssd.loss(sgt, dict(gamma=2.0, alpha=0.25)))
    replays = []
    for _ in range(4):
        r = step.replay()
        replays.append([float(r['localization_loss']), float(r['classification_loss'])])
    step.release()                                   # teardown order: graphs first, then the process group (graph.CapturedStep.release)
    # head layout (per-level tower outputs) on the sharded batch, with the peer all-reduce
    from oracle import box_predictor as obp
    shapes = obp.level_shapes(H, W, gen.strides)
    n_loc = gen.num_anchors_per_location
    hssd = pkg.SSD.from_head_outputs(H, W, [cuda(t) for t in obp.split_to_levels(codes[lo:hi], shapes, n_loc)],
                                     [cuda(t) for t in obp.split_to_levels(logits[lo:hi], shapes, n_loc)], gen, C)
    hssd.process_group, hssd.peer_all_reduce = True, True
    res_h = hssd.loss(sgt, dict(gamma=2.0, alpha=0.25))
    l_hfb, _ = hssd.loss_with_gradients(sgt, dict(gamma=2.0, alpha=0.25))
    peer_out.update(head=[float(res_h['localization_loss']), float(res_h['classification_loss'])],
                    head_fb=[float(l_hfb['localization_loss']), float(l_hfb['classification_loss'])])
    peer_out.update(vals=vals, loc=float(res_p['localization_loss']), cls=float(res_p['classification_loss']),
                    fb=[float(l_fb['localization_loss']), float(l_fb['classification_loss'])],
                    fb_nccl=[float(l_nc['localization_loss']), float(l_nc['classification_loss'])],
                    grads_equal=bool(torch.equal(g_fb['class_predictions'], g_nc['class_predictions'])
                                     and torch.equal(g_fb['encoded_boxes'], g_nc['encoded_boxes'])),
                    replays=replays, error=pkg.parallel.peer_error())
out = dict(rank=rank, peer=peer_out, lo=lo, hi=hi, loc=float(res['localization_loss']), cls=float(res['classification_loss']),
           num_matches=float(ssd.num_matches), matches=matches.cpu().numpy().tolist(),
           num_boxes=pred['num_boxes'].cpu().numpy().tolist(), labels=pred['labels'].cpu().numpy().tolist())
json.dump(out, open(os.path.join({out_dir!r}, 'result_%d.json' % rank), 'w'))
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.timeout(600)
def test_two_gpu_sharded_equals_single(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    import importlib
    from oracle import losses as olosses, nms as onms, ssd as ossd
    from oracle.anchor_generator import AnchorGenerator as OracleGen
    syn = importlib.import_module('single-shot-detector_b200.synthetic')
    script = tmp_path / 'worker.py'
    script.write_text(WORKER.format(root=ROOT, out_dir=str(tmp_path)))
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        # results go through files, logs to files: a rank blocked on a full stdout pipe would deadlock the barrier
        log = open(str(tmp_path / ('log_%d.txt' % rank)), 'w')
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=log, stderr=subprocess.STDOUT))
    for rank, p in enumerate(procs):
        try:
            p.wait(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise AssertionError('rank %d hung: %s' % (rank, open(str(tmp_path / ('log_%d.txt' % rank))).read()[-2000:]))
        assert p.returncode == 0, open(str(tmp_path / ('log_%d.txt' % rank))).read()[-3000:]
    res = [json.load(open(str(tmp_path / ('result_%d.json' % rank)))) for rank in range(2)]
    r0, r1 = sorted(res, key=lambda r: r['rank'])
    H, W, C, B, G = 256, 320, 12, 6, 6
    anchors = OracleGen()(H, W)
    A = anchors.shape[0]
    gt = syn.make_groundtruth(91, B, G, H, W, C, vary_count=True)
    logits = syn.make_logits('realistic', 91, B, A, C, anchors, gt)
    codes = (syn.make_codes(91, B, A) * np.float32(0.5)).astype(np.float32)
    full = ossd.loss(anchors, codes, logits, gt, dict(gamma=2.0, alpha=0.25), C, return_all=True)
    assert (r0['lo'], r0['hi'], r1['lo'], r1['hi']) == (0, 3, 3, 6)
    for r in (r0, r1):                                                    # every rank holds the GLOBAL losses
        assert r['num_matches'] == float(full['num_matches'])
        assert abs(r['loc'] - float(full['localization_loss'])) <= 1e-5 * abs(float(full['localization_loss']))
        assert abs(r['cls'] - float(full['classification_loss'])) <= 1e-5 * abs(float(full['classification_loss']))
    assert r0['loc'] == r1['loc'] and r0['cls'] == r1['cls']
    # the NVLink peer-memory all-reduce: same values on both ranks, equal to the NCCL results
    assert r0['peer']['connected'] and r1['peer']['connected'], 'peer mailboxes could not be mapped on a 2-GPU box'
    for it in range(5):
        want_vals = [(1.0 + it) + (2.0 + it), 30.0, 1.0, -2.0]
        assert r0['peer']['vals'][it] == want_vals and r1['peer']['vals'][it] == want_vals
    for r in (r0, r1):
        p = r['peer']
        assert p['error'] == 0
        assert (p['loc'], p['cls']) == (r['loc'], r['cls'])               # sums are added in rank order on every rank
        assert p['fb'] == p['fb_nccl'] and p['grads_equal']
        assert all(rep == [p['loc'], p['cls']] for rep in p['replays'])
        for got in (p['head'], p['head_fb']):                             # head-layout path on the shards: the global losses
            assert abs(got[0] - float(full['localization_loss'])) <= 1e-5 * abs(float(full['localization_loss']))
            assert abs(got[1] - float(full['classification_loss'])) <= 1e-5 * abs(float(full['classification_loss']))
    assert r0['peer']['head'] == r1['peer']['head'] and r0['peer']['head_fb'] == r1['peer']['head_fb']
    got = np.concatenate([np.array(r0['matches'], np.int32), np.array(r1['matches'], np.int32)])
    assert np.array_equal(got, full['matches'])                           # bit-exact across the shard boundary
    want = onms.batch_multiclass_non_max_suppression(codes, anchors, olosses.sigmoid(logits), 0.05, 0.5, 10)
    assert r0['num_boxes'] + r1['num_boxes'] == want[3].tolist()
    assert np.array_equal(np.array(r0['labels'] + r1['labels'], np.int32), want[2])


def test_library_calls_leave_the_current_device_alone():
    """ADVICE (round 1): a call on tensors of cuda:1 must not switch the thread's current device (one process driving several
    GPUs); every entry point restores the caller's device (SsdkDeviceGuard)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    import importlib
    pkg = importlib.import_module('single-shot-detector_b200')
    torch.cuda.set_device(0)
    gen = pkg.AnchorGenerator()
    a1 = gen(128, 160, device=torch.device('cuda', 1))
    assert a1.device.index == 1 and torch.cuda.current_device() == 0
    b = torch.rand([5, 4], device='cuda:1')
    got = pkg.iou(b, a1[:7])
    assert got.device.index == 1 and torch.cuda.current_device() == 0
    x = torch.empty(3, device='cuda')
    assert x.device.index == 0
    a0 = gen(128, 160, device=torch.device('cuda', 0))
    assert torch.equal(a0.cpu(), a1.cpu())
