"""bench.py -- images/sec of the per-anchor detection hot path at 896x640 (BASELINE.json metric).

One "step" per GPU = the training-side path (target assignment + focal / smooth-L1 loss, SSD.loss) over one
cfg2 batch (16 images, 107,415 anchors, 90 classes, 20 GT boxes/image) PLUS the inference-side path
(sigmoid + decode + per-class NMS, SSD.get_predictions, 0.05 / 0.5 / 100) over one cfg3 batch (32 images).
value = images processed by all ranks per second ((16 + 32) * n_gpus / step time); the two sub-paths are also
reported separately under "breakdown".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

--impl reference times the reference's CPU semantics (the NumPy/C oracle port; TensorFlow cannot be installed
here) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = 'single-shot-detector_b200'

TRAIN_CFG, INFER_CFG = 2, 3
SCORE_THR, IOU_THR, K_PER_CLASS = 0.05, 0.5, 100
PARAMS = {'gamma': 2.0, 'alpha': 0.25}
UPSTREAM = (1.0, 1.0)          # localization_loss_weight, classification_loss_weight of the shipped configs


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--e2e-steps', type=int, default=0, help='steps of the host-buffer (e2e) loop; 0 = min(steps, 10)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of CUDA-graph replays')
    ap.add_argument('--no-overlap', action='store_true', help='headline: replay the two sub-paths back to back on one stream')
    ap.add_argument('--nccl-allreduce', action='store_true', help='N>1: all-reduce the loss sums with NCCL instead of the peer-memory kernel')
    ap.add_argument('--cpu-sample', type=int, default=2, help='cpu_baseline sample: this many train images + 2x infer images')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workload
def algorithmic_bytes(A, C, G, K):
    """bytes per image, each tensor counted once."""
    train = 4 * A * C + 56 * A + 20 * G
    infer = 4 * A * C + 32 * A + 24 * C * K + 4
    loss_kernel = 4 * A * C + 16 * A + 16 * A + 8 * A          # logits + codes + reg_targets + (cls, matches)
    filter_kernel = 4 * A * C                                   # scores read once (+ the few candidates written)
    return train, infer, loss_kernel, filter_kernel


def make_host_inputs(syn, rank, pin):
    import torch
    out = {}
    for name, cfg_id, kind in (('train', TRAIN_CFG, 'train'), ('infer', INFER_CFG, 'realistic')):
        cfg = syn.CONFIGS[cfg_id]
        B, H, W, C, G = cfg['B'], cfg['H'], cfg['W'], cfg['C'], cfg['G']
        first = rank * B
        out[name] = dict(cfg=cfg, B=B, H=H, W=W, C=C, G=G, first=first, kind=kind)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed regions run."""
    Q = ('timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix='clocks_', suffix='.csv')
        self.proc = None
        self.windows = []
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def finish(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        import datetime
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                rows.append((ts, float(f[2]), float(f[3]), f[5:]))
            except Exception:
                continue
        try:
            os.unlink(self.path)
        except OSError:
            pass
        inside = [r for r in rows if any(t0 - 0.05 <= r[0] <= t1 + 0.05 for t0, t1 in self.windows)]
        use = inside if inside else rows
        if not use:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        names = ['active', 'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = set()
        for r in use:
            for n, v in zip(names[1:], r[3][1:]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median([r[1] for r in use])), 'sm_max_mhz': use[0][2],
                'reasons': sorted(reasons), 'samples': len(use), 'samples_in_timed_region': len(inside)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(n_train, n_infer, steps, warmup, threads):
    """Reference semantics on the host cores: the oracle port (NumPy float32 op for op + C NonMaxSuppressionV3),
    one image per task on a thread pool, as the reference's tf.map_fn does with parallel_iterations."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ssd as ossd
    from oracle.anchor_generator import AnchorGenerator as OracleGen
    from oracle import nms as onms
    syn = importlib.import_module(PKG + '.synthetic')
    onms._lib()
    tc, ic = syn.CONFIGS[TRAIN_CFG], syn.CONFIGS[INFER_CFG]
    anchors = OracleGen(scale_multipliers=tc['scale_multipliers'])(tc['H'], tc['W'])
    A = anchors.shape[0]
    gt = syn.make_groundtruth(TRAIN_CFG, n_train, tc['G'], tc['H'], tc['W'], tc['C'])
    t_logits = syn.make_logits('train', TRAIN_CFG, n_train, A, tc['C'])
    t_codes = syn.make_codes(TRAIN_CFG, n_train, A)
    igt = syn.make_groundtruth(INFER_CFG, n_infer, ic['G'], ic['H'], ic['W'], ic['C'])
    i_logits = syn.make_logits('realistic', INFER_CFG, n_infer, A, ic['C'], anchors, igt)
    i_codes = syn.make_codes(INFER_CFG, n_infer, A)

    def train_image(b):
        g = {k: v[b:b + 1] for k, v in gt.items()}
        r = ossd.loss(anchors, t_codes[b:b + 1], t_logits[b:b + 1], g, PARAMS, tc['C'], return_all=True)
        return r['loc_sum64'], r['cls_sum64'], float(r['num_matches'])

    def infer_image(b):
        p = ossd.get_predictions(anchors, i_codes[b:b + 1], i_logits[b:b + 1], SCORE_THR, IOU_THR, K_PER_CLASS)
        return int(p['num_boxes'][0])

    def step(pool):
        futs = [pool.submit(train_image, b) for b in range(n_train)] + [pool.submit(infer_image, b) for b in range(n_infer)]
        res = [f.result() for f in futs]
        tr = np.array(res[:n_train], np.float64).sum(axis=0)
        norm = max(tr[2], 1.0)
        return tr[0] / norm, tr[1] / norm, sum(res[n_train:])

    with ThreadPoolExecutor(max_workers=threads) as pool:
        for _ in range(warmup):
            step(pool)
        t0 = time.perf_counter()
        for _ in range(steps):
            out = step(pool)
        dt = time.perf_counter() - t0
    return (n_train + n_infer) * steps / dt, dt / steps, out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = host_threads()
    n_train, n_infer = 4, 8          # bounded sample of the step (same 1:2 train:infer mix as the GPU arm)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # keep the whole run within a few minutes whatever K the driver asks for
    ips, sec, _ = cpu_reference(n_train, n_infer, 1, 0, threads)
    budget = 150.0
    if sec * (steps + warmup) > budget:
        scale = max(1, int(budget / max(sec, 1e-3)))
        warmup = min(warmup, max(1, scale // 10))
        steps_run = max(1, scale - warmup)
    else:
        steps_run = steps
    ips, sec, _ = cpu_reference(n_train, n_infer, steps_run, warmup, threads)
    syn = importlib.import_module(PKG + '.synthetic')
    line = {
        'impl': 'reference', 'metric': 'images_per_sec_target_assign_focal_loss_and_decode_nms_896x640',
        'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'steps_timed': steps_run, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(syn, sample='%d train + %d infer images per step (bounded sample of the 16 + 32 step)' % (n_train, n_infer)),
        'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d steps x (%d train + %d infer images), oracle port (NumPy f32 + C NMS), one image per thread-pool task'
                                   % (steps_run, n_train, n_infer)},
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(syn, sample=None, all_reduce=None):
    tc, ic = syn.CONFIGS[TRAIN_CFG], syn.CONFIGS[INFER_CFG]
    cfg = {
        'workload': 'per step and GPU: SSD.loss (targets + focal/smooth-L1) on cfg2 batch %d  +  SSD.get_predictions '
                    '(sigmoid, decode, per-class NMS %.2f/%.1f/%d) on cfg3 batch %d; 640x896, 5 FPN levels x 9 anchors '
                    '= 107415 anchors, 90 classes, 20 GT boxes/image' % (tc['B'], SCORE_THR, IOU_THR, K_PER_CLASS, ic['B']),
        'train_batch_per_gpu': tc['B'], 'infer_batch_per_gpu': ic['B'],
        'logits': 'train: N(-4.595,1) prior-bias init; infer: N(-7,1) background + N(1.5,1.5) on anchors with IoU>=0.4 to a GT',
        'l2': 'inputs per step (0.62 GB + 1.24 GB of logits) exceed the 126 MB L2; no flush needed',
        'parallelism': 'image-sharded, one all-reduce of 3 doubles per step',
        'launch': 'one CUDA graph per step; the training-side and the inference-side sub-path are independent and are issued on '
                  'two streams inside it (breakdown.sequential_graph_* = the same graph on one stream)',
    }
    if sample:
        cfg['sample'] = sample
    if all_reduce:
        cfg['all_reduce'] = all_reduce
    return cfg


# ------------------------------------------------------------------------------------------------ GPU arm
def finish_rank(world):
    """Multi-rank teardown.  Destroying an NCCL communicator that live CUDA graphs still reference can block for
    minutes; all results are out by now, so flush and leave the process without running that teardown."""
    if world > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    pkg = importlib.import_module(PKG)
    syn = importlib.import_module(PKG + '.synthetic')
    lib = pkg._lib

    tc, ic = syn.CONFIGS[TRAIN_CFG], syn.CONFIGS[INFER_CFG]
    H, W, C = tc['H'], tc['W'], tc['C']
    gen = pkg.AnchorGenerator(scale_multipliers=tc['scale_multipliers'])
    anchors = gen(H, W, device=dev)
    A = anchors.shape[0]
    anchors_np = anchors.cpu().numpy()
    Bt, Bi, G = tc['B'], ic['B'], tc['G']

    # ---- synthetic inputs in pinned host memory (what the e2e loop copies from), then resident copies in HBM
    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    h_tlog, h_tcod = pinned([Bt, A, C], torch.float32), pinned([Bt, A, 4], torch.float32)
    h_ilog, h_icod = pinned([Bi, A, C], torch.float32), pinned([Bi, A, 4], torch.float32)
    gt = syn.make_groundtruth(TRAIN_CFG, Bt, G, H, W, C, first_image=rank * Bt)
    igt = syn.make_groundtruth(INFER_CFG, Bi, G, H, W, C, first_image=rank * Bi)
    syn.make_logits('train', TRAIN_CFG, Bt, A, C, first_image=rank * Bt, out=h_tlog.numpy())
    syn.make_codes(TRAIN_CFG, Bt, A, first_image=rank * Bt, out=h_tcod.numpy())
    syn.make_logits('realistic', INFER_CFG, Bi, A, C, anchors_np, igt, first_image=rank * Bi, out=h_ilog.numpy())
    syn.make_codes(INFER_CFG, Bi, A, first_image=rank * Bi, out=h_icod.numpy())
    d_tlog, d_tcod, d_ilog, d_icod = (t.to(dev) for t in (h_tlog, h_tcod, h_ilog, h_icod))
    d_gt = {k: torch.from_numpy(v).to(dev) for k, v in gt.items()}

    raw_t = {'encoded_boxes': d_tcod, 'class_predictions': d_tlog}
    raw_i = {'encoded_boxes': d_icod, 'class_predictions': d_ilog}
    ssd_t = pkg.SSD.from_predictions(H, W, raw_t, gen, C)
    ssd_i = pkg.SSD.from_predictions(H, W, raw_i, gen, C)
    peer = False
    if world > 1:
        ssd_t.process_group = True
        if not args.nccl_allreduce:
            peer = bool(pkg.parallel.connect_peers())          # NVLink peer-memory all-reduce (csrc/comm.cu); NCCL if it cannot connect
        ssd_t.peer_all_reduce = peer

    def step_resident():
        losses = ssd_t.loss(d_gt, PARAMS)
        pred = ssd_i.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS)
        return losses, pred

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        out = step_resident()
    barrier()

    def timed_loop(run_step, n):
        """n steps bracketed by barrier + synchronize, CUDA events on the launching stream; returns (ms, wall window, out)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.time()
        ev0.record()
        for _ in range(n):
            o = run_step()
        ev1.record()
        barrier()
        w1 = time.time()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), (w0, w1), o

    # ---- timed region 1a: inputs resident in HBM, one Python call per library entry point (eager launches)
    l0 = lib.launch_count(local_rank)
    ms_eager, win, out = timed_loop(step_resident, args.steps)
    launches = lib.launch_count(local_rank) - l0
    eager_ms_per_step = ms_eager / args.steps

    # ---- timed region 1b (the headline `value`): the same step captured once into a CUDA graph and replayed --
    #      identical kernels and inputs, one launch per step, so the host cannot starve the GPU
    mode = 'eager'
    ms_per_step = eager_ms_per_step
    ms_sequential_graph = None
    if not args.no_graph:
        try:
            captured = pkg.graph.capture(step_resident, warmup=2)
            for _ in range(3):
                captured.replay()
            ms_graph, win, out = timed_loop(captured.replay, args.steps)
            ms_per_step = ms_sequential_graph = ms_graph / args.steps
            launches = captured.launches_per_replay * args.steps
            mode = 'cuda_graph'
        except Exception as e:                                       # keep the eager number, say why
            mode = 'eager (graph capture failed: %s)' % str(e)[:200]
            torch.cuda.synchronize()
    # ---- timed region 1c: the same step with its two independent sub-paths on two streams inside the graph (same
    #      kernels, same inputs, same results): matching / sorting / NMS / packing hide behind the two HBM-bound passes
    if mode == 'cuda_graph' and not args.no_overlap:
        try:
            step_overlapped = pkg.graph.concurrent(lambda: ssd_t.loss(d_gt, PARAMS),
                                                   lambda: ssd_i.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS))
            captured_o = pkg.graph.capture(step_overlapped, warmup=2)
            for _ in range(3):
                captured_o.replay()
            ms_o, win_o, out_o = timed_loop(captured_o.replay, args.steps)
            same = (float(out_o[0]['localization_loss']) == float(out[0]['localization_loss'])
                    and float(out_o[0]['classification_loss']) == float(out[0]['classification_loss'])
                    and all(torch.equal(out_o[1][k], out[1][k]) for k in ('boxes', 'labels', 'scores', 'num_boxes')))
            if same and ms_o / args.steps < ms_per_step:
                ms_per_step, win, out = ms_o / args.steps, win_o, out_o
                launches = captured_o.launches_per_replay * args.steps
                mode = 'cuda_graph, sub-paths on two streams'
        except Exception:
            torch.cuda.synchronize()
    if sampler:
        sampler.window(*win)
    value = (Bt + Bi) * world / (ms_per_step * 1e-3)
    losses, pred = out
    check = {'localization_loss': float(losses['localization_loss']), 'classification_loss': float(losses['classification_loss']),
             'detections_image0': int(pred['num_boxes'][0])}

    # ---- sub-path timings (same resident inputs), each its own event-timed loop
    eager_fallbacks = []
    def timed(fn, n):
        """ms per call of one sub-path, launched the same way as the headline number (graph replay when available)."""
        run = fn
        if mode.startswith('cuda_graph'):
            try:
                run = pkg.graph.capture(fn, warmup=2).replay
            except Exception as e:
                eager_fallbacks.append(str(e)[:120])
                torch.cuda.synchronize()
        for _ in range(3):
            run()
        return timed_loop(run, n)[0] / n
    ms_train = timed(lambda: ssd_t.loss(d_gt, PARAMS), args.steps)
    ms_infer = timed(lambda: ssd_i.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS), args.steps)
    # forward + backward of the training side: targets, losses, all-reduce, gradients w.r.t. both heads
    ms_train_fb = timed(lambda: ssd_t.loss_with_gradients(d_gt, PARAMS, upstream=UPSTREAM), args.steps)
    # the same fused step when the targets were assigned earlier (SSD.assign_targets needs only anchors + ground truth, so a
    # training loop can run it on a side stream during the network's forward pass): the streaming pass alone.  NOT the
    # headline -- the matching work is outside this timed region; it shows what remains on the critical path.
    pre_targets = pkg.SSD.assign_targets(ssd_t.anchors, d_gt)
    ms_train_fb_pre = timed(lambda: ssd_t.loss_with_gradients(None, PARAMS, upstream=UPSTREAM, targets=pre_targets), args.steps)
    ms_train_nccl = None
    if world > 1 and peer:                       # the same training sub-path with the library collective, for comparison
        ssd_t.peer_all_reduce = False
        ms_train_nccl = timed(lambda: ssd_t.loss(d_gt, PARAMS), args.steps)
        ssd_t.peer_all_reduce = True

    # ---- per-kernel durations (library-side CUDA events on the launching stream) for the roofline object
    lib.set_profiling(True, local_rank)
    for _ in range(min(args.steps, 10)):
        step_resident()
    prof_step = lib.profile_read(local_rank)              # exactly the kernels of min(steps, 10) steps
    ssd_t._loss_forward(d_gt, PARAMS, keep_targets=True)
    sv = ssd_t._saved
    sums_tmp = torch.empty([3], dtype=torch.float64, device=dev)
    for _ in range(min(args.steps, 10)):
        ssd_t.loss_backward(UPSTREAM)
        # the row-tiled forward kernel (ssdk_ssd_loss: targets given), not part of the step any more
        lib.check(lib.load().ssdk_ssd_loss(lib.context(local_rank), sv['logits'].data_ptr(), sv['codes'].data_ptr(), sv['reg_targets'].data_ptr(),
                                           sv['cls_targets'].data_ptr(), sv['matches'].data_ptr(), Bt, A, C, PARAMS['gamma'], PARAMS['alpha'],
                                           sums_tmp.data_ptr(), None, None))
    prof = lib.profile_read(local_rank)
    for k_, v_ in prof_step.items():                      # roofline entries use every launch seen, the per-step table only the steps
        prof[k_] = (prof[k_][0] + v_[0], prof[k_][1] + v_[1])
    lib.set_profiling(False, local_rank)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'MEASURED_PEAKS.json hbm_gbs (of measured)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (of fallback)'
    b_train, b_infer, b_loss, b_filter = algorithmic_bytes(A, C, G, K_PER_CLASS)

    # DRAM traffic per launch from the committed `ncu --set full` capture of this same command (profiles/ncu_traffic.json,
    # written by scripts/ncu_traffic.py); null for kernels that capture does not contain
    try:
        ncu_traffic = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
    except Exception:
        ncu_traffic = {}

    def kernel_roof(name, bytes_per_launch, traffic_key=None):
        tot, n = prof[name]
        if n == 0:
            return None
        avg_ms = tot / n
        ach = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        return {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                'traffic': (ncu_traffic.get(traffic_key or (name + '_kernel')) or {}).get('dram_bytes_per_launch'), 'avg_launch_ms': avg_ms, 'algorithmic_bytes_per_launch': bytes_per_launch, 'peak_source': peak_src}
    # SSD.loss = matcher (side stream) || flat pass over the logits, then the matched-anchor corrections (csrc/head.cu); the
    # row-tiled ssd_loss_kernel serves the per-anchor-output API and ssdk_ssd_loss, and is profiled apart below
    roof_flat = kernel_roof('head_flat', 4 * A * C * Bt, 'head_flat_forward_kernel')
    roof_loss = kernel_roof('ssd_loss', b_loss * Bt)
    roof_filter = kernel_roof('filter', b_filter * Bi)
    b_backward = 8 * A * C + 56 * A                 # logits read + grad written; codes, reg_targets, grad_codes, cls, matches
    roof_backward = kernel_roof('ssd_loss_backward', b_backward * Bt)
    step_kernel_ms = {k: (v[0] / max(1, min(args.steps, 10))) for k, v in prof_step.items() if v[1]}
    step_kernel_ms['note'] = ('eager launches with event pairs; `match` runs on the side stream concurrently with `head_flat` '
                              '(its wall time while co-running, 0.04 ms alone)')
    dominant = max((r for r in (roof_loss, roof_flat, roof_filter) if r), key=lambda r: r['avg_launch_ms'])
    dominant = dict(dominant)
    dominant['share_of_step_kernel_time'] = dominant['avg_launch_ms'] / max(1e-9, sum(v for k, v in step_kernel_ms.items() if k not in ('note', 'match')))

    # ---- head-layout path: the same two sub-paths fed with the per-level tower outputs
    #      [B, n*C, h, w] / [B, n*4, h, w] (channels_first, as the reference's box predictor emits them) instead of the
    #      concatenated [B,A,C] / [B,A,4]; plus the cost of reshape_and_concatenate itself, which this path removes
    head = None
    try:
        n_loc = gen.num_anchors_per_location
        shapes = [(-(-H // s_), -(-W // s_)) for s_ in gen.strides]

        def to_levels(t, D):
            out_l, off = [], 0
            for h_, w_ in shapes:
                cnt = h_ * w_ * n_loc
                out_l.append(t[:, off:off + cnt].reshape(t.shape[0], h_, w_, n_loc * D).permute(0, 3, 1, 2).contiguous())
                off += cnt
            return out_l
        lv_tlog, lv_tcod, lv_ilog, lv_icod = to_levels(d_tlog, C), to_levels(d_tcod, 4), to_levels(d_ilog, C), to_levels(d_icod, 4)
        ssd_th = pkg.SSD.from_head_outputs(H, W, lv_tcod, lv_tlog, gen, C)
        ssd_ih = pkg.SSD.from_head_outputs(H, W, lv_icod, lv_ilog, gen, C)
        if world > 1:
            ssd_th.process_group = True
            ssd_th.peer_all_reduce = peer
        up_dev = torch.tensor(UPSTREAM, dtype=torch.float32, device=dev)
        hl = ssd_th.loss(d_gt, PARAMS)
        hp = ssd_ih.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS)
        head_check = {'localization_loss': float(hl['localization_loss']), 'classification_loss': float(hl['classification_loss']),
                      'detections_image0': int(hp['num_boxes'][0]),
                      'detections_identical_to_anchor_major': bool(all(torch.equal(hp[k], pred[k]) for k in ('boxes', 'labels', 'scores', 'num_boxes')))}
        ms_h_train = timed(lambda: ssd_th.loss(d_gt, PARAMS), args.steps)
        ms_h_fb = timed(lambda: ssd_th.loss_with_gradients(d_gt, PARAMS, upstream=up_dev), args.steps)
        ms_h_infer = timed(lambda: ssd_ih.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS), args.steps)
        ms_concat_t = timed(lambda: pkg.reshape_and_concatenate(lv_tcod, lv_tlog, C, n_loc, lazy=False), args.steps)
        ms_concat_i = timed(lambda: pkg.reshape_and_concatenate(lv_icod, lv_ilog, C, n_loc, lazy=False), args.steps)
        nprof = min(args.steps, 10)
        lib.set_profiling(True, local_rank)
        lib.profile_read(local_rank)
        for _ in range(nprof):
            ssd_th.loss(d_gt, PARAMS)
        prof_f = lib.profile_read(local_rank)
        for _ in range(nprof):
            ssd_th.loss_with_gradients(d_gt, PARAMS, upstream=up_dev)
        prof_fb = lib.profile_read(local_rank)
        for _ in range(nprof):
            ssd_ih.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS)
        prof_i = lib.profile_read(local_rank)
        lib.set_profiling(False, local_rank)

        def roof_of(p, name, nbytes, traffic_key=None):
            tot, cnt = p[name]
            if cnt == 0:
                return None
            avg = tot / cnt
            ach = nbytes / (avg * 1e-3) / 1e9
            return {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                    'avg_launch_ms': avg, 'algorithmic_bytes_per_launch': nbytes, 'peak_source': peak_src,
                    'traffic': (ncu_traffic.get(traffic_key or (name + '_kernel')) or {}).get('dram_bytes_per_launch')}
        bh_train = 4 * A * C + 4 * A + 40 * A + 20 * G          # logits + matches read by the loss; targets written/read by the matcher side
        head = {
            'what': 'same sub-paths on per-level channels_first tower outputs (no reshape_and_concatenate); results checked against the anchor-major path',
            'check': head_check,
            'train_ms_per_step': ms_h_train, 'train_images_per_sec': Bt * world / (ms_h_train * 1e-3),
            'train_fwd_bwd_ms_per_step': ms_h_fb, 'train_fwd_bwd_images_per_sec': Bt * world / (ms_h_fb * 1e-3),
            'infer_ms_per_step': ms_h_infer, 'infer_images_per_sec': Bi * world / (ms_h_infer * 1e-3),
            'reshape_and_concatenate_ms': {'train_batch': ms_concat_t, 'infer_batch': ms_concat_i,
                                           'note': 'cost of the copy the reference makes before the anchor-major path (8AC+32A bytes per image)'},
            'unfused_train_ms_per_step': ms_concat_t + ms_train, 'unfused_train_fwd_bwd_ms_per_step': ms_concat_t + ms_train_fb,
            'unfused_infer_ms_per_step': ms_concat_i + ms_infer,
            'roofline_head_flat_forward': roof_of(prof_f, 'head_flat', 4 * A * C * Bt, 'head_flat_forward_kernel'),
            'roofline_head_flat_forward_backward': roof_of(prof_fb, 'head_flat', 8 * A * C * Bt, 'head_flat_forward_backward_kernel'),
            'roofline_head_filter': roof_of(prof_i, 'filter', 4 * A * C * Bi),
            'kernel_ms': {'forward': {k: v[0] / nprof for k, v in prof_f.items() if v[1]},
                          'forward_backward': {k: v[0] / nprof for k, v in prof_fb.items() if v[1]},
                          'infer': {k: v[0] / nprof for k, v in prof_i.items() if v[1]}},
            'train_frac_of_hbm_roofline': (bh_train * Bt / (ms_h_train * 1e-3) / 1e9) / peak,
            'infer_frac_of_hbm_roofline': (b_infer * Bi / (ms_h_infer * 1e-3) / 1e9) / peak,
        }
        del lv_tlog, lv_tcod, lv_ilog, lv_icod, ssd_th, ssd_ih
    except Exception as e:                                            # the headline numbers do not depend on this section
        head = {'error': str(e)[:300]}
        torch.cuda.synchronize()

    # ---- BASELINE.json configs[0] (the reference's own CPU-runnable case: config_mobilenet anchor set, one 640x640 image,
    #      20 GT boxes, matching + focal loss) and single-image inference latency (inference/detector.py serves one image)
    small = None
    try:
        c1 = syn.CONFIGS[1]
        gen1 = pkg.AnchorGenerator(scale_multipliers=c1['scale_multipliers'])
        anc1 = gen1(c1['H'], c1['W'], device=dev)
        A1, C1 = anc1.shape[0], c1['C']
        gt1 = syn.make_groundtruth(1, 1, c1['G'], c1['H'], c1['W'], C1)
        log1 = syn.make_logits('train', 1, 1, A1, C1)
        cod1 = syn.make_codes(1, 1, A1)
        ssd1 = pkg.SSD.from_predictions(c1['H'], c1['W'], {'encoded_boxes': torch.from_numpy(cod1).to(dev),
                                                           'class_predictions': torch.from_numpy(log1).to(dev)}, gen1, C1)
        d_gt1 = {k: torch.from_numpy(v).to(dev) for k, v in gt1.items()}
        ms_cfg1 = timed(lambda: ssd1.loss(d_gt1, PARAMS), args.steps)
        l1 = ssd1.loss(d_gt1, PARAMS)
        ssd_b1 = pkg.SSD.from_predictions(H, W, {'encoded_boxes': d_icod[:1].contiguous(), 'class_predictions': d_ilog[:1].contiguous()}, gen, C)
        ms_b1 = timed(lambda: ssd_b1.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS), args.steps)
        small = {'cfg1_train_one_image_640x640_ms': ms_cfg1, 'cfg1_anchors': int(A1), 'cfg1_classes': int(C1),
                 'cfg1_check': {'localization_loss': float(l1['localization_loss']), 'classification_loss': float(l1['classification_loss'])},
                 'infer_latency_one_image_640x896_ms': ms_b1, '_inputs': (anc1.cpu().numpy(), cod1, log1, gt1, C1)}
    except Exception as e:
        small = {'error': str(e)[:300]}
        torch.cuda.synchronize()

    # ---- timed region 2 (e2e): the same step through the public API with HOST buffers; every step copies its inputs
    #      from pinned host memory to the device and reads the results back (ssdk_*_host entry points)
    h_raw_t = {'encoded_boxes': h_tcod.numpy(), 'class_predictions': h_tlog.numpy()}
    h_raw_i = {'encoded_boxes': h_icod.numpy(), 'class_predictions': h_ilog.numpy()}
    ssd_ht = pkg.SSD.from_predictions(H, W, h_raw_t, gen, C)
    ssd_hi = pkg.SSD.from_predictions(H, W, h_raw_i, gen, C)
    if world > 1:
        ssd_ht.process_group = True
    M = C * K_PER_CLASS
    h_out = {'boxes': pinned([Bi, M, 4], torch.float32).numpy(), 'scores': pinned([Bi, M], torch.float32).numpy(),
             'labels': pinned([Bi, M], torch.int32).numpy(), 'num_boxes': pinned([Bi], torch.int32).numpy()}

    def step_host():
        lo = ssd_ht.loss(gt, PARAMS)
        pr = ssd_hi.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS, out=h_out)
        return lo, pr
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    for _ in range(2):
        eo = step_host()
    e2e_total_ms, e2e_win, eo = timed_loop(step_host, e2e_steps)
    if sampler:
        sampler.window(*e2e_win)
    e2e_ms = e2e_total_ms / e2e_steps
    h2d = sum(t.numel() * t.element_size() for t in (h_tlog, h_tcod, h_ilog, h_icod)) + 2 * anchors_np.nbytes + \
        sum(v.nbytes for v in gt.values())
    d2h = sum(v.nbytes for v in h_out.values()) + 32
    e2e = {'value': (Bt + Bi) * world / (e2e_ms * 1e-3), 'unit': 'images/s', 'h2d_bytes_per_step': int(h2d),
           'd2h_bytes_per_step': int(d2h), 'ms_per_step': e2e_ms, 'steps': e2e_steps,
           'api': 'SSD.loss / SSD.get_predictions with NumPy (pinned) buffers -> ssdk_ssd_targets_and_loss_host, ssdk_postprocess_host',
           'check': {'localization_loss': float(eo[0]['localization_loss']), 'classification_loss': float(eo[0]['classification_loss']),
                     'detections_image0': int(eo[1]['num_boxes'][0])}}
    clocks = sampler.finish() if sampler else None

    if world > 1:
        dist.barrier()
    if rank != 0:
        finish_rank(world)
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        n_t, n_i = args.cpu_sample, 2 * args.cpu_sample
        ips, sec, _ = cpu_reference(n_t, n_i, 2, 1, threads)
        cpu = {'value': ips, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
               'sample': '2 steps x (%d train + %d infer images) of the same workload, oracle port (NumPy f32 op-for-op + C '
                         'NonMaxSuppressionV3), one image per thread-pool task' % (n_t, n_i)}

    if small and '_inputs' in small:
        anc1_np, cod1, log1, gt1, C1 = small.pop('_inputs')
        if world == 1 and not args.no_cpu_baseline:            # the same single image through the CPU port
            from oracle import ssd as ossd
            t0 = time.perf_counter()
            o1 = ossd.loss(anc1_np, cod1, log1, gt1, PARAMS, C1)
            small['cfg1_cpu_port_ms'] = (time.perf_counter() - t0) * 1e3
            small['cfg1_cpu_port_check'] = {'localization_loss': float(o1['localization_loss']), 'classification_loss': float(o1['classification_loss'])}
    line = {
        'metric': 'images_per_sec_target_assign_focal_loss_and_decode_nms_896x640',
        'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(syn, all_reduce=(
            None if world == 1 else ('NVLink peer-memory kernel fused with the loss finalisation (csrc/comm.cu)' if peer else 'NCCL'))),
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'launch_mode': mode,
        'roofline': dominant,
        'cpu_baseline': cpu,
        'breakdown': {
            'eager_ms_per_step': eager_ms_per_step, 'eager_images_per_sec': (Bt + Bi) * world / (eager_ms_per_step * 1e-3),
            'sequential_graph_ms_per_step': ms_sequential_graph,
            'sequential_graph_images_per_sec': None if not ms_sequential_graph else (Bt + Bi) * world / (ms_sequential_graph * 1e-3),
            'train_images_per_sec': Bt * world / (ms_train * 1e-3), 'train_ms_per_step': ms_train,
            'train_frac_of_hbm_roofline': (b_train * Bt / (ms_train * 1e-3) / 1e9) / peak,
            'infer_images_per_sec': Bi * world / (ms_infer * 1e-3), 'infer_ms_per_step': ms_infer,
            'infer_frac_of_hbm_roofline': (b_infer * Bi / (ms_infer * 1e-3) / 1e9) / peak,
            'algorithmic_bytes_per_image': {'train': b_train, 'infer': b_infer},
            'kernel_ms_per_step': step_kernel_ms,
            'roofline_loss_flat_pass': roof_flat, 'roofline_ssd_loss': roof_loss, 'roofline_filter': roof_filter,
            'roofline_ssd_loss_backward': roof_backward,
            'train_fwd_bwd_images_per_sec': Bt * world / (ms_train_fb * 1e-3), 'train_fwd_bwd_ms_per_step': ms_train_fb,
            'train_fwd_bwd_frac_of_hbm_roofline': ((b_train + 8 * A * C // 2 + 16 * A) * Bt / (ms_train_fb * 1e-3) / 1e9) / peak,
            'train_fwd_bwd_ms_per_step_with_targets_assigned_earlier': ms_train_fb_pre,
            'train_fwd_bwd_with_targets_assigned_earlier_frac_of_hbm_roofline': ((8 * A * C + 56 * A) * Bt / (ms_train_fb_pre * 1e-3) / 1e9) / peak,
            'head_layout': head,
            'small_cases': small,
            'sub_path_timings_that_fell_back_to_eager_launches': eager_fallbacks,
            'train_ms_per_step_with_nccl_all_reduce': ms_train_nccl,
        },
        'check': check,
    }
    print(json.dumps(line), flush=True)
    finish_rank(world)


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
