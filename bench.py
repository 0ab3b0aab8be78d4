"""bench.py -- images/sec of the per-anchor detection hot path at 896x640 (BASELINE.json metric).

One "step" per GPU = the training-side path (target assignment + focal / smooth-L1 loss, SSD.loss) over one
cfg2 batch (16 images, 107,415 anchors, 90 classes, 20 GT boxes/image) PLUS the inference-side path
(sigmoid + decode + per-class NMS, SSD.get_predictions, 0.05 / 0.5 / 100) over one cfg3 batch (32 images).
value = images processed by all ranks per second ((16 + 32) * n_gpus / step time); the two sub-paths are also
reported separately under "breakdown", together with BASELINE.json's other configurations: configs[3] (batch 256 sharded over
the ranks, strong scaling, `breakdown.cfg4_strong`), configs[4] (stress, `breakdown.stress`) and configs[0] (`small_cases`).
At N > 1 the line carries `check.sharded_equals_single`: rank 0 gathers every rank's inputs and results over NCCL, recomputes
the whole batch on its own GPU and compares (matches / detections bit for bit, losses to 1e-6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

--impl reference times the reference's CPU semantics (the NumPy/C oracle port; TensorFlow cannot be installed
here) on the host cores, on a bounded sample of the same workload -- the same sample, threads and code as the GPU arm's
`cpu_baseline` leg.
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
    ap.add_argument('--no-extras', action='store_true', help='skip the cfg4 / stress / head-layout / small-case sections (profiling runs)')
    ap.add_argument('--no-numa', action='store_true', help='do not bind the rank to the CPU cores / memory of its GPU\'s NUMA node')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workload
def algorithmic_bytes(A, C, G, K):
    """bytes per image, each tensor counted once."""
    train = 4 * A * C + 56 * A + 20 * G
    infer = 4 * A * C + 32 * A + 24 * C * K + 4
    loss_kernel = 4 * A * C + 16 * A + 16 * A + 8 * A          # logits + codes + reg_targets + (cls, matches)
    filter_kernel = 4 * A * C                                   # scores read once (+ the few candidates written)
    return train, infer, loss_kernel, filter_kernel


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
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_size(threads):
    """The bounded sample both arms time: whole steps (16 train + 32 infer images each), as many as give at least
    2 x `threads` one-image tasks, so that the thread pool is never under-subscribed."""
    steps_worth = max(1, -(-2 * threads // 48))
    return 16 * steps_worth, 32 * steps_worth


def cpu_reference(n_train, n_infer, steps, warmup, threads, budget_s=None):
    """Reference semantics on the host cores: the oracle port (NumPy float32 op for op + C NonMaxSuppressionV3),
    one image per task on a thread pool of `threads` workers, as the reference's tf.map_fn does with parallel_iterations.
    Returns (images/s, seconds per pass, last result, passes timed)."""
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

    def one_pass(pool):
        futs = [pool.submit(train_image, b) for b in range(n_train)] + [pool.submit(infer_image, b) for b in range(n_infer)]
        res = [f.result() for f in futs]
        tr = np.array(res[:n_train], np.float64).sum(axis=0)
        norm = max(tr[2], 1.0)
        # per-image results too: the GPU arm checks its first batch against them (same seeded inputs)
        return tr[0] / norm, tr[1] / norm, sum(res[n_train:]), np.array(res[:n_train], np.float64), np.array(res[n_train:], np.int64)

    with ThreadPoolExecutor(max_workers=threads) as pool:
        t0 = time.perf_counter()
        out = one_pass(pool)                                   # first pass: also the estimate for the time budget
        first = time.perf_counter() - t0
        if budget_s is not None and first * (steps + warmup) > budget_s:
            total = max(2, int(budget_s / max(first, 1e-3)))
            warmup = min(warmup, max(1, total // 5))
            steps = max(1, total - warmup)
        for _ in range(max(0, warmup - 1)):
            one_pass(pool)
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            out = one_pass(pool)
            times.append(time.perf_counter() - t0)
        dt = float(np.median(times))                           # the median pass: a pass disturbed by another tenant of the host does not count
    return (n_train + n_infer) / dt, dt, out, steps


def cpu_baseline_object(ips, threads, n_train, n_infer, passes):
    return {'value': ips, 'unit': 'images/s', 'cores': threads, 'kind': 'port', 'images_per_s_per_core': ips / max(1, threads),
            'sample': 'median of %d timed passes over %d train + %d infer images (= whole steps of the workload; %d one-image tasks on %d '
                      'threads), oracle port: NumPy float32 op for op + C NonMaxSuppressionV3' % (passes, n_train, n_infer,
                                                                                                 n_train + n_infer, threads)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = host_threads()
    n_train, n_infer = cpu_sample_size(threads)
    ips, sec, _, passes = cpu_reference(n_train, n_infer, max(1, args.steps), max(1, args.warmup), threads, budget_s=150.0)
    syn = importlib.import_module(PKG + '.synthetic')
    # a "step" of this arm = 16 + 32 images, as in the GPU arm: ms_per_step is scaled to that
    line = {
        'impl': 'reference', 'metric': 'images_per_sec_target_assign_focal_loss_and_decode_nms_896x640',
        'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'steps_timed': passes, 'ms_per_step': 48.0 / ips * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(syn),
        'cpu_baseline': cpu_baseline_object(ips, threads, n_train, n_infer, passes),
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(syn):
    """Identical in both arms (the driver compares it)."""
    tc, ic = syn.CONFIGS[TRAIN_CFG], syn.CONFIGS[INFER_CFG]
    return {
        'workload': 'per step and GPU: SSD.loss (targets + focal/smooth-L1) on cfg2 batch %d  +  SSD.get_predictions '
                    '(sigmoid, decode, per-class NMS %.2f/%.1f/%d) on cfg3 batch %d; 640x896, 5 FPN levels x 9 anchors '
                    '= 107415 anchors, 90 classes, 20 GT boxes/image' % (tc['B'], SCORE_THR, IOU_THR, K_PER_CLASS, ic['B']),
        'train_batch_per_gpu': tc['B'], 'infer_batch_per_gpu': ic['B'],
        'logits': 'train: N(-4.595,1) prior-bias init; infer: N(-7,1) background + N(1.5,1.5) on anchors with IoU>=0.4 to a GT',
        'l2': 'inputs per step (0.62 GB + 1.24 GB of logits) exceed the 126 MB L2; no flush needed',
        'parallelism': 'image-sharded, one all-reduce of 3 doubles per step',
        'launch': 'one CUDA graph per step; the training-side and the inference-side sub-path are independent and are issued on '
                  'two streams inside it (breakdown.sequential_graph_* = the same graph on one stream; launch_mode says which arrangement was fastest)',
    }


# ------------------------------------------------------------------------------------------------ NUMA placement
def bind_to_gpu_numa_node(gpu_index):
    """Pin this process (and with it the pinned host buffers it allocates afterwards: first touch) to the CPU cores of the
    NUMA node the GPU hangs off.  Returns a description for the JSON line."""
    info = {'bound': False}
    try:
        bus = subprocess.run(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=pci.bus_id', '--format=csv,noheader'],
                             stdout=subprocess.PIPE, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith('00000000:'):
            bus = bus[4:]
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read().strip())
        info.update(pci_bus_id=bus, numa_node=node)
        if node < 0:
            return info
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as e:
        info['error'] = str(e)[:120]
    return info


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    all_cpus = set(os.sched_getaffinity(0))
    numa = {'bound': False} if (args.no_numa or world == 1) else bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    pkg = importlib.import_module(PKG)
    syn = importlib.import_module(PKG + '.synthetic')
    lib = pkg._lib
    captured_steps = []

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

    # private workspace the inference sub-path needs (bounded candidate regions: 32 KB per (image, class) segment), measured before
    # anything else touches the context
    ws_before = lib.workspace_bytes(local_rank)
    ssd_i.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS)
    torch.cuda.synchronize()
    ws_post = lib.workspace_bytes(local_rank) - ws_before

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

    def capture(fn):
        cap = pkg.graph.capture(fn, warmup=2)
        captured_steps.append(cap)
        return cap

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
            captured = capture(step_resident)
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
    #      kernels, same inputs, same results): sorting / NMS / packing hide behind the other sub-path's HBM-bound pass
    if mode == 'cuda_graph' and not args.no_overlap:
        try:
            step_overlapped = pkg.graph.concurrent(lambda: ssd_t.loss(d_gt, PARAMS),
                                                   lambda: ssd_i.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS))
            captured_o = capture(step_overlapped)
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
    # ---- timed region 1d: the same step with the post-processing split in its two phases: the HBM-bound score scan first (alone on
    #      the GPU), then the latency-bound rest (sort / NMS / pack) next to the training-step kernel.  Same kernels, same results.
    if mode.startswith('cuda_graph') and not args.no_overlap:
        try:
            rest_and_train = pkg.graph.concurrent(lambda: ssd_i.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS, phase='finish'),
                                                  lambda: ssd_t.loss(d_gt, PARAMS))

            def step_split():
                ssd_i.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS, phase='scan')
                p_, l_ = rest_and_train()
                return l_, p_
            captured_s = capture(step_split)
            for _ in range(3):
                captured_s.replay()
            ms_s, win_s, out_s = timed_loop(captured_s.replay, args.steps)
            same = (float(out_s[0]['localization_loss']) == float(out[0]['localization_loss'])
                    and float(out_s[0]['classification_loss']) == float(out[0]['classification_loss'])
                    and all(torch.equal(out_s[1][k], out[1][k]) for k in ('boxes', 'labels', 'scores', 'num_boxes')))
            if same and ms_s / args.steps < ms_per_step:
                ms_per_step, win, out = ms_s / args.steps, win_s, out_s
                launches = captured_s.launches_per_replay * args.steps
                mode = 'cuda_graph, score scan first, then the training step next to the NMS stages (two streams)'
        except Exception:
            torch.cuda.synchronize()
    if sampler:
        sampler.window(*win)
    value = (Bt + Bi) * world / (ms_per_step * 1e-3)
    losses, pred = out
    check = {'localization_loss': float(losses['localization_loss']), 'classification_loss': float(losses['classification_loss']),
             'num_matches_all_ranks': float(ssd_t.num_matches), 'detections_image0': int(pred['num_boxes'][0]),
             'detections_this_rank': int(pred['num_boxes'].sum())}

    # ---- sub-path timings (same resident inputs), each its own event-timed loop
    eager_fallbacks = []

    def timed(fn, n):
        """ms per call of one sub-path, launched the same way as the headline number (graph replay when available)."""
        run, cap = fn, None
        if mode.startswith('cuda_graph'):
            try:
                cap = pkg.graph.capture(fn, warmup=2)
                run = cap.replay
            except Exception as e:
                eager_fallbacks.append(str(e)[:120])
                torch.cuda.synchronize()
        for _ in range(3):
            run()
        ms = timed_loop(run, n)[0] / n
        if cap is not None:
            cap.release()
        return ms
    ms_train = timed(lambda: ssd_t.loss(d_gt, PARAMS), args.steps)
    ms_infer = timed(lambda: ssd_i.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS), args.steps)
    # forward + backward of the training side: targets, losses, all-reduce, gradients w.r.t. both heads
    ms_train_fb = timed(lambda: ssd_t.loss_with_gradients(d_gt, PARAMS, upstream=UPSTREAM), args.steps)
    # the same fused step when the targets were assigned earlier (SSD.assign_targets needs only anchors + ground truth, so a
    # training loop can run it on a side stream during the network's forward pass): the streaming pass alone.  NOT the
    # headline -- the matching work is outside this timed region; it shows what remains on the critical path.
    pre_targets = pkg.SSD.assign_targets(ssd_t.anchors, d_gt)
    ms_train_fb_pre = timed(lambda: ssd_t.loss_with_gradients(None, PARAMS, upstream=UPSTREAM, targets=pre_targets), args.steps)
    # the training sub-path as separate launches (matcher, flat pass, matched-anchor pass, finalisation) instead of the one fused kernel
    lib.set_option(lib.SSDK_OPT_FUSED_TRAIN_STEP, 0, local_rank)
    ms_train_unfused = timed(lambda: ssd_t.loss(d_gt, PARAMS), args.steps)
    lib.set_option(lib.SSDK_OPT_FUSED_TRAIN_STEP, 1, local_rank)
    # the fused kernel with the static split of the chunk list (double accumulation in a fixed order) instead of dynamically handed-out chunks
    lib.set_option(lib.SSDK_OPT_TRAIN_DYNAMIC_CHUNKS, 0, local_rank)
    ms_train_static = timed(lambda: ssd_t.loss(d_gt, PARAMS), args.steps)
    l_static = ssd_t.loss(d_gt, PARAMS)
    lib.set_option(lib.SSDK_OPT_TRAIN_DYNAMIC_CHUNKS, 1, local_rank)
    l_dyn = ssd_t.loss(d_gt, PARAMS)
    static_vs_dynamic = max(abs(float(l_static[k]) - float(l_dyn[k])) / abs(float(l_dyn[k])) for k in ('localization_loss', 'classification_loss'))
    ms_train_nccl = None
    if world > 1 and peer:                       # the same training sub-path with the library collective, for comparison
        ssd_t.peer_all_reduce = False
        ms_train_nccl = timed(lambda: ssd_t.loss(d_gt, PARAMS), args.steps)
        ssd_t.peer_all_reduce = True

    # ---- per-kernel durations (library-side CUDA events on the launching stream) for the roofline object
    nprof = min(args.steps, 10)
    lib.set_profiling(True, local_rank)
    lib.profile_read(local_rank)
    for _ in range(nprof):
        step_resident()
    prof_step = lib.profile_read(local_rank)              # exactly the kernels of min(steps, 10) steps
    ssd_t._loss_forward(d_gt, PARAMS, keep_targets=True)
    sv = ssd_t._saved
    sums_tmp = torch.empty([3], dtype=torch.float64, device=dev)
    import ctypes
    flat_head = pkg.HeadPredictions([d_tcod.reshape(Bt, A, 1, 4)], [d_tlog.reshape(Bt, A, 1, C)], C, 1, 'channels_last')
    lib.profile_read(local_rank)
    for _ in range(nprof):
        ssd_t.loss_backward(UPSTREAM)
        ctx_h = lib.context(local_rank)
        # the row-tiled forward kernel (ssdk_ssd_loss: targets given) and the stand-alone flat pass + matched-anchor pass
        # (ssdk_head_ssd_loss: targets given): the un-fused building blocks, not part of the step any more
        lib.check(lib.load().ssdk_ssd_loss(ctx_h, sv['logits'].data_ptr(), sv['codes'].data_ptr(), sv['reg_targets'].data_ptr(),
                                           sv['cls_targets'].data_ptr(), sv['matches'].data_ptr(), Bt, A, C, PARAMS['gamma'], PARAMS['alpha'],
                                           sums_tmp.data_ptr(), None, None))
        d_flat = flat_head.descriptor()
        lib.check(lib.load().ssdk_head_ssd_loss(ctx_h, ctypes.byref(d_flat), sv['reg_targets'].data_ptr(), sv['cls_targets'].data_ptr(),
                                                sv['matches'].data_ptr(), Bt, A, C, PARAMS['gamma'], PARAMS['alpha'], sums_tmp.data_ptr()))
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

    # DRAM traffic per launch: NOT measured by this run -- read from the committed `ncu --set full` capture of the same kernels
    # (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py on the GPU box); null for kernels that capture lacks
    try:
        ncu_traffic = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
    except Exception:
        ncu_traffic = {}
    traffic_source = ('profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full '
                      'capture "%s" of this command; a constant of that capture, not a measurement of this run' % ncu_traffic.get('_capture', 'unknown'))

    def kernel_roof(p, name, bytes_per_launch, traffic_key=None, label=None):
        tot, n = p[name]
        if n == 0:
            return None
        avg_ms = tot / n
        ach = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        tr = (ncu_traffic.get(traffic_key or (name + '_kernel')) or {}).get('dram_bytes_per_launch')
        return {'kernel': label or name, 'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                'traffic': tr, 'traffic_source': traffic_source if tr is not None else None, 'avg_launch_ms': avg_ms,
                'algorithmic_bytes_per_launch': bytes_per_launch, 'peak_source': peak_src}
    # SSD.loss = ONE kernel (train_step): matcher CTAs + streaming CTAs + final reduction; its algorithmic bytes are the whole
    # training-side figure of SURVEY.md 8(d) (4AC + 56A + 20G per image).  head_flat: the flat pass alone (targets given).
    roof_step = kernel_roof(prof, 'train_step', b_train * Bt, 'train_step_kernel')
    roof_flat = kernel_roof(prof, 'head_flat', 4 * A * C * Bt, 'head_flat_forward_kernel', 'head_flat (flat pass alone, targets given)')
    roof_loss = kernel_roof(prof, 'ssd_loss', b_loss * Bt)
    roof_filter = kernel_roof(prof, 'filter', b_filter * Bi)
    b_backward = 8 * A * C + 56 * A                 # logits read + grad written; codes, reg_targets, grad_codes, cls, matches
    roof_backward = kernel_roof(prof, 'ssd_loss_backward', b_backward * Bt)
    step_kernel_ms = {k: (v[0] / max(1, nprof)) for k, v in prof_step.items() if v[1]}
    step_kernel_ms['note'] = 'eager launches with event pairs around every kernel (a few microseconds more than inside the replayed graph)'
    dominant = max((r for r in (roof_step, roof_filter) if r), key=lambda r: r['avg_launch_ms'])
    dominant = dict(dominant)
    dominant['share_of_step_kernel_time'] = dominant['avg_launch_ms'] / max(1e-9, sum(v for k, v in step_kernel_ms.items() if k != 'note'))

    extras = {}
    if not args.no_extras:
        extras = run_extras(args, pkg, syn, lib, dev, local_rank, rank, world, peer, gen, anchors, A, C, H, W, G, Bt, Bi, d_tlog, d_tcod, d_ilog,
                            d_icod, d_gt, pred, timed, peak, peak_src, ncu_traffic, traffic_source, ms_train, ms_train_fb, ms_infer,
                            b_infer)

    # ---- N > 1: do the shards reproduce the single-GPU result?  Rank 0 gathers every rank's inputs and results (NCCL over
    #      NVLink), recomputes the whole batch alone and compares.
    sharded = None
    if world > 1:
        sharded = sharded_equals_single(pkg, dist, dev, rank, world, gen, H, W, C, A, Bt, Bi, d_tlog, d_tcod, d_ilog, d_icod, d_gt, ssd_t,
                                        losses, pred)

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
           'h2d_gb_per_s_per_gpu': h2d / (e2e_ms * 1e-3) / 1e9,
           'host_placement': numa,
           'api': 'SSD.loss / SSD.get_predictions with NumPy (pinned) buffers -> ssdk_ssd_targets_and_loss_host, ssdk_postprocess_host',
           'check': {'localization_loss': float(eo[0]['localization_loss']), 'classification_loss': float(eo[0]['classification_loss']),
                     'detections_image0': int(eo[1]['num_boxes'][0])}}
    clocks = sampler.finish() if sampler else None

    # ---- teardown: graphs first, then the process group (graph.CapturedStep.release documents the order)
    for cap in captured_steps:
        cap.release()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)
        threads = host_threads()
        n_t, n_i = cpu_sample_size(threads)
        ips, sec, cpu_out, passes = cpu_reference(n_t, n_i, 3, 1, threads)
        cpu = cpu_baseline_object(ips, threads, n_t, n_i, passes)
        # the CPU port ran the same seeded images as this rank's batches: compare whole-batch results (the checker role of oracle/)
        tr = cpu_out[3][:Bt].sum(axis=0)
        o_loc, o_cls, o_n = tr[0] / max(tr[2], 1.0), tr[1] / max(tr[2], 1.0), tr[2]
        o_det = cpu_out[4][:Bi]
        rel = max(abs(check['localization_loss'] - o_loc) / abs(o_loc), abs(check['classification_loss'] - o_cls) / abs(o_cls))
        det_gpu = pred['num_boxes'].cpu().numpy().astype(np.int64)
        check['vs_cpu_port'] = {'train_images': int(Bt), 'infer_images': int(Bi), 'losses_max_rel_diff': float(rel),
                                'num_matches_equal': bool(o_n == check['num_matches_all_ranks']),
                                'detections_per_image_equal': bool(np.array_equal(det_gpu, o_det)),
                                'agrees': bool(rel <= 1e-5 and o_n == check['num_matches_all_ranks'] and np.array_equal(det_gpu, o_det))}

    small = extras.get('small_cases')
    if small and '_inputs' in small:
        anc1_np, cod1, log1, gt1, C1 = small.pop('_inputs')
        if world == 1 and not args.no_cpu_baseline:            # the same single image through the CPU port
            from oracle import ssd as ossd
            t0 = time.perf_counter()
            o1 = ossd.loss(anc1_np, cod1, log1, gt1, PARAMS, C1)
            small['cfg1_cpu_port_ms'] = (time.perf_counter() - t0) * 1e3
            small['cfg1_cpu_port_check'] = {'localization_loss': float(o1['localization_loss']), 'classification_loss': float(o1['classification_loss'])}
    if sharded is not None:
        check['sharded_equals_single'] = sharded.pop('equal')
        check['sharded_vs_single'] = sharded
    line = {
        'metric': 'images_per_sec_target_assign_focal_loss_and_decode_nms_896x640',
        'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(syn),
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'launch_mode': mode,
        'roofline': dominant,
        'cpu_baseline': cpu,
        'breakdown': {
            'all_reduce': None if world == 1 else ('NVLink peer-memory exchange inside the training-step kernel\'s last CTA (csrc/train_step.cu, comm.cuh)'
                                                   if peer else 'NCCL'),
            'eager_ms_per_step': eager_ms_per_step, 'eager_images_per_sec': (Bt + Bi) * world / (eager_ms_per_step * 1e-3),
            'sequential_graph_ms_per_step': ms_sequential_graph,
            'sequential_graph_images_per_sec': None if not ms_sequential_graph else (Bt + Bi) * world / (ms_sequential_graph * 1e-3),
            'train_images_per_sec': Bt * world / (ms_train * 1e-3), 'train_ms_per_step': ms_train,
            'train_frac_of_hbm_roofline': (b_train * Bt / (ms_train * 1e-3) / 1e9) / peak,
            'train_static_split_ms_per_step': ms_train_static, 'train_static_vs_dynamic_losses_rel_diff': static_vs_dynamic,
            'train_unfused_ms_per_step': ms_train_unfused,
            'train_unfused_frac_of_hbm_roofline': (b_train * Bt / (ms_train_unfused * 1e-3) / 1e9) / peak,
            'infer_images_per_sec': Bi * world / (ms_infer * 1e-3), 'infer_ms_per_step': ms_infer,
            'infer_frac_of_hbm_roofline': (b_infer * Bi / (ms_infer * 1e-3) / 1e9) / peak,
            'algorithmic_bytes_per_image': {'train': b_train, 'infer': b_infer},
            'kernel_ms_per_step': step_kernel_ms,
            'roofline_train_step': roof_step, 'roofline_loss_flat_pass': roof_flat, 'roofline_ssd_loss': roof_loss,
            'roofline_filter': roof_filter, 'roofline_ssd_loss_backward': roof_backward,
            'train_fwd_bwd_images_per_sec': Bt * world / (ms_train_fb * 1e-3), 'train_fwd_bwd_ms_per_step': ms_train_fb,
            'train_fwd_bwd_frac_of_hbm_roofline': ((b_train + 8 * A * C // 2 + 16 * A) * Bt / (ms_train_fb * 1e-3) / 1e9) / peak,
            'train_fwd_bwd_ms_per_step_with_targets_assigned_earlier': ms_train_fb_pre,
            'train_fwd_bwd_with_targets_assigned_earlier_frac_of_hbm_roofline': ((8 * A * C + 56 * A) * Bt / (ms_train_fb_pre * 1e-3) / 1e9) / peak,
            'workspace_bytes': lib.workspace_bytes(local_rank),
            'postprocess_workspace_bytes_cfg3': int(ws_post), 'postprocess_workspace_over_logits_bytes': ws_post / float(d_ilog.numel() * 4),
            'workspace_note': 'workspace_bytes = everything this context ever needed in this run (incl. the staging buffers of the e2e loop and the batch-256 targets of cfg4)',
            'cfg4_strong': extras.get('cfg4_strong'),
            'stress': extras.get('stress'),
            'head_layout': extras.get('head_layout'),
            'small_cases': small,
            'sub_path_timings_that_fell_back_to_eager_launches': eager_fallbacks,
            'train_ms_per_step_with_nccl_all_reduce': ms_train_nccl,
        },
        'check': check,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sharded_equals_single(pkg, dist, dev, rank, world, gen, H, W, C, A, Bt, Bi, d_tlog, d_tcod, d_ilog, d_icod, d_gt, ssd_t, losses, pred):
    """Every rank's inputs, matches and detections go to rank 0 (dist.gather over NCCL); rank 0 runs the whole batch (16N train,
    32N infer images) on its own GPU and compares: matches and detections bit for bit, num_matches exactly, losses to 1e-6."""
    import torch
    ssd_t._loss_forward(d_gt, PARAMS, keep_targets=True)
    my_matches = ssd_t._saved['matches']

    def gather(t):
        t = t.contiguous()
        parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, parts, dst=0)
        return torch.cat(parts) if rank == 0 else None
    g_tlog, g_tcod, g_ilog, g_icod = gather(d_tlog), gather(d_tcod), gather(d_ilog), gather(d_icod)
    g_gt = {k: gather(v) for k, v in d_gt.items()}
    g_matches = gather(my_matches)
    g_pred = {k: gather(pred[k]) for k in ('boxes', 'labels', 'scores', 'num_boxes')}
    g_loss = gather(torch.stack([losses['localization_loss'], losses['classification_loss']]).reshape(1, 2))
    if rank != 0:
        return None
    single_t = pkg.SSD.from_predictions(H, W, {'encoded_boxes': g_tcod, 'class_predictions': g_tlog}, gen, C)
    s_loss = single_t._loss_forward(g_gt, PARAMS, keep_targets=True)
    s_matches = single_t._saved['matches']
    single_i = pkg.SSD.from_predictions(H, W, {'encoded_boxes': g_icod, 'class_predictions': g_ilog}, gen, C)
    s_pred = single_i.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS)
    s_l = torch.stack([s_loss['localization_loss'], s_loss['classification_loss']]).double()
    rel = ((g_loss.double() - s_l[None]).abs() / s_l[None].abs().clamp_min(1e-30)).max().item()
    det_equal = all(bool(torch.equal(g_pred[k], s_pred[k])) for k in g_pred)
    matches_equal = bool(torch.equal(g_matches, s_matches))
    count_equal = float(single_t.num_matches) == float(ssd_t.num_matches)
    return {'equal': bool(det_equal and matches_equal and count_equal and rel <= 1e-6),
            'images': {'train': Bt * world, 'infer': Bi * world}, 'matches_bit_identical': matches_equal,
            'detections_bit_identical': det_equal, 'num_matches_identical': count_equal,
            'losses_max_rel_diff_over_ranks': rel, 'losses_bit_identical': bool((g_loss.double() == s_l[None]).all()),
            'all_ranks_returned_the_same_losses': bool((g_loss == g_loss[0:1]).all()),
            'single_gpu_losses': [float(s_l[0]), float(s_l[1])], 'single_gpu_detections': int(s_pred['num_boxes'].sum())}


def run_extras(args, pkg, syn, lib, dev, local_rank, rank, world, peer, gen, anchors, A, C, H, W, G, Bt, Bi, d_tlog, d_tcod, d_ilog, d_icod,
               d_gt, pred, timed, peak, peak_src, ncu_traffic, traffic_source, ms_train, ms_train_fb, ms_infer, b_infer):
    """BASELINE.json's other configurations and the head-layout path; none of them feeds the headline."""
    import torch
    import torch.distributed as dist
    out = {}

    # ---- configs[3]: batch 256 at 640x896 sharded over the ranks (128 / 64 / 32 images per GPU at 2 / 4 / 8), one all-reduce of
    #      the loss sums per step: STRONG scaling.  Inputs are generated on the device from per-image seeds, so rank 0 can
    #      also run the whole batch alone: single-GPU time (same run, same GPU model) and a sharded == single check.
    try:
        c4 = syn.CONFIGS[4]
        B4 = c4['B']
        lo, hi = pkg.parallel.shard_range(B4, rank, world)

        def gen_images(first, last):
            lg = torch.empty([last - first, A, C], dtype=torch.float32, device=dev)
            cd = torch.empty([last - first, A, 4], dtype=torch.float32, device=dev)
            g_ = torch.Generator(device=dev)
            for i in range(first, last):
                g_.manual_seed(4_000_000 + i)
                lg[i - first].normal_(-4.595, 1.0, generator=g_)
                cd[i - first].normal_(0.0, 1.0, generator=g_)
            return lg, cd
        gt4_all = syn.make_groundtruth(4, B4, c4['G'], H, W, C)
        lg4, cd4 = gen_images(lo, hi)
        gt4 = {k: torch.from_numpy(v[lo:hi]).to(dev) for k, v in gt4_all.items()}
        ssd4 = pkg.SSD.from_predictions(H, W, {'encoded_boxes': cd4, 'class_predictions': lg4}, gen, C)
        if world > 1:
            ssd4.process_group = True
            ssd4.peer_all_reduce = peer
        ms4 = timed(lambda: ssd4.loss(gt4, PARAMS), args.steps)
        ms4_fb = timed(lambda: ssd4.loss_with_gradients(gt4, PARAMS, upstream=UPSTREAM), max(3, args.steps // 3))
        l4 = ssd4._loss_forward(gt4, PARAMS, keep_targets=True)
        b_train4 = 4 * A * C + 56 * A + 20 * c4['G']
        res = {'what': 'BASELINE.json configs[3]: batch %d sharded over %d rank(s), %d images per GPU, all-reduce of the loss sums inside the step' % (B4, world, hi - lo),
               'images': B4, 'images_per_gpu': hi - lo, 'ms_per_step': ms4, 'images_per_sec': B4 / (ms4 * 1e-3),
               'frac_of_hbm_roofline_per_gpu': (b_train4 * (hi - lo) / (ms4 * 1e-3) / 1e9) / peak,
               'fwd_bwd_ms_per_step': ms4_fb, 'fwd_bwd_images_per_sec': B4 / (ms4_fb * 1e-3),
               'localization_loss': float(l4['localization_loss']), 'classification_loss': float(l4['classification_loss']),
               'num_matches': float(ssd4.num_matches)}
        if world > 1:
            my_m = ssd4._saved['matches'].contiguous()
            parts = [torch.empty_like(my_m) for _ in range(world)] if rank == 0 else None
            dist.gather(my_m, parts, dst=0)
            del ssd4, lg4, cd4
            torch.cuda.empty_cache()
            if rank == 0:                                         # the whole batch on one GPU, while the other ranks wait
                lgf, cdf = gen_images(0, B4)
                gtf = {k: torch.from_numpy(v).to(dev) for k, v in gt4_all.items()}
                full = pkg.SSD.from_predictions(H, W, {'encoded_boxes': cdf, 'class_predictions': lgf}, gen, C)
                lf = full._loss_forward(gtf, PARAMS, keep_targets=True)
                cap1 = pkg.graph.capture(lambda: full.loss(gtf, PARAMS), warmup=2)      # launched like the sharded step: one graph replay
                for _ in range(3):
                    cap1.replay()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                ev0.record()
                for _ in range(args.steps):
                    cap1.replay()
                ev1.record()
                torch.cuda.synchronize()
                ms_single = ev0.elapsed_time(ev1) / args.steps
                cap1.release()
                rel = max(abs(float(lf[k]) - float(l4[k])) / max(abs(float(lf[k])), 1e-30) for k in lf)
                res.update(single_gpu_ms_per_step_same_run=ms_single, speedup_vs_single_gpu=ms_single / ms4,
                           strong_scaling_efficiency=ms_single / ms4 / world,
                           sharded_equals_single=bool(torch.equal(torch.cat(parts), full._saved['matches']) and rel <= 1e-6
                                                      and float(full.num_matches) == float(res['num_matches'])),
                           losses_rel_diff_vs_single=rel)
                del full, lgf, cdf
            dist.barrier()
        else:
            del ssd4, lg4, cd4
        torch.cuda.empty_cache()
        out['cfg4_strong'] = res
    except Exception as e:
        out['cfg4_strong'] = {'error': str(e)[:300]}
        torch.cuda.synchronize()

    # ---- configs[4] (stress): 1344x896, 300 GT boxes per image, shipped 6-anchor set (150,402 anchors, 80 classes), dense scores
    try:
        if rank == 0:
            out['stress'] = stress_section(pkg, syn, lib, dev, local_rank, peak, args)
        if world > 1:
            dist.barrier()
    except Exception as e:
        out['stress'] = {'error': str(e)[:300]}
        torch.cuda.synchronize()

    # ---- head-layout path: the same two sub-paths fed with the per-level tower outputs
    #      [B, n*C, h, w] / [B, n*4, h, w] (channels_first, as the reference's box predictor emits them) instead of the
    #      concatenated [B,A,C] / [B,A,4]; plus the cost of reshape_and_concatenate itself, which this path removes
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
            tr = (ncu_traffic.get(traffic_key or (name + '_kernel')) or {}).get('dram_bytes_per_launch')
            return {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                    'avg_launch_ms': avg, 'algorithmic_bytes_per_launch': nbytes, 'peak_source': peak_src,
                    'traffic': tr, 'traffic_source': traffic_source if tr is not None else None}
        bh_train = 4 * A * C + 56 * A + 20 * G                  # the same figure as the anchor-major path
        out['head_layout'] = {
            'what': 'same sub-paths on per-level channels_first tower outputs (no reshape_and_concatenate); results checked against the anchor-major path',
            'check': head_check,
            'train_ms_per_step': ms_h_train, 'train_images_per_sec': Bt * world / (ms_h_train * 1e-3),
            'train_fwd_bwd_ms_per_step': ms_h_fb, 'train_fwd_bwd_images_per_sec': Bt * world / (ms_h_fb * 1e-3),
            'infer_ms_per_step': ms_h_infer, 'infer_images_per_sec': Bi * world / (ms_h_infer * 1e-3),
            'reshape_and_concatenate_ms': {'train_batch': ms_concat_t, 'infer_batch': ms_concat_i,
                                           'note': 'cost of the copy the reference makes before the anchor-major path (8AC+32A bytes per image)'},
            'unfused_train_ms_per_step': ms_concat_t + ms_train, 'unfused_train_fwd_bwd_ms_per_step': ms_concat_t + ms_train_fb,
            'unfused_infer_ms_per_step': ms_concat_i + ms_infer,
            'roofline_train_step': roof_of(prof_f, 'train_step', bh_train * Bt, 'train_step_kernel'),
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
        out['head_layout'] = {'error': str(e)[:300]}
        torch.cuda.synchronize()

    # ---- BASELINE.json configs[0] (the reference's own CPU-runnable case: config_mobilenet anchor set, one 640x640 image,
    #      20 GT boxes, matching + focal loss) and single-image inference latency (inference/detector.py serves one image)
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
        out['small_cases'] = {'cfg1_train_one_image_640x640_ms': ms_cfg1, 'cfg1_anchors': int(A1), 'cfg1_classes': int(C1),
                              'cfg1_check': {'localization_loss': float(l1['localization_loss']), 'classification_loss': float(l1['classification_loss'])},
                              'infer_latency_one_image_640x896_ms': ms_b1, '_inputs': (anc1.cpu().numpy(), cod1, log1, gt1, C1)}
    except Exception as e:
        out['small_cases'] = {'error': str(e)[:300]}
        torch.cuda.synchronize()
    return out


def stress_section(pkg, syn, lib, dev, local_rank, peak, args):
    """BASELINE.json configs[4]: the matcher is FP32 / ALU-bound at 300 boxes per image (reported as IoU pairs per second, SURVEY.md
    8d), the post-processing sees ~73 % of all (anchor, class) pairs above the threshold: every segment overflows its candidate
    region and goes through the rounds of nms_rounds_kernel."""
    import torch
    cfg = syn.CONFIGS[5]
    H, W, C, G, B = cfg['H'], cfg['W'], cfg['C'], cfg['G'], cfg['B']
    gen5 = pkg.AnchorGenerator(scale_multipliers=cfg['scale_multipliers'])
    anc5 = gen5(H, W, device=dev)
    A = anc5.shape[0]
    gt5 = {k: torch.from_numpy(v).to(dev) for k, v in syn.make_groundtruth(5, B, G, H, W, C).items()}
    g5 = torch.Generator(device=dev).manual_seed(5)
    log5 = torch.randn([B, A, C], device=dev, generator=g5) * 1.5 - 2.0          # 'dense'
    cod5 = torch.randn([B, A, 4], device=dev, generator=g5)
    s5 = pkg.SSD.from_predictions(H, W, {'encoded_boxes': cod5, 'class_predictions': log5}, gen5, C)
    n = max(3, args.steps // 3)

    def timed_simple(fn):
        cap = pkg.graph.capture(fn, warmup=2)
        for _ in range(2):
            cap.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            o = cap.replay()
        e1.record()
        torch.cuda.synchronize()
        cap.release()
        return e0.elapsed_time(e1) / n, o
    ms_m, tg = timed_simple(lambda: pkg.SSD.assign_targets(anc5, gt5))
    ms_t, ls = timed_simple(lambda: s5.loss(gt5, PARAMS))
    ms_fb, _ = timed_simple(lambda: s5.loss_with_gradients(gt5, PARAMS))
    ms_pp, p5 = timed_simple(lambda: s5.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS))
    lib.set_profiling(True, local_rank)
    lib.profile_read(local_rank)
    for _ in range(3):
        s5.get_predictions(SCORE_THR, IOU_THR, K_PER_CLASS)
    prof = lib.profile_read(local_rank)
    lib.set_profiling(False, local_rank)
    rt = lib.round_times(local_rank)
    phases = None
    if rt[0] > 0 and rt[7] > 0:
        names = ['init', 'round1_start', 'histogram_pass', 'plan', 'collect_pass', 'nms']
        phases = {names[i]: (rt[i + 2] - rt[i + 1]) * 1e-6 for i in range(6)}
    cand = float((torch.sigmoid(log5[0]) > SCORE_THR).float().mean())
    res = {
        'config': 'stress (BASELINE.json configs[4]): %dx%d, %d anchors (6 per location), %d classes, batch %d, %d GT boxes/image, dense logits N(-2,1.5)'
                  % (H, W, A, C, B, G),
        'matcher': {'ms': ms_m, 'iou_pairs_per_s': B * G * A / (ms_m * 1e-3), 'images_per_s': B / (ms_m * 1e-3),
                    'positives': int((tg['matches'] >= 0).sum()), 'bound': 'FP32/ALU issue (SURVEY.md 8d; ncu: profiles/)'},
        'train_forward': {'ms': ms_t, 'images_per_s': B / (ms_t * 1e-3),
                          'frac_of_hbm_roofline': ((4 * A * C + 56 * A + 20 * G) * B / (ms_t * 1e-3) / 1e9) / peak,
                          'localization_loss': float(ls['localization_loss']), 'classification_loss': float(ls['classification_loss'])},
        'train_forward_backward': {'ms': ms_fb, 'images_per_s': B / (ms_fb * 1e-3),
                                   'frac_of_hbm_roofline': ((8 * A * C + 72 * A + 20 * G) * B / (ms_fb * 1e-3) / 1e9) / peak},
        'postprocess_dense': {'ms': ms_pp, 'images_per_s': B / (ms_pp * 1e-3), 'candidate_fraction_image0': cand,
                              'candidates_per_image': cand * A * C, 'detections_image0': int(p5['num_boxes'][0]),
                              'frac_of_hbm_roofline': ((4 * A * C + 32 * A + 24 * C * K_PER_CLASS + 4) * B / (ms_pp * 1e-3) / 1e9) / peak,
                              'rounds': rt[0], 'round1_phase_ms': phases, 'async_error': lib.async_error(local_rank)},
        'kernel_ms_per_call': {k: v[0] / 3 for k, v in prof.items() if v[1]},
    }
    del s5, log5, cod5
    torch.cuda.empty_cache()
    return res


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
