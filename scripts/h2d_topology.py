"""Why does the end-to-end (host-buffer) number not scale with the number of GPUs?  

One process, one thread per GPU: every GPU copies a pinned 1 GiB host buffer to its HBM over and over (cudaMemcpyAsync, CUDA events),
1 / 2 / 4 / 8 GPUs at the same time, with the pinned buffers placed
  * "default": allocated and first touched by the main thread wherever the kernel puts them, and
  * "numa_local": allocated and first touched by a thread bound to the CPU cores of the GPU's own NUMA node
    (/sys/bus/pci/devices/<bus id>/numa_node), which is what bench.py does per rank.
Prints the table as JSON, preceded by `nvidia-smi topo -m` and the NUMA layout.   python scripts/h2d_topology.py > profiles/<tag>_h2d_topology.txt"""
import json
import os
import subprocess
import threading

import torch


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60).stdout
    except Exception as e:
        return 'failed: %s' % e


def gpu_numa(i):
    bus = sh('nvidia-smi -i %d --query-gpu=pci.bus_id --format=csv,noheader' % i).strip().lower()
    if bus.startswith('00000000:'):
        bus = bus[4:]
    try:
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read())
    except Exception:
        node = -1
    cpus = set()
    if node >= 0:
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
    return bus, node, cpus


def main():
    n = torch.cuda.device_count()
    all_cpus = set(os.sched_getaffinity(0))
    print(sh('nvidia-smi topo -m'))
    print(sh('lscpu | grep -i -E "model name|socket|numa|^cpu\\(s\\)"'))
    info = [gpu_numa(i) for i in range(n)]
    print(json.dumps({'gpus': [{'index': i, 'pci_bus_id': b, 'numa_node': nd, 'node_cpus_allowed': len(c & all_cpus)} for i, (b, nd, c) in enumerate(info)],
                      'cpus_allowed': len(all_cpus)}))
    nbytes = 1 << 30
    reps = 6
    table = {}
    for placement in ('default', 'numa_local'):
        bufs = [None] * n

        def alloc(i):
            if placement == 'numa_local' and info[i][2] & all_cpus:
                os.sched_setaffinity(0, info[i][2] & all_cpus)                 # this thread only
            t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            t.fill_(1)                                                         # first touch
            bufs[i] = t
        for i in range(n):
            th = threading.Thread(target=alloc, args=(i,))
            th.start()
            th.join()
        dst = [torch.empty(nbytes, dtype=torch.uint8, device='cuda:%d' % i) for i in range(n)]
        for k in [k for k in (1, 2, 4, 8) if k <= n]:
            for group in ([list(range(k))] + ([list(range(n - k, n))] if k < n else [])):
                res = {}
                barrier = threading.Barrier(len(group))

                def run(i):
                    if placement == 'numa_local' and info[i][2] & all_cpus:
                        os.sched_setaffinity(0, info[i][2] & all_cpus)
                    torch.cuda.set_device(i)
                    s = torch.cuda.Stream(device=i)
                    with torch.cuda.stream(s):
                        dst[i].copy_(bufs[i], non_blocking=True)
                        s.synchronize()
                        barrier.wait()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(s)
                        for _ in range(reps):
                            dst[i].copy_(bufs[i], non_blocking=True)
                        e1.record(s)
                        s.synchronize()
                        res[i] = nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
                ths = [threading.Thread(target=run, args=(i,)) for i in group]
                for th in ths:
                    th.start()
                for th in ths:
                    th.join()
                table['%s/gpus_%s' % (placement, ','.join(map(str, group)))] = {
                    'per_gpu_GBps': {str(i): round(res[i], 1) for i in group}, 'sum_GBps': round(sum(res.values()), 1)}
        del bufs, dst
    print(json.dumps(table, indent=1))


if __name__ == '__main__':
    main()
