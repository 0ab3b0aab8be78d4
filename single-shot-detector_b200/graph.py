"""CUDA-graph replay of a step of the hot path.

A training or inference step is a fixed sequence of ~14 short kernels; launched one by one from Python the host
(ctypes + tensor bookkeeping, ~0.2 ms) can be slower than the GPU (~0.45 ms per 48-image step), and on a busy host
the GPU then idles between kernels.  `capture(fn)` records everything `fn` enqueues on the current stream -- the
library's kernels, its memsets, the NCCL all-reduce of the loss sums -- into one CUDA graph; `replay()` launches
the whole step with a single call.  The tensors `fn` read are static inputs: refill them in place
(`tensor.copy_(...)`) before a replay; its return value (`outputs`) is overwritten by every replay.

The library side needs nothing special: it launches on the caller's stream, never synchronises in the steady state
and keeps its workspace across calls (run `fn` once eagerly first so that the workspace has its final size --
`capture` does that)."""
import torch

from . import _lib


class CapturedStep:
    def __init__(self, graph, outputs, launches):
        self.graph = graph
        self.outputs = outputs
        self.launches_per_replay = launches

    def replay(self):
        if self.graph is None:
            raise _lib.SsdkError('this CapturedStep has been released')
        self.graph.replay()
        return self.outputs

    def release(self):
        """Destroys the CUDA graph (and the references it holds to communicators, streams and tensors).  Teardown order for
        multi-GPU programs: synchronize -> release() every CapturedStep -> torch.distributed.destroy_process_group().
        A captured NCCL all-reduce keeps its communicator busy as long as the graph exists, and destroying the process group
        first can block for minutes; the library's own peer-memory exchange has no such tie (plain kernels)."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None
            self.outputs = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()
        return False


def concurrent(*fns, device=None, priorities=None, train_ctas_per_sm=None):
    """Returns a function that runs the independent `fns` on one side stream each, forked from and joined to the current
    stream.  Inside `capture` this becomes parallel branches of the CUDA graph, so the ALU- / latency-bound kernels of one
    sub-path (matching, sorting, NMS, packing) hide behind the HBM-bound kernels of another (the loss pass, the score scan).
    The library keeps separate workspace for the training-side and the inference-side sub-paths, so `SSD.loss` and
    `SSD.get_predictions` may overlap; two calls of the SAME sub-path must not."""
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    # `priorities`: one CUDA stream priority per function (lower number = scheduled first; -1 is the usual high priority)
    streams = [torch.cuda.Stream(device=dev, priority=0 if priorities is None else int(priorities[i])) for i in range(len(fns))]

    def run():
        cur = torch.cuda.current_stream(dev)
        outs = []
        for st in streams:
            st.wait_stream(cur)
        # train_ctas_per_sm: the fused training-step kernel is persistent and takes every CTA slot of the GPU (six per SM); a smaller
        # value leaves room for the other branch's short kernels.  Measured for the bench step with the round-2 kernels
        # (profiles/r3a_overlap.json): six 0.340 ms, five 0.344 ms, four 0.371 ms -- the default leaves it alone.
        if train_ctas_per_sm:
            _lib.set_option(_lib.SSDK_OPT_TRAIN_CTAS_PER_SM, int(train_ctas_per_sm), dev.index)
        try:
            for st, fn in zip(streams, fns):
                with torch.cuda.stream(st):
                    outs.append(fn())
        finally:
            if train_ctas_per_sm:
                _lib.set_option(_lib.SSDK_OPT_TRAIN_CTAS_PER_SM, 0, dev.index)
        for st in streams:
            cur.wait_stream(st)
        return tuple(outs)
    return run


def capture(fn, warmup=2, device=None):
    """Run `fn` `warmup` times eagerly on a side stream (grows the workspace, initialises NCCL), then capture one call."""
    if not torch.cuda.is_available():
        raise _lib.SsdkError('no CUDA device: CUDA graphs need the GPU (no CPU fallback)')
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(max(1, warmup)):
            fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    before = _lib.launch_count(dev.index)
    with torch.cuda.graph(graph):
        outputs = fn()
    launches = _lib.launch_count(dev.index) - before
    return CapturedStep(graph, outputs, launches)
