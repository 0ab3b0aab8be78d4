"""Image-sharded data parallelism (one process per GPU, torch.distributed for the plumbing).

Every image is independent through matching, encoding, the per-anchor losses, decoding and NMS (the reference's
tf.map_fn over images at ssd.py:193 and nms.py:96), so a batch is split into contiguous image ranges with no
data-path collective.  The only coupling is the loss normaliser (ssd.py:121-123): one all-reduce(sum) of the
float64 triple (sum loc, sum cls, num_matches) -- 24 bytes, pure latency."""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def shard_range(batch_size, rank, world_size):
    """Contiguous image range [lo, hi) of `rank`; the first (batch_size % world_size) ranks get one extra image."""
    base, rem = divmod(int(batch_size), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_groundtruth(groundtruth, rank, world_size):
    lo, hi = shard_range(groundtruth['num_boxes'].shape[0], rank, world_size)
    return {k: v[lo:hi] for k, v in groundtruth.items()}


def all_reduce_sums(sums, group=None):
    """In-place sum over ranks of the [3] float64 tensor produced by SSD.loss_sums (NCCL on GPUs, gloo on CPU)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def connect_peers(group=None):
    """Sets up the hand-written NVLink all-reduce (csrc/comm.cu) between the ranks of `group` (default: WORLD), for the
    library context of the calling thread on the current CUDA device.  torch.distributed only carries the 64-byte CUDA IPC
    handles of the per-rank mailboxes.  Returns True when EVERY rank mapped every peer (then `SSD.peer_all_reduce = True`
    replaces the NCCL all-reduce of the loss sums by a fused peer-memory kernel); False -- on all ranks alike -- when the
    process group is absent / single-rank or some peer cannot be mapped (different boxes, no P2P): NCCL stays in use."""
    if not (dist.is_available() and dist.is_initialized()):
        return False
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world < 2 or world > 16 or not torch.cuda.is_available():
        return False
    lib = _lib.load()
    dev = torch.cuda.current_device()
    ctx = _lib.context(dev)
    if lib.ssdk_comm_world(ctx) == world:
        return True
    buf = (ctypes.c_ubyte * 64)()
    status = lib.ssdk_comm_local_handle(ctx, ctypes.cast(buf, ctypes.c_void_p))
    mine = torch.tensor(list(buf), dtype=torch.uint8, device=torch.device('cuda', dev))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    handles = torch.stack(gathered).cpu().numpy().tobytes()
    if status == 0:
        status = lib.ssdk_comm_connect(ctx, rank, world, handles)
    ok = torch.tensor([1 if status == 0 else 0], dtype=torch.int32, device=torch.device('cuda', dev))
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        lib.ssdk_comm_disconnect(ctx)
        return False
    dist.barrier(group=group)              # every mailbox is mapped everywhere before the first exchange
    return True


def peer_all_reduce_sum(values):
    """In-place sum over the connected ranks of a float64 CUDA tensor with at most 8 elements (rank order: every rank gets
    bit-identical results).  Asynchronous on the current stream; a collective."""
    assert values.is_cuda and values.dtype == torch.float64 and values.is_contiguous() and 1 <= values.numel() <= 8
    dev = values.device
    ctx = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
    _lib.check(_lib.load().ssdk_ctx_set_stream(ctx, torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(_lib.load().ssdk_comm_all_reduce_sum(ctx, values.data_ptr(), values.numel()))
    return values


def peer_error(device_index=None):
    """0, or the number of the exchange in which a peer failed to arrive within ~15 s (its outputs are NaN).  Synchronises."""
    dev = torch.cuda.current_device() if device_index is None else device_index
    e = ctypes.c_int64(0)
    _lib.check(_lib.load().ssdk_comm_error(_lib.context(dev), ctypes.byref(e)))
    return int(e.value)


def finalize_losses(sums):
    """(localization_loss, classification_loss) from the (all-reduced) sums: ssd.py:123,131-133.  Host/torch
    version of ssdk_loss_finalize for places where the sums live on the CPU (gloo tests)."""
    norm = torch.clamp(sums[2], min=1.0)
    return (sums[0] / norm).to(torch.float32), (sums[1] / norm).to(torch.float32)
