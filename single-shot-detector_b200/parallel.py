"""Image-sharded data parallelism (one process per GPU, torch.distributed for the plumbing).

Every image is independent through matching, encoding, the per-anchor losses, decoding and NMS (the reference's
tf.map_fn over images at ssd.py:193 and nms.py:96), so a batch is split into contiguous image ranges with no
data-path collective.  The only coupling is the loss normaliser (ssd.py:121-123): one all-reduce(sum) of the
float64 triple (sum loc, sum cls, num_matches) -- 24 bytes, pure latency."""
import torch
import torch.distributed as dist


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


def finalize_losses(sums):
    """(localization_loss, classification_loss) from the (all-reduced) sums: ssd.py:123,131-133.  Host/torch
    version of ssdk_loss_finalize for places where the sums live on the CPU (gloo tests)."""
    norm = torch.clamp(sums[2], min=1.0)
    return (sums[0] / norm).to(torch.float32), (sums[1] / norm).to(torch.float32)
