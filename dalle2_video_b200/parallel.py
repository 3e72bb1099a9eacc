"""Multi-GPU sampling = batch sharding with no data-path collective.

Every op of the denoiser and of the sampler is per-sample (GroupNorm statistics are per sample,
attention is per sample, conditional dropout is per sample), so the rows of `video_embed` /
`text_encodings` are split contiguously across ranks; each rank holds a full weight replica and its
own generator seeded `seed + rank`.  Results stay sharded unless the caller gathers them.
"""
from __future__ import annotations


def shard_bounds(n_rows: int, rank: int, world_size: int):
    """Contiguous [lo, hi) slice of `n_rows` owned by `rank`; the first `n_rows % world_size` ranks get one more."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_rows, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(tensor, rank: int, world_size: int):
    """Rows of `tensor` (dim 0) owned by `rank`; None passes through."""
    if tensor is None:
        return None
    lo, hi = shard_bounds(tensor.shape[0], rank, world_size)
    return tensor[lo:hi]


def allreduce_mean_(flat):
    """Data-parallel gradient averaging of the training config: ONE all-reduce of the flat gradient
    buffer, then divide by the world size (accelerate-DDP semantics, /root/reference/dalle2_video/trainer.py:117-124,360).
    NCCL for CUDA tensors, gloo in the CPU tests; a no-op without an initialised process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return flat
    dist.all_reduce(flat)
    flat.div_(dist.get_world_size())
    return flat
