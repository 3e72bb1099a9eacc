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


class GradBuckets:
    """The gradient all-reduce of the training config split into contiguous buckets of the flat gradient buffer, each
    launched (async) as soon as the backward segment that finishes its last parameter has run, so the transfers ride
    under the rest of the backward ("buckets >= 25 MB, overlapped with backward"; the reference
    gets this from accelerate-DDP's bucketed hooks, /root/reference/dalle2_video/trainer.py:117-124,360).

    `items` = (offset, numel, segment) per parameter in flat-buffer order: `segment` is the index of the backward
    segment after which that slice of `flat` is final (use `n_segments` for parameters whose gradient arrives later,
    e.g. through torch autograd after the native backward).  Consecutive parameters with the same segment form a run;
    runs smaller than `min_bytes` join the neighbouring run that becomes ready next (never delaying a large run when
    it can be avoided; a small run that is only ready with the last segment or later stays a bucket of its own), so there are few,
    large collectives.  launch(s) starts the buckets of segment s on the process group's own stream; finish() launches
    whatever is left, waits, and divides by the world size (sum-then-divide, as allreduce_mean_ does)."""

    def __init__(self, flat, items, n_segments, min_bytes=1 << 20, group=None):
        self.flat, self.group, self.n_segments = flat, group, n_segments
        esz = flat.element_size()
        runs = []  # [lo, hi, seg]
        pos = 0
        for off, n, seg in sorted(items):
            if off < pos:
                raise ValueError("overlapping gradient slices")
            if off > pos:  # a gap nobody claimed: reduce it at the end
                runs.append([pos, off, n_segments])
            seg = min(max(int(seg), 0), n_segments)
            if runs and runs[-1][2] == seg and runs[-1][1] == off:
                runs[-1][1] = off + n
            else:
                runs.append([off, off + n, seg])
            pos = off + n
        if pos < flat.numel():
            runs.append([pos, flat.numel(), n_segments])
        kept = set()  # small runs that stay alone: ready only at the very end and no later neighbour to join
        while len(runs) > 1:
            small = [i for i, r in enumerate(runs) if (r[1] - r[0]) * esz < min_bytes and (r[0], r[2]) not in kept]
            if not small:
                break
            i = min(small, key=lambda j: runs[j][1] - runs[j][0])
            nb = [j for j in (i - 1, i + 1) if 0 <= j < len(runs)]
            later = [j for j in nb if runs[j][2] >= runs[i][2]]
            if later:      # join the neighbour that is ready soonest after this run: delays only the small run
                j = min(later, key=lambda k: runs[k][2])
            elif runs[i][2] >= n_segments - 1:   # ready only with the last segment (the conditioning prologue's parameters,
                                                   # the first layers) or after it: never delays an earlier neighbour
                kept.add((runs[i][0], runs[i][2]))
                continue
            else:          # a local maximum in the middle of the backward: delay the closer neighbour
                j = min(nb, key=lambda k: abs(runs[k][2] - runs[i][2]))
            a, b = min(i, j), max(i, j)
            runs[a] = [runs[a][0], runs[b][1], max(runs[a][2], runs[b][2])]
            del runs[b]
            k = a  # merging may have made two equal-segment neighbours adjacent
            while k + 1 < len(runs) and runs[k + 1][2] == runs[k][2]:
                runs[k][1] = runs[k + 1][1]
                del runs[k + 1]
            while k > 0 and runs[k - 1][2] == runs[k][2]:
                runs[k - 1][1] = runs[k][1]
                del runs[k]
                k -= 1
        self.buckets = [tuple(r) for r in runs]
        self._by_seg = {}
        for lo, hi, seg in self.buckets:
            self._by_seg.setdefault(seg, []).append((lo, hi))
        self._pending, self._done = [], set()

    def launch(self, segment):
        """Start the all-reduce of every bucket that became final with backward segment `segment`."""
        import torch.distributed as dist
        for lo, hi in self._by_seg.get(segment, ()):
            if (lo, hi) in self._done:
                continue
            self._done.add((lo, hi))
            self._pending.append(dist.all_reduce(self.flat[lo:hi], group=self.group, async_op=True))

    def finish(self):
        """Launch the buckets no segment claimed, wait for all of them, and turn the sums into means."""
        import torch.distributed as dist
        for seg in sorted(self._by_seg):
            self.launch(seg)
        for work in self._pending:
            work.wait()
        self._pending, self._done = [], set()
        self.flat.div_(dist.get_world_size(self.group))
        return self.flat
