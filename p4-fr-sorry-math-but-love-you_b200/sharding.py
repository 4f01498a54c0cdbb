"""Host-side data parallelism.

Inference: the image batch is split over ranks -- one process per GPU, weights replicated, KV cache private -- with
NO collective on the data path; the only communication is the optional gathering of the (small) int64 token matrix
and the max-over-ranks of timings.

Training: every rank runs the library's forward/backward on its shard of the batch and the flat fp32 gradient buffer
(108.9 MB for EfficientSATRN) is summed over ranks with NCCL all-reduce, bucket by bucket, while the backward pass is
still being enqueued (GradBucketReducer); the optimiser kernel applies 1 / world.  Both go through torch.distributed
(NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world_size):
    """Contiguous, balanced split: the first (n % world) ranks get one extra item."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(tensor, rank=None, world_size=None):
    """This rank's slice of a batch-major tensor."""
    if rank is None:
        rank, world_size = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    lo, hi = shard_bounds(tensor.size(0), rank, world_size)
    return tensor[lo:hi]


def gather_tokens(local_tokens, n_total):
    """All ranks' token rows in the original batch order ([n_total, steps] int64).
    Shards may be ragged, so rows are padded to the largest shard for all_gather."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_tokens
    world = dist.get_world_size()
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(max_rows, local_tokens.size(1), dtype=local_tokens.dtype, device=local_tokens.device)
    pad[: local_tokens.size(0)] = local_tokens
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def max_over_ranks(value, device="cpu"):
    """Timing convention of bench.py: the job takes as long as its slowest rank."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


class GradBucketReducer:
    """Sum-all-reduce of one flat gradient buffer in the buckets the backward pass reports.

    ``on_bucket(offset, count)`` is what the library calls (frx_train_set_bucket_callback) as soon as it has ENQUEUED
    every kernel that writes ``flat[offset:offset + count]``: the bucket's all-reduce is launched asynchronously
    (NCCL orders it after the work already on the calling stream) and runs under the rest of the backward pass.
    ``finish()`` makes the calling stream wait for all of them (no host synchronisation on CUDA) and reduces whatever
    the callbacks did not cover, so the result always equals one all-reduce of the whole buffer."""

    def __init__(self, flat, process_group=None):
        self.flat, self.group = flat, process_group
        self.works, self.covered = [], []

    def begin(self):
        del self.works[:]
        del self.covered[:]

    def on_bucket(self, offset, count):
        offset, count = int(offset), int(count)
        if count <= 0:
            return
        self.covered.append((offset, offset + count))
        self.works.append(dist.all_reduce(self.flat[offset:offset + count], group=self.group, async_op=True))

    def uncovered(self):
        """Ranges of the buffer no bucket has reported (sorted, disjoint)."""
        gaps, pos = [], 0
        for lo, hi in sorted(self.covered):
            if lo > pos:
                gaps.append((pos, lo))
            pos = max(pos, hi)
        if pos < self.flat.numel():
            gaps.append((pos, self.flat.numel()))
        return gaps

    def finish(self):
        for lo, hi in self.uncovered():
            self.works.append(dist.all_reduce(self.flat[lo:hi], group=self.group, async_op=True))
        for w in self.works:
            w.wait()
        n = len(self.works)
        del self.works[:]
        return n
