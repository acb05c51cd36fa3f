"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The hot path shards by video:
  * sampling: the batch of videos is partitioned across ranks, each rank holds a full weight replica and its own
    RNG stream; there is NO collective on the data path (`gather_ids` only collects results for output);
  * training: data parallel — the one exchange step is the bucketed gradient reduction (the reference's Lightning
    `DDPStrategy`, train_transformer.py:41), issued bucket by bucket as backward finishes chunks of blocks: either an
    all-reduce (every rank then runs the whole optimizer), or - with the fused flat AdamW - a reduce-scatter, the
    update of this rank's 1/N shard, and an all-gather of the updated bf16 tensor-core operands (same bytes on the
    wire as the all-reduce when counted in fp32, 3/4 of them here; the optimizer's HBM traffic divided by N).
Below one video the stack does not shard (256 latents couple every token in every block): replicas only.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced partition of `n_items` videos: ranks < n_items % world get one extra."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def rank_seed(base_seed: int, rank: int) -> int:
    """Per-rank RNG stream (`seeds base+rank`, SURVEY.md §8(d) cfg #3)."""
    return int(base_seed) + int(rank)


def allreduce_mean_(buf: torch.Tensor, group=None, async_op: bool = False):
    """In-place mean over ranks of one gradient bucket.  NCCL averages in the collective; gloo (CPU tests) sums and
    divides.  Returns the work handle when async_op."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    if dist.get_backend(group) == "nccl":
        return dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
    work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=False)
    buf.div_(dist.get_world_size(group))
    return work


def shard_bounds(lo: int, hi: int, rank: int, world: int):
    """This rank's 1/world slice of the bucket [lo, hi) (equal shards: (hi - lo) % world == 0)."""
    n = hi - lo
    if n % world:
        raise ValueError(f"bucket of {n} elements does not split into {world} equal shards")
    k = n // world
    return lo + rank * k, lo + (rank + 1) * k


def reduce_scatter_mean_(bucket: torch.Tensor, group=None):
    """In-place reduce-scatter of one gradient bucket: on return this rank's shard of `bucket` (shard_bounds) holds the
    mean over ranks; the rest of the bucket is unspecified.  NCCL: one ncclReduceScatter (AVG) writing into its own
    input slice; gloo (CPU tests) has no reduce-scatter: all-reduce and keep the shard."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(0, bucket.numel(), rank, world)
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(bucket[lo:hi], bucket, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
        bucket[lo:hi].div_(world)
    return bucket[lo:hi]


def all_gather_shards_(bucket: torch.Tensor, group=None):
    """In-place all-gather: every rank contributes its shard of `bucket` and receives all the others."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(0, bucket.numel(), rank, world)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(bucket, bucket[lo:hi], group=group)
    else:
        parts = [torch.empty_like(bucket[lo:hi]) for _ in range(world)]
        dist.all_gather(parts, bucket[lo:hi].clone(), group=group)
        for r, part in enumerate(parts):
            a, b = shard_bounds(0, bucket.numel(), r, world)
            bucket[a:b].copy_(part)
    return bucket


def bucket_slices(block_slices, chunks, head_slice, emb_slice):
    """Gradient buckets of the flat buffer in the order backward completes them: block chunks from the last to the
    first (the head's ln_f + weight bucket right after the last chunk), the embeddings at the end."""
    out = []
    n = len(block_slices)
    for lb, le in reversed(chunks):
        out.append((block_slices[lb][0], block_slices[le - 1][1]))
        if le == n:
            out.append(tuple(head_slice))
    out.append(tuple(emb_slice))
    return out


def gather_ids(ids: torch.Tensor, dst: int = 0, group=None):
    """Collect the sampled token grids of every rank on `dst` (output only; not on the timed data path)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [ids]
    world = dist.get_world_size(group)
    out = [torch.empty_like(ids) for _ in range(world)] if dist.get_rank(group) == dst else None
    dist.gather(ids, out, dst=dst, group=group)
    return out
