"""Multi-GPU plumbing of the path: test samples are independent (fresh LoRA + AdamW state per sample, ttl.py:338-344),
so rank r simply takes samples r, r+W, ... of the seeded evaluation order; the only collective is one all-reduce(sum)
of {top1_correct, top5_correct, n} at the end of a dataset (AverageMeter semantics, utils/tools.py:40-44)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_samples: int, rank: int, world: int) -> range:
    return range(rank, n_samples, world)


def reduce_counts(counts: torch.Tensor, world: int) -> List[int]:
    """counts int64[3] on this rank's device -> global sums as Python ints (NCCL on GPUs, gloo on CPU)."""
    c = counts.clone()
    if world > 1 and dist.is_initialized():
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return [int(x) for x in c.tolist()]


def gather_predictions(pred: torch.Tensor, world: int) -> torch.Tensor:
    """Optional: all-gather per-sample int32 predictions (rank-major) for agreement checks."""
    if world == 1 or not dist.is_initialized():
        return pred
    out = [torch.empty_like(pred) for _ in range(world)]
    dist.all_gather(out, pred)
    return torch.stack(out, 1).reshape(-1)


def merge_sharded(per_rank: Sequence[Sequence[int]]) -> List[int]:
    """Inverse of shard_indices for equally sized shards: interleave rank-major lists back into sample order."""
    world = len(per_rank)
    n = sum(len(p) for p in per_rank)
    out = [0] * n
    for r, p in enumerate(per_rank):
        for j, v in enumerate(p):
            out[r + j * world] = v
    return out
