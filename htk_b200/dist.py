"""Multi-GPU plumbing of the E-step: utterances shard across ranks, accumulators are combined
by ONE FP64 all-reduce per pass.

This replaces the reference's only form of parallelism -- N `HERest -p k` processes writing
HER<k>.acc files that `HERest -p 0` sums (HTKTools/HERest.c:366-367, :514-521;
HTKLib/HTrain.c:1626-1687).  `torch.distributed` is plumbing only: the buffer that is reduced
is the library's resident accumulator block (ForwardBackward.acc_tensor()).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .flat import Batch


def shard_indices(T: Sequence[int], Q: Sequence[int], world: int) -> List[np.ndarray]:
    """Greedy longest-processing-time split of utterances over `world` ranks, balancing the
    recursion work sum(T*Q), deterministic for a given input."""
    T = np.asarray(T, dtype=np.int64); Q = np.asarray(Q, dtype=np.int64)
    cost = T * np.maximum(Q, 1)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        parts[r].append(int(i)); load[r] += cost[i]
    return [np.sort(np.asarray(p, dtype=np.int64)) for p in parts]


def shard_batch(batch: Batch, rank: int, world: int) -> Batch:
    """The sub-batch of `batch` this rank owns (features are copied; use once per pass)."""
    T = np.diff(batch.frameOff); Q = np.diff(batch.labOff)
    idx = shard_indices(T, Q, world)[rank]
    feats = [batch.feat[batch.frameOff[i]:batch.frameOff[i + 1]] for i in idx]
    labs = [batch.lab[batch.labOff[i]:batch.labOff[i + 1]] for i in idx]
    sub = Batch(feats, labs, batch.feat.shape[1])
    sub.index = idx
    return sub


def allreduce_accs(acc_tensor) -> None:
    """In-place SUM over ranks of the flat FP64 accumulator buffer (device or host tensor)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc_tensor, op=dist.ReduceOp.SUM)
