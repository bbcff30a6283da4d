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


def merge_parity(device: int, rank: int, world: int, name: str = "merge_p0_tied_m4") -> dict:
    """The multi-GPU exchange against the reference's multi-process merge, on whatever process group is up
    (world = 1: no collective, same comparison).  Every rank runs the CUDA E-step on its share of the golden
    utterances (tests/golden/merge_p0_*.npz), the resident FP64 buffers are summed with ONE all-reduce, and
    rank 0 compares

      * the summed accumulators with the stock single-process dump (HTrain.c:1454-1505), and
      * the device M-step on them with the MMF that the stock `HERest -p 0` wrote after merging four per-process
        dumps (HERest.c:514-521, HTrain.c:1626-1687) -- means / variances within 1e-4 (north_star).

    Returns {"max_err", "mean_err", "var_err", "ok", ...} on rank 0, {} elsewhere."""
    import torch
    from .compare import acc_errors, load_merge_golden, reestimated_errors
    from .estep import ForwardBackward
    z, fm, b = load_merge_golden(name)
    fb = ForwardBackward(fm, device=device)
    sub = shard_batch(b, rank, world)
    res, _ = fb.FBFile(sub)
    bad = sum(1 for r in res if r.status != 0)
    t = fb.acc_tensor()
    torch.cuda.synchronize()
    allreduce_accs(t)
    torch.cuda.synchronize()
    out = {}
    if rank == 0:
        acc = fb.GetAccs()
        e = acc_errors(acc, z["ref_acc"], fm)
        new, info = fb.MStep()
        r = reestimated_errors(new.mean, info["var"], z["merged_mean"].astype(np.float64), z["merged_var"].astype(np.float64))
        w_err = float(np.max(np.abs(np.exp(new.mixLogWt.astype(np.float64)) - np.exp(z["merged_mixLogWt"].astype(np.float64)))))
        L = fm.layout
        out = {"fixture": name, "world": world, "max_err": max(e.values()), "mean_err": r["mean"], "var_err": r["var"],
               "weight_err": w_err, "frames": int(acc[L.totalT]), "frames_ref": int(z["ref_acc"][L.totalT]),
               "reference": "stock HERest -p 1 dump (accumulators); stock HERest -p 1..%d + -p 0 merged MMF (parameters)" % int(z["nproc"])}
        # MMF text carries 7 significant digits: one unit of that on top of the 1e-4 bar
        out["ok"] = bool(bad == 0 and out["frames"] == out["frames_ref"] and out["max_err"] < 1e-4 and
                         r["mean"] < 1e-4 + 2e-6 and r["var"] < 1e-4 + 2e-6 and w_err < 1e-4)
    fb.close()
    return out
