#!/usr/bin/env python
"""torchrun worker: shard a golden batch across ranks, run the CUDA E-step per rank, combine with
ONE NCCL all-reduce of the resident FP64 accumulator buffer, and compare (rank 0) with the
single-process reference dump -- the multi-GPU result must equal what `HERest -p 0` merges."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from conftest import acc_errors, load_golden
from htk_b200.dist import allreduce_accs, shard_batch
from htk_b200.estep import ForwardBackward


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    worst = 0.0
    for name in sys.argv[1:] or ["synth_tied_m4", "synth_tee_m2"]:
        z, fm, b, kw = load_golden(name)
        fb = ForwardBackward(fm, device=lr, **kw)
        sub = shard_batch(b, rank, world)
        res, _ = fb.FBFile(sub)
        t = fb.acc_tensor()
        torch.cuda.synchronize()
        allreduce_accs(t)
        torch.cuda.synchronize()
        acc = fb.GetAccs()
        if rank == 0:
            e = acc_errors(acc, z["ref_acc"], fm)
            worst = max(worst, max(e.values()))
            L = fm.layout
            assert acc[L.totalT] == z["ref_acc"][L.totalT], (acc[L.totalT], z["ref_acc"][L.totalT])
            print("%s: world=%d max normalised accumulator error %.2e" % (name, world, max(e.values())))
        fb.close()
    from htk_b200.dist import merge_parity
    mp = merge_parity(lr, rank, world)                 # vs the stock -p 1..4 + -p 0 merge
    if rank == 0:
        print("merge parity:", mp)
        assert mp["ok"], mp
        assert worst < 1e-4, worst
        print("MGPU_PARITY_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
