"""The N>1 path on CPU (gloo, world_size 2): utterance sharding + one all-reduce of the flat
accumulator buffer must equal the single-process result, i.e. what `HERest -p 0` computes when it
merges per-process dumps.  The per-rank compute here is the CPU oracle (test infrastructure); on the
GPU box bench.py runs the same plumbing with the CUDA library and NCCL."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, acc_errors, load_golden


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, name, out_path):
    import sys
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import load_golden as lg
    from htk_b200.dist import allreduce_accs, shard_batch
    from htk_b200.flat import make_options
    from oracle import oracle_lib as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z, fm, b, kw = lg(name)
    sub = shard_batch(b, rank, world)
    acc, res, _ = O.accumulate(fm, make_options(**kw), sub, acc_double=True, want_beams=False)
    t = torch.from_numpy(acc)
    allreduce_accs(t)
    ok = torch.tensor([sum(1 for r in res if r[0] == 0)], dtype=torch.int64)
    dist.all_reduce(ok)
    if rank == 0:
        np.savez(out_path, acc=t.numpy(), n_ok=int(ok[0]), mine=len(sub.index))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["synth_tee_m2", "htkdemo_t2000"])
def test_two_rank_allreduce_equals_single_process(tmp_path, name):
    import torch.multiprocessing as mp
    from htk_b200.flat import make_options
    from oracle import oracle_lib as O
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), name, out), nprocs=2, join=True)
    got = np.load(out)
    z, fm, b, kw = load_golden(name)
    whole, res, _ = O.accumulate(fm, make_options(**kw), b, acc_double=True, want_beams=False)
    assert got["n_ok"] == sum(1 for r in res if r[0] == 0)
    assert 0 < got["mine"] < b.numUtt
    assert np.allclose(got["acc"], whole, rtol=1e-11, atol=1e-9)
    # and against the reference's own single-process dump (float accumulators): its noise floor
    e = acc_errors(got["acc"], z["ref_acc"], fm)
    assert max(e.values()) < 2e-5, e


def test_shard_indices_balanced_and_complete():
    from htk_b200.dist import shard_indices
    rng = np.random.default_rng(0)
    T = rng.integers(200, 1500, size=101); Q = rng.integers(20, 150, size=101)
    for world in (1, 2, 4, 8):
        parts = shard_indices(T, Q, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(101))
        loads = np.array([np.sum(T[p] * Q[p]) for p in parts], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.1


def _merge_worker(rank, world, port, out_path):
    import sys
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from htk_b200.compare import load_merge_golden
    from htk_b200.dist import allreduce_accs, shard_batch
    from htk_b200.flat import make_options
    from oracle import oracle_lib as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z, fm, b = load_merge_golden()
    sub = shard_batch(b, rank, world)
    acc, res, _ = O.accumulate(fm, make_options(), sub, acc_double=True, want_beams=False)
    t = torch.from_numpy(acc)
    allreduce_accs(t)
    if rank == 0:
        np.savez(out_path, acc=t.numpy())
    dist.destroy_process_group()


def test_allreduce_equals_stock_p0_merge(tmp_path):
    """Sharded accumulators + ONE all-reduce against what the stock `HERest -p 0` produced from FOUR per-process
    dumps (tests/golden/merge_p0_tied_m4.npz, HERest.c:514-521 / HTrain.c:1626-1687): the accumulators against the
    single-process dump, and the means / variances re-estimated from them (HERest.c:974-1012, :1045-1122) against the
    merged MMF."""
    import torch.multiprocessing as mp
    from htk_b200.compare import load_merge_golden, reestimated_errors
    out = str(tmp_path / "m.npz")
    mp.spawn(_merge_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    acc = np.load(out)["acc"]
    z, fm, b = load_merge_golden()
    e = acc_errors(acc, z["ref_acc"], fm)
    assert max(e.values()) < 2e-5, e
    L, D = fm.layout, fm.D
    occ = acc[L.muOcc:L.vaSum]
    mu = acc[L.muSum:L.muOcc].reshape(-1, D); va = acc[L.vaSum:L.vaOcc].reshape(-1, D)
    upd = np.any(z["merged_mean"] != fm.mean, axis=1)              # Gaussians the reference re-estimated (-m 3)
    assert upd.sum() > fm.G // 2
    dm = mu[upd] / occ[upd, None]
    mean = fm.mean[upd].astype(np.float64) + dm
    var = va[upd] / occ[upd, None] - dm * dm
    r = reestimated_errors(mean, var, z["merged_mean"][upd].astype(np.float64), z["merged_var"][upd].astype(np.float64))
    assert r["mean"] < 1e-4 and r["var"] < 1e-4, r
