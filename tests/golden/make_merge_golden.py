#!/usr/bin/env python
"""Golden vector for the multi-process merge: what `HERest -p 0` computes from per-process dumps.

The reference's only form of parallelism is N `HERest -p k` processes over disjoint utterance lists followed by
`HERest -p 0 ... HER1.acc .. HERN.acc`, which sums the (float) dumps and re-estimates the models
(HTKTools/HERest.c:514-521, HTKLib/HTrain.c:1626-1687).  This script RUNS THAT with the unmodified reference
binaries of oracle/_ref/bin (container only) on a small tied-state set and stores, next to the inputs,

  * ref_acc        the accumulators of ONE process over all utterances (decoded HER1.acc),
  * merged_*       means / variances / mixture weights / transition probabilities of the MMF that `-p 0` wrote after
                   merging FOUR per-process dumps, in the flat model's own order.

The library's per-GPU accumulators + ONE all-reduce (NCCL on the box, gloo in the CPU test) + its device M-step must
reproduce both: tests/test_dist_gloo.py (oracle per rank), tools/mgpu_parity.py and bench.py's `allreduce_parity`.

    python tests/golden/make_merge_golden.py
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from htk_b200 import htkio, synth                      # noqa: E402
from htk_b200.flat import flatten                      # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
OUT = os.path.join(ROOT, "tests", "golden")
WORK = os.environ.get("GOLDEN_WORK", "/tmp/hfb_golden_work")
MODEL_KEYS = ("mean", "ivar", "gConst", "meanId", "varId", "stateMixOff", "mixGauss", "mixLogWt",
              "hmmNumStates", "hmmStateOff", "hmmState", "hmmTrans", "transN", "transOff", "transLogA")
NPROC = 4


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("command failed: " + " ".join(cmd))
    return p.stdout


def main():
    name = "merge_p0_tied_m4"
    w = os.path.join(WORK, name)
    shutil.rmtree(w, ignore_errors=True)
    os.makedirs(os.path.join(w, "feat"))
    hs = synth.make_tied_triphone_set(n_states=48, M=4, n_phys=32, n_logical=48, n_centre=6, seed=23, spread=0.2)
    htkio.write_mmf(os.path.join(w, "mmf"), hs)
    htkio.write_hmm_list(os.path.join(w, "hmmlist"), hs)
    names = open(os.path.join(w, "hmmlist")).read().splitlines()
    hs2 = htkio.read_mmf([os.path.join(w, "mmf")], hmm_list=names)
    fm = flatten(hs2)
    feats, labs = synth.sample_corpus(fm, 12, 240, 24, seed=17, T_jitter=20)
    mlf, scp = {}, []
    for i, (f, l) in enumerate(zip(feats, labs)):
        fn = os.path.join(w, "feat", "u%04d.mfc" % i)
        htkio.write_htk_features(fn, f, hs.parm_kind)
        mlf["u%04d" % i] = [fm.names[j] for j in l]
        scp.append(fn)
    htkio.write_mlf(os.path.join(w, "labs.mlf"), mlf)
    herest = os.path.join(BIN, "HERest")
    base = ["-u", "tmvw", "-H", "mmf", "-I", "labs.mlf"]
    # one process over everything: the accumulators the summed per-rank buffers must equal
    os.makedirs(os.path.join(w, "all"))
    open(os.path.join(w, "all.scp"), "w").write("\n".join(scp) + "\n")
    run([herest] + base + ["-p", "1", "-S", "all.scp", "-M", "all", "hmmlist"], w)
    acc, tp, tt = htkio.read_acc_dump(os.path.join(w, "all", "HER1.acc"), hs2, fm, 15)
    # NPROC processes over disjoint lists, then the -p 0 merge + update
    dumps = []
    for k in range(NPROC):
        d = "p%d" % (k + 1)
        os.makedirs(os.path.join(w, d))
        open(os.path.join(w, d + ".scp"), "w").write("\n".join(scp[k::NPROC]) + "\n")
        run([herest] + base + ["-p", str(k + 1), "-S", d + ".scp", "-M", d, "hmmlist"], w)
        dumps.append(os.path.join(d, "HER%d.acc" % (k + 1)))
    os.makedirs(os.path.join(w, "out"))
    run([herest, "-u", "tmvw", "-p", "0", "-H", "mmf", "-M", "out", "hmmlist"] + dumps, w)
    hsO = htkio.read_mmf([os.path.join(w, "out", "mmf")], hmm_list=names)
    fmO = flatten(hsO)
    assert np.array_equal(fmO.hmmState, fm.hmmState) and np.array_equal(fmO.mixGauss, fm.mixGauss)
    T = np.array([f.shape[0] for f in feats], dtype=np.int64)
    d = dict(D=fm.D, names=np.array(fm.names), prune=np.array([1e20, 0, 1e20]), minFrwdP=np.float32(10.0),
             feat=np.concatenate(feats, 0).astype(np.float32), frameOff=np.concatenate([[0], np.cumsum(T)]),
             lab=np.concatenate(labs).astype(np.int32),
             labOff=np.concatenate([[0], np.cumsum([len(l) for l in labs])]).astype(np.int32),
             ref_acc=acc, ref_totalPr=tp, ref_totalT=tt, nproc=NPROC,
             merged_mean=fmO.mean, merged_var=(1.0 / fmO.ivar.astype(np.float64)).astype(np.float32),
             merged_mixLogWt=fmO.mixLogWt, merged_transLogA=fmO.transLogA)
    for k in MODEL_KEYS:
        d["m_" + k] = getattr(fm, k)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote %s.npz: %d utts, %d frames" % (name, len(feats), int(T.sum())))


if __name__ == "__main__":
    main()
