#!/usr/bin/env python
"""Golden vectors for two-model re-estimation (HERest `ALIGNMODELMMF`, HFB.c:296-333 UseAlignHMMSet, :1518-1547).

The stock reference binary (oracle/_ref/bin/HERest, container only) aligns every utterance with one model set and
collects the statistics of ANOTHER set with the same logical names: occupancies alpha_j beta_j / pr come from the
alignment set, component posteriors, centred sums and accumulator layout from the update set.  Stored per case: both
flat models as the reference loaded them, the features, the label sequences resolved in EACH set (the logical -> physical
maps differ), and the reference's outputs (decoded HER1.acc of the update set, per-utterance log-likelihoods).

    python tests/golden/make_two_model_golden.py
"""
from __future__ import annotations

import os
import re
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


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("command failed: " + " ".join(cmd))
    return p.stdout


def load(w, tag):
    names = open(os.path.join(w, tag + ".list")).read().splitlines()
    hs = htkio.read_mmf([os.path.join(w, tag + ".mmf")], hmm_list=names)
    return hs, flatten(hs)


def case(name, hs_al, hs_up, n_utts, T, Q, prune, seed, uflags="mvw", T_jitter=0, comp_level=False):
    w = os.path.join(WORK, name)
    shutil.rmtree(w, ignore_errors=True)
    os.makedirs(os.path.join(w, "feat")); os.makedirs(os.path.join(w, "accs"))
    for tag, hs in (("al", hs_al), ("up", hs_up)):
        htkio.write_mmf(os.path.join(w, tag + ".mmf"), hs)
        htkio.write_hmm_list(os.path.join(w, tag + ".list"), hs)
    hs_al, fa = load(w, "al")
    hs_up, fu = load(w, "up")
    # the utterances are drawn from the ALIGNMENT set over labels both sets know; transcriptions use logical names too
    common = [n for n in fa.hmm_index if n in fu.hmm_index]
    rng = np.random.default_rng(seed)
    feats, labs_al, labs_up, mlf, scp = [], [], [], {}, []
    for i in range(n_utts):
        names = [common[int(k)] for k in rng.integers(0, len(common), size=Q)]
        la = np.array([fa.hmm_index[n] for n in names], dtype=np.int32)
        lu = np.array([fu.hmm_index[n] for n in names], dtype=np.int32)
        Ti = T + (int(rng.integers(-T_jitter, T_jitter + 1)) if T_jitter else 0)
        f = synth.sample_utterance(fa, la, Ti, rng)
        fn = os.path.join(w, "feat", "u%04d.mfc" % i)
        htkio.write_htk_features(fn, f, hs_up.parm_kind)
        feats.append(f); labs_al.append(la); labs_up.append(lu); mlf["u%04d" % i] = names; scp.append(fn)
    htkio.write_mlf(os.path.join(w, "labs.mlf"), mlf)
    open(os.path.join(w, "train.scp"), "w").write("\n".join(scp) + "\n")
    open(os.path.join(w, "two.cfg"), "w").write("ALIGNMODELMMF = al.mmf\nALIGNHMMLIST = al.list\n" +
                                                ("HFB: ALIGNCOMPLEVEL = T\n" if comp_level else ""))
    targs = [] if prune is None else (["-t"] + ["%.1f" % x for x in prune])
    out = run([os.path.join(BIN, "HERest"), "-C", "two.cfg", "-T", "1", "-u", uflags] + targs +
              ["-p", "1", "-H", "up.mmf", "-I", "labs.mlf", "-S", "train.scp", "-M", "accs", "up.list"], w)
    assert "2-model re-estimation enabled" in out, out[:2000]
    ppf = [float(v) for v in re.findall(r"Utterance prob per frame = (\S+)", out)]
    uf = sum({"t": 4, "m": 1, "v": 2, "w": 8}[c] for c in uflags)
    acc, tp, tt = htkio.read_acc_dump(os.path.join(w, "accs/HER1.acc"), hs_up, fu, uf)
    Ts = np.array([f.shape[0] for f in feats], dtype=np.int64)
    d = dict(D=fu.D, names_al=np.array(fa.names), names_up=np.array(fu.names), uflags=uf, comp_level=bool(comp_level),
             prune=np.array(prune if prune else [1e20, 0, 1e20], dtype=np.float64), minFrwdP=np.float32(10.0),
             feat=np.concatenate(feats, 0).astype(np.float32), frameOff=np.concatenate([[0], np.cumsum(Ts)]),
             lab=np.concatenate(labs_up).astype(np.int32), labAlign=np.concatenate(labs_al).astype(np.int32),
             labOff=np.concatenate([[0], np.cumsum([len(l) for l in labs_up])]).astype(np.int32),
             ref_acc=acc, ref_totalPr=tp, ref_totalT=tt, ref_pr_per_frame=np.array(ppf))
    for k in MODEL_KEYS:
        d["m_" + k] = getattr(fu, k)
        d["a_" + k] = getattr(fa, k)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote %s.npz: %d utts, %d frames, totalPr/T=%.6f, %d of %d utterances processed" %
          (name, n_utts, int(Ts.sum()), tp / max(tt, 1), len(ppf), n_utts))


def main():
    os.makedirs(WORK, exist_ok=True)
    # tied-state sets with the same physical / logical names, different tying, component counts (1, 3, 4 with dead
    # components) and parameters
    al = synth.make_tied_triphone_set(n_states=60, M=2, n_phys=40, n_logical=60, n_centre=8, seed=13, spread=0.2)
    up = synth.make_tied_triphone_set(n_states=50, M=[4, 1, 3], n_phys=40, n_logical=60, n_centre=8, seed=31, spread=0.2)
    case("two_model_tied", al, up, n_utts=5, T=280, Q=28, prune=None, seed=41, T_jitter=20)
    # single-Gaussian alignment set (fbInfo->maxM = 1, HFB.c:323), two-component update set, pruning on
    al = synth.make_monophone_set(n_phones=10, M=1, seed=17, spread=0.2)
    up = synth.make_monophone_set(n_phones=10, M=2, seed=18, spread=0.2)
    case("two_model_mono", al, up, n_utts=4, T=240, Q=24, prune=(250.0, 150.0, 1000.0), seed=42, uflags="tmvw")
    # HFB: ALIGNCOMPLEVEL = T (HFB.c:1521-1530): component posteriors from the ALIGNMENT set's states (same component counts)
    al = synth.make_tied_triphone_set(n_states=60, M=3, n_phys=40, n_logical=60, n_centre=8, seed=13, spread=0.2)
    up = synth.make_tied_triphone_set(n_states=50, M=3, n_phys=40, n_logical=60, n_centre=8, seed=31, spread=0.2)
    case("two_model_complevel", al, up, n_utts=4, T=260, Q=26, prune=None, seed=43, T_jitter=20, comp_level=True)


if __name__ == "__main__":
    main()
