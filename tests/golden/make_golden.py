#!/usr/bin/env python
"""Generate the committed golden fixtures by RUNNING THE REFERENCE ITSELF.

Needs /root/reference (HTKDemo data, perl MakeProtoHMMSet) and the unmodified reference
binaries built by ``make -C oracle ref`` into oracle/_ref/bin.  Neither exists on the GPU
box, so the outputs (``tests/golden/*.npz``) are committed and this script is only re-run
here, in the build container:

    python tests/golden/make_golden.py

Each fixture holds everything a parity test needs and nothing of the reference's source:
the flat model exactly as the reference loaded it from the text MMF, the expanded
features, the label sequences, and the reference's OUTPUTS -- decoded ``HER1.acc``
accumulators, per-utterance log-likelihoods, the pruning threshold finally used, and the
per-frame beta / alpha beams printed by ``HFB: TRACE = 5``.
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

REF = os.environ.get("HTK_REFERENCE", "/root/reference")
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
OUT = os.path.join(ROOT, "tests", "golden")
WORK = os.environ.get("GOLDEN_WORK", "/tmp/hfb_golden_work")

MODEL_KEYS = ("mean", "ivar", "gConst", "meanId", "varId", "stateMixOff", "mixGauss", "mixLogWt",
              "hmmNumStates", "hmmStateOff", "hmmState", "hmmTrans", "transN", "transOff", "transLogA")


def run(cmd, cwd=None, ok=(0,)):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode not in ok:
        sys.stderr.write(p.stdout)
        raise RuntimeError("command failed (%d): %s" % (p.returncode, " ".join(cmd)))
    return p.stdout


def parse_trace(text: str, utt_T):
    """Per-utterance pr/T, final threshold, and beams from HERest -T 1 + HFB:TRACE=5."""
    utts = []
    cur = None
    for line in text.splitlines():
        if line.startswith(" Processing Data:"):
            cur = {"beta": {}, "alpha": {}, "thresh": None, "pr_per_frame": None, "retries": 0}
            utts.append(cur)
        elif cur is None:
            continue
        elif line.startswith("Retrying Beta pass at"):
            cur["thresh"] = float(line.split()[-1]); cur["retries"] += 1
            cur["beta"] = {}
        elif "Beta Beam" in line:
            m = re.match(r"\s*(\d+): Beta Beam (\d+)->(\d+)", line)
            cur["beta"][int(m.group(1))] = (int(m.group(2)), int(m.group(3)))
        elif "Alpha Beam" in line:
            m = re.match(r"\s*(\d+): Alpha Beam (\d+)->(\d+)", line)
            cur["alpha"][int(m.group(1))] = (int(m.group(2)), int(m.group(3)))
        elif "Utterance prob per frame" in line:
            cur["pr_per_frame"] = float(line.split("=")[1])
    return utts


def pack_case(name, hs, fm, feats, labs, acc, total_pr, total_t, tr, prune, min_frwd=10.0, extra=None):
    T = np.array([f.shape[0] for f in feats], dtype=np.int64)
    totalT = int(T.sum())
    qLo = np.zeros(totalT, np.int16); qHi = np.zeros(totalT, np.int16)
    sq = np.zeros(totalT, np.int16); eq = np.zeros(totalT, np.int16)
    off = 0
    have_beams = all(len(u["beta"]) > 0 for u in tr)
    for u, Tu in zip(tr, T):
        if have_beams:
            for t in range(1, Tu + 1):
                if t in u["beta"]:
                    qLo[off + t - 1], qHi[off + t - 1] = u["beta"][t]
                if t in u["alpha"]:
                    sq[off + t - 1], eq[off + t - 1] = u["alpha"][t]
        off += Tu
    d = dict(
        D=fm.D, names=np.array(fm.names), prune=np.array(prune if prune else [1e20, 0, 1e20], dtype=np.float64),
        minFrwdP=np.float32(min_frwd),
        feat=np.concatenate(feats, 0).astype(np.float32), frameOff=np.concatenate([[0], np.cumsum(T)]),
        lab=np.concatenate(labs).astype(np.int32),
        labOff=np.concatenate([[0], np.cumsum([len(l) for l in labs])]).astype(np.int32),
        ref_acc=acc, ref_totalPr=total_pr, ref_totalT=total_t,
        ref_pr_per_frame=np.array([u["pr_per_frame"] if u["pr_per_frame"] is not None else np.nan for u in tr]),
        ref_thresh=np.array([u["thresh"] if u["thresh"] is not None else (prune[0] if prune else 1e20) for u in tr]),
        ref_retries=np.array([u["retries"] for u in tr]),
        have_beams=have_beams, ref_qLo=qLo, ref_qHi=qHi, ref_sq=sq, ref_eq=eq,
    )
    for k in MODEL_KEYS:
        d["m_" + k] = getattr(fm, k)
    if extra:
        d.update(extra)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote %s.npz: %d utts, %d frames, totalPr/T=%.6f" % (name, len(feats), totalT, total_pr / max(total_t, 1)))


def herest(args, cwd, trace_cfg=True):
    cmd = [os.path.join(BIN, "HERest")]
    if trace_cfg:
        with open(os.path.join(cwd, "tr.cfg"), "w") as f:
            f.write("HFB: TRACE = 5\n")
        cmd += ["-C", "tr.cfg"]
    return run(cmd + args, cwd=cwd)


# --------------------------------------------------------------------------- HTKDemo

def htkdemo():
    w = os.path.join(WORK, "demo")
    shutil.rmtree(w, ignore_errors=True)
    os.makedirs(w)
    demo = os.path.join(REF, "HTKDemo")
    for d in ("data", "labels", "lists", "toolconfs", "protoconfs", "MakeProtoHMMSet"):
        os.symlink(os.path.join(demo, d), os.path.join(w, d))
    for d in ("proto", "hmms/hmm.0", "hmms/hmm.1", "hmms/hmm.2", "accs", "d26"):
        os.makedirs(os.path.join(w, d))
    run(["perl", "MakeProtoHMMSet", "protoconfs/proto_s1_m1_dc.pcf"], cwd=w)
    train = sorted(os.path.join("data/train", f) for f in os.listdir(os.path.join(demo, "data/train"))
                   if f.endswith(".mfc"))
    phones = open(os.path.join(demo, "lists/bcplist")).read().split()
    for m in phones:
        run([os.path.join(BIN, "HInit"), "-i", "10", "-L", "labels/bcplabs/mon", "-l", m, "-o", m,
             "-C", "toolconfs/hinit.conf", "-M", "hmms/hmm.0", "proto/" + m] + train, cwd=w)
        run([os.path.join(BIN, "HRest"), "-u", "tmvw", "-w", "3", "-v", "0.05", "-i", "10",
             "-L", "labels/bcplabs/mon", "-l", m, "-C", "toolconfs/hrest.conf", "-M", "hmms/hmm.1",
             "hmms/hmm.0/" + m] + train, cwd=w)
    # pre-expanded 26-dim features (same HParm code path HERest would run on the fly)
    with open(os.path.join(w, "tk.cfg"), "w") as f:
        f.write("TARGETKIND = MFCC_E_D\n")
    d26 = []
    for t in train:
        o = os.path.join("d26", os.path.basename(t))
        run([os.path.join(BIN, "HCopy"), "-C", "tk.cfg", t, o], cwd=w)
        d26.append(o)
    hs = htkio.read_mmf([os.path.join(w, "hmms/hmm.1", m) for m in phones], hmm_list=phones)
    fm = flatten(hs)
    feats = [htkio.read_htk_features(os.path.join(w, f))[0] for f in d26]
    labs = [np.array([fm.hmm_index[l] for l in htkio.read_label_file(
        os.path.join(demo, "labels/bcplabs/mon", os.path.basename(f).replace(".mfc", ".lab")))], dtype=np.int32)
        for f in d26]
    for name, prune in (("htkdemo_t2000", (2000.0, 0.0, 2000.0)), ("htkdemo_t20_15_200", (20.0, 15.0, 200.0))):
        targs = ["-t", "%.1f" % prune[0]] if prune[1] == 0 else ["-t"] + ["%.1f" % x for x in prune]
        shutil.rmtree(os.path.join(w, "accs")); os.makedirs(os.path.join(w, "accs"))
        out = herest(["-u", "tmvw", "-d", "hmms/hmm.1", "-L", "labels/bcplabs/mon"] + targs +
                     ["-T", "1", "-p", "1", "-M", "accs", "lists/bcplist"] + d26, cwd=w)
        tr = parse_trace(out, [f.shape[0] for f in feats])
        acc, tp, tt = htkio.read_acc_dump(os.path.join(w, "accs/HER1.acc"), hs, fm, 15)
        pack_case(name, hs, fm, feats, labs, acc, tp, tt, tr, prune)
    # M-step golden: re-estimated means / variances / transitions from the stock -p 0 merge
    run([os.path.join(BIN, "HERest"), "-w", "3", "-v", "0.05", "-u", "tmvw", "-d", "hmms/hmm.1", "-p", "0",
         "-M", "hmms/hmm.2", "lists/bcplist", "accs/HER1.acc"], cwd=w)


# --------------------------------------------------------------------------- synthetic

def synthetic(name, hs, n_utts, T, Q, prune, seed, tee=False, T_jitter=0, uflags="tmvw", add_short=False):
    w = os.path.join(WORK, name)
    shutil.rmtree(w, ignore_errors=True)
    os.makedirs(os.path.join(w, "feat")); os.makedirs(os.path.join(w, "accs"))
    htkio.write_mmf(os.path.join(w, "mmf"), hs)
    htkio.write_hmm_list(os.path.join(w, "hmmlist"), hs)
    # the reference reads the 6-digit text; re-parse so the flat model holds the same floats
    hs2 = htkio.read_mmf([os.path.join(w, "mmf")], hmm_list=open(os.path.join(w, "hmmlist")).read().splitlines())
    fm = flatten(hs2)
    tee_index = fm.hmm_index["sp"] if tee else None
    feats, labs = synth.sample_corpus(fm, n_utts, T, Q, seed=seed, tee_index=tee_index, T_jitter=T_jitter)
    if add_short:   # features drawn from a DIFFERENT transcription: forces beam retries
        rng = np.random.default_rng(seed + 2000)
        for i in (0, 5):
            other = [p for p in labs[i][::-1]]
            feats[i] = synth.sample_utterance(fm, np.asarray(other, dtype=np.int32), feats[i].shape[0], rng)
    if add_short:   # an utterance with fewer frames than its minimum duration: skipped with -7324
        rng = np.random.default_rng(seed + 1000)
        cand = [p for p in range(fm.P) if p != tee_index]
        labs.insert(2, rng.choice(cand, size=Q).astype(np.int32))
        feats.insert(2, rng.standard_normal((Q, fm.D)).astype(np.float32))
    mlf = {}
    scp = []
    for i, (f, l) in enumerate(zip(feats, labs)):
        fn = os.path.join(w, "feat", "u%04d.mfc" % i)
        htkio.write_htk_features(fn, f, hs.parm_kind)
        mlf["u%04d" % i] = [fm.names[j] for j in l]
        scp.append(fn)
    htkio.write_mlf(os.path.join(w, "labs.mlf"), mlf)
    with open(os.path.join(w, "train.scp"), "w") as f:
        f.write("\n".join(scp) + "\n")
    targs = [] if prune is None else (["-t"] + ["%.1f" % x for x in prune])
    out = herest(["-T", "1", "-u", uflags] + targs + ["-p", "1", "-H", "mmf", "-I", "labs.mlf", "-S", "train.scp",
                 "-M", "accs", "hmmlist"], cwd=w)
    tr = parse_trace(out, [f.shape[0] for f in feats])
    uf = sum({"t": 4, "m": 1, "v": 2, "w": 8}[c] for c in uflags)
    acc, tp, tt = htkio.read_acc_dump(os.path.join(w, "accs/HER1.acc"), hs2, fm, uf)
    pack_case(name, hs2, fm, feats, labs, acc, tp, tt, tr, prune, extra={"uflags": uf})


PRUNE_TEE = tuple(float(x) for x in os.environ.get("PRUNE_TEE", "15,10,200").split(","))


def main():
    os.makedirs(WORK, exist_ok=True)
    which = sys.argv[1:] or ["demo", "mono1", "tee2", "tied4", "long"]
    if "demo" in which:
        htkdemo()
    if "mono1" in which:   # config #2 shape, small: single Gaussian monophones, pruning on
        hs = synth.make_monophone_set(n_phones=12, M=1, seed=11, spread=0.2)
        synthetic("synth_mono_m1", hs, n_utts=6, T=300, Q=30, prune=(250.0, 150.0, 1000.0), seed=5, T_jitter=40)
    if "tee2" in which:    # M=2, entry-skip arcs and a tee sp model, tight pruning with retries
        hs = synth.make_monophone_set(n_phones=12, M=2, seed=12, tee_model=True, entry_skip=0.1, spread=0.15)
        synthetic("synth_tee_m2", hs, n_utts=8, T=260, Q=20, prune=PRUNE_TEE, seed=6, tee=True, T_jitter=30,
                  add_short=True)
    if "tied4" in which:   # configs #3/#4 shape, small: tied states, shared transitions, logical names
        hs = synth.make_tied_triphone_set(n_states=60, M=4, n_phys=40, n_logical=80, n_centre=8, seed=13, spread=0.2)
        synthetic("synth_tied_m4", hs, n_utts=6, T=320, Q=32, prune=None, seed=7, T_jitter=20)
    if "long" in which:    # config #5 shape, small: long utterance, beam active
        hs = synth.make_tied_triphone_set(n_states=90, M=3, n_phys=60, n_logical=60, n_centre=6, seed=14, spread=0.25)
        synthetic("synth_long_m3", hs, n_utts=2, T=1500, Q=160, prune=(250.0, 150.0, 1000.0), seed=8)


if __name__ == "__main__":
    main()
