"""The reference's fatal error paths on the hot path (missing #8):

  7332  CreateInsts: tee model first / last / twice in a row          HFB.c:557-565
  7323  SetBeta: beta prune failed                                    HFB.c:1257  (unreachable, see below)
  7390  StepAlpha: alpha prune failed                                 HFB.c:706, :718

The stock tool dies on them (HError with a positive code); the library reports the same number as the status of
the offending utterance and keeps the rest of the batch.  Deterministic triggers: a transcription that starts with /
ends with / repeats the tee model `sp`; a negative beam (PRUNEINIT = -1: every model is "outside"); a negative
MINFORPROB (minFrwdP = -1: every model fails the alpha test).

CPU part: the C oracle's codes are pinned to the STOCK binary's messages.  GPU part: library == oracle.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from htk_b200 import htkio
from htk_b200.flat import Batch, make_options

HEREST = os.path.join(ROOT, "oracle", "_ref", "bin", "HERest")


def _tee_batches():
    """synth_tee_m2's model has a tee `sp` (N = 3, a_13 > 0): three bad transcriptions + one good utterance."""
    z, fm, b, kw = load_golden("synth_tee_m2")
    sp = [i for i in range(fm.P) if fm.hmmNumStates[i] == 3]
    assert len(sp) == 1
    sp = sp[0]
    f0, f1 = int(z["frameOff"][0]), int(z["frameOff"][1])
    l0, l1 = int(z["labOff"][0]), int(z["labOff"][1])
    feat = z["feat"][f0:f1]
    good = np.array([q for q in z["lab"][l0:l1] if q != sp], dtype=np.int32)
    cases = {"first": np.concatenate([[sp], good]), "last": np.concatenate([good, [sp]]),
             "twice": np.concatenate([good[:3], [sp, sp], good[3:]]), "good": good}
    return fm, kw, feat, cases, sp


def _stock(tmp, fm_names, hs_path_args, feat, lab_names, extra):
    htkio.write_htk_features(os.path.join(tmp, "u.mfc"), feat)
    htkio.write_mlf(os.path.join(tmp, "l.mlf"), {"u": lab_names})
    open(os.path.join(tmp, "scp"), "w").write(os.path.join(tmp, "u.mfc") + "\n")
    os.makedirs(os.path.join(tmp, "o"), exist_ok=True)
    return subprocess.run([HEREST, "-u", "tmvw"] + extra + ["-p", "1"] + hs_path_args + ["-I", "l.mlf", "-S", "scp", "-M", "o", "list"],
                          cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def test_oracle_error_codes_are_the_stock_tools(tmp_path):
    """Pins the oracle's 7332 / 7323 / 7390 to the unmodified reference binary."""
    if not os.path.exists(HEREST):
        pytest.skip("reference HERest not built")
    from htk_b200 import synth
    from htk_b200.flat import flatten
    from oracle import oracle_lib as O
    tmp = str(tmp_path)
    hs = synth.make_monophone_set(n_phones=8, M=2, seed=12, tee_model=True, entry_skip=0.1, spread=0.15)
    htkio.write_mmf(os.path.join(tmp, "mmf"), hs)
    htkio.write_hmm_list(os.path.join(tmp, "list"), hs)
    names = open(os.path.join(tmp, "list")).read().splitlines()
    fm = flatten(htkio.read_mmf([os.path.join(tmp, "mmf")], hmm_list=names))
    sp = fm.hmm_index["sp"]
    feats, labs = synth.sample_corpus(fm, 1, 200, 14, seed=3, tee_index=sp)
    feat = feats[0]
    good = np.array([q for q in labs[0] if q != sp], dtype=np.int32)

    def both(lab, extra, **okw):
        p = _stock(tmp, names, ["-H", "mmf"], feat, [fm.names[q] for q in lab], extra)
        b = Batch([feat], [np.asarray(lab, np.int32)], fm.D)
        _, res, _ = O.accumulate(fm, make_options(**okw), b, acc_double=True, want_beams=False)
        return p, res[0][0]

    for lab in (np.concatenate([[sp], good]), np.concatenate([good, [sp]]), np.concatenate([good[:3], [sp, sp], good[3:]])):
        p, st = both(lab, [])
        assert p.returncode != 0 and "7332" in p.stdout, p.stdout[-500:]
        assert st == 7332
    # the command line range-checks -t and -c (HERest.c:353, :411); the HFB configuration variables do not (HFB.c:221-233)
    open(os.path.join(tmp, "beam.cfg"), "w").write("PRUNEINIT = -1.0\nPRUNEINC = 0.0\nPRUNELIM = -1.0\n")
    open(os.path.join(tmp, "frwd.cfg"), "w").write("MINFORPROB = -1.0\n")
    # 7323 cannot be reached: with a threshold >= 0 the model that holds gMax always survives both loops of
    # HFB.c:1254-1263, and a negative threshold (config file only) empties the beam from the top instead, which is the
    # "over pruning" skip -7324 (HFB.c:1266-1268, :1354).  Pinned as such.
    p, st = both(good, ["-C", "beam.cfg"], prune=(-1.0, 0.0, -1.0))
    assert p.returncode == 0 and "7324" in p.stdout and "7323" not in p.stdout, p.stdout[-500:]
    assert st == 7324
    p, st = both(good, ["-C", "frwd.cfg"], min_frwd_p=-1.0)
    assert p.returncode != 0 and "7390" in p.stdout, p.stdout[-500:]
    assert st == 7390
    p, st = both(good, [])
    assert p.returncode == 0 and st == 0


@pytest.mark.gpu
def test_tee_misuse_is_reported_per_utterance():
    from htk_b200.estep import ForwardBackward
    from oracle import oracle_lib as O
    fm, kw, feat, cases, sp = _tee_batches()
    order = ["first", "good", "last", "twice", "good"]
    b = Batch([feat] * len(order), [cases[k] for k in order], fm.D)
    fb = ForwardBackward(fm, **kw)
    res, _ = fb.FBFile(b)
    acc = fb.GetAccs()
    fb.close()
    oacc, ores, _ = O.accumulate(fm, make_options(**kw), b, acc_double=True, want_beams=False)
    assert [r.status for r in res] == [7332, 0, 7332, 7332, 0] == [o[0] for o in ores]
    assert abs(res[1].pr - ores[1][2]) <= 1e-6 * abs(ores[1][2])
    L = fm.layout
    assert acc[L.numOk] == 2 and acc[L.totalT] == 2 * feat.shape[0]
    from conftest import acc_errors
    assert max(acc_errors(acc, oacc, fm).values()) < 1e-4            # the bad utterances left nothing behind


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["synth_tee_m2", "synth_tied_m4"])      # general kernels / standard-topology kernels
def test_beta_and_alpha_prune_failures(name):
    from htk_b200.estep import ForwardBackward
    from oracle import oracle_lib as O
    z, fm, b, kw = load_golden(name)
    for okw, code in ((dict(prune=(-1.0, 0.0, -1.0)), 7324), (dict(min_frwd_p=-1.0), 7390)):
        k2 = dict(kw); k2.update(okw)
        fb = ForwardBackward(fm, **k2)
        res, _ = fb.FBFile(b)
        acc = fb.GetAccs()
        fb.close()
        _, ores, _ = O.accumulate(fm, make_options(**k2), b, acc_double=True, want_beams=False)
        assert [r.status for r in res] == [o[0] for o in ores]
        assert code in [r.status for r in res]
        L = fm.layout
        assert acc[L.numOk] == sum(1 for r in res if r.status == 0)
        # a failed utterance contributes nothing (the reference would have died before accumulating it)
        if all(r.status != 0 for r in res):
            assert not np.any(acc[:L.numEgs]) and acc[L.totalT] == 0
