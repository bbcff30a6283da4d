"""The CPU oracle (oracle/hfb_oracle.c) against fixtures produced by the reference's own
HERest (tests/golden/make_golden.py).  Accumulators are float sums made in the same order
with the same float/double choices, so the bar is bit-exact, beams included."""
import numpy as np
import pytest

from conftest import (GOLDEN_CASES, QUALIFIER_CASES, TWO_MODEL_CASES, acc_errors, load_golden, load_qualifier_golden,
                      load_two_model_golden)
from htk_b200.flat import make_options
from oracle import oracle_lib as O


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_bit_exact(name):
    z, fm, b, kw = load_golden(name)
    acc, res, beams = O.accumulate(fm, make_options(**kw), b)
    ref = z["ref_acc"]
    L = fm.layout
    # every accumulator the reference dumps (HER1.acc) must be identical
    assert np.array_equal(acc[:L.totalT].astype(np.float32), ref[:L.totalT].astype(np.float32))
    assert acc[L.totalT] == ref[L.totalT]
    assert np.float32(acc[L.totalPr]) == np.float32(ref[L.totalPr])
    T = np.diff(z["frameOff"])
    for (st, retries, pr, thr), t, rpf, rthr, rret in zip(res, T, z["ref_pr_per_frame"], z["ref_thresh"],
                                                          z["ref_retries"]):
        if np.isnan(rpf):
            assert st == 7324                       # reference skipped it with warning -7324
            continue
        assert st == 0
        assert float("%e" % (pr / t)) == pytest.approx(rpf, rel=0, abs=0)   # printed with %e
        assert thr == rthr and retries == rret
    if bool(z["have_beams"]):
        for k, a in (("qLo", beams.qLo), ("qHi", beams.qHi), ("sq", beams.sq), ("eq", beams.eq)):
            r = z["ref_" + k]
            m = r > 0
            assert np.array_equal(a[m], r[m]), k


@pytest.mark.parametrize("name", TWO_MODEL_CASES)
def test_oracle_two_model_matches_reference_bit_exact(name):
    """Two-model re-estimation (HFB.c:296-333, :1518-1547): the stock HERest with ALIGNMODELMMF / ALIGNHMMLIST aligned with
    one set and dumped the accumulators of another; the oracle, given both, reproduces the dump bit for bit."""
    z, fu, fa, b, kw = load_two_model_golden(name)
    acc, res, _ = O.accumulate(fu, make_options(**kw), b)
    ref = z["ref_acc"]
    L = fu.layout
    assert np.all(ref[L.tran:L.wtC] == 0)              # HFB.c:313-316: no transition statistics on a 2-model alignment
    assert np.array_equal(acc[:L.totalT].astype(np.float32), ref[:L.totalT].astype(np.float32))
    assert acc[L.totalT] == ref[L.totalT]
    assert np.float32(acc[L.totalPr]) == np.float32(ref[L.totalPr])
    for (st, retries, pr, thr), t, rpf in zip(res, np.diff(z["frameOff"]), z["ref_pr_per_frame"]):
        assert st == 0 and float("%e" % (pr / t)) == rpf
    # and it is not the one-model answer: the update set alone gives other occupancies
    kw1 = dict(kw); kw1.pop("align_model")
    one, _, _ = O.accumulate(fu, make_options(**kw1), b)
    assert not np.allclose(one[L.muOcc:L.vaSum], acc[L.muOcc:L.vaSum], rtol=1e-3, atol=1e-3)


def test_oracle_double_accumulators_close_to_float():
    z, fm, b, kw = load_golden("synth_tied_m4")
    a32, _, _ = O.accumulate(fm, make_options(**kw), b)
    a64, _, _ = O.accumulate(fm, make_options(**kw), b, acc_double=True)
    e = acc_errors(a64, a32, fm)
    assert max(e.values()) < 5e-6, e


def test_oracle_threads_equal_sequential():
    z, fm, b, kw = load_golden("synth_tee_m2")
    a1, r1, _ = O.accumulate(fm, make_options(**kw), b, acc_double=True, threads=1)
    a4, r4, _ = O.accumulate(fm, make_options(**kw), b, acc_double=True, threads=4)
    assert [x[0] for x in r1] == [x[0] for x in r4]
    assert np.allclose(a1, a4, rtol=1e-12, atol=1e-12)


def test_oracle_update_flags():
    z, fm, b, kw = load_golden("synth_tied_m4")
    L = fm.layout
    kw2 = dict(kw); kw2["uflags"] = 1          # means only
    a, _, _ = O.accumulate(fm, make_options(**kw2), b, acc_double=True)
    assert np.all(a[L.vaSum:L.numEgs] == 0) and np.all(a[L.tran:L.wtC] == 0) and np.all(a[L.wtC:L.wtOcc] == 0)
    assert a[L.muOcc:L.vaSum].sum() > 0 and a[L.wtOcc:L.muSum].sum() > 0


def test_min_durs():
    z, fm, b, kw = load_golden("synth_tee_m2")
    md = O.min_durs(fm)
    names = list(z["names"])
    tee = fm.hmmTrans[names.index("sp")]
    assert md[tee] == 0
    assert all(md[i] >= 1 for i in range(fm.numTrans) if i != tee)


@pytest.mark.parametrize("name", QUALIFIER_CASES)
def test_qualifier_oracle_bit_identical_to_hcopy(name):
    """oracle/hparm_oracle.py (HParm.c AddQualifiers / HSigP.c Regress, FZeroMean restated) against what the
    unmodified reference's HCopy wrote for the target kind, utterances of 1..333 frames."""
    from oracle import hparm_oracle as H
    q, static, expanded = load_qualifier_golden(name)
    for x, y in zip(static, expanded):
        o = H.expand(x, q.del_win, q.acc_win, q.third_win, q.simple_diffs, q.zero_mean_cols)
        assert o.shape == y.shape == (x.shape[0], q.vec_size)
        assert np.array_equal(o.view(np.uint32), y.view(np.uint32))
