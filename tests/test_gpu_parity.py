"""Parity of the CUDA path (through the libhfbgpu C ABI) with the reference.

Three kinds of evidence:
  * golden fixtures produced by the reference's own HERest (tests/golden/*.npz);
  * the CPU oracle (pinned bit-exact to those fixtures) on the same seeded inputs;
  * size-independent properties at larger sizes (occupancies sum to T, batch linearity,
    idempotence of zeroing, transition counts vs state occupancies).

Tolerances (BASELINE.json north_star): log-likelihoods, occupancies and accumulators within
1e-4 relative; pruned-frame sets identical except for counted beam-boundary ties.
"""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, QUALIFIER_CASES, acc_errors, load_golden, load_qualifier_golden
from htk_b200.flat import Batch, make_options

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _fb(fm, **kw):
    from htk_b200.estep import ForwardBackward
    return ForwardBackward(fm, **kw)


def _oracle(fm, b, kw, acc_double=True):
    from oracle import oracle_lib as O
    return O.accumulate(fm, make_options(**kw), b, acc_double=acc_double)


@pytest.mark.parametrize("gmm_kernel", [1, 2])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_accumulators_and_beams(name, gmm_kernel):
    z, fm, b, kw = load_golden(name)
    try:
        fb = _fb(fm, gmm_kernel=gmm_kernel, **kw)
    except Exception as e:
        if gmm_kernel == 2 and getattr(e, "code", 0) == -5:
            pytest.skip("tcgen05 path not built")
        raise
    res, beams = fb.FBFile(b, want_beams=True)
    acc = fb.GetAccs()
    ref = z["ref_acc"]
    T = np.diff(z["frameOff"])
    n_ok = 0
    for r, t, rpf, rthr, rret in zip(res, T, z["ref_pr_per_frame"], z["ref_thresh"], z["ref_retries"]):
        if np.isnan(rpf):
            assert r.status == 7324
            continue
        assert r.status == 0
        n_ok += 1
        assert abs(r.pr / t - rpf) <= RTOL * abs(rpf) + 1e-6      # reference prints 7 digits
        assert r.pruneThresh == rthr and r.retries == rret
    L = fm.layout
    assert acc[L.numOk] == n_ok and acc[L.numSkipped] == len(res) - n_ok
    e = acc_errors(acc, ref, fm)
    assert max(e.values()) < RTOL, e
    if bool(z["have_beams"]):
        ties = 0
        total = 0
        for k, a in (("qLo", beams.qLo), ("qHi", beams.qHi), ("sq", beams.sq), ("eq", beams.eq)):
            r = z["ref_" + k]
            m = r > 0
            total += int(m.sum())
            ties += int(np.sum(a[m] != r[m]))
        # identical pruned-frame sets except counted beam-boundary ties
        assert ties <= max(1, total // 2000), "beam mismatches: %d of %d" % (ties, total)
    fb.close()


@pytest.mark.parametrize("name", ["synth_tee_m2", "synth_tied_m4"])
def test_against_oracle_fp64(name):
    """Same inputs through the oracle with FP64 accumulators: tighter than the float dump."""
    z, fm, b, kw = load_golden(name)
    fb = _fb(fm, **kw)
    res, beams = fb.FBFile(b, want_beams=True)
    acc = fb.GetAccs()
    oacc, ores, obeams = _oracle(fm, b, kw)
    for r, o in zip(res, ores):
        assert r.status == o[0]
        if o[0] == 0:
            assert abs(r.pr - o[2]) <= 1e-6 * abs(o[2])
            assert r.pruneThresh == o[3]
    e = acc_errors(acc, oacc, fm)
    assert max(e.values()) < RTOL, e
    assert np.array_equal(beams.qLo, obeams.qLo) and np.array_equal(beams.qHi, obeams.qHi)
    assert np.array_equal(beams.sq, obeams.sq) and np.array_equal(beams.eq, obeams.eq)
    fb.close()


@pytest.mark.parametrize("gmm_kernel", [1, 2])
def test_outp_matches_oracle(gmm_kernel):
    from oracle import oracle_lib as O
    names = ("htkdemo_t2000", "synth_tied_m4") if gmm_kernel == 1 else ("synth_tied_m4", "synth_tee_m2", "synth_long_m3")
    for name in names:
        z, fm, b, kw = load_golden(name)
        fb = _fb(fm, gmm_kernel=gmm_kernel, **kw)      # kernel 2 = tcgen05: must exist for M > 1 sets
        feat = z["feat"][:300]
        states = np.arange(fm.J, dtype=np.int32)
        got = fb.OutP(feat, states)
        want = O.state_loglik(fm, feat, states)
        # absolute error of a log-likelihood of magnitude ~60-100 (float eps there is ~8e-6)
        assert np.max(np.abs(got - want)) < 2e-4, np.max(np.abs(got - want))
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-5
        fb.close()


@pytest.mark.parametrize("variant", ["f16_tmem", "f16_fused", "f16_pair", "tf32_pair", "tf32_single"])
def test_outp_tensor_core_variants_on_badly_scaled_features(variant, monkeypatch):
    """The five tcgen05 GMM kernels (gmm_tc4 = default: A operand in tensor memory; gmm_tc3: the same with A in shared
    memory, both with fused expansion + taper skipping; CTA pair 3xFP16 with the
    pre-expanded operand, CTA pair 3xTF32, single CTA 3xTF32) against a
    float64 evaluation on features whose dimensions span five decades of scale with offsets of hundreds --
    what a real front end delivers; the FP16 split relies on its per-dimension power-of-two scaling here."""
    from htk_b200 import synth
    from htk_b200.flat import flatten
    if variant == "f16_fused":                     # gmm_tc3_kernel: operand A expanded into shared memory
        monkeypatch.setenv("HFBGPU_GMM_V3", "1")
    elif variant == "f16_pair":                    # round-1 kernel: pre-expanded operand in HBM, no taper skipping
        monkeypatch.setenv("HFBGPU_GMM_V2", "1")
    elif variant == "tf32_pair":
        monkeypatch.setenv("HFBGPU_TC_TF32", "1")
    elif variant == "tf32_single":
        monkeypatch.setenv("HFBGPU_NO_PAIR", "1")
    hs = synth.make_tied_triphone_set(n_states=200, M=16, n_phys=120, n_logical=120, n_centre=10, seed=23, spread=0.2)
    fm = flatten(hs)
    rs = np.random.default_rng(5)
    sc = (10.0 ** rs.uniform(-2, 3, fm.D)).astype(np.float32); of = rs.uniform(-500, 500, fm.D).astype(np.float32)
    fm.mean[:, :fm.D] = fm.mean[:, :fm.D] * sc + of
    fm.ivar[:, :fm.D] = fm.ivar[:, :fm.D] / (sc * sc)
    fm.gConst[:] = (fm.D * np.log(2 * np.pi) - np.sum(np.log(fm.ivar[:, :fm.D].astype(np.float64)), axis=1)).astype(np.float32)
    feats, _ = synth.sample_corpus(fm, n_utts=1, T=700, Q=60, seed=3)
    x = feats[0].astype(np.float64)
    states = np.arange(fm.J, dtype=np.int32)
    mean = fm.mean.astype(np.float64); iv = fm.ivar.astype(np.float64); gc = fm.gConst.astype(np.float64)
    ex = np.zeros((len(x), fm.J))
    for s_ in range(fm.J):
        o, e = fm.stateMixOff[s_], fm.stateMixOff[s_ + 1]
        g = fm.mixGauss[o:e]
        d = x[:, None, :] - mean[g][None][:, :, :fm.D]
        lp = -0.5 * (gc[g][None] + np.sum(d * d * iv[g][None][:, :, :fm.D], axis=2)) + fm.mixLogWt[o:e].astype(np.float64)[None]
        m = lp.max(1); ex[:, s_] = m + np.log(np.exp(lp - m[:, None]).sum(1))
    fb = _fb(fm, gmm_kernel=2)
    got = fb.OutP(feats[0], states).astype(np.float64)
    fb.close()
    err = np.abs(got - ex)
    assert err.mean() < 1e-5 and err.max() < 1e-4, (err.mean(), err.max())


def test_update_flags_and_zero():
    z, fm, b, kw = load_golden("synth_tied_m4")
    L = fm.layout
    for uf in (1, 2, 4, 8, 3, 15):
        kw2 = dict(kw); kw2["uflags"] = uf
        fb = _fb(fm, **kw2)
        fb.FBFile(b)
        acc = fb.GetAccs()
        oacc, _, _ = _oracle(fm, b, kw2)
        e = acc_errors(acc, oacc, fm)
        assert max(e.values()) < RTOL, (uf, e)
        if not uf & 4:
            assert np.all(acc[L.tran:L.wtC] == 0)
        if not uf & 1:
            assert np.all(acc[L.muSum:L.vaSum] == 0)
        fb.ZeroAccs()
        assert np.all(fb.GetAccs() == 0)
        fb.close()


def test_batch_linearity_and_empty():
    """acc(batch A + batch B) == acc(A) + acc(B); an empty batch is a no-op."""
    z, fm, b, kw = load_golden("synth_mono_m1")
    fo, lo = z["frameOff"], z["labOff"]
    fb = _fb(fm, **kw)
    fb.FBFile(b)
    whole = fb.GetAccs()
    fb.ZeroAccs()
    k = 3
    bA = Batch.from_arrays(z["feat"], fo[:k + 1], z["lab"], lo[:k + 1])
    bB = Batch.from_arrays(z["feat"], fo[k:], z["lab"], lo[k:])
    fb.FBFile(bA); fb.FBFile(bB)
    parts = fb.GetAccs()
    # the statistics kernel sums FP32 fragments over the positions of a tied state in list order, which
    # depends on how the utterances are grouped into calls: agreement to FP32 summation noise, not bitwise
    e = acc_errors(parts, whole, fm)
    assert max(e.values()) < 1e-5, e
    empty = Batch([], [], fm.D)
    r, _ = fb.FBFile(empty)
    assert r == [] and np.array_equal(parts, fb.GetAccs())
    fb.close()


def test_small_workspace_waves_equal_single_wave():
    z, fm, b, kw = load_golden("synth_tied_m4")
    fb = _fb(fm, **kw); fb.FBFile(b); a1 = fb.GetAccs(); fb.close()
    fb = _fb(fm, workspace_bytes=3 << 20, **kw); fb.FBFile(b); a2 = fb.GetAccs(); fb.close()
    e = acc_errors(a2, a1, fm)           # FP32 summation order differs between wave splits (see above)
    assert max(e.values()) < 1e-5, e


def test_properties_at_scale():
    """Config-#3-shaped workload at a size the oracle would take minutes on: check the
    invariants the domain offers instead."""
    from htk_b200 import synth
    from htk_b200.flat import flatten
    hs = synth.make_tied_triphone_set(n_states=600, M=8, n_phys=400, n_logical=400, n_centre=20, seed=21, spread=0.2)
    fm = flatten(hs)
    feats, labs = synth.sample_corpus(fm, n_utts=24, T=600, Q=60, seed=3)
    b = Batch(feats, labs, fm.D)
    fb = _fb(fm)
    res, _ = fb.FBFile(b)
    acc = fb.GetAccs()
    L = fm.layout
    assert all(r.status == 0 for r in res)
    T = b.totalT
    assert acc[L.totalT] == T
    # every frame distributes one unit of occupancy over emitting states (minus what the
    # minimum-occupancy rule drops, < e^-10 per component)
    assert abs(acc[L.wtOcc:L.muSum].sum() - T) < 2e-3 * T
    assert abs(acc[L.muOcc:L.vaSum].sum() - acc[L.wtOcc:L.muSum].sum()) < 1e-6 * T
    assert abs(acc[L.wtC:L.wtOcc].sum() - acc[L.wtOcc:L.muSum].sum()) < 1e-6 * T
    # second-order sums are non-negative
    assert acc[L.vaSum:L.vaOcc].min() >= 0.0
    # each model instance is entered and left exactly once per utterance
    N = 5
    tr = acc[L.tran:L.tranOcc].reshape(-1, N, N)
    n_inst = sum(len(l) for l in labs)
    assert abs(tr[:, 0, 1:].sum() - n_inst) < 1e-3 * n_inst
    assert abs(tr[:, 1:N - 1, N - 1].sum() - n_inst) < 1e-3 * n_inst
    assert acc[L.numEgs:L.totalT].sum() == n_inst
    # total log-likelihood equals the sum of per-utterance values
    assert abs(acc[L.totalPr] - sum(r.pr for r in res)) < 1e-6 * abs(acc[L.totalPr])
    # against the oracle on a 4-utterance subset
    sub = Batch(feats[:4], labs[:4], fm.D)
    fb.ZeroAccs(); fb.FBFile(sub); a = fb.GetAccs()
    oacc, ores, _ = _oracle(fm, sub, dict(prune=None))
    e = acc_errors(a, oacc, fm)
    assert max(e.values()) < RTOL, e
    fb.close()


def test_exact_ladd_env(monkeypatch):
    """FP64 exp/log in every log-add (HFBGPU_EXACT_LADD) vs the FP32 correction term."""
    z, fm, b, kw = load_golden("htkdemo_t20_15_200")
    fb = _fb(fm, **kw); r1, _ = fb.FBFile(b); a1 = fb.GetAccs(); fb.close()
    monkeypatch.setenv("HFBGPU_EXACT_LADD", "1")
    fb = _fb(fm, **kw); r2, _ = fb.FBFile(b); a2 = fb.GetAccs(); fb.close()
    for x, y in zip(r1, r2):
        assert abs(x.pr - y.pr) < 1e-4
    e = acc_errors(a1, a2, fm)
    assert max(e.values()) < 2e-5, e


@pytest.mark.parametrize("hook", ["HFBGPU_NO_FAST", "HFBGPU_FORCE_REDO", "HFBGPU_NO_L2R"])
@pytest.mark.parametrize("name", ["htkdemo_t20_15_200", "synth_tee_m2", "synth_tied_m4", "synth_long_m3"])
def test_generic_kernels_and_redo_path(name, hook, monkeypatch):
    """The generic (any N, any Q) recursion kernels and the fast->generic alpha fallback give the
    same beams and accumulators as the register-resident fast path."""
    z, fm, b, kw = load_golden(name)
    fb = _fb(fm, **kw); r1, b1 = fb.FBFile(b, want_beams=True); a1 = fb.GetAccs(); st1 = fb.stats(); fb.close()
    monkeypatch.setenv(hook, "1")
    fb = _fb(fm, **kw); r2, b2 = fb.FBFile(b, want_beams=True); a2 = fb.GetAccs(); st2 = fb.stats(); fb.close()
    # the standard-topology kernels (hfb_l2r.cuh) run exactly on the sets without tee models / skips
    assert (st1.launchesL2R > 0) == (name != "synth_tee_m2"), (name, st1.launchesL2R)
    assert hook == "HFBGPU_FORCE_REDO" or st2.launchesL2R == 0
    for x, y in zip(r1, r2):
        assert x.status == y.status and x.pruneThresh == y.pruneThresh
        if x.status == 0:
            assert abs(x.pr - y.pr) <= 1e-9 * abs(x.pr)
    for k in ("qLo", "qHi", "sq", "eq"):
        assert np.array_equal(getattr(b1, k), getattr(b2, k)), k
    e = acc_errors(a2, a1, fm)             # identical alpha/beta; FP32 summation order of the statistics varies
    assert max(e.values()) < 1e-5, e
    e = acc_errors(a2, z["ref_acc"], fm)
    assert max(e.values()) < RTOL, e


@pytest.mark.parametrize("prune", [None, (2.0, 6.0, 200.0), (60.0, 30.0, 300.0)])
def test_beta_warp_kernel_equals_block_kernel(prune, monkeypatch):
    """beta_l2r_warp_kernel (one warp per utterance, four models per lane, no barriers) against beta_l2r_kernel
    (one thread per model, HFBGPU_NO_BETA_WARP) and the oracle: transcriptions of 1..128 labels, with and without
    pruning (tight beam: StepBack retries), identical beams, thresholds and log-likelihoods."""
    from htk_b200 import synth
    from htk_b200.flat import flatten
    hs = synth.make_tied_triphone_set(n_states=60, M=4, n_phys=40, n_logical=40, n_centre=6, seed=5, spread=0.3)
    fm = flatten(hs)
    feats, labs = [], []
    for i, Q in enumerate([1, 2, 5, 31, 32, 33, 63, 64, 65, 97, 127, 128]):
        f, l = synth.sample_corpus(fm, n_utts=1, T=max(12, 4 * Q + 7 * (i % 3)), Q=Q, seed=100 + i)
        feats += f; labs += l
    b = Batch(feats, labs, fm.D)
    kw = dict(prune=prune)
    fb = _fb(fm, **kw); r1, b1 = fb.FBFile(b, want_beams=True); a1 = fb.GetAccs(); fb.close()
    monkeypatch.setenv("HFBGPU_NO_BETA_WARP", "1")
    fb = _fb(fm, **kw); r2, b2 = fb.FBFile(b, want_beams=True); a2 = fb.GetAccs(); fb.close()
    assert any(r.status == 0 for r in r1)
    if prune and prune[0] < 30:
        assert any(r.retries > 0 for r in r1), [r.retries for r in r1]
    for x, y in zip(r1, r2):
        assert (x.status, x.retries, x.pruneThresh) == (y.status, y.retries, y.pruneThresh)
        if x.status == 0:
            assert x.pr == y.pr
    for k in ("qLo", "qHi", "sq", "eq"):
        assert np.array_equal(getattr(b1, k), getattr(b2, k)), k
    e = acc_errors(a1, a2, fm)
    assert max(e.values()) < 1e-5, e
    oacc, ores, obeams = _oracle(fm, b, kw)
    for x, y in zip(r1, ores):
        assert x.status == y[0] and (y[0] != 0 or x.pruneThresh == y[3])
    for k in ("qLo", "qHi", "sq", "eq"):
        assert np.array_equal(getattr(b1, k), getattr(obeams, k)), k
    e = acc_errors(a1, oacc, fm)
    assert max(e.values()) < RTOL, e


@pytest.mark.parametrize("name", ["synth_tied_m4", "synth_long_m3", "synth_tee_m2"])
def test_tensor_core_statistics_equal_fp32_statistics(name, monkeypatch):
    """stats4_kernel (occupancy-weighted sums as 3xTF32 mma products about the state centre) against
    stats3_kernel (FP32 sums about each component's own mean, HFBGPU_STATS3) and the reference."""
    z, fm, b, kw = load_golden(name)
    fb = _fb(fm, **kw); fb.FBFile(b); a1 = fb.GetAccs(); fb.close()
    monkeypatch.setenv("HFBGPU_STATS3", "1")
    fb = _fb(fm, **kw); fb.FBFile(b); a2 = fb.GetAccs(); fb.close()
    e = acc_errors(a1, a2, fm)             # tcgen05 statistics: posteriors from the 3xFP16-split product (~4e-6 on log N,
    assert max(e.values()) < 5e-5, e       # times |o - mu| / sigma ~ 3 in the centred sums); stats3: the reference's order of operations
    e = acc_errors(a1, z["ref_acc"], fm)
    assert max(e.values()) < RTOL, e


@pytest.mark.parametrize("name", ["synth_tied_m4", "synth_long_m3", "synth_tee_m2"])
@pytest.mark.parametrize("cap", [None, "700"])
def test_statistics_front_kernel_equals_inline_front(name, cap, monkeypatch):
    """stats_pre_kernel (transitions + list of valid frames per position) against the inline front of
    stats5_kernel (HFBGPU_NO_STATS_PRE); with a list too small for the wave (HFBGPU_STATS_PRE_CAP) some positions
    take one front and some the other inside the same launch."""
    z, fm, b, kw = load_golden(name)
    monkeypatch.setenv("HFBGPU_STATS5", "1")           # same back end (mma.sync sums) behind both fronts
    if cap:
        monkeypatch.setenv("HFBGPU_STATS_PRE_CAP", cap)
    fb = _fb(fm, **kw); fb.FBFile(b); a1 = fb.GetAccs(); fb.close()
    monkeypatch.setenv("HFBGPU_NO_STATS_PRE", "1")
    fb = _fb(fm, **kw); fb.FBFile(b); a2 = fb.GetAccs(); fb.close()
    e = acc_errors(a1, a2, fm)             # chunks group other frames: FP32 summation order differs
    assert max(e.values()) < 5e-6, e
    e = acc_errors(a1, z["ref_acc"], fm)
    assert max(e.values()) < RTOL, e


@pytest.mark.parametrize("shape", [
    dict(D=39, M=[1, 3, 4, 7, 16], parm="MFCC_0_D_A"),      # mixed component counts incl. single Gaussians, dead components
    dict(D=39, M=20, parm="MFCC_0_D_A"),                    # two 16-component tiles in the statistics kernel, MP = 32
    dict(D=26, M=5, parm="MFCC_E_D"),                       # 7 of 12 K steps, 4 right-hand-side tiles
    dict(D=13, M=4, parm="MFCC_0"),                         # one K chunk
    # MP = 64: a state spans two 32-column chunks of the K1 epilogue (carry).  2 400 Gaussians on 6 utterances: the centred
    # first-order sums of the barely occupied ones sit at 0.8-1.0e-4 from the FP64 oracle with EVERY kernel combination
    # (FP32 CUDA-core GMM 8.2e-5, round-1 tensor-core kernel 9.8e-5, FP32 statistics kernel 1.0e-4) -- the
    # cancellation SURVEY 8a warns about, not a kernel's rounding -- hence the wider bound on that block alone
    dict(D=39, M=40, parm="MFCC_0_D_A", tol_centred=2e-4),
    dict(D=63, M=4, parm="USER"),                           # 8 K steps: the A operand fills all 256 TMEM columns left of the accumulators
    dict(D=50, M=1, parm="USER"),                           # single-Gaussian set on the wide (DP = 64) kernel variants
])
def test_model_shapes_against_oracle(shape):
    """Kernel template / tiling variants the golden fixtures do not reach, against the C oracle."""
    from htk_b200 import synth
    from htk_b200.flat import flatten
    hs = synth.make_tied_triphone_set(n_states=60, M=shape["M"], n_phys=40, n_logical=40, n_centre=6, D=shape["D"],
                                      seed=77, spread=0.2, parm_kind=shape["parm"])
    fm = flatten(hs)
    feats, labs = synth.sample_corpus(fm, n_utts=6, T=260, Q=24, seed=8)
    b = Batch(feats, labs, fm.D)
    kw = dict(prune=(60.0, 30.0, 300.0))
    fb = _fb(fm, **kw)
    res, beams = fb.FBFile(b, want_beams=True)
    acc = fb.GetAccs(); st = fb.stats(); fb.close()
    oacc, ores, obeams = _oracle(fm, b, kw)
    assert st.launchesL2R > 0
    for r, o in zip(res, ores):
        assert r.status == o[0] and r.pruneThresh == o[3]
        if o[0] == 0:
            assert abs(r.pr - o[2]) <= 1e-6 * abs(o[2])
    for k in ("qLo", "qHi", "sq", "eq"):
        assert np.array_equal(getattr(beams, k), getattr(obeams, k)), k
    e = acc_errors(acc, oacc, fm)
    for k, v in e.items():
        assert v < (shape.get("tol_centred", RTOL) if k in ("muSum", "vaSum") else RTOL), (k, v, e)


def test_shared_variance_vectors_against_oracle():
    """Variance vectors tied below the component level (~v macros: one VaAcc for several components,
    HTrain.c:1001-1017): accumulators keyed by varId, against the C oracle."""
    from htk_b200 import synth
    from htk_b200.flat import flatten
    hs = synth.make_tied_triphone_set(n_states=40, M=4, n_phys=30, n_logical=30, n_centre=5, seed=91, spread=0.2)
    seen = set()
    for h in hs.hmms:
        for st in h.states:
            if id(st) in seen:
                continue
            seen.add(id(st))
            v = st.mixes[0][1].var
            for _, g in st.mixes[1:]:
                g.var = v                                  # all components of the state share one variance vector
                g.gconst = None
    fm = flatten(hs)
    assert fm.numVarAcc == fm.J and fm.numMeanAcc == fm.G
    feats, labs = synth.sample_corpus(fm, n_utts=5, T=240, Q=20, seed=14)
    b = Batch(feats, labs, fm.D)
    kw = dict(prune=None)
    fb = _fb(fm, **kw); res, _ = fb.FBFile(b); acc = fb.GetAccs(); fb.close()
    oacc, ores, _ = _oracle(fm, b, kw)
    for r, o in zip(res, ores):
        assert r.status == o[0] and abs(r.pr - o[2]) <= 1e-6 * abs(o[2])
    e = acc_errors(acc, oacc, fm)
    assert max(e.values()) < RTOL, e


@pytest.mark.parametrize("prune", [(120.0, 60.0, 600.0), (5000.0, 0.0, 5000.0)])
def test_long_transcription_sliding_beta_window(prune, monkeypatch):
    """Transcriptions longer than one CTA's 256 threads under a beam: beta_l2r_slide_kernel (256-model sliding
    window) against the C oracle; with a beam too wide for the window every utterance takes the redo path through
    the one-thread-per-label kernel (second parameter set); both against HFBGPU_NO_SLIDE."""
    from htk_b200 import synth
    from htk_b200.flat import flatten
    hs = synth.make_tied_triphone_set(n_states=90, M=4, n_phys=60, n_logical=60, n_centre=8, seed=61, spread=0.2)
    fm = flatten(hs)
    feats, labs = synth.sample_corpus(fm, n_utts=3, T=2600, Q=330, seed=12)
    b = Batch(feats, labs, fm.D)
    kw = dict(prune=prune)
    fb = _fb(fm, **kw); res, beams = fb.FBFile(b, want_beams=True); acc = fb.GetAccs(); st = fb.stats(); fb.close()
    assert st.launchesL2R > 0 and st.launchesBeta == 2          # slide kernel + redo pass
    oacc, ores, obeams = _oracle(fm, b, kw)
    for r, o in zip(res, ores):
        assert r.status == o[0] and r.pruneThresh == o[3] and r.retries == o[1]
        if o[0] == 0:
            assert abs(r.pr - o[2]) <= 1e-6 * abs(o[2])
    for k in ("qLo", "qHi", "sq", "eq"):
        assert np.array_equal(getattr(beams, k), getattr(obeams, k)), k
    e = acc_errors(acc, oacc, fm)
    assert max(e.values()) < RTOL, e
    monkeypatch.setenv("HFBGPU_NO_SLIDE", "1")
    fb = _fb(fm, **kw); res2, beams2 = fb.FBFile(b, want_beams=True); fb.close()
    for x, y in zip(res, res2):
        assert x.status == y.status and x.pruneThresh == y.pruneThresh and abs(x.pr - y.pr) <= 1e-12 * abs(y.pr)
    for k in ("qLo", "qHi"):
        assert np.array_equal(getattr(beams, k), getattr(beams2, k)), k


def test_submit_wait_equals_blocking_call():
    """hfbgpu_submit / hfbgpu_wait (batches overlapping on two streams) == hfbgpu_accumulate."""
    z, fm, b, kw = load_golden("synth_tied_m4")
    fb = _fb(fm, **kw)
    r0, _ = fb.FBFile(b); r0b, _ = fb.FBFile(b); r0c, _ = fb.FBFile(b)
    a0 = fb.GetAccs()
    fb.ZeroAccs()
    tickets = [fb.Submit(b, want_beams=(i == 1)) for i in range(3)]
    a1 = fb.GetAccs()                      # waits implicitly
    for tk in tickets:
        for x, y in zip(tk.results(), r0):
            assert x.status == y.status and abs(x.pr - y.pr) <= 1e-12 * abs(y.pr)
    e = acc_errors(a1, a0, fm)
    assert max(e.values()) < 1e-5, e
    assert tickets[1].beams.qHi.max() > 0
    fb.close()


# ---- parameter-kind qualifiers on the device ---------------------------------------------------

def _model_of_width(D, M=2):
    from htk_b200 import synth
    from htk_b200.flat import flatten
    return flatten(synth.make_monophone_set(n_phones=6, M=M, D=D, seed=5, spread=0.3))


@pytest.mark.parametrize("name", QUALIFIER_CASES)
def test_qualifier_expansion_bit_identical_to_hcopy(name):
    """feat_regress_kernel / feat_zeromean_kernel against the files the unmodified reference's HCopy wrote
    (tests/golden/make_qualifier_golden.py): every float identical, utterances of 1..333 frames."""
    q, static, expanded = load_qualifier_golden(name)
    fb = _fb(_model_of_width(q.vec_size))
    fb.SetQualifiers(q)
    out = fb.ExpandFeatures(static)
    fb.close()
    for o, y in zip(out, expanded):
        assert o.shape == y.shape
        assert np.array_equal(o.view(np.uint32), y.view(np.uint32))


def test_qualifier_description_is_checked():
    from htk_b200 import capi
    from htk_b200.flat import Qualifiers
    fb = _fb(_model_of_width(39))
    for bad in (Qualifiers(13, 2, 0, 0), Qualifiers(13, 0, 2, 0), Qualifiers(13, 2, 2, 0, zero_mean_cols=14),
                Qualifiers(12, 2, 2, 0)):
        with pytest.raises(capi.HfbError):
            fb.SetQualifiers(bad)
    fb.SetQualifiers(Qualifiers(13, 2, 2, 0, zero_mean_cols=13))
    fb.SetQualifiers(None)
    fb.close()


@pytest.mark.parametrize("device_feat", [False, True])
def test_estep_on_static_features_equals_estep_on_expanded_features(device_feat):
    """With hfbgpu_set_qualifiers the batch holds 13 static columns; accumulators, log-likelihoods and beams must
    equal those of the same utterances expanded beforehand (oracle/hparm_oracle.py, pinned to HCopy), through the
    host-feature and the device-feature entry, and agree with the C oracle on the expanded features."""
    import torch
    from htk_b200 import synth
    from htk_b200.flat import Qualifiers, flatten
    from oracle import hparm_oracle as H
    hs = synth.make_tied_triphone_set(n_states=40, M=4, n_phys=30, n_logical=30, n_centre=5, D=39, seed=12, spread=0.25)
    fm = flatten(hs)
    feats, labs = synth.sample_corpus(fm, n_utts=7, T=180, Q=14, seed=21, T_jitter=30)
    q = Qualifiers.from_kinds("MFCC_0", "MFCC_0_D_A_Z", 13)
    static = [np.ascontiguousarray(f[:, :13]) for f in feats]
    full = [H.expand(x, q.del_win, q.acc_win, q.third_win, q.simple_diffs, q.zero_mean_cols) for x in static]
    kw = dict(prune=None)
    b_full, b_static = Batch(full, labs, 39), Batch(static, labs, 13)
    fb = _fb(fm, **kw)
    res0, beams0 = fb.FBFile(b_full, want_beams=True); a0 = fb.GetAccs()
    fb.ZeroAccs(); fb.SetQualifiers(q)
    if device_feat:
        d = torch.from_numpy(b_static.feat).cuda()
        res1, beams1 = fb.FBFile(b_static, want_beams=True, device_feat_ptr=d.data_ptr())
    else:
        res1, beams1 = fb.FBFile(b_static, want_beams=True)
    a1 = fb.GetAccs()
    st = fb.stats()
    fb.close()
    assert [r.status for r in res0] == [r.status for r in res1] and any(r.status == 0 for r in res0)
    for x, y in zip(res0, res1):
        if x.status == 0:
            assert abs(x.pr - y.pr) <= 1e-9 * abs(x.pr)
    for k in ("qLo", "qHi", "sq", "eq"):
        assert np.array_equal(getattr(beams0, k), getattr(beams1, k)), k
    e = acc_errors(a1, a0, fm)                 # identical features; atomic / FP32 fragment order only
    assert max(e.values()) < 1e-5, e
    ao = _oracle(fm, b_full, kw)[0]
    e = acc_errors(a1, ao, fm)
    assert max(e.values()) < RTOL, e


def test_outlier_frames_get_the_references_values():
    """the FP16 operands of the tensor-core kernel have a finite range.  Frames far outside it
    (a coordinate 300 sigma off, a frame of zeros, a frame 80 sigma off in EVERY dimension so that every component of
    every state is below -50 000) must come out as the reference computes them (IDOutP in float, HModel.c:5420-5431):
    the true, hugely negative, value -- never a clamped one, never log zero -- and the utterance is still processed."""
    from oracle import oracle_lib as O
    z, fm, b, kw = load_golden("synth_tied_m4")
    feat = np.array(z["feat"], dtype=np.float32)
    f0 = int(z["frameOff"][1])                       # corrupt utterance 1
    sd = np.sqrt(1.0 / fm.ivar.astype(np.float64)).mean(0).astype(np.float32)
    feat[f0 + 10, 3] += 300.0 * sd[3]
    feat[f0 + 11, :] = 0.0
    feat[f0 + 12, :] += 80.0 * sd
    feat[f0 + 13, 7] -= 3000.0 * sd[7]               # beyond the FP16 range even after scaling
    bad = [f0 + 10, f0 + 11, f0 + 12, f0 + 13]
    rows = np.r_[f0:f0 + 40]
    states = np.arange(fm.J, dtype=np.int32)
    want = O.state_loglik(fm, feat[rows], states)
    assert want[12].max() < -50000.0                 # the case the old epilogue turned into log zero
    for gk in (2, 1):
        fb = _fb(fm, gmm_kernel=gk, **kw)
        got = fb.OutP(feat[rows], states)
        assert np.all(got > -1e9)
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-5, np.max(np.abs(got - want) / np.abs(want))
        b2 = Batch.from_arrays(feat, z["frameOff"], z["lab"], z["labOff"])
        res, beams = fb.FBFile(b2, want_beams=True)
        acc = fb.GetAccs()
        oacc, ores, obeams = _oracle(fm, b2, kw)
        for r, o in zip(res, ores):
            assert r.status == o[0] == 0
            assert abs(r.pr - o[2]) <= 1e-6 * abs(o[2])
        assert res[1].pr < ores[0][2] - 1e4          # the corrupt frames are in the likelihood
        from htk_b200.compare import acc_errors_ties
        e, ties = acc_errors_ties(acc, oacc, fm)      # a component posterior exactly at the exp(-minFrwdP) cut may flip
        assert max(e.values()) < RTOL and ties <= 2, (e, ties)
        assert np.array_equal(beams.sq, obeams.sq) and np.array_equal(beams.qLo, obeams.qLo)
        fb.close()


@pytest.mark.parametrize("name", ["synth_tied_m4", "synth_long_m3", "synth_tee_m2", "synth_mono_m1", "htkdemo_t20_15_200"])
def test_taper_skipping_changes_nothing(name, monkeypatch):
    """gmm_tc3_kernel leaves out the (tile, frame block) combinations outside the beam taper, as the reference's
    Setotprob does (HFB.c:1014-1016).  Everything downstream must be identical to evaluating every frame x every
    state, and to the round-1 kernel: same likelihoods, thresholds and beams, accumulators equal up to the order of the
    FP64 atomics."""
    z, fm, b, kw = load_golden(name)
    outs = []
    for env in (None, "HFBGPU_NO_TAPER_SKIP", "HFBGPU_GMM_V2", "HFBGPU_GMM_V3"):
        if env:
            monkeypatch.setenv(env, "1")
        fb = _fb(fm, **kw)
        res, beams = fb.FBFile(b, want_beams=True)
        outs.append((res, beams, fb.GetAccs(), fb.stats().gmmPairs))
        fb.close()
        if env:
            monkeypatch.delenv(env)
    (r0, b0, a0, p0), (r1, b1, a1, p1), (r2, b2, a2, p2), (r3, b3, a3, p3) = outs
    assert [tuple(r) for r in r0] == [tuple(r) for r in r1]             # bit-identical likelihoods
    # A in tensor memory (default) vs A in shared memory: the same operands through the same MMAs in the same order; the
    # epilogues add the 16 exponentials of a state in another order (packed pairs), i.e. log b differs in the last bits
    assert p0 == p3
    for x, y in zip(r0, r3):
        assert (x.status, x.retries, x.pruneThresh) == (y.status, y.retries, y.pruneThresh)
        assert abs(x.pr - y.pr) <= 1e-7 * abs(y.pr)
    for k in ("qLo", "qHi", "sq", "eq"):
        assert np.array_equal(getattr(b0, k), getattr(b3, k))
    e = acc_errors(a0, a3, fm)
    assert max(e.values()) < 2e-5, e
    for k in ("qLo", "qHi", "sq", "eq"):
        assert np.array_equal(getattr(b0, k), getattr(b1, k)) and np.array_equal(getattr(b0, k), getattr(b2, k))
    # same b inside the beams -> same posteriors; the statistics kernels sum FP32 fragments in an order that depends on
    # the atomic cursor of the frame lists, so two runs agree to FP32 summation noise, not bitwise
    e = acc_errors(a0, a1, fm)
    assert max(e.values()) < 1e-5, e
    e = acc_errors(a0, a2, fm)                                           # other kernel, other rounding of log b: ~4e-6 each
    assert max(e.values()) < 5e-5, e
    assert p0 <= p1                                                       # fewer (frame, state) pairs evaluated
    if name in ("synth_tied_m4", "synth_long_m3"):
        assert p0 < 0.9 * p1, (p0, p1)


def test_ring_window_beta_equals_sliding_block_kernel(monkeypatch):
    """Transcriptions of more than 128 labels under a beam: the warp kernel with a 128-model ring window (default)
    against the 256-thread sliding block kernel (HFBGPU_NO_RING) and the one-thread-per-label kernel (HFBGPU_NO_SLIDE):
    same recursion, same order of operations -- identical likelihoods, thresholds and beams."""
    z, fm, b, kw = load_golden("synth_long_m3")                      # Q = 160, -t 250 150 1000
    outs = []
    for env in (None, "HFBGPU_NO_RING", "HFBGPU_NO_SLIDE"):
        if env:
            monkeypatch.setenv("HFBGPU_NO_RING", "1")
            monkeypatch.setenv(env, "1")
        fb = _fb(fm, **kw)
        res, beams = fb.FBFile(b, want_beams=True)
        outs.append((res, beams, fb.GetAccs()))
        fb.close()
    for res, beams, acc in outs[1:]:
        assert [tuple(r) for r in res] == [tuple(r) for r in outs[0][0]]
        for k in ("qLo", "qHi", "sq", "eq"):
            assert np.array_equal(getattr(beams, k), getattr(outs[0][1], k))
        e = acc_errors(acc, outs[0][2], fm)
        assert max(e.values()) < 1e-5, e
    # a tight beam on the same data: retries (the window is placed at the top again), narrower windows
    kw2 = dict(kw); kw2["prune"] = (60.0, 40.0, 400.0)
    outs = []
    for env in (None, "HFBGPU_NO_RING"):
        if env:
            monkeypatch.setenv(env, "1")
        else:
            monkeypatch.delenv("HFBGPU_NO_RING", raising=False); monkeypatch.delenv("HFBGPU_NO_SLIDE", raising=False)
        fb = _fb(fm, **kw2)
        res, beams = fb.FBFile(b, want_beams=True)
        outs.append((res, beams))
        fb.close()
    assert [tuple(r) for r in outs[0][0]] == [tuple(r) for r in outs[1][0]]
    assert np.array_equal(outs[0][1].qLo, outs[1][1].qLo) and np.array_equal(outs[0][1].qHi, outs[1][1].qHi)
    oacc, ores, obeams = _oracle(fm, b, kw2)
    assert [(r.status, r.retries, r.pruneThresh) for r in outs[0][0]] == [(o[0], o[1], o[3]) for o in ores]
    assert np.array_equal(outs[0][1].qLo, obeams.qLo) and np.array_equal(outs[0][1].qHi, obeams.qHi)


@pytest.mark.parametrize("name", ["synth_tied_m4", "synth_long_m3", "synth_tee_m2"])
def test_tcgen05_statistics_equal_mma_sync_statistics(name, monkeypatch):
    """K4 on tcgen05 (hfb_stats_tc.cuh: component posteriors and occupancy-weighted sums as two UMMA contractions per
    tied state, the default: rows copied from K1's expanded operand into the compact, unswizzled tile) against its
    other forms -- 128-byte-swizzled tile (HFBGPU_ST_SWZ), rows expanded by the kernel itself (HFBGPU_NO_EXPA) --,
    stats5_kernel (FP32 posteriors + mma.sync sums, HFBGPU_STATS5) and the per-position FP32 kernel (HFBGPU_STATS3);
    all against the oracle elsewhere."""
    z, fm, b, kw = load_golden(name)
    outs = []
    for env in (None, "HFBGPU_ST_SWZ", "HFBGPU_NO_EXPA", "HFBGPU_STATS5", "HFBGPU_STATS3"):
        if env:
            monkeypatch.setenv(env, "1")
        fb = _fb(fm, **kw)
        fb.FBFile(b)
        outs.append(fb.GetAccs())
        fb.close()
        if env:
            monkeypatch.delenv(env)
    for other in outs[1:]:
        e = acc_errors(outs[0], other, fm)
        assert max(e.values()) < 5e-5, e
    oacc, _, _ = _oracle(fm, b, kw)
    e = acc_errors(outs[0], oacc, fm)
    assert max(e.values()) < RTOL, e


def test_tcgen05_statistics_leave_overflowing_waves_to_stats5(monkeypatch):
    """When the frame lists of stats_pre_kernel overflow (HFBGPU_STATS_PRE_CAP), the tcgen05 statistics kernel steps
    aside for the whole wave and stats5_kernel (with its inline front for the positions without a list) does it."""
    z, fm, b, kw = load_golden("synth_tied_m4")
    fb = _fb(fm, **kw); fb.FBFile(b); a1 = fb.GetAccs(); fb.close()
    monkeypatch.setenv("HFBGPU_STATS_PRE_CAP", "700")
    fb = _fb(fm, **kw); fb.FBFile(b); a2 = fb.GetAccs(); fb.close()
    e = acc_errors(a1, a2, fm)
    assert max(e.values()) < 5e-5, e
    oacc, _, _ = _oracle(fm, b, kw)
    e = acc_errors(a2, oacc, fm)
    assert max(e.values()) < RTOL, e


# ------------------------------------------------------------------------------------------
# two-model re-estimation (UseAlignHMMSet, HFB.c:296-333; UpMixParms :1518-1547)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("gmm_kernel", [1, 2])
@pytest.mark.parametrize("name", ["two_model_tied", "two_model_mono", "two_model_complevel"])
def test_two_model_reestimation_matches_stock_herest(name, gmm_kernel):
    """hfb_options.alignModel: aligned with one set, statistics of another -- against the dump the stock HERest wrote
    with ALIGNMODELMMF / ALIGNHMMLIST (tests/golden/make_two_model_golden.py) and against the oracle."""
    from conftest import load_two_model_golden
    z, fu, fa, b, kw = load_two_model_golden(name)
    fb = _fb(fu, gmm_kernel=gmm_kernel, **kw)
    res, beams = fb.FBFile(b, want_beams=True)
    acc = fb.GetAccs()
    L = fu.layout
    for r, t, rpf in zip(res, np.diff(z["frameOff"]), z["ref_pr_per_frame"]):
        assert r.status == 0
        assert abs(r.pr / t - rpf) <= RTOL * abs(rpf) + 1e-6
    ref = z["ref_acc"]
    assert np.all(acc[L.tran:L.wtC] == 0)                         # HFB.c:313-316: UPTRANS dropped
    assert np.array_equal(acc[L.numEgs:L.totalT], ref[L.numEgs:L.totalT])      # up_hmm->hook
    assert acc[L.totalT] == ref[L.totalT] and acc[L.numOk] == b.numUtt
    e = acc_errors(acc, ref, fu)
    assert max(e.values()) < RTOL, e
    oacc, ores, obeams = _oracle(fu, b, kw)
    e = acc_errors(acc, oacc, fu)
    assert max(e.values()) < RTOL, e
    for k in ("qLo", "qHi", "sq", "eq"):
        assert np.array_equal(getattr(beams, k), getattr(obeams, k)), k
    # twice the batch = twice the accumulators; zeroing clears them (and the numEgs tail of the alignment kernels)
    fb.FBFile(b)
    acc2 = fb.GetAccs()
    assert np.allclose(acc2, 2 * acc, rtol=1e-9, atol=1e-9)
    fb.ZeroAccs()
    assert not fb.GetAccs().any()
    # M-step = the update set re-estimated (the device M-step against the oracle-fed one)
    fb.FBFile(b)
    new, info = fb.MStep(min_egs=1)
    fb.close()
    if name == "two_model_complevel":
        # HFB: ALIGNCOMPLEVEL = T needs the same number of components in both sets: refused before anything is launched
        from htk_b200 import capi
        z2, fu2, fa2, b2, kw2 = load_two_model_golden("two_model_tied")
        fb2 = _fb(fu2, align_comp_level=True, **kw2)
        with pytest.raises(capi.HfbError):
            fb2.FBFile(b2)
        assert not fb2.GetAccs().any()
        fb2.close()
    assert new.mean.shape == fu.mean.shape and np.isfinite(new.mean).all()
    occ = acc[L.muOcc:L.vaSum]
    g = int(np.argmax(occ[fu.meanId]))
    want = fu.mean[g] + acc[L.muSum + fu.meanId[g] * fu.D: L.muSum + (fu.meanId[g] + 1) * fu.D] / occ[fu.meanId[g]]
    assert np.allclose(new.mean[g], want, rtol=1e-5, atol=1e-5)             # HERest.c:974-1012


def test_two_model_rejects_bad_input():
    from conftest import load_two_model_golden
    from htk_b200 import capi
    z, fu, fa, b, kw = load_two_model_golden("two_model_mono")
    fb = _fb(fu, **kw)
    plain = Batch.from_arrays(z["feat"], z["frameOff"], z["lab"], z["labOff"])          # no alignment labels
    with pytest.raises(capi.HfbError):
        fb.FBFile(plain)
    assert not fb.GetAccs().any()
    fb.close()
