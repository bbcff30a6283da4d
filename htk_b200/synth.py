"""Seeded synthetic model sets and utterances for the BASELINE.json configs
(SURVEY.md 8d).  There is no network, so every workload except HTKDemo is synthetic:
39-dim "MFCC_0_D_A"-shaped features, left-to-right 3-emitting-state HMMs, random-init
diagonal GMMs whose state distributions OVERLAP (means N(0, spread^2), spread << 1 sigma
apart would be separable and every occupancy 0/1 -- see SURVEY.md 8a "lessons").
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .htkio import HMM, Gaussian, HMMSetDef, State, TransMat

LZERO = -1.0e10


def _lr_trans(n_emit: int, self_loop: float, skip: float = 0.0, tee: float = 0.0) -> np.ndarray:
    """Left-to-right transition matrix with optional entry-skip and tee (1->N) arcs."""
    N = n_emit + 2
    A = np.zeros((N, N), dtype=np.float64)
    A[0, 1] = 1.0 - skip - tee
    if skip > 0 and n_emit > 1:
        A[0, 2] = skip
    elif skip > 0:
        A[0, 1] += skip
    if tee > 0:
        A[0, N - 1] = tee
    for i in range(1, N - 1):
        A[i, i] = self_loop
        A[i, i + 1] = 1.0 - self_loop
    return A


def _make_state(rng, D, M, centre, mix_spread, name=None, dead_component: bool = False) -> State:
    if M == 1:
        w = np.ones(1)
    else:
        w = rng.dirichlet(np.ones(M))
        w = np.maximum(w, 1e-3)
        if dead_component:
            w[M - 1] = 5.0e-6         # below MINMIX = 1e-5: HTK skips the component everywhere (HFB.c:953, :1573)
        w = w / w.sum()
    mixes = []
    for m in range(M):
        mean = (centre + mix_spread * rng.standard_normal(D)).astype(np.float32)
        var = rng.uniform(0.6, 1.6, D).astype(np.float32)
        mixes.append((float(w[m]), Gaussian(mean, var)))
    return State(mixes, name)


def make_monophone_set(n_phones: int = 40, n_emit: int = 3, M: int = 1, D: int = 39, seed: int = 1234,
                       spread: float = 0.35, mix_spread: float = 0.25, self_loop: float = 0.6,
                       tee_model: bool = False, entry_skip: float = 0.0,
                       parm_kind: str = "MFCC_0_D_A") -> HMMSetDef:
    """Config #2 style: plain monophones.  With ``tee_model`` an extra 1-emitting-state
    'sp' model with a 1->N arc is added (exercises every tee term of HFB.c)."""
    rng = np.random.default_rng(seed)
    hs = HMMSetDef(D, parm_kind)
    for p in range(n_phones):
        states = [_make_state(rng, D, M, spread * rng.standard_normal(D), mix_spread) for _ in range(n_emit)]
        sl = float(np.clip(self_loop + 0.1 * rng.standard_normal(), 0.3, 0.85))
        hs.hmms.append(HMM("p%02d" % p, states, TransMat(_lr_trans(n_emit, sl, entry_skip))))
    if tee_model:
        st = [_make_state(rng, D, M, spread * rng.standard_normal(D), mix_spread)]
        hs.hmms.append(HMM("sp", st, TransMat(_lr_trans(1, 0.5, 0.0, tee=0.4))))
    return hs


def make_tied_triphone_set(n_states: int = 5000, M: int = 16, n_phys: int = 8000, n_logical: int = 20000,
                           n_centre: int = 40, D: int = 39, seed: int = 1234, spread: float = 0.35,
                           mix_spread: float = 0.25, self_loop: float = 0.6,
                           parm_kind: str = "MFCC_0_D_A") -> HMMSetDef:
    """Configs #3-#5 style: ``n_phys`` physical triphones whose 3 states index a pool of
    ``n_states`` tied states (~s), transition matrices tied per centre phone (~t), and
    ``n_logical`` logical names mapped onto the physical ones."""
    rng = np.random.default_rng(seed)
    hs = HMMSetDef(D, parm_kind)
    # M may be a sequence: per-state component counts, cycled; every 5th multi-component state then also
    # gets one component below the weight floor
    Ms = [M] * n_states if np.isscalar(M) else [int(M[j % len(M)]) for j in range(n_states)]
    pool = [_make_state(rng, D, Ms[j], spread * rng.standard_normal(D), mix_spread, "ST_%d" % j,
                        dead_component=(not np.isscalar(M) and Ms[j] > 1 and j % 5 == 0))
            for j in range(n_states)]
    tms = [TransMat(_lr_trans(3, float(np.clip(self_loop + 0.1 * rng.standard_normal(), 0.3, 0.85))),
                    "T_c%d" % c) for c in range(n_centre)]
    # make sure every tied state is used at least once when there is room
    picks = rng.integers(0, n_states, size=(n_phys, 3))
    flat_cover = rng.permutation(n_states)[: min(n_states, n_phys * 3)]
    picks.reshape(-1)[: len(flat_cover)] = flat_cover
    for p in range(n_phys):
        c = p % n_centre
        name = "l%d-c%d+r%d" % (p // n_centre, c, p)
        hs.hmms.append(HMM(name, [pool[int(j)] for j in picks[p]], tms[c]))
    for l in range(max(0, n_logical - n_phys)):
        hs.logical.append(("x%d-y+z" % l, hs.hmms[int(rng.integers(0, n_phys))].name))
    return hs


def sample_utterance(flat, lab: Sequence[int], T: int, rng) -> np.ndarray:
    """Frames drawn from the labelled state sequence: random segment durations (>=1 frame
    per emitting state), mixture picked by weight, x = mean + sigma * N(0,1)."""
    pos_states = np.concatenate([flat.hmmState[flat.hmmStateOff[p]:flat.hmmStateOff[p + 1]] for p in lab])
    n = len(pos_states)
    if n > T:
        raise ValueError("utterance too short for its transcription")
    cuts = np.sort(rng.choice(np.arange(1, T), size=n - 1, replace=False)) if n > 1 else np.zeros(0, int)
    bounds = np.concatenate([[0], cuts, [T]])
    durs = np.diff(bounds)
    st_of_t = np.repeat(pos_states, durs)
    mo = flat.stateMixOff[st_of_t]
    Mn = flat.stateMixOff[st_of_t + 1] - mo
    w = np.exp(np.minimum(flat.mixLogWt.astype(np.float64), 0.0))
    u = rng.random(T)
    comp = np.zeros(T, dtype=np.int64)
    for t in range(T):
        ww = w[mo[t]:mo[t] + Mn[t]]
        comp[t] = mo[t] + min(int(np.searchsorted(np.cumsum(ww) / ww.sum(), u[t])), Mn[t] - 1)
    g = flat.mixGauss[comp]
    x = flat.mean[g] + rng.standard_normal((T, flat.D)).astype(np.float32) / np.sqrt(flat.ivar[g])
    return x.astype(np.float32)


def sample_corpus(flat, n_utts: int, T: int, Q: int, seed: int = 99, tee_index: Optional[int] = None,
                  T_jitter: int = 0) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """``n_utts`` utterances of ~T frames with Q labels drawn uniformly from the physical
    HMMs.  ``tee_index``: physical index of a tee model inserted between some labels
    (never first, last or twice in a row -- HFB.c:557-565)."""
    rng = np.random.default_rng(seed)
    feats, labs = [], []
    cand = [p for p in range(flat.P) if p != tee_index]
    for _ in range(n_utts):
        base = rng.choice(cand, size=Q)
        lab: List[int] = []
        for i, p in enumerate(base):
            lab.append(int(p))
            if tee_index is not None and i < Q - 1 and rng.random() < 0.3:
                lab.append(int(tee_index))
        Tu = T + (int(rng.integers(-T_jitter, T_jitter + 1)) if T_jitter else 0)
        lab_arr = np.asarray(lab, dtype=np.int32)
        # tee models may be skipped entirely: sample their frames anyway (>=1) for simplicity
        feats.append(sample_utterance(flat, lab_arr, Tu, rng))
        labs.append(lab_arr)
    return feats, labs


# --------------------------------------------------------------------------- large sets

def make_flat_tied(n_states: int = 5000, M: int = 16, n_phys: int = 8000, n_centre: int = 40, D: int = 39,
                   seed: int = 1234, spread: float = 0.2, mix_spread: float = 0.25, self_loop: float = 0.6):
    """Same construction as make_tied_triphone_set but straight into flat arrays (no Python
    object per Gaussian), for the BASELINE-size sets: G = n_states * M Gaussians."""
    from .flat import flat_from_arrays
    rng = np.random.default_rng(seed)
    G = n_states * M
    centre = (spread * rng.standard_normal((n_states, 1, D))).astype(np.float32)
    mean = (centre + mix_spread * rng.standard_normal((n_states, M, D)).astype(np.float32)).reshape(G, D)
    var = rng.uniform(0.6, 1.6, (G, D)).astype(np.float32)
    ivar = (np.float32(1.0) / var).astype(np.float32)
    gconst = (np.float32(D * np.log(2 * np.pi)) + np.log(var.astype(np.float64)).sum(1)).astype(np.float32)
    if M > 1:
        w = rng.dirichlet(np.ones(M), size=n_states)
        w = np.maximum(w, 1e-3)
        w = w / w.sum(1, keepdims=True)
    else:
        w = np.ones((n_states, 1))
    picks = rng.integers(0, n_states, size=(n_phys, 3))
    cover = rng.permutation(n_states)[: min(n_states, n_phys * 3)]
    picks.reshape(-1)[: len(cover)] = cover
    N = 5
    A = np.full((n_centre, N, N), -1.0e10, dtype=np.float32)
    for c in range(n_centre):
        sl = float(np.clip(self_loop + 0.1 * rng.standard_normal(), 0.3, 0.85))
        A[c, 0, 1] = 0.0
        for i in range(1, N - 1):
            A[c, i, i] = np.float32(np.log(sl))
            A[c, i, i + 1] = np.float32(np.log(1.0 - sl))
    fm = flat_from_arrays(
        D=D, mean=mean, ivar=ivar, gConst=gconst, meanId=np.arange(G), varId=np.arange(G),
        stateMixOff=np.arange(n_states + 1) * M, mixGauss=np.arange(G),
        mixLogWt=np.log(w).astype(np.float32).reshape(-1),
        hmmNumStates=np.full(n_phys, N), hmmStateOff=np.arange(n_phys + 1) * 3, hmmState=picks.reshape(-1),
        hmmTrans=np.arange(n_phys) % n_centre, transN=np.full(n_centre, N),
        transOff=np.arange(n_centre + 1) * N * N, transLogA=A.reshape(-1),
        names=["l%d-c%d+r%d" % (p // n_centre, p % n_centre, p) for p in range(n_phys)])
    fm.var = var
    return fm


def make_flat_mono(n_phones: int = 40, M: int = 1, D: int = 39, seed: int = 1234, spread: float = 0.2,
                   mix_spread: float = 0.25):
    """Config #2: plain monophones, 3 emitting states, M Gaussians per state."""
    fm = make_flat_tied(n_states=3 * n_phones, M=M, n_phys=n_phones, n_centre=n_phones, D=D, seed=seed,
                        spread=spread, mix_spread=mix_spread)
    fm.hmmState[:] = np.arange(3 * n_phones)
    fm.names = ["p%02d" % p for p in range(n_phones)]
    return fm


def write_flat_as_mmf(path: str, list_path: str, fm, parm_kind: str = "MFCC_0_D_A") -> None:
    """Text MMF + HMM list for a make_flat_* set (5-state models, tied states / transitions),
    so the unmodified reference HERest can load the very same model."""
    D = fm.D
    var = getattr(fm, "var", None)
    if var is None:
        var = (1.0 / fm.ivar.astype(np.float64)).astype(np.float32)
    with open(path, "w") as f:
        f.write("~o\n<STREAMINFO> 1 %d\n<VECSIZE> %d<NULLD><%s><DIAGC>\n" % (D, D, parm_kind))
        for c in range(fm.numTrans):
            N = int(fm.transN[c])
            A = fm.transLogA[fm.transOff[c]:fm.transOff[c + 1]].reshape(N, N).astype(np.float64)
            P = np.where(A > -0.5e10, np.exp(A), 0.0)
            f.write('~t "T_%d"\n<TRANSP> %d\n' % (c, N))
            for i in range(N):
                f.write(" " + " ".join("%.6e" % x for x in P[i]) + "\n")
        for s in range(fm.J):
            o, e = int(fm.stateMixOff[s]), int(fm.stateMixOff[s + 1])
            f.write('~s "ST_%d"\n' % s)
            if e - o > 1:
                f.write("<NUMMIXES> %d\n" % (e - o))
            for m in range(o, e):
                g = int(fm.mixGauss[m])
                if e - o > 1:
                    f.write("<MIXTURE> %d %.6e\n" % (m - o + 1, float(np.exp(fm.mixLogWt[m]))))
                f.write("<MEAN> %d\n " % D + " ".join("%.6e" % x for x in fm.mean[g]) + "\n")
                f.write("<VARIANCE> %d\n " % D + " ".join("%.6e" % x for x in var[g]) + "\n")
        for p in range(fm.P):
            N = int(fm.hmmNumStates[p])
            f.write('~h "%s"\n<BEGINHMM>\n<NUMSTATES> %d\n' % (fm.names[p], N))
            for j in range(N - 2):
                f.write('<STATE> %d\n~s "ST_%d"\n' % (j + 2, int(fm.hmmState[fm.hmmStateOff[p] + j])))
            f.write('~t "T_%d"\n<ENDHMM>\n' % int(fm.hmmTrans[p]))
    with open(list_path, "w") as f:
        f.write("\n".join(fm.names) + "\n")


def corpus_plan(fm, n_utts: int, T: int, Q: int, seed: int):
    """Labels and the per-frame (state, mixture-component) plan of a synthetic corpus:
    returns (lab[n_utts, Q] int32, gauss_of_frame[n_utts*T] int64)."""
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, fm.P, size=(n_utts, Q)).astype(np.int32)
    n = 3 * Q
    states = fm.hmmState.reshape(fm.P, 3)[lab].reshape(n_utts, n)          # [U, 3Q]
    # random composition of T into n positive parts
    cuts = np.sort(rng.random((n_utts, n - 1)), axis=1)
    edges = np.concatenate([np.zeros((n_utts, 1)), cuts, np.ones((n_utts, 1))], axis=1)
    durs = np.diff(np.floor(edges * (T - n)).astype(np.int64), axis=1) + 1
    durs[:, -1] += T - durs.sum(1)
    st_of_t = np.repeat(states.reshape(-1), durs.reshape(-1))
    Mn = int(fm.stateMixOff[1] - fm.stateMixOff[0])
    w = np.exp(fm.mixLogWt.astype(np.float64)).reshape(fm.J, Mn)
    cdf = np.cumsum(w, axis=1)
    u = rng.random(len(st_of_t))
    comp = (u[:, None] > cdf[st_of_t]).sum(1).clip(0, Mn - 1)
    gauss = fm.mixGauss[fm.stateMixOff[st_of_t] + comp]
    return lab, gauss.astype(np.int64)
