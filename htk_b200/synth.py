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


def _make_state(rng, D, M, centre, mix_spread, name=None) -> State:
    if M == 1:
        w = np.ones(1)
    else:
        w = rng.dirichlet(np.ones(M))
        w = np.maximum(w, 1e-3)
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
    pool = [_make_state(rng, D, M, spread * rng.standard_normal(D), mix_spread, "ST_%d" % j)
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
