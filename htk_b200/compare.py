"""Comparison helpers shared by the tests, tools/mgpu_parity.py and bench.py's `allreduce_parity`:
normalised deviations between two flat accumulator buffers and between two sets of re-estimated
parameters.  Pure numpy; nothing here computes the E-step."""
from __future__ import annotations

import os

import numpy as np


def acc_errors(acc, ref, fm):
    """Normalised deviations per accumulator block.

    Occupancy-like blocks: |a-b| / max(|b|, 1e-2).  Centred first/second-order sums are
    near-cancelling (SURVEY.md 8a "do not compare the centred mean sums element-wise"): they
    are compared as the re-estimation formulae use them, mu/occ and var/occ in units of sigma
    (sigma ~ 1 on all fixtures), with the occupancy floored at one frame -- below that the
    reference's own float rounding of log b_j(o_t) (~1e-5 absolute) exceeds the tolerance."""
    L = fm.layout
    D = fm.D
    out = {}

    def rel(a, b, floor=1e-2):
        a = np.asarray(a); b = np.asarray(b)
        return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0

    out["tran"] = rel(acc[L.tran:L.tranOcc], ref[L.tran:L.tranOcc])
    out["tranOcc"] = rel(acc[L.tranOcc:L.wtC], ref[L.tranOcc:L.wtC])
    out["wtC"] = rel(acc[L.wtC:L.wtOcc], ref[L.wtC:L.wtOcc])
    out["wtOcc"] = rel(acc[L.wtOcc:L.muSum], ref[L.wtOcc:L.muSum])
    out["muOcc"] = rel(acc[L.muOcc:L.vaSum], ref[L.muOcc:L.vaSum])
    out["vaOcc"] = rel(acc[L.vaOcc:L.numEgs], ref[L.vaOcc:L.numEgs])
    mocc = np.repeat(np.maximum(ref[L.muOcc:L.vaSum], 1.0), D)
    vocc = np.repeat(np.maximum(ref[L.vaOcc:L.numEgs], 1.0), D)
    out["muSum"] = float(np.max(np.abs(acc[L.muSum:L.muOcc] - ref[L.muSum:L.muOcc]) / mocc)) if mocc.size else 0.0
    out["vaSum"] = float(np.max(np.abs(acc[L.vaSum:L.vaOcc] - ref[L.vaSum:L.vaOcc]) /
                                np.maximum(vocc, np.abs(ref[L.vaSum:L.vaOcc])))) if vocc.size else 0.0
    out["numEgs"] = float(np.max(np.abs(acc[L.numEgs:L.totalT] - ref[L.numEgs:L.totalT]))) if fm.P else 0.0
    out["totalT"] = abs(acc[L.totalT] - ref[L.totalT])
    out["totalPr"] = abs(acc[L.totalPr] - ref[L.totalPr]) / max(abs(ref[L.totalPr]), 1.0)
    return out




def acc_errors_ties(acc, ref, fm, min_frwd_p=10.0, max_flips=3):
    """acc_errors with the minimum-occupancy ties counted instead of charged.

    UpMixParms keeps a component's contribution only if its log posterior exceeds -minFrwdP (HFB.c:1606), i.e. if
    Lr > exp(-minFrwdP) = 4.5e-5.  A posterior that sits exactly at the cut is kept by one evaluation and dropped by
    another that differs in the last float bit of log b_j(o_t) -- the stock tool against its own exact-arithmetic
    restatement included.  One such flip moves a component occupancy by 4.5e-5, which is 4.5e-3 of the 1e-2 floor the
    occupancy blocks are normalised with.  north_star: "the set of pruned frames identical, except for counted ...
    ties": Gaussians whose occupancy differs by no more than `max_flips` such quanta (and fails the plain test) are
    counted and left out of the component-level blocks; everything else is compared as in acc_errors.
    Returns (errors, number of tied Gaussians)."""
    L, D = fm.layout, fm.D
    acc = np.array(acc, np.float64); ref = np.asarray(ref, np.float64)
    quantum = float(np.exp(-min_frwd_p))
    do, ro = acc[L.muOcc:L.vaSum], ref[L.muOcc:L.vaSum]
    d = np.abs(do - ro)
    tied = (d > 1e-4 * np.maximum(np.abs(ro), 1e-2)) & (d <= max_flips * 1.2 * quantum)
    n = int(tied.sum())
    if n:
        gm = np.asarray(fm.meanId)                                  # Gaussian -> mean accumulator
        gt = tied[gm]                                               # per Gaussian
        a2 = acc.copy()
        idx = np.nonzero(tied)[0]
        a2[L.muOcc + idx] = ref[L.muOcc + idx]
        for i in idx:
            a2[L.muSum + i * D:L.muSum + (i + 1) * D] = ref[L.muSum + i * D:L.muSum + (i + 1) * D]
        vi = np.unique(np.asarray(fm.varId)[gt])
        a2[L.vaOcc + vi] = ref[L.vaOcc + vi]
        for i in vi:
            a2[L.vaSum + i * D:L.vaSum + (i + 1) * D] = ref[L.vaSum + i * D:L.vaSum + (i + 1) * D]
        mg = np.asarray(fm.mixGauss)
        wt = np.nonzero(gt[mg])[0]                                  # mixture slots of the tied Gaussians
        a2[L.wtC + wt] = ref[L.wtC + wt]
        st = np.unique(np.searchsorted(np.asarray(fm.stateMixOff), wt, side="right") - 1)
        a2[L.wtOcc + st] = ref[L.wtOcc + st]
        acc = a2
    return acc_errors(acc, ref, fm), n


def reestimated_errors(mean, var, ref_mean, ref_var):
    """north_star: "the re-estimated MMF after one pass within 1e-4 relative on means and variances" --
    means in units of the reference's standard deviation, variances relative."""
    ref_var = np.asarray(ref_var, np.float64)
    return {"mean": float(np.max(np.abs(np.asarray(mean, np.float64) - ref_mean) / np.sqrt(ref_var))),
            "var": float(np.max(np.abs(np.asarray(var, np.float64) - ref_var) / ref_var))}


def load_merge_golden(name="merge_p0_tied_m4"):
    """tests/golden/merge_p0_*.npz (made by tests/golden/make_merge_golden.py from the stock `HERest -p 1..4` +
    `-p 0`): -> (npz, FlatModel, Batch)."""
    from .flat import Batch, flat_from_arrays
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests", "golden", name + ".npz"), allow_pickle=False)
    fm = flat_from_arrays(D=int(z["D"]), **{k[2:]: z[k] for k in z.files if k.startswith("m_")})
    b = Batch.from_arrays(z["feat"], z["frameOff"], z["lab"], z["labOff"])
    return z, fm, b
