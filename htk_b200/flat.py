"""Flat model / batch / accumulator-layout structures of the libhfbgpu C ABI.

``flatten()`` turns an :class:`htk_b200.htkio.HMMSetDef` into the arrays of
``hfb_model`` (include/hfbgpu.h), applying exactly the conversions the reference
applies between an MMF on disk and what the E-step reads:

* transition probabilities -> float logs, ``<= MINLARG`` -> ``LZERO``
  (HTKLib/HModel.c:2031-2036);
* ``gConst = D*log(2*pi) + sum log var`` as a sequential float sum
  (``FixDiagGConst``, HTKLib/HModel.c:5641-5654) unless the MMF gave one (:1720);
* variances -> inverse variances in float (``ConvDiagC``, HTKLib/HUtil.c:413-440);
* mixture weights -> float logs, ``< MINMIX`` -> ``LZERO``
  (``ConvLogWt``/``MixLogWeight``, HTKLib/HUtil.c:474-485, HModel.c:5288-5295).

In production the C bridge (bridge/hfbgpu_bridge.c) reads the same numbers straight
out of HTK's memory; this module is what the tests and bench.py use.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from .htkio import HMMSetDef, LZERO, MINLARG, MINMIX, TPI

MINVAR, MAXVAR = 1e-30, 1e30


class hfb_model(C.Structure):
    _fields_ = [
        ("vecSize", C.c_int32),
        ("numGauss", C.c_int32),
        ("mean", C.c_void_p), ("ivar", C.c_void_p), ("gConst", C.c_void_p),
        ("meanId", C.c_void_p), ("varId", C.c_void_p),
        ("numMeanAcc", C.c_int32), ("numVarAcc", C.c_int32),
        ("numStates", C.c_int32),
        ("stateMixOff", C.c_void_p), ("mixGauss", C.c_void_p), ("mixLogWt", C.c_void_p),
        ("numHmm", C.c_int32),
        ("hmmNumStates", C.c_void_p), ("hmmStateOff", C.c_void_p), ("hmmState", C.c_void_p),
        ("hmmTrans", C.c_void_p),
        ("numTrans", C.c_int32),
        ("transN", C.c_void_p), ("transOff", C.c_void_p), ("transLogA", C.c_void_p),
    ]


class hfb_options(C.Structure):
    _fields_ = [
        ("pruneInit", C.c_double), ("pruneInc", C.c_double), ("pruneLim", C.c_double),
        ("minFrwdP", C.c_float), ("uFlags", C.c_int32), ("device", C.c_int32),
        ("gmmKernel", C.c_int32), ("flags", C.c_int32), ("workspaceBytes", C.c_size_t),
        ("alignModel", C.c_void_p),
    ]


class hfb_mstep_options(C.Structure):
    _fields_ = [("minEgs", C.c_int32), ("mixWeightFloor", C.c_float), ("varFloor", C.c_void_p)]


class hfb_mstep_result(C.Structure):
    _fields_ = [("mean", C.c_void_p), ("var", C.c_void_p), ("gConst", C.c_void_p), ("mixWeight", C.c_void_p),
                ("transP", C.c_void_p), ("nFloorVar", C.c_int32), ("nFloorVarMix", C.c_int32),
                ("nCopied", C.c_int32), ("nNoOcc", C.c_int32)]


class hfb_qualifiers(C.Structure):
    _fields_ = [("numStatic", C.c_int32), ("delWin", C.c_int32), ("accWin", C.c_int32), ("thirdWin", C.c_int32),
                ("simpleDiffs", C.c_int32), ("zeroMeanCols", C.c_int32), ("suppressEnergy", C.c_int32)]


class Qualifiers:
    """What HParm's configuration says about the step from the files' kind to the target kind
    (TARGETKIND, DELTAWINDOW, ACCWINDOW, THIRDWINDOW, SIMPLEDIFFS; HParm.c:838-871)."""

    def __init__(self, num_static: int, del_win: int = 0, acc_win: int = 0, third_win: int = 0,
                 simple_diffs: bool = False, zero_mean_cols: int = 0, suppress_energy: bool = False):
        self.num_static, self.del_win, self.acc_win, self.third_win = num_static, del_win, acc_win, third_win
        self.simple_diffs, self.zero_mean_cols = bool(simple_diffs), zero_mean_cols
        self.suppress_energy = bool(suppress_energy)     # _N: last static column (energy / c0) left out of the observation

    @classmethod
    def from_kinds(cls, source_kind: str, target_kind: str, num_static: int, del_win: int = 2, acc_win: int = 2,
                   third_win: int = 2, simple_diffs: bool = False) -> "Qualifiers":
        """source_kind: kind of the files (e.g. "MFCC_0"), target_kind: TARGETKIND (e.g. "MFCC_0_D_A_Z");
        num_static: width of the files' vectors."""
        sq, tq = set(source_kind.split("_")[1:]), set(target_kind.split("_")[1:])
        if source_kind.split("_")[0] != target_kind.split("_")[0] or (sq - {"K", "C"}) - tq or sq & {"D", "A", "T", "Z"}:
            raise ValueError("cannot go from %s to %s on the device" % (source_kind, target_kind))
        if (tq - sq) - {"D", "A", "T", "Z", "N"}:
            raise ValueError("only _D _A _T _Z _N can be added on the device (%s -> %s)" % (source_kind, target_kind))
        if "N" in tq and ("D" not in tq or not (tq & {"E", "0"}) or {"E", "0"} <= tq):
            raise ValueError("_N needs _D and exactly one of _E / _0 (HParm.c:1415-1420)")
        zc = 0
        if "Z" in tq:
            zc = num_static - (1 if "E" in tq else 0)       # cepstra and c0, not the energy (HParm.c:1709-1712)
            if "N" in tq:
                zc = min(zc, num_static - 1)                # the suppressed column is never delivered
        return cls(num_static, del_win if "D" in tq else 0, acc_win if "A" in tq else 0,
                   third_win if "T" in tq else 0, simple_diffs, zc, "N" in tq)

    @property
    def vec_size(self) -> int:
        return self.num_static * (1 + (self.del_win > 0) + (self.acc_win > 0) + (self.third_win > 0)) - (1 if self.suppress_energy else 0)

    def c_struct(self) -> "hfb_qualifiers":
        return hfb_qualifiers(self.num_static, self.del_win, self.acc_win, self.third_win,
                              1 if self.simple_diffs else 0, self.zero_mean_cols, 1 if self.suppress_energy else 0)


class hfb_batch(C.Structure):
    _fields_ = [
        ("numUtt", C.c_int32),
        ("frameOff", C.c_void_p), ("feat", C.c_void_p),
        ("labOff", C.c_void_p), ("lab", C.c_void_p), ("labAlign", C.c_void_p),
    ]


class hfb_compressed(C.Structure):
    _fields_ = [("feat", C.c_void_p), ("scaleA", C.c_void_p), ("scaleB", C.c_void_p)]


class hfb_utt_result(C.Structure):
    _fields_ = [("status", C.c_int32), ("retries", C.c_int32),
                ("pr", C.c_double), ("pruneThresh", C.c_double)]


class hfb_beams(C.Structure):
    _fields_ = [("qLo", C.c_void_p), ("qHi", C.c_void_p), ("sq", C.c_void_p), ("eq", C.c_void_p)]


class hfb_acc_layout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "tran", "tranOcc", "wtC", "wtOcc", "muSum", "muOcc", "vaSum", "vaOcc", "numEgs",
        "totalT", "totalPr", "numOk", "numSkipped", "count", "tranOccStride")]


class hfb_stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "launches", "launchesGmm", "launchesBeta", "launchesAlpha", "launchesStats", "launchesMisc")] + \
        [(n, C.c_double) for n in ("msGmm", "msBeta", "msAlpha", "msStats")] + \
        [(n, C.c_int64) for n in ("betaCells", "alphaCells", "gmmPairs", "h2dBytes", "d2hBytes", "launchesL2R")] + \
        [("msExpand", C.c_double)]


NOPRUNE = 1.0e20


def make_options(prune=None, min_frwd_p: float = 10.0, uflags: int = 15, device: int = 0,
                 gmm_kernel: int = 0, workspace_bytes: int = 0, align_model=None, align_comp_level: bool = False) -> hfb_options:
    """``prune`` = None (off, HFB.c:83) or (init, inc, lim) as HERest -t takes them.  ``align_model`` = a FlatModel:
    two-model re-estimation (HFB.c:296-333), the model the options are used with is then the update set."""
    o = hfb_options()
    if align_model is not None:
        o._align_model = align_model                    # keeps the arrays and the struct alive with the options
        o._align_keep = align_model.c_struct()
        o.alignModel = C.addressof(o._align_keep)
    if prune is None:
        o.pruneInit, o.pruneInc, o.pruneLim = NOPRUNE, 0.0, NOPRUNE
    else:
        o.pruneInit, o.pruneInc, o.pruneLim = [float(x) for x in prune]
    o.minFrwdP = min_frwd_p
    o.uFlags = uflags
    o.device = device
    o.gmmKernel = gmm_kernel
    o.flags = 1 if align_comp_level else 0          # HFB_OPT_ALIGN_COMP_LEVEL (HFB: ALIGNCOMPLEVEL = T)
    o.workspaceBytes = workspace_bytes
    return o


@dataclass
class AccLayout:
    tran: int
    tranOcc: int
    wtC: int
    wtOcc: int
    muSum: int
    muOcc: int
    vaSum: int
    vaOcc: int
    numEgs: int
    totalT: int
    totalPr: int
    numOk: int
    numSkipped: int
    count: int


class FlatModel:
    """Owns the numpy arrays behind an ``hfb_model`` and the index maps back to the
    Python model objects (used to decode/encode accumulator dumps)."""

    def __init__(self):
        self.hs: Optional[HMMSetDef] = None

    # -- C view ------------------------------------------------------------
    def c_struct(self) -> hfb_model:
        m = hfb_model()
        m.vecSize = self.D
        m.numGauss = self.G
        for name in ("mean", "ivar", "gConst", "meanId", "varId", "stateMixOff", "mixGauss", "mixLogWt",
                     "hmmNumStates", "hmmStateOff", "hmmState", "hmmTrans", "transN", "transOff", "transLogA"):
            setattr(m, name, getattr(self, name).ctypes.data)
        m.numMeanAcc = self.numMeanAcc
        m.numVarAcc = self.numVarAcc
        m.numStates = self.J
        m.numHmm = self.P
        m.numTrans = self.numTrans
        return m

    def compute_layout(self) -> AccLayout:
        o = 0
        nn = int(np.sum(self.transN.astype(np.int64) ** 2))
        n = int(np.sum(self.transN))
        vals = {}
        for name, size in (("tran", nn), ("tranOcc", n), ("wtC", int(self.stateMixOff[-1])), ("wtOcc", self.J),
                           ("muSum", self.numMeanAcc * self.D), ("muOcc", self.numMeanAcc),
                           ("vaSum", self.numVarAcc * self.D), ("vaOcc", self.numVarAcc),
                           ("numEgs", self.P), ("totalT", 1), ("totalPr", 1), ("numOk", 1), ("numSkipped", 1)):
            vals[name] = o
            o += size
        vals["count"] = o
        return AccLayout(**vals)


def _log_trans(prob: np.ndarray) -> np.ndarray:
    p = np.asarray(prob, dtype=np.float32)      # MMF text -> float (ReadFloat)
    out = np.full(p.shape, LZERO, dtype=np.float32)
    nz = p > MINLARG
    out[nz] = np.log(p[nz].astype(np.float64)).astype(np.float32)
    return out


def _gconst(var: np.ndarray) -> np.float32:
    s = np.float32(len(var) * np.log(TPI))
    for v in var:
        z = np.float32(LZERO) if v <= MINLARG else np.float32(np.log(np.float64(v)))
        s = np.float32(s + z)
    return s


def flatten(hs: HMMSetDef, order: Optional[Sequence[str]] = None) -> FlatModel:
    """Physical HMMs are numbered in ``order`` (default: ``hs.hmms`` order)."""
    fm = FlatModel()
    fm.hs = hs
    D = hs.vec_size
    byname = {h.name: h for h in hs.hmms}
    names = list(order) if order is not None else hs.physical_names()
    fm.names = names
    fm.hmm_index = {n: i for i, n in enumerate(names)}
    fm.state_index: Dict[int, int] = {}
    fm.gauss_index: Dict[int, int] = {}
    fm.mean_index: Dict[int, int] = {}
    fm.var_index: Dict[int, int] = {}
    fm.trans_index: Dict[int, int] = {}
    state_objs, gauss_objs, trans_objs = [], [], []
    means, ivars, gconsts, meanId, varId = [], [], [], [], []
    stateMixOff, mixGauss, mixLogWt = [0], [], []
    hmmN, hmmStateOff, hmmState, hmmTrans = [], [0], [], []
    for n in names:
        h = byname[n]
        hmmN.append(len(h.states) + 2)
        for st in h.states:
            if id(st) not in fm.state_index:
                fm.state_index[id(st)] = len(state_objs)
                state_objs.append(st)
                for w, g in st.mixes:
                    if id(g) not in fm.gauss_index:
                        fm.gauss_index[id(g)] = len(gauss_objs)
                        gauss_objs.append(g)
                        if id(g.mean) not in fm.mean_index:
                            fm.mean_index[id(g.mean)] = len(fm.mean_index)
                        if id(g.var) not in fm.var_index:
                            fm.var_index[id(g.var)] = len(fm.var_index)
                        var = np.asarray(g.var, dtype=np.float32)
                        means.append(np.asarray(g.mean, dtype=np.float32))
                        ivars.append((np.float32(1.0) / np.clip(var, np.float32(MINVAR), np.float32(MAXVAR))).astype(np.float32))
                        gconsts.append(np.float32(g.gconst) if g.gconst is not None else _gconst(var))
                        meanId.append(fm.mean_index[id(g.mean)])
                        varId.append(fm.var_index[id(g.var)])
                    mixGauss.append(fm.gauss_index[id(g)])
                    w32 = np.float32(w)
                    mixLogWt.append(np.float32(LZERO) if w32 < MINMIX else np.float32(np.log(np.float64(w32))))
                stateMixOff.append(len(mixGauss))
            hmmState.append(fm.state_index[id(st)])
        hmmStateOff.append(len(hmmState))
        if id(h.trans) not in fm.trans_index:
            fm.trans_index[id(h.trans)] = len(trans_objs)
            trans_objs.append(h.trans)
        hmmTrans.append(fm.trans_index[id(h.trans)])
    fm.D = D
    fm.G = len(gauss_objs)
    fm.J = len(state_objs)
    fm.P = len(names)
    fm.numTrans = len(trans_objs)
    fm.numMeanAcc = len(fm.mean_index)
    fm.numVarAcc = len(fm.var_index)
    fm.mean = np.ascontiguousarray(np.stack(means), dtype=np.float32)
    fm.ivar = np.ascontiguousarray(np.stack(ivars), dtype=np.float32)
    fm.gConst = np.asarray(gconsts, dtype=np.float32)
    fm.meanId = np.asarray(meanId, dtype=np.int32)
    fm.varId = np.asarray(varId, dtype=np.int32)
    fm.stateMixOff = np.asarray(stateMixOff, dtype=np.int32)
    fm.mixGauss = np.asarray(mixGauss, dtype=np.int32)
    fm.mixLogWt = np.asarray(mixLogWt, dtype=np.float32)
    fm.hmmNumStates = np.asarray(hmmN, dtype=np.int32)
    fm.hmmStateOff = np.asarray(hmmStateOff, dtype=np.int32)
    fm.hmmState = np.asarray(hmmState, dtype=np.int32)
    fm.hmmTrans = np.asarray(hmmTrans, dtype=np.int32)
    fm.transN = np.asarray([t.prob.shape[0] for t in trans_objs], dtype=np.int32)
    offs = np.concatenate([[0], np.cumsum(fm.transN.astype(np.int64) ** 2)]).astype(np.int32)
    fm.transOff = offs
    fm.transLogA = np.concatenate([_log_trans(t.prob).reshape(-1) for t in trans_objs]).astype(np.float32)
    fm.tranAccOff = offs[:-1].astype(np.int64)
    fm.tranOccOff = np.concatenate([[0], np.cumsum(fm.transN)])[:-1].astype(np.int64)
    fm.layout = fm.compute_layout()
    fm.state_objs, fm.gauss_objs, fm.trans_objs = state_objs, gauss_objs, trans_objs
    return fm


def flat_from_arrays(**kw) -> FlatModel:
    """Build a FlatModel directly from arrays (synthetic models too large for the
    object representation).  Required keys = hfb_model array fields + D."""
    fm = FlatModel()
    fm.D = int(kw["D"])
    for k in ("mean", "ivar", "gConst", "mixLogWt", "transLogA"):
        setattr(fm, k, np.ascontiguousarray(kw[k], dtype=np.float32))
    for k in ("meanId", "varId", "stateMixOff", "mixGauss", "hmmNumStates", "hmmStateOff", "hmmState",
              "hmmTrans", "transN", "transOff"):
        setattr(fm, k, np.ascontiguousarray(kw[k], dtype=np.int32))
    fm.G = fm.mean.shape[0]
    fm.J = len(fm.stateMixOff) - 1
    fm.P = len(fm.hmmNumStates)
    fm.numTrans = len(fm.transN)
    fm.numMeanAcc = int(fm.meanId.max()) + 1
    fm.numVarAcc = int(fm.varId.max()) + 1
    fm.tranAccOff = fm.transOff[:-1].astype(np.int64)
    fm.tranOccOff = np.concatenate([[0], np.cumsum(fm.transN)])[:-1].astype(np.int64)
    fm.layout = fm.compute_layout()
    fm.names = kw.get("names")
    return fm


class Batch:
    """Host-side batch of loaded utterances (what HERest's file loop would buffer)."""

    def __init__(self, feats: Sequence[np.ndarray], labs: Sequence[Sequence[int]], D: int):
        self.numUtt = len(feats)
        T = np.array([f.shape[0] for f in feats], dtype=np.int64)
        self.frameOff = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
        self.feat = (np.ascontiguousarray(np.concatenate(feats, axis=0), dtype=np.float32)
                     if len(feats) else np.zeros((0, D), np.float32))
        assert self.feat.shape[1] == D
        Q = np.array([len(l) for l in labs], dtype=np.int64)
        self.labOff = np.concatenate([[0], np.cumsum(Q)]).astype(np.int32)
        self.lab = (np.ascontiguousarray(np.concatenate([np.asarray(l, dtype=np.int32) for l in labs]))
                    if len(labs) else np.zeros(0, np.int32))
        self.totalT = int(self.frameOff[-1])

    @classmethod
    def from_arrays(cls, feat: np.ndarray, frameOff: np.ndarray, lab: np.ndarray, labOff: np.ndarray):
        b = cls.__new__(cls)
        b.numUtt = len(frameOff) - 1
        b.frameOff = np.ascontiguousarray(frameOff, dtype=np.int64)
        b.feat = feat
        b.labOff = np.ascontiguousarray(labOff, dtype=np.int32)
        b.lab = np.ascontiguousarray(lab, dtype=np.int32)
        b.totalT = int(b.frameOff[-1])
        return b

    def with_align_labels(self, lab_align) -> "Batch":
        """Two-model re-estimation: ``lab`` indexes the update set, ``lab_align`` the alignment set (same length)."""
        self.labAlign = np.ascontiguousarray(lab_align, dtype=np.int32)
        assert self.labAlign.shape == self.lab.shape
        return self

    def c_struct(self, feat_ptr: Optional[int] = None) -> hfb_batch:
        b = hfb_batch()
        b.numUtt = self.numUtt
        b.frameOff = self.frameOff.ctypes.data
        b.feat = feat_ptr if feat_ptr is not None else self.feat.ctypes.data
        b.labOff = self.labOff.ctypes.data
        b.lab = self.lab.ctypes.data
        la = getattr(self, "labAlign", None)
        b.labAlign = la.ctypes.data if la is not None else None
        return b


class CompressedFeatures:
    """The feature side of a batch whose files are `_C` compressed (include/hfbgpu.h hfb_compressed): the files' 16-bit
    integers [totalT][cols] in the batch's frame order and the vectors A, B of every file ([numUtt][cols])."""

    def __init__(self, shorts: Sequence[np.ndarray], A: Sequence[np.ndarray], B: Sequence[np.ndarray]):
        self.feat = np.ascontiguousarray(np.concatenate(shorts, axis=0), dtype=np.int16)
        self.scaleA = np.ascontiguousarray(np.stack(A), dtype=np.float32)
        self.scaleB = np.ascontiguousarray(np.stack(B), dtype=np.float32)
        self.frameOff = np.concatenate([[0], np.cumsum([x.shape[0] for x in shorts])]).astype(np.int64)
        assert self.scaleA.shape == self.scaleB.shape == (len(shorts), self.feat.shape[1])

    @classmethod
    def from_arrays(cls, feat, scaleA, scaleB, feat_ptr=None):
        """Arrays already in place (e.g. views of pinned memory); feat_ptr overrides the address of the integers."""
        c = cls.__new__(cls)
        c.feat, c.scaleA, c.scaleB, c._ptr = feat, scaleA, scaleB, feat_ptr
        return c

    def c_struct(self) -> hfb_compressed:
        c = hfb_compressed()
        p = getattr(self, "_ptr", None)
        c.feat = p if p is not None else self.feat.ctypes.data
        c.scaleA = self.scaleA.ctypes.data
        c.scaleB = self.scaleB.ctypes.data
        return c


class Beams:
    def __init__(self, totalT: int):
        self.qLo = np.zeros(totalT, np.int16)
        self.qHi = np.zeros(totalT, np.int16)
        self.sq = np.zeros(totalT, np.int16)
        self.eq = np.zeros(totalT, np.int16)

    def c_struct(self) -> hfb_beams:
        b = hfb_beams()
        b.qLo, b.qHi, b.sq, b.eq = (self.qLo.ctypes.data, self.qHi.ctypes.data,
                                    self.sq.ctypes.data, self.eq.ctypes.data)
        return b
