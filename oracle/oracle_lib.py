"""ctypes binding of oracle/libhfboracle.so -- TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under htk_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from htk_b200.flat import (Batch, Beams, FlatModel, hfb_batch, hfb_beams, hfb_model, hfb_options,
                           hfb_utt_result)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libhfboracle.so")
    src = os.path.join(_HERE, "hfb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src,
                               "-lm", "-lpthread"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.hfbo_accumulate.argtypes = [C.POINTER(hfb_model), C.POINTER(hfb_options), C.POINTER(hfb_batch),
                                         C.POINTER(hfb_utt_result), C.POINTER(hfb_beams), C.c_void_p,
                                         C.c_int, C.c_int]
        _LIB.hfbo_state_loglik.argtypes = [C.POINTER(hfb_model), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                           C.c_void_p, C.c_void_p]
        _LIB.hfbo_min_durs.argtypes = [C.POINTER(hfb_model), C.c_void_p]
        _LIB.hfbo_utt_occupancy.argtypes = [C.POINTER(hfb_model), C.POINTER(hfb_options), C.c_void_p, C.c_int32,
                                            C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(hfb_utt_result)]
    return _LIB


def accumulate(fm: FlatModel, opt: hfb_options, batch: Batch, acc_double: bool = False, threads: int = 1,
               want_beams: bool = True):
    """Returns (acc float64[count], results list, Beams)."""
    m = fm.c_struct()
    b = batch.c_struct()
    res = (hfb_utt_result * max(1, batch.numUtt))()
    beams = Beams(batch.totalT) if want_beams else None
    bs = beams.c_struct() if beams is not None else None
    acc = np.zeros(fm.layout.count, dtype=np.float64 if acc_double else np.float32)
    rc = lib().hfbo_accumulate(C.byref(m), C.byref(opt), C.byref(b), res,
                               C.byref(bs) if bs is not None else None,
                               acc.ctypes.data, 1 if acc_double else 0, threads)
    if rc != 0:
        raise RuntimeError("hfbo_accumulate failed: %d" % rc)
    return acc.astype(np.float64), [(r.status, r.retries, r.pr, r.pruneThresh) for r in res[:batch.numUtt]], beams


def state_loglik(fm: FlatModel, feat: np.ndarray, states: np.ndarray, want_mix: bool = False):
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    states = np.ascontiguousarray(states, dtype=np.int32)
    T, n = feat.shape[0], len(states)
    out = np.zeros((T, n), np.float32)
    sumM = int(np.sum(fm.stateMixOff[states + 1] - fm.stateMixOff[states]))
    mix = np.zeros((T, sumM), np.float32) if want_mix else None
    m = fm.c_struct()
    lib().hfbo_state_loglik(C.byref(m), feat.ctypes.data, T, states.ctypes.data, n, out.ctypes.data,
                            mix.ctypes.data if mix is not None else None)
    return (out, mix) if want_mix else out


def min_durs(fm: FlatModel) -> np.ndarray:
    out = np.zeros(fm.numTrans, np.int32)
    m = fm.c_struct()
    lib().hfbo_min_durs(C.byref(m), out.ctypes.data)
    return out


def utt_occupancy(fm: FlatModel, opt: hfb_options, feat: np.ndarray, lab: np.ndarray):
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    lab = np.ascontiguousarray(lab, dtype=np.int32)
    P = int(np.sum(fm.hmmNumStates[lab] - 2))
    occ = np.zeros((feat.shape[0], P), np.float32)
    r = hfb_utt_result()
    m = fm.c_struct()
    lib().hfbo_utt_occupancy(C.byref(m), C.byref(opt), feat.ctypes.data, feat.shape[0], lab.ctypes.data,
                             len(lab), occ.ctypes.data, C.byref(r))
    return occ, (r.status, r.retries, r.pr, r.pruneThresh)
