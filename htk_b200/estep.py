"""Host-side mirror of the reference's E-step interface, on top of the libhfbgpu C ABI.

Names follow HTKLib/HFB.h so that the parity tests read like the reference's call sites:

    fb = ForwardBackward(flat_model, prune=(250, 150, 1000))   # InitialiseForBack, HFB.h:117
    results, beams = fb.FBFile(batch)                          # FBFile per utterance, HFB.h:143
    acc = fb.GetAccs()                                         # TrAcc/WtAcc/MuAcc/VaAcc contents

The compute is the CUDA library; this module only marshals numpy arrays / device pointers.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from . import capi
from .flat import (Batch, Beams, CompressedFeatures, FlatModel, hfb_mstep_options, hfb_mstep_result, hfb_stats, hfb_utt_result,
                   make_options)


class UttResult(tuple):
    """(status, retries, pr, pruneThresh)"""
    status = property(lambda s: s[0])
    retries = property(lambda s: s[1])
    pr = property(lambda s: s[2])
    pruneThresh = property(lambda s: s[3])


class _Ticket:
    """Keeps the buffers of an in-flight batch alive."""

    def __init__(self, batch, cb, res, beams, bs):
        self.batch, self._cb, self._res, self.beams, self._bs = batch, cb, res, beams, bs

    def results(self) -> List[UttResult]:
        return [UttResult((r.status, r.retries, r.pr, r.pruneThresh)) for r in self._res[:self.batch.numUtt]]


class ForwardBackward:
    def __init__(self, fm: FlatModel, prune=None, min_frwd_p: float = 10.0, uflags: int = 15,
                 device: int = 0, gmm_kernel: int = 0, workspace_bytes: int = 0, devices=None, align_model=None,
                 align_comp_level: bool = False):
        """`devices`: list of CUDA ordinals -> one context driving all of them (hfbgpu_create_multi).
        `align_model`: a FlatModel = two-model re-estimation (UseAlignHMMSet, HFB.c:296-333): `fm` is then the update set
        and every batch carries both label arrays (Batch.with_align_labels)."""
        self.lib = capi.load()
        self.fm = fm
        self.opt = make_options(prune, min_frwd_p, uflags, device if not devices else devices[0], gmm_kernel, workspace_bytes,
                                align_model=align_model, align_comp_level=align_comp_level)
        self._m = fm.c_struct()
        h = C.c_void_p()
        if devices:
            arr = (C.c_int32 * len(devices))(*devices)
            rc = self.lib.hfbgpu_create_multi(C.byref(h), C.byref(self._m), C.byref(self.opt), arr, len(devices))
        else:
            rc = self.lib.hfbgpu_create(C.byref(h), C.byref(self._m), C.byref(self.opt))
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_create")
        self.h = h
        self.layout = fm.layout
        self._inflight = []
        self.qualifiers = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.hfbgpu_wait(self.h)
            self._inflight = []
            self.lib.hfbgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the FBFile loop over a batch ------------------------------------------------------
    def FBFile(self, batch: Batch, want_beams: bool = False, device_feat_ptr: Optional[int] = None
               ) -> Tuple[List[UttResult], Optional[Beams]]:
        res = (hfb_utt_result * max(1, batch.numUtt))()
        beams = Beams(batch.totalT) if want_beams else None
        bs = beams.c_struct() if beams is not None else None
        b = batch.c_struct(device_feat_ptr)
        fn = self.lib.hfbgpu_accumulate if device_feat_ptr is None else self.lib.hfbgpu_accumulate_device
        rc = fn(self.h, C.byref(b), res, C.byref(bs) if bs is not None else None)
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_accumulate")
        return [UttResult((r.status, r.retries, r.pr, r.pruneThresh)) for r in res[:batch.numUtt]], beams

    def FBFileCompressed(self, batch: Batch, cf: CompressedFeatures, want_beams: bool = False):
        """FBFile over utterances whose parameter files are `_C` compressed: the files' 16-bit integers and A / B vectors go
        to the device as they are and are decoded there exactly as HParm does (HParm.c:3489-3494); batch.feat is ignored."""
        res = (hfb_utt_result * max(1, batch.numUtt))()
        beams = Beams(batch.totalT) if want_beams else None
        bs = beams.c_struct() if beams is not None else None
        b, c = batch.c_struct(), cf.c_struct()
        rc = self.lib.hfbgpu_accumulate_compressed(self.h, C.byref(b), C.byref(c), res, C.byref(bs) if bs is not None else None)
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_accumulate_compressed")
        return [UttResult((r.status, r.retries, r.pr, r.pruneThresh)) for r in res[:batch.numUtt]], beams

    def SubmitCompressed(self, batch: Batch, cf: CompressedFeatures):
        """Asynchronous form of FBFileCompressed (hfbgpu_submit_compressed); ticket as for Submit."""
        res = (hfb_utt_result * max(1, batch.numUtt))()
        b, c = batch.c_struct(), cf.c_struct()
        rc = self.lib.hfbgpu_submit_compressed(self.h, C.byref(b), C.byref(c), res, None)
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_submit_compressed")
        tk = _Ticket(batch, (b, c, cf), res, None, None)
        self._inflight.append(tk)
        return tk

    def DecompressFeatures(self, cf: CompressedFeatures) -> np.ndarray:
        """The decoding alone: [totalT][cols] floats, what ReadAsTable yields for those files."""
        cols = cf.feat.shape[1]
        dst = np.zeros(cf.feat.shape, np.float32)
        c = cf.c_struct()
        rc = self.lib.hfbgpu_decompress_features(self.h, C.byref(c), cf.frameOff.ctypes.data, len(cf.frameOff) - 1, cols,
                                                 dst.ctypes.data)
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_decompress_features")
        return dst

    def FBFileRetrain(self, batch: Batch, feat2: np.ndarray, want_beams: bool = False):
        """HERest -r (single-pass retraining): alignment on batch.feat, mean / variance statistics on feat2
        ([totalT][D], the second parameterisation of the same frames)."""
        feat2 = np.ascontiguousarray(feat2, dtype=np.float32)
        assert feat2.shape == (batch.totalT, self.fm.D)
        res = (hfb_utt_result * max(1, batch.numUtt))()
        beams = Beams(batch.totalT) if want_beams else None
        bs = beams.c_struct() if beams is not None else None
        b = batch.c_struct()
        rc = self.lib.hfbgpu_accumulate_retrain(self.h, C.byref(b), feat2.ctypes.data, res,
                                                C.byref(bs) if bs is not None else None, 0)
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_accumulate_retrain")
        return [UttResult((r.status, r.retries, r.pr, r.pruneThresh)) for r in res[:batch.numUtt]], beams

    # -- asynchronous form: consecutive batches overlap on the library's two streams --------
    def Submit(self, batch: Batch, device_feat_ptr: Optional[int] = None, want_beams: bool = False):
        """Enqueue a batch; returns a ticket whose .results() is valid after Wait()."""
        res = (hfb_utt_result * max(1, batch.numUtt))()
        beams = Beams(batch.totalT) if want_beams else None
        bs = beams.c_struct() if beams is not None else None
        b = batch.c_struct(device_feat_ptr)
        rc = self.lib.hfbgpu_submit(self.h, C.byref(b), res, C.byref(bs) if bs is not None else None,
                                    0 if device_feat_ptr is None else 1)
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_submit")
        tk = _Ticket(batch, b, res, beams, bs)
        self._inflight.append(tk)          # the library writes into these buffers until Wait()
        return tk

    def Wait(self):
        rc = self.lib.hfbgpu_wait(self.h)
        self._inflight = []
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_wait")

    def ZeroAccs(self):
        rc = self.lib.hfbgpu_zero_accs(self.h)
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_zero_accs")

    def GetAccs(self) -> np.ndarray:
        out = np.zeros(self.layout.count, np.float64)
        rc = self.lib.hfbgpu_get_accs(self.h, out.ctypes.data)
        self._inflight = []
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_get_accs")
        return out

    def SetAccs(self, acc: np.ndarray):
        acc = np.ascontiguousarray(acc, np.float64)
        rc = self.lib.hfbgpu_set_accs(self.h, acc.ctypes.data)
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_set_accs")

    def MStep(self, min_egs: int = 3, min_var: float = 0.0, mix_weight_floor: float = 0.0,
              var_floor: Optional[np.ndarray] = None) -> Tuple[FlatModel, dict]:
        """MLUpdateModels (HTKTools/HERest.c:1262-1321) on the device, from the resident accumulators.

        ``min_egs`` / ``min_var`` / ``mix_weight_floor`` are HERest's -m / -v / -w; ``var_floor`` is the
        per-dimension floor of a ``~v varFloor1`` macro (overrides ``min_var``, HERest.c:623 SetVFloor).
        Returns the re-estimated model in the flat layout (ready for a new ForwardBackward) and the
        counters HERest reports."""
        import copy
        from .htkio import LZERO, MINLARG, MINMIX
        fm = self.fm
        vf = np.full(fm.D, min_var, np.float32) if var_floor is None else np.ascontiguousarray(var_floor, np.float32)
        o = hfb_mstep_options(int(min_egs), float(mix_weight_floor) * MINMIX, vf.ctypes.data)
        mean = np.zeros((fm.G, fm.D), np.float32); var = np.zeros((fm.G, fm.D), np.float32)
        gc = np.zeros(fm.G, np.float32); w = np.zeros(int(fm.stateMixOff[-1]), np.float32)
        tp = np.zeros(len(fm.transLogA), np.float32)
        r = hfb_mstep_result(mean.ctypes.data, var.ctypes.data, gc.ctypes.data, w.ctypes.data, tp.ctypes.data, 0, 0, 0, 0)
        rc = self.lib.hfbgpu_mstep(self.h, C.byref(o), C.byref(r))
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_mstep")
        new = copy.copy(fm)
        new.mean = mean
        new.ivar = (np.float32(1.0) / var).astype(np.float32)
        new.gConst = gc
        lw = np.full(w.shape, LZERO, np.float32)
        nz = w >= MINMIX
        lw[nz] = np.log(w[nz].astype(np.float64)).astype(np.float32)
        new.mixLogWt = lw
        lt = np.full(tp.shape, LZERO, np.float32)
        nz = tp > MINLARG
        lt[nz] = np.log(tp[nz].astype(np.float64)).astype(np.float32)
        new.transLogA = lt
        info = dict(nFloorVar=r.nFloorVar, nFloorVarMix=r.nFloorVarMix, nCopied=r.nCopied, nNoOcc=r.nNoOcc,
                    var=var, mixWeight=w, transP=tp)
        return new, info

    # -- parameter-kind qualifiers on the device (HParm.c AddQualifiers) ----------------------
    def SetQualifiers(self, q=None):
        """q: flat.Qualifiers or None.  Afterwards every Batch holds static coefficients only
        (q.num_static columns); deltas / accelerations / thirds and _Z are formed on the device."""
        cs = q.c_struct() if q is not None else None
        rc = self.lib.hfbgpu_set_qualifiers(self.h, C.byref(cs) if cs is not None else None)
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_set_qualifiers")
        self.qualifiers = q

    def ExpandFeatures(self, feats) -> list:
        """The expansion alone (what HCopy with that TARGETKIND writes): list of [T][num_static] -> list of [T][D]."""
        ns, D = self.qualifiers.num_static, self.fm.D
        T = np.array([f.shape[0] for f in feats], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
        src = np.ascontiguousarray(np.concatenate(feats, axis=0), dtype=np.float32)
        assert src.shape[1] == ns
        dst = np.zeros((int(off[-1]), D), np.float32)
        rc = self.lib.hfbgpu_expand_features(self.h, src.ctypes.data, off.ctypes.data, len(feats), dst.ctypes.data)
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_expand_features")
        return [dst[off[i]:off[i + 1]] for i in range(len(feats))]

    def acc_device_ptr(self) -> int:
        return int(self.lib.hfbgpu_acc_device_ptr(self.h))

    def OutP(self, feat: np.ndarray, states: np.ndarray) -> np.ndarray:
        """log b_j(o_t) for the listed tied states (HModel.c OutP / HFB.c ShStrP)."""
        feat = np.ascontiguousarray(feat, np.float32)
        states = np.ascontiguousarray(states, np.int32)
        out = np.zeros((feat.shape[0], len(states)), np.float32)
        rc = self.lib.hfbgpu_state_loglik(self.h, feat.ctypes.data, feat.shape[0], states.ctypes.data,
                                          len(states), out.ctypes.data, None)
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_state_loglik")
        return out

    def MinDurs(self) -> np.ndarray:
        out = np.zeros(self.fm.numTrans, np.int32)
        self.lib.hfbgpu_get_min_durs(self.h, out.ctypes.data)
        return out

    def stats(self) -> hfb_stats:
        s = hfb_stats()
        self.lib.hfbgpu_get_stats(self.h, C.byref(s))
        return s

    def reset_stats(self):
        self.lib.hfbgpu_reset_stats(self.h)

    def set_stream(self, cuda_stream_handle: int):
        rc = self.lib.hfbgpu_set_stream(self.h, C.c_void_p(cuda_stream_handle))
        if rc != 0:
            raise capi.HfbError(rc, "hfbgpu_set_stream")

    def acc_tensor(self):
        """The resident FP64 accumulator buffer as a torch CUDA tensor (no copy), for the
        per-pass NCCL all-reduce (torch.distributed is plumbing, the buffer is the library's)."""
        import torch

        class _Wrap:
            pass
        w = _Wrap()
        w.__cuda_array_interface__ = {"shape": (int(self.layout.count),), "typestr": "<f8",
                                      "data": (self.acc_device_ptr(), False), "version": 2}
        t = torch.as_tensor(w, device="cuda:%d" % self.opt.device)
        t._hfb_owner = self
        return t

    def set_timing(self, on: bool):
        self.lib.hfbgpu_set_timing(self.h, 1 if on else 0)
