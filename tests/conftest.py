import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["htkdemo_t2000", "htkdemo_t20_15_200", "synth_mono_m1", "synth_tee_m2", "synth_tied_m4",
                "synth_long_m3"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """-> (npz, FlatModel, Batch, hfb_options)"""
    from htk_b200.flat import Batch, flat_from_arrays, make_options
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    fm = flat_from_arrays(D=int(z["D"]), **{k[2:]: z[k] for k in z.files if k.startswith("m_")})
    b = Batch.from_arrays(z["feat"], z["frameOff"], z["lab"], z["labOff"])
    pr = z["prune"]
    prune = None if pr[0] >= 1e19 else tuple(float(x) for x in pr)
    uf = int(z["uflags"]) if "uflags" in z.files else 15
    return z, fm, b, dict(prune=prune, min_frwd_p=float(z["minFrwdP"]), uflags=uf)


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


def have_gpu():
    try:
        from htk_b200 import capi
        return capi.load().hfbgpu_device_count() > 0
    except Exception:
        return False


QUALIFIER_CASES = ["qualifiers_0_D_A", "qualifiers_0_D_A_Z", "qualifiers_E_D_A_Z_w3", "qualifiers_0_D_A_T",
                   "qualifiers_D_simple", "qualifiers_Z_only"]


def load_qualifier_golden(name):
    """-> (Qualifiers, [static per utterance], [expanded per utterance as written by the reference's HCopy])"""
    from htk_b200.flat import Qualifiers
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    cfg = dict(l.split(" = ") for l in z["config"].tolist())
    ns = z["static"].shape[1]
    q = Qualifiers.from_kinds(str(z["src_kind"]), str(z["tgt_kind"]), ns, int(cfg.get("DELTAWINDOW", 2)),
                              int(cfg.get("ACCWINDOW", 2)), int(cfg.get("THIRDWINDOW", 2)),
                              cfg.get("SIMPLEDIFFS", "F") == "T")
    off = np.concatenate([[0], np.cumsum(z["lengths"])])
    n = len(z["lengths"])
    return (q, [z["static"][off[i]:off[i + 1]] for i in range(n)], [z["expanded"][off[i]:off[i + 1]] for i in range(n)])
