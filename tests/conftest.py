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


TWO_MODEL_CASES = ["two_model_tied", "two_model_mono", "two_model_complevel"]


def load_two_model_golden(name):
    """-> (npz, update FlatModel, alignment FlatModel, Batch with both label arrays, options kwargs)
    Fixtures of tests/golden/make_two_model_golden.py: the stock HERest run with ALIGNMODELMMF."""
    from htk_b200.flat import Batch, flat_from_arrays
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    fu = flat_from_arrays(D=int(z["D"]), names=list(z["names_up"]), **{k[2:]: z[k] for k in z.files if k.startswith("m_")})
    fa = flat_from_arrays(D=int(z["D"]), names=list(z["names_al"]), **{k[2:]: z[k] for k in z.files if k.startswith("a_")})
    b = Batch.from_arrays(z["feat"], z["frameOff"], z["lab"], z["labOff"]).with_align_labels(z["labAlign"])
    pr = z["prune"]
    prune = None if pr[0] >= 1e19 else tuple(float(x) for x in pr)
    kw = dict(prune=prune, min_frwd_p=float(z["minFrwdP"]), uflags=int(z["uflags"]), align_model=fa)
    if "comp_level" in z.files and bool(z["comp_level"]):
        kw["align_comp_level"] = True
    return z, fu, fa, b, kw


from htk_b200.compare import acc_errors  # noqa: E402,F401  (re-exported: the tests import it from here)


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
