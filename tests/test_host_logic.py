"""Host-side logic that needs no GPU: file formats, flattening, the C-ABI library loading
and exporting every symbol include/hfbgpu.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from htk_b200 import htkio, synth
from htk_b200.flat import flatten, hfb_acc_layout


def test_feature_file_roundtrip(tmp_path):
    x = np.random.default_rng(0).standard_normal((17, 39)).astype(np.float32)
    p = str(tmp_path / "a.mfc")
    htkio.write_htk_features(p, x, "MFCC_0_D_A")
    y, period, kind = htkio.read_htk_features(p)
    assert np.array_equal(x, y) and period == 100000 and kind == 0o21406 - 0o20000 + 0o20000


def test_mmf_roundtrip_and_flatten(tmp_path):
    hs = synth.make_tied_triphone_set(n_states=20, M=3, n_phys=12, n_logical=20, n_centre=4, seed=3)
    p = str(tmp_path / "mmf")
    htkio.write_mmf(p, hs)
    htkio.write_hmm_list(str(tmp_path / "list"), hs)
    hs2 = htkio.read_mmf([p], hmm_list=open(str(tmp_path / "list")).read().splitlines())
    f1, f2 = flatten(hs), flatten(hs2)
    assert f1.J == f2.J == 20 and f1.G == f2.G == 60 and f1.P == f2.P == 12 and f1.numTrans == f2.numTrans == 4
    assert np.allclose(f1.mean, f2.mean, rtol=1e-6, atol=1e-6)
    assert np.array_equal(f1.hmmState, f2.hmmState) and np.array_equal(f1.hmmTrans, f2.hmmTrans)
    assert len(hs2.logical) == 8
    # log transitions: absent arcs are LZERO, rows of emitting states normalise
    A = f1.transLogA[:25].reshape(5, 5)
    assert A[0, 0] == np.float32(-1e10) and abs(np.exp(A[1, 1]) + np.exp(A[1, 2]) - 1) < 1e-6


def test_gconst_and_logweights():
    hs = synth.make_monophone_set(n_phones=2, M=2, seed=1)
    fm = flatten(hs)
    g = hs.hmms[0].states[0].mixes[0][1]
    want = 39 * np.log(2 * np.pi) + np.sum(np.log(g.var.astype(np.float64)))
    assert abs(fm.gConst[0] - want) < 1e-4
    assert np.allclose(fm.ivar[0], 1.0 / g.var, rtol=1e-6)
    w = hs.hmms[0].states[0].mixes[0][0]
    assert abs(fm.mixLogWt[0] - np.log(w)) < 1e-6


def test_scan_order_is_hash_order():
    names = ["S", "C", "V", "N", "L"]
    order = htkio.scan_order(names)
    assert sorted(order) == sorted(names)
    assert [htkio.htk_hash(n) for n in order] == sorted(htkio.htk_hash(n) for n in names)


def test_acc_dump_roundtrip(tmp_path):
    hs = synth.make_tied_triphone_set(n_states=10, M=2, n_phys=6, n_logical=6, n_centre=2, seed=4)
    fm = flatten(hs)
    rng = np.random.default_rng(1)
    acc = rng.random(fm.layout.count) * 10
    acc[fm.layout.numEgs:fm.layout.totalT] = rng.integers(0, 50, fm.P)
    acc[fm.layout.totalT] = 1234
    acc[fm.layout.numOk] = acc[fm.layout.numSkipped] = 0
    p = str(tmp_path / "HER1.acc")
    htkio.write_acc_dump(p, hs, fm, acc)
    back, tp, tt = htkio.read_acc_dump(p, hs, fm)
    assert tt == 1234
    assert np.allclose(back[:fm.layout.numOk], acc[:fm.layout.numOk].astype(np.float32), rtol=1e-7)


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "hfbgpu.h")).read()
    return sorted(set(re.findall(r"\b(hfbgpu_[a-z_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from htk_b200 import capi
    lib = capi.load()
    syms = _declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) == set(capi.EXPORTS)
    assert lib.hfbgpu_abi_version() == 3


def test_layout_agrees_between_library_and_python():
    from htk_b200 import capi
    lib = capi.load()
    for name in ("htkdemo_t2000", "synth_tied_m4"):
        z, fm, b, kw = load_golden(name)
        L = hfb_acc_layout()
        m = fm.c_struct()
        assert lib.hfbgpu_acc_layout(C.byref(m), C.byref(L)) == 0
        for k in ("tran", "tranOcc", "wtC", "wtOcc", "muSum", "muOcc", "vaSum", "vaOcc", "numEgs", "totalT",
                  "totalPr", "numOk", "numSkipped", "count"):
            assert getattr(L, k) == getattr(fm.layout, k), k


def test_no_cpu_fallback_without_device():
    """Without a GPU, create must fail loudly with HFB_ENODEVICE -- never compute on the CPU."""
    from htk_b200 import capi
    lib = capi.load()
    if lib.hfbgpu_device_count() > 0:
        pytest.skip("a device is present")
    from htk_b200.estep import ForwardBackward
    z, fm, b, kw = load_golden("htkdemo_t2000")
    with pytest.raises(capi.HfbError) as e:
        ForwardBackward(fm, **kw)
    assert e.value.code == -2


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "htk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "hfb_oracle" not in txt and "libhfboracle" not in txt, f


def test_qualifier_description_from_kinds():
    """Qualifiers.from_kinds mirrors what HParm derives from SOURCEKIND / TARGETKIND (FindSpans, HParm.c:1430-1455;
    AddQualifiers :1707-1712): static width, which orders to add, which leading columns _Z zero-means."""
    from htk_b200.flat import Qualifiers
    q = Qualifiers.from_kinds("MFCC_0", "MFCC_0_D_A_Z", 13)
    assert (q.num_static, q.del_win, q.acc_win, q.third_win, q.zero_mean_cols, q.vec_size) == (13, 2, 2, 0, 13, 39)
    q = Qualifiers.from_kinds("MFCC_E", "MFCC_E_D_A_Z", 13, del_win=3, acc_win=1)
    assert (q.del_win, q.acc_win, q.zero_mean_cols, q.vec_size) == (3, 1, 12, 39)      # the energy is not zero-meaned
    q = Qualifiers.from_kinds("MFCC", "MFCC_D", 12, simple_diffs=True)
    assert (q.vec_size, q.simple_diffs, q.zero_mean_cols) == (24, True, 0)
    q = Qualifiers.from_kinds("MFCC_0_K", "MFCC_0_D_A_T", 13)                          # _K (CRC) is a file property
    assert q.vec_size == 52
    q = Qualifiers.from_kinds("MFCC_E", "MFCC_E_D_A_N", 13)                            # _N: energy out of the observation
    assert (q.suppress_energy, q.vec_size, q.c_struct().suppressEnergy) == (True, 38, 1)
    q = Qualifiers.from_kinds("MFCC_0", "MFCC_0_D_N_Z", 13)
    assert (q.vec_size, q.zero_mean_cols) == (25, 12)
    # _N needs _D and exactly one of _E / _0 (ValidConversion, HParm.c:1415-1420)
    for src, tgt in (("MFCC_0", "PLP_0_D"), ("MFCC_0_D", "MFCC_0_D_A"), ("MFCC_0", "MFCC_0_N"), ("MFCC", "MFCC_D_N"),
                     ("MFCC_E", "MFCC_0_D")):
        with pytest.raises(ValueError):
            Qualifiers.from_kinds(src, tgt, 13)
    c = Qualifiers(13, 2, 2, 0, True, 13).c_struct()
    assert (c.numStatic, c.delWin, c.accWin, c.thirdWin, c.simpleDiffs, c.zeroMeanCols) == (13, 2, 2, 0, 1, 13)


def test_qualifier_oracle_properties():
    """Size-independent properties of the regression restatement (oracle/hparm_oracle.py): a constant signal has zero
    differentials, a ramp has slope-1 deltas away from the ends and clamped ones at the ends (first / last frame
    replicated, HSigP.c:844-845), zero-meaning removes the column mean."""
    from oracle import hparm_oracle as H
    T = 50
    const = np.full((T, 3), 2.5, np.float32)
    out = H.expand(const, 2, 2)
    assert np.all(out[:, 3:] == 0)
    ramp = np.arange(T, dtype=np.float32)[:, None].repeat(2, 1)
    d = H.regress(ramp, 2, False)
    assert np.allclose(d[2:-2], 1.0) and np.allclose(d[0], (1 * 1 + 2 * 2) / 10.0) and np.allclose(d[-1], d[0])
    assert np.allclose(H.regress(ramp, 2, True)[2:-2], 1.0)
    x = np.random.default_rng(0).standard_normal((T, 4)).astype(np.float32) + 7
    z = H.expand(x, 0, 0, 0, False, 3)
    assert np.all(np.abs(z[:, :3].mean(0)) < 1e-6) and np.array_equal(z[:, 3], x[:, 3])
