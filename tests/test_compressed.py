"""HTK `_C` compressed parameter files (HASCOMPX, optionally `_K`): the data format on the input side of the path.

The reference's loader decodes them on the host (HTKLib/HParm.c:3489-3494: v = ((float)s + B) / A); the library takes the
files' 16-bit integers and decodes them on the device (feat_decompress_kernel).  Evidence:
  * CPU: the numpy restatement (oracle/hparm_oracle.py) and the host-side file reader / writer (htk_b200/htkio.py) against
    files written AND read back by the unmodified reference's HCopy (tests/golden/compressed_*.npz) -- bit-exact;
  * GPU: hfbgpu_decompress_features bit-identical to those vectors; the E-step fed with the integers equals the E-step
    fed with the decoded floats (same log-likelihoods and beams), alone, with device qualifiers, asynchronously.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, acc_errors
from htk_b200 import htkio
from htk_b200.flat import Batch, CompressedFeatures

CASES = ["compressed_0_D_A_K", "compressed_0_D_A", "compressed_0_static_K"]


def _load(name, tmp_path):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    off = np.concatenate([[0], np.cumsum(z["lengths"])])
    foff = np.concatenate([[0], np.cumsum(z["file_sizes"])])
    n = len(z["lengths"])
    files = []
    for i in range(n):
        p = str(tmp_path / ("f%d.mfc" % i))
        open(p, "wb").write(z["file_bytes"][foff[i]:foff[i + 1]].tobytes())
        files.append(p)
    return (z, files, [z["original"][off[i]:off[i + 1]] for i in range(n)], [z["decoded"][off[i]:off[i + 1]] for i in range(n)])


@pytest.mark.parametrize("name", CASES)
def test_oracle_decoding_and_check_sum_match_hcopy(name, tmp_path):
    from oracle import hparm_oracle as H
    z, files, orig, dec = _load(name, tmp_path)
    for f, x, y in zip(files, orig, dec):
        s, A, B, period, kind, ok = htkio.read_htk_compressed(f)
        assert s.shape == x.shape and kind & 0o2000
        assert ok is (True if bool(z["crc"]) else None)
        if ok:                                         # the oracle's restatement of UpdateCRCC on the same words
            raw = open(f, "rb").read()
            assert H.crc(np.frombuffer(raw[12:-2], dtype=">u2")) == int.from_bytes(raw[-2:], "big")
        v = H.decompress(s, A, B)
        assert np.array_equal(v.view(np.uint32), y.view(np.uint32))          # what the reference's loader decoded
        assert np.array_equal(htkio.decompress_params(s, A, B).view(np.uint32), y.view(np.uint32))
        # compression is lossy by half a quantisation step per column at most
        assert np.all(np.abs(v - x) <= 0.5001 / A + 1e-6 * np.abs(x))


@pytest.mark.parametrize("name", CASES)
def test_writer_reproduces_hcopy_files_byte_for_byte(name, tmp_path):
    """htkio.write_htk_compressed restates CalcCompress / CompressPBlock (HParm.c:4892-4960) and the check sum."""
    z, files, orig, _ = _load(name, tmp_path)
    for i, (f, x) in enumerate(zip(files, orig)):
        p = str(tmp_path / ("w%d.mfc" % i))
        htkio.write_htk_compressed(p, x, str(z["kind"]), with_crc=bool(z["crc"]))
        assert open(p, "rb").read() == open(f, "rb").read()


def test_corrupt_file_fails_the_check_sum(tmp_path):
    x = np.random.default_rng(3).standard_normal((20, 13)).astype(np.float32)
    p = str(tmp_path / "a.mfc")
    htkio.write_htk_compressed(p, x, "MFCC_0")
    assert htkio.read_htk_compressed(p)[5] is True
    raw = bytearray(open(p, "rb").read())
    raw[200] ^= 0x10
    open(p, "wb").write(bytes(raw))
    assert htkio.read_htk_compressed(p)[5] is False


# ------------------------------------------------------------------------------------------ GPU

def _fb(fm, **kw):
    from htk_b200.estep import ForwardBackward
    return ForwardBackward(fm, **kw)


def _model_of_width(D, M=2):
    from htk_b200 import synth
    from htk_b200.flat import flatten
    return flatten(synth.make_monophone_set(n_phones=6, M=M, D=D, seed=5, spread=0.3))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_decoding_bit_identical_to_hcopy(name, tmp_path):
    z, files, _, dec = _load(name, tmp_path)
    parts = [htkio.read_htk_compressed(f) for f in files]
    cf = CompressedFeatures([p[0] for p in parts], [p[1] for p in parts], [p[2] for p in parts])
    fb = _fb(_model_of_width(cf.feat.shape[1]))
    out = fb.DecompressFeatures(cf)
    fb.close()
    want = np.concatenate(dec)
    assert out.shape == want.shape
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def _corpus(static_only):
    from htk_b200 import synth
    from htk_b200.flat import flatten
    hs = synth.make_tied_triphone_set(n_states=40, M=4, n_phys=30, n_logical=30, n_centre=5, D=39, seed=12, spread=0.25)
    fm = flatten(hs)
    feats, labs = synth.sample_corpus(fm, n_utts=9, T=170, Q=13, seed=33, T_jitter=40)
    if static_only:
        feats = [np.ascontiguousarray(f[:, :13]) for f in feats]
    comp = [htkio.compress_params(f) for f in feats]
    return fm, labs, comp


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["blocking", "async", "small_waves"])
def test_estep_on_compressed_features_equals_estep_on_decoded_floats(mode):
    """The integers of `_C` files in, decoded on the device: log-likelihoods, retry thresholds and beams identical to the
    same utterances decoded beforehand (oracle/hparm_oracle.py, pinned to HCopy), accumulators equal up to the order of
    the atomic additions, and within 1e-4 of the C oracle on the decoded floats."""
    from oracle import hparm_oracle as H
    from oracle import oracle_lib as O
    from htk_b200.flat import make_options
    fm, labs, comp = _corpus(False)
    dec = [H.decompress(s, A, B) for s, A, B in comp]
    cf = CompressedFeatures([c[0] for c in comp], [c[1] for c in comp], [c[2] for c in comp])
    b = Batch(dec, labs, 39)
    kw = dict(prune=(60.0, 30.0, 300.0))
    if mode == "small_waves":
        kw["workspace_bytes"] = 1 << 20                # several waves per call: A / B rows must follow the utterances
    fb = _fb(fm, **kw)
    res0, beams0 = fb.FBFile(b, want_beams=True); a0 = fb.GetAccs()
    fb.ZeroAccs()
    bc = Batch.from_arrays(None, b.frameOff, b.lab, b.labOff)          # no float table at all
    bc.feat = np.zeros((0, 39), np.float32)
    if mode == "async":
        tk = fb.SubmitCompressed(bc, cf); fb.Wait(); res1 = tk.results(); beams1 = None
    else:
        res1, beams1 = fb.FBFileCompressed(bc, cf, want_beams=True)
    a1 = fb.GetAccs()
    st = fb.stats()
    fb.close()
    assert [r.status for r in res0] == [r.status for r in res1] and any(r.status == 0 for r in res0)
    for x, y in zip(res0, res1):
        assert x.pr == y.pr and x.pruneThresh == y.pruneThresh and x.retries == y.retries
    if beams1 is not None:
        for k in ("qLo", "qHi", "sq", "eq"):
            assert np.array_equal(getattr(beams0, k), getattr(beams1, k)), k
    e = acc_errors(a1, a0, fm)
    assert max(e.values()) < 1e-5, e
    assert st.h2dBytes > 0
    ao = O.accumulate(fm, make_options(**{k: v for k, v in kw.items() if k != "workspace_bytes"}), b, acc_double=True)[0]
    e = acc_errors(a1, ao, fm)
    assert max(e.values()) < 1e-4, e


@pytest.mark.gpu
def test_compressed_static_files_with_device_qualifiers():
    """The HTK book's own recipe: compressed static coefficients on disk (MFCC_0_C_K), TARGETKIND = MFCC_0_D_A_Z in the
    training configuration.  Integers -> floats -> differentials + mean normalisation, all on the device."""
    from oracle import hparm_oracle as H
    from htk_b200.flat import Qualifiers
    fm, labs, comp = _corpus(True)
    q = Qualifiers.from_kinds("MFCC_0", "MFCC_0_D_A_Z", 13)
    static = [H.decompress(s, A, B) for s, A, B in comp]
    full = [H.expand(x, q.del_win, q.acc_win, q.third_win, q.simple_diffs, q.zero_mean_cols) for x in static]
    b_full = Batch(full, labs, 39)
    fb = _fb(fm)
    res0, beams0 = fb.FBFile(b_full, want_beams=True); a0 = fb.GetAccs()
    fb.ZeroAccs(); fb.SetQualifiers(q)
    cf = CompressedFeatures([c[0] for c in comp], [c[1] for c in comp], [c[2] for c in comp])
    bc = Batch.from_arrays(np.zeros((0, 13), np.float32), b_full.frameOff, b_full.lab, b_full.labOff)
    res1, beams1 = fb.FBFileCompressed(bc, cf, want_beams=True); a1 = fb.GetAccs()
    fb.close()
    assert [r.status for r in res0] == [r.status for r in res1] and any(r.status == 0 for r in res0)
    for x, y in zip(res0, res1):
        if x.status == 0:
            assert abs(x.pr - y.pr) <= 1e-9 * abs(x.pr)
    for k in ("qLo", "qHi", "sq", "eq"):
        assert np.array_equal(getattr(beams0, k), getattr(beams1, k)), k
    e = acc_errors(a1, a0, fm)
    assert max(e.values()) < 1e-5, e


@pytest.mark.gpu
def test_compressed_entry_rejects_bad_arguments():
    import ctypes as C
    from htk_b200 import capi
    from htk_b200.flat import hfb_utt_result
    fm, labs, comp = _corpus(False)
    cf = CompressedFeatures([c[0] for c in comp], [c[1] for c in comp], [c[2] for c in comp])
    b = Batch.from_arrays(np.zeros((0, 39), np.float32), cf.frameOff, np.concatenate(labs).astype(np.int32),
                          np.concatenate([[0], np.cumsum([len(l) for l in labs])]).astype(np.int32))
    fb = _fb(fm)
    res = (hfb_utt_result * b.numUtt)()
    bs = b.c_struct()
    assert fb.lib.hfbgpu_accumulate_compressed(fb.h, C.byref(bs), None, res, None) == -1
    c = cf.c_struct(); c.scaleA = None
    assert fb.lib.hfbgpu_accumulate_compressed(fb.h, C.byref(bs), C.byref(c), res, None) == -1
    a = fb.GetAccs()
    assert not a.any()                                   # nothing was accumulated by the rejected calls
    fb.close()
