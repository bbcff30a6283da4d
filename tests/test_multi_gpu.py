"""NCCL path on real GPUs (skipped with fewer than 2 devices): see tools/mgpu_parity.py."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpu_allreduce_matches_reference_dump():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(ROOT, "tools", "mgpu_parity.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0 and "MGPU_PARITY_OK" in p.stdout, p.stdout[-3000:]


def test_device_group_behind_the_c_abi_equals_single_device():
    """hfbgpu_create_multi: ONE context / one host thread driving two GPUs, accumulators combined by the library's own
    peer-memory reduction -- against the single-device result, the stock single-process dump and the stock
    `-p 1..4` + `-p 0` merged MMF (tests/golden/merge_p0_tied_m4.npz)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import acc_errors
    from htk_b200.compare import load_merge_golden, reestimated_errors
    from htk_b200.estep import ForwardBackward
    z, fm, b = load_merge_golden()
    one = ForwardBackward(fm); r1, _ = one.FBFile(b); a1 = one.GetAccs(); one.close()
    grp = ForwardBackward(fm, devices=[0, 1])
    assert grp.lib.hfbgpu_num_devices(grp.h) == 2
    r2, beams = grp.FBFile(b, want_beams=True)
    tk = grp.Submit(b); grp.Wait()                      # asynchronous form, second pass over the same data
    a2 = grp.GetAccs()
    assert [tuple(r) for r in r1] == [tuple(r) for r in r2] == [tuple(r) for r in tk.results()]
    e = acc_errors(a2 / 2.0, a1, fm)
    assert max(e.values()) < 1e-5, e
    e = acc_errors(a2 / 2.0, z["ref_acc"], fm)
    assert max(e.values()) < 1e-4, e
    grp.ZeroAccs(); grp.FBFile(b)
    new, info = grp.MStep()
    grp.close()
    r = reestimated_errors(new.mean, info["var"], z["merged_mean"].astype(np.float64), z["merged_var"].astype(np.float64))
    assert r["mean"] < 1e-4 + 2e-6 and r["var"] < 1e-4 + 2e-6, r


def test_herest_gpu_on_two_devices_equals_stock(tmp_path, monkeypatch):
    """The drop-in tool with HFBGPU_DEVICES=0,1 (one process, two GPUs) against the stock HERest dump."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import test_herest_dropin as D
    if not (os.path.exists(D.HEREST) and os.path.exists(D.HEREST_GPU)):
        pytest.skip("reference binaries not built")
    from htk_b200 import htkio, synth
    tmp = str(tmp_path)
    hs = synth.make_tied_triphone_set(n_states=50, M=4, n_phys=30, n_logical=45, n_centre=6, seed=31, spread=0.2)
    hs2, fm = D._setup(tmp, hs, n_utts=12, T=300, Q=30, seed=4)
    base = ["-T", "1", "-u", "tmvw", "-p", "1", "-H", "mmf", "-I", "labs.mlf", "-S", "scp"]
    os.makedirs(os.path.join(tmp, "accA")); os.makedirs(os.path.join(tmp, "accB"))
    D._run([D.HEREST] + base + ["-M", "accA", "list"], tmp)
    monkeypatch.setenv("HFBGPU_DEVICES", "0,1")
    out = D._run([D.HEREST_GPU] + base + ["-M", "accB", "list"], tmp)
    assert "on 2 GPU(s)" in out
    a, prA, tA = htkio.read_acc_dump(os.path.join(tmp, "accA", "HER1.acc"), hs2, fm)
    b, prB, tB = htkio.read_acc_dump(os.path.join(tmp, "accB", "HER1.acc"), hs2, fm)
    assert tA == tB and abs(prA - prB) <= 1e-6 * abs(prA)
    from conftest import acc_errors
    e = acc_errors(b, a, fm)
    e.pop("totalPr"); e.pop("totalT")
    assert max(e.values()) < 1e-4, e


def test_two_model_reestimation_on_a_device_group():
    """hfb_options.alignModel through hfbgpu_create_multi: every child context aligns with one set and accumulates the other;
    the summed buffers equal the single-device result and the stock HERest dump (ALIGNMODELMMF fixture)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import acc_errors, load_two_model_golden
    from htk_b200.estep import ForwardBackward
    z, fu, fa, b, kw = load_two_model_golden("two_model_tied")
    one = ForwardBackward(fu, **kw); r1, _ = one.FBFile(b); a1 = one.GetAccs(); one.close()
    grp = ForwardBackward(fu, devices=[0, 1], **kw)
    r2, _ = grp.FBFile(b)
    a2 = grp.GetAccs()
    grp.close()
    assert [tuple(r) for r in r1] == [tuple(r) for r in r2]
    e = acc_errors(a2, a1, fu)
    assert max(e.values()) < 1e-5, e
    e = acc_errors(a2, z["ref_acc"], fu)
    assert max(e.values()) < 1e-4, e
    L = fu.layout
    assert np.array_equal(a2[L.numEgs:L.totalT], z["ref_acc"][L.numEgs:L.totalT])


def test_compressed_features_on_a_device_group():
    """hfbgpu_accumulate_compressed / hfbgpu_submit_compressed through hfbgpu_create_multi: the batch is cut into one
    utterance range per GPU and every range must take its own files' A / B vectors along (uneven utterance lengths, so
    the cut is not in the middle) -- same log-likelihoods as one device on the decoded floats, same accumulators."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import acc_errors
    from htk_b200 import htkio, synth
    from htk_b200.estep import ForwardBackward
    from htk_b200.flat import Batch, CompressedFeatures, flatten
    from oracle import hparm_oracle as H
    hs = synth.make_tied_triphone_set(n_states=40, M=4, n_phys=30, n_logical=30, n_centre=5, D=39, seed=12, spread=0.25)
    fm = flatten(hs)
    feats, labs = synth.sample_corpus(fm, n_utts=11, T=170, Q=13, seed=35, T_jitter=60)
    comp = [htkio.compress_params(f) for f in feats]
    dec = [H.decompress(s, A, B) for s, A, B in comp]
    cf = CompressedFeatures([c[0] for c in comp], [c[1] for c in comp], [c[2] for c in comp])
    b = Batch(dec, labs, 39)
    one = ForwardBackward(fm); r1, _ = one.FBFile(b); a1 = one.GetAccs(); one.close()
    grp = ForwardBackward(fm, devices=[0, 1])
    r2, _ = grp.FBFileCompressed(b, cf)
    tk = grp.SubmitCompressed(b, cf); grp.Wait()
    a2 = grp.GetAccs()
    grp.close()
    assert [tuple(r) for r in r1] == [tuple(r) for r in r2] == [tuple(r) for r in tk.results()]
    e = acc_errors(a2 / 2.0, a1, fm)
    assert max(e.values()) < 1e-5, e
