"""Timeline of the CTA-pair GMM kernel (pair 0) from HFBGPU_TC_TRACE.

    HFBGPU_TC_TRACE=gpurun_out/tc_trace.bin python tools/tc_trace.py run     # on the GPU box
    python tools/tc_trace.py show gpurun_out/tc_trace.bin [first_tile] [n]    # anywhere
"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
NT = 512

if sys.argv[1] == "run":
    import torch
    from bench import WORKLOADS, make_model, make_batch
    from htk_b200.estep import ForwardBackward
    cfg = WORKLOADS["cfg3"]
    fm = make_model(cfg); dev = torch.device("cuda", 0)
    fb = ForwardBackward(fm)
    b, df = make_batch(fm, cfg, 512, 1000, dev)
    tr = os.environ.pop("HFBGPU_TC_TRACE")
    for _ in range(2):
        fb.FBFile(b, device_feat_ptr=df.data_ptr())
    os.environ["HFBGPU_TC_TRACE"] = tr
    os.environ["HFBGPU_STREAMS"] = "1"
    fb.FBFile(b, device_feat_ptr=df.data_ptr())
    print("trace written to", tr)
else:
    a = np.fromfile(sys.argv[2], dtype=np.int64).reshape(3, 2, NT, 16)
    t0 = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    mma, epi, prod = a[0], a[1], a[2]
    base = mma[0, t0, 0]
    def r(x): return "%6d" % (x - base) if x else "     -"
    print("cycles relative to the MMA thread starting tile %d (leader CTA clock)" % t0)
    print("tile | MMA: waitE gotE | waitB0 gotB0 waitB1 gotB1 waitB2 gotB2 | commit || EPI0: wait got done || PROD0 k0: wait got  k1: wait got  k2: wait got")
    for t in range(t0, t0 + n):
        m = mma[0, t]; e = epi[0, t]; pr = prod[0, t]
        print("%4d | %s %s | %s %s %s %s %s %s | %s || %s %s %s || %s %s  %s %s  %s %s" % (
            t, r(m[0]), r(m[1]), r(m[2]), r(m[3]), r(m[4]), r(m[5]), r(m[6]), r(m[7]), r(m[8]),
            r(e[0]), r(e[1]), r(e[2]), r(pr[0]), r(pr[1]), r(pr[2]), r(pr[3]), r(pr[4]), r(pr[5])))
    d = np.diff(mma[0, 50:450, 0])
    print("cycles per tile (MMA thread, tiles 50..450): mean %.0f median %.0f" % (d.mean(), np.median(d)))
    w = lambda x, y: float(np.mean(x[50:450] - y[50:450]))
    print("MMA thread waits: tmemEmpty %.0f, B chunk0 %.0f, chunk1 %.0f, chunk2 %.0f; issue+rest %.0f" % (
        w(mma[0, :, 1], mma[0, :, 0]), w(mma[0, :, 3], mma[0, :, 2]), w(mma[0, :, 5], mma[0, :, 4]), w(mma[0, :, 7], mma[0, :, 6]),
        d.mean() - w(mma[0, :, 1], mma[0, :, 0]) - w(mma[0, :, 3], mma[0, :, 2]) - w(mma[0, :, 5], mma[0, :, 4]) - w(mma[0, :, 7], mma[0, :, 6])))
    print("epilogue warp (leader CTA): wait tmemFull %.0f, work %.0f" % (w(epi[0, :, 1], epi[0, :, 0]), w(epi[0, :, 2], epi[0, :, 1])))
    print("producer (leader CTA) waits on emptyB: k0 %.0f k1 %.0f k2 %.0f" % (
        w(prod[0, :, 1], prod[0, :, 0]), w(prod[0, :, 3], prod[0, :, 2]), w(prod[0, :, 5], prod[0, :, 4])))
