"""Compare the GMM kernel variants (CTA pair 3xFP16 = default, CTA pair 3xTF32, single CTA 3xTF32) with each
other and with an FP64 evaluation (OutP only)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from htk_b200.estep import ForwardBackward
from htk_b200 import synth
from htk_b200.flat import flatten

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 600
hs = synth.make_tied_triphone_set(n_states=600, M=M, n_phys=400, n_logical=400, n_centre=20, seed=21, spread=0.2)
fm = flatten(hs)
if len(sys.argv) > 3 and sys.argv[3] == "scaled":
    # real front ends are not unit-variance: per-dimension scales over five decades and offsets of hundreds
    rs = np.random.default_rng(5)
    sc = (10.0 ** rs.uniform(-2, 3, fm.D)).astype(np.float32); of = rs.uniform(-500, 500, fm.D).astype(np.float32)
    fm.mean[:, :fm.D] = fm.mean[:, :fm.D] * sc + of
    fm.ivar[:, :fm.D] = fm.ivar[:, :fm.D] / (sc * sc)
    fm.gConst[:] = (fm.D * np.log(2 * np.pi) - np.sum(np.log(fm.ivar[:, :fm.D].astype(np.float64)), axis=1)).astype(np.float32)
feats, labs = synth.sample_corpus(fm, n_utts=1, T=T, Q=max(3, T // 10), seed=3)
feat = feats[0]; states = np.arange(fm.J, dtype=np.int32)
x = feat.astype(np.float64)
mean = fm.mean.astype(np.float64); iv = fm.ivar.astype(np.float64); gc = fm.gConst.astype(np.float64)
ex = np.zeros((len(x), fm.J))
for s in range(fm.J):
    o, e = fm.stateMixOff[s], fm.stateMixOff[s + 1]
    g = fm.mixGauss[o:e]
    d = x[:, None, :] - mean[g][None]
    lp = -0.5 * (gc[g][None] + np.sum(d * d * iv[g][None], axis=2)) + fm.mixLogWt[o:e].astype(np.float64)[None]
    m = lp.max(1); ex[:, s] = m + np.log(np.exp(lp - m[:, None]).sum(1))
res = {}
for name, env in (("f16", None), ("pair", "HFBGPU_TC_TF32"), ("single", "HFBGPU_NO_PAIR")):
    if env: os.environ[env] = "1"
    fb = ForwardBackward(fm, gmm_kernel=2); got = fb.OutP(feat, states).astype(np.float64); fb.close()
    if env: del os.environ[env]
    res[name] = got
    print("%-6s vs exact: max %.2e mean %.2e" % (name, np.abs(got - ex).max(), np.abs(got - ex).mean()), flush=True)
print("pair vs single: max %.2e" % np.abs(res["pair"] - res["single"]).max())
print("f16 vs pair(tf32): max %.2e mean %.2e" % (np.abs(res["f16"] - res["pair"]).max(), np.abs(res["f16"] - res["pair"]).mean()))
