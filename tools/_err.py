import os, sys
sys.path.insert(0, "tests")
from test_gpu_parity import load_golden, _fb, acc_errors
for name in ["synth_tied_m4", "synth_long_m3", "synth_tee_m2"]:
    z, fm, b, kw = load_golden(name)
    out = {}
    for tag, env in (("stats5", None), ("stats3", "HFBGPU_STATS3")):
        if env: os.environ[env] = "1"
        fb = _fb(fm, **kw); fb.FBFile(b); a = fb.GetAccs(); fb.close()
        if env: del os.environ[env]
        e = acc_errors(a, z["ref_acc"], fm)
        out[tag] = a
        print(name, tag, "vs ref", {k: float("%.2g" % v) for k, v in e.items()})
    print(name, "5 vs 3", {k: float("%.2g" % v) for k, v in acc_errors(out["stats5"], out["stats3"], fm).items()})
