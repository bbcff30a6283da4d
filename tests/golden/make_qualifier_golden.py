"""Golden vectors for the qualifier expansion: static coefficients in, what the UNMODIFIED reference's HCopy
(oracle/_ref/bin/HCopy, built by oracle/Makefile from /root/reference) writes for the target kind out.
Run in the build container: python tests/golden/make_qualifier_golden.py"""
import os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from htk_b200 import htkio

HCOPY = os.path.join(ROOT, "oracle", "_ref", "bin", "HCopy")
CASES = {
    # name: (source kind, target kind, static width, config lines)
    "qualifiers_0_D_A": ("MFCC_0", "MFCC_0_D_A", 13, []),
    "qualifiers_0_D_A_Z": ("MFCC_0", "MFCC_0_D_A_Z", 13, []),
    "qualifiers_E_D_A_Z_w3": ("MFCC_E", "MFCC_E_D_A_Z", 13, ["DELTAWINDOW = 3", "ACCWINDOW = 1"]),
    "qualifiers_0_D_A_T": ("MFCC_0", "MFCC_0_D_A_T", 13, ["THIRDWINDOW = 2"]),
    "qualifiers_D_simple": ("MFCC", "MFCC_D", 12, ["SIMPLEDIFFS = T"]),
    "qualifiers_Z_only": ("MFCC_0", "MFCC_0_Z", 13, []),
}
LENGTHS = [1, 2, 3, 4, 5, 9, 40, 333]          # incl. utterances shorter than the windows (AddDiffs' n <= 0 branch)


def main():
    rng = np.random.default_rng(20240917)
    for name, (src_kind, tgt_kind, ns, extra) in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            cfg = os.path.join(d, "c.cfg")
            open(cfg, "w").write("\n".join(["TARGETKIND = " + tgt_kind] + extra) + "\n")
            ins, outs = [], []
            for i, T in enumerate(LENGTHS):
                x = (rng.standard_normal((T, ns)) * rng.uniform(0.5, 20.0, ns) + rng.uniform(-30, 30, ns)).astype(np.float32)
                a, b = os.path.join(d, "a%d.mfc" % i), os.path.join(d, "b%d.mfc" % i)
                htkio.write_htk_features(a, x, src_kind)
                r = subprocess.run([HCOPY, "-C", cfg, a, b], capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError(r.stdout + r.stderr)
                y, _, _ = htkio.read_htk_features(b)
                ins.append(x); outs.append(y)
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"),
                                src_kind=src_kind, tgt_kind=tgt_kind, config=np.array(extra, dtype=object).astype(str),
                                lengths=np.array(LENGTHS), static=np.concatenate(ins), expanded=np.concatenate(outs))
            print(name, outs[0].shape[1])


if __name__ == "__main__":
    main()
