"""Golden vectors for `_C` compressed parameter files: float tables in, the file the UNMODIFIED reference's HCopy
(oracle/_ref/bin/HCopy, built by oracle/Makefile from /root/reference) writes with SAVECOMPRESSED = T (with and without
SAVEWITHCRC) out, and what the reference's own loader decodes from that file (a second HCopy back to floats).
Run in the build container: python tests/golden/make_compressed_golden.py"""
import os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from htk_b200 import htkio

HCOPY = os.path.join(ROOT, "oracle", "_ref", "bin", "HCopy")
CASES = {
    # name: (kind, columns, SAVEWITHCRC)
    "compressed_0_D_A_K": ("MFCC_0_D_A", 39, True),
    "compressed_0_D_A": ("MFCC_0_D_A", 39, False),
    "compressed_0_static_K": ("MFCC_0", 13, True),
}
LENGTHS = [1, 2, 7, 40, 333]


def run(args):
    r = subprocess.run(args, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout + r.stderr)


def main():
    rng = np.random.default_rng(20261018)
    for name, (kind, D, crc) in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            c1, c2 = os.path.join(d, "c1.cfg"), os.path.join(d, "c2.cfg")
            open(c1, "w").write("SAVECOMPRESSED = T\nSAVEWITHCRC = %s\n" % ("T" if crc else "F"))
            open(c2, "w").write("SAVECOMPRESSED = F\nSAVEWITHCRC = F\n")
            xs, files, decs = [], [], []
            for i, T in enumerate(LENGTHS):
                a, b, c = (os.path.join(d, "%s%d.mfc" % (k, i)) for k in "abc")
                while True:
                    x = (rng.standard_normal((T, D)) * rng.uniform(0.05, 20.0, D) + rng.uniform(-30, 30, D)).astype(np.float32)
                    if T >= 7:
                        x[:, 3] = x[0, 3]                  # a constant column: CalcCompress' A = 1, B = max branch
                    htkio.write_htk_features(a, x, kind)
                    try:
                        run([HCOPY, "-C", c1, a, b])
                        break
                    except RuntimeError as e:
                        # the reference's own float rounding can push the column maximum to 32768 (HError 6393,
                        # CompressPBlock): such a table cannot be saved compressed at all -- draw another one
                        if "6393" not in str(e):
                            raise
                run([HCOPY, "-C", c2, b, c])
                y, _, k2 = htkio.read_htk_features(c)
                assert y.shape == x.shape, (y.shape, x.shape)
                xs.append(x); decs.append(y); files.append(np.frombuffer(open(b, "rb").read(), np.uint8))
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), kind=kind, crc=crc,
                                lengths=np.array(LENGTHS), original=np.concatenate(xs), decoded=np.concatenate(decs),
                                file_bytes=np.concatenate(files), file_sizes=np.array([len(f) for f in files]))
            print(name, [len(f) for f in files])


if __name__ == "__main__":
    main()
