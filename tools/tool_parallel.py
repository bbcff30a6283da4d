"""Throughput of the drop-in TOOL when the reference's own parallel mode is used on ONE GPU:
P concurrent `HERest_gpu -p k` processes (k = 1..P, each with its share of the script list, exactly how HERest is run
on a compute farm -- HERest.c:514-521), all bound to the same B200, followed by the stock `HERest -p 0` merge.

One HERest process feeds the GPU at ~12 M frames/s: its file loop is HTK's own single-threaded label loading
(`LoadLabs`, ~27 us per utterance).  The GPU takes ~200 M frames/s, so the way to fill it from files is the
reference's own `-p` parallelism -- this script measures how far that goes.

The corpus is `--files` distinct cfg3 utterances (1000 frames, 100 labels); every script list names each of its files
`--repeats` times (HERest does not care; the files stay in the page cache, the label lookups are the real ones), so a run
lasts seconds instead of milliseconds without tens of GB on disk.  Rate = marginal: (frames of the long run - frames of the
short run) / (wall of the long run - wall of the short run), start-up (CUDA context, MMF load, MLF index) cancels.

    python tools/tool_parallel.py [--procs 1,4,8,16] [--files 1024] [--repeats 8,32] [--compressed]
"""
import argparse, json, os, re, shutil, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from htk_b200 import htkio, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", default="1,4,8,16")
    ap.add_argument("--files", type=int, default=1024)
    ap.add_argument("--repeats", default="8,32")
    ap.add_argument("--compressed", action="store_true")
    args = ap.parse_args()
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "HERest_gpu")
    cfg = dict(bench.WORKLOADS["cfg3"])
    fm = bench.make_model(cfg)
    T, Q = cfg["T"], cfg["Q"]
    w = tempfile.mkdtemp(prefix="hfb_par_")
    try:
        synth.write_flat_as_mmf(os.path.join(w, "mmf"), os.path.join(w, "list"), fm)
        rng = np.random.default_rng(11)
        lab, gauss = synth.corpus_plan(fm, args.files, T, Q, seed=900)
        mlf, scp = {}, []
        for u0 in range(0, args.files, 64):
            g = gauss[u0 * T:(u0 + 64) * T]
            x = (fm.mean[g] + rng.standard_normal((len(g), fm.D)).astype(np.float32) / np.sqrt(fm.ivar[g])).astype(np.float32)
            for i in range(len(g) // T):
                fn = os.path.join(w, "u%d.mfc" % (u0 + i))
                if args.compressed:
                    htkio.write_htk_compressed(fn, x[i * T:(i + 1) * T], "MFCC_0_D_A", with_crc=True)
                else:
                    htkio.write_htk_features(fn, x[i * T:(i + 1) * T])
                mlf["u%d" % (u0 + i)] = [fm.names[j] for j in lab[u0 + i]]
                scp.append(fn)
        htkio.write_mlf(os.path.join(w, "labs.mlf"), mlf)
        reps = [int(r) for r in args.repeats.split(",")]
        out = {"files": args.files, "frames_per_file": T, "compressed": args.compressed, "runs": []}
        for P in [int(p) for p in args.procs.split(",")]:
            walls, frames, profs, utils = {}, {}, {}, {}
            for R in reps:
                cmds = []
                for k in range(P):
                    mine = scp[k::P] * R
                    tag = "p%d_r%d_k%d" % (P, R, k)
                    open(os.path.join(w, tag + ".scp"), "w").write("\n".join(mine) + "\n")
                    os.makedirs(os.path.join(w, tag), exist_ok=True)
                    cmds.append(([exe, "-T", "1", "-u", "tmvw", "-p", str(k + 1), "-H", os.path.join(w, "mmf"), "-I",
                                  os.path.join(w, "labs.mlf"), "-S", os.path.join(w, tag + ".scp"), "-M", os.path.join(w, tag),
                                  os.path.join(w, "list")], len(mine)))
                util = open(os.path.join(w, "util.csv"), "w+")
                smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=utilization.gpu", "--format=csv,noheader,nounits",
                                        "-lms", "100"], stdout=util, stderr=subprocess.DEVNULL)
                time.sleep(0.5)
                t0 = time.time()
                ps = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for c, _ in cmds]
                outs = [p.communicate()[0] for p in ps]
                walls[R] = time.time() - t0
                smi.terminate(); smi.wait()
                util.seek(0)
                v = [float(x) for x in util.read().split() if x.strip().isdigit()]
                util.close()
                for p, o in zip(ps, outs):
                    if p.returncode != 0:
                        raise RuntimeError(o[-1500:])
                frames[R] = sum(n for _, n in cmds) * T
                loops = []
                for o in outs:
                    m = re.search(r"file loop ([0-9.]+) .*final flush ([0-9.]+); download \+ scatter ([0-9.]+)", o)
                    loops.append(sum(float(x) for x in m.groups()) if m else None)
                profs[R] = loops
                busy = [x for x in v if x > 0]
                utils[R] = {"mean_while_busy_pct": float(np.mean(busy)) if busy else 0.0, "busy_samples": len(busy), "samples": len(v)}
            lo, hi = min(reps), max(reps)
            rate = (frames[hi] - frames[lo]) / max(walls[hi] - walls[lo], 1e-6)
            rec = {"procs": P, "marginal_frames_per_s": rate, "wall_s": walls, "frames": frames,
                   "slowest_file_loop_s": {r: max(x for x in profs[r] if x is not None) for r in reps}, "gpu_utilization": utils,
                   "loop_frames_per_s": frames[hi] / max(x for x in profs[hi] if x is not None)}
            out["runs"].append(rec)
            print(json.dumps(rec), flush=True)
        print(json.dumps(out))
    finally:
        shutil.rmtree(w, ignore_errors=True)


if __name__ == "__main__":
    main()
