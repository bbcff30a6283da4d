"""A/B of the end-to-end pass on one GPU: FP32 host tables against the 16-bit integers of `_C` compressed files
(hfbgpu_submit against hfbgpu_submit_compressed), alternating, cfg3 by default.  python tools/e2e_ab.py [workload] [K]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from htk_b200.estep import ForwardBackward

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = dict(bench.WORKLOADS[name])
T, Q = cfg["T"], cfg["Q"]
n = cfg.get("utts") or 1024
fm = bench.make_model(cfg)
dev = torch.device("cuda", 0)
fb = ForwardBackward(fm, prune=cfg.get("prune"), device=0, workspace_bytes=int(cfg.get("workspace_gb", 0)) << 30)
stream = torch.cuda.current_stream()
fb.set_stream(stream.cuda_stream)
batch, dfeat = bench.make_batch(fm, cfg, n, seed=1000, device=dev)
cf = bench.compress_batch(batch, dfeat, n, T, fm.D)
legs = {"device": lambda: fb.Submit(batch, device_feat_ptr=dfeat.data_ptr()), "host": lambda: fb.Submit(batch),
        "compressed": lambda: fb.SubmitCompressed(batch, cf)}
for f in legs.values():
    for _ in range(5):
        f()
    fb.Wait()
# host cost of a submit when a wave slot is free (nothing to wait for): three submits after an idle point
for k, f in legs.items():
    fb.Wait(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); f(); ts.append(1e3 * (time.perf_counter() - t0))
    fb.Wait()
    print("%-10s host ms of three unblocked submits: %s" % (k, " ".join("%.3f" % t for t in ts)), flush=True)
for rep in range(5):
    sampler = None
    if rep >= 3:                      # the last two repetitions with bench.py's nvidia-smi sampler running beside them
        sampler = bench.ClockSampler(0)
        time.sleep(1.5 if rep == 3 else 0.0)
        print("-- nvidia-smi -lms sampler on%s" % (" (1.5 s head start)" if rep == 3 else " (started just now)"), flush=True)
    for k, f in legs.items():
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        tk = [f() for _ in range(K)]
        t1 = time.time()
        fb.Wait()
        e1.record(stream)
        torch.cuda.synchronize()
        print("%-10s ms/step %.3f   (host time in the %d submits: %.3f ms each; wall %.3f)" % (k, e0.elapsed_time(e1) / K, K, 1e3 * (t1 - t0) / K, 1e3 * (time.time() - t0) / K), flush=True)
    if sampler is not None:
        print("   sampler:", sampler.stop(), flush=True)
fb.close()
