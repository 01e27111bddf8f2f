import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
import apss_b200
from apss_b200 import native, synth
cfg = synth.CONFIGS["C3"]; D, t, B = cfg["D"], cfg["threshold"], cfg["batch"]
N = 300000
data = synth.generate(N + 8 * B, D, cfg["nnz_mean"], seed=cfg["seed"], device="cuda"); torch.cuda.synchronize()
def rows(lo, hi):
    b = data.rows(lo, hi); r = (b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous()); torch.cuda.synchronize(); return r
g = native.Index(D, t)
for lo in range(0, N, B): g.insert_batch(*rows(lo, min(N, lo + B)), index_only=True)
cur = N
for mode in ["dev", "dev", "host", "host", "dev", "host"]:
    r_ = rows(cur, cur + B)
    if mode == "host":
        hb = data.rows(cur, cur + B).pin(); args = (hb.indptr.numpy(), hb.indices.numpy(), hb.values.numpy())
    else:
        args = r_
    t0 = time.perf_counter(); r = g.insert_batch(*args); t1 = time.perf_counter()
    q, c, s = g.fetch_pairs(); t2 = time.perf_counter()
    print(mode, "wall_ms %.1f fetch_ms %.1f score_ms %.1f device_ms %.1f pairs %d pf %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, r.score_ms, r.device_ms, r.n_pairs, r.n_prefilter), flush=True)
    cur += B
