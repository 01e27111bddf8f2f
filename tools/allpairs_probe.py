#!/usr/bin/env python
"""The all-pairs job from an empty index through one engine, per-batch timings (debug aid for bench.py's `allpairs`)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from apss_b200 import native, synth
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
cfg = synth.CONFIGS["C3"]; B, D, t = cfg["batch"], cfg["D"], cfg["threshold"]
data = synth.generate(N, D, cfg["nnz_mean"], seed=cfg["seed"], device="cuda"); torch.cuda.synchronize()
eng = native.Index(D, t, pruning=mode, reserve_vectors=int(N * 1.1) + 2 * B, reserve_nnz=int(data.nnz * 1.15) + (1 << 20))
t0 = time.time(); tot = 0
for lo in range(0, N, B):
    b = data.rows(lo, min(N, lo + B))
    w = time.time()
    r = eng.insert_batch(b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous())
    tot += r.n_pairs
    print("batch %3d wall %.2f ms device %.2f score %.2f prefilter %d pairs %d segs %d" % (lo // B, (time.time() - w) * 1e3, r.device_ms, r.score_ms, r.n_prefilter, r.n_pairs, eng.stats()["n_tiles"]), flush=True)
print("total %.3f s pairs %d" % (time.time() - t0, tot))
