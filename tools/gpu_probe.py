"""First-contact GPU probe: accumulator micro-benchmarks + scoring-kernel variant sweep.
Usage: python tools/gpu_probe.py [N] [out.json]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import apss_b200
from apss_b200 import native, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
out_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/probe.json"
res = {"gpu": torch.cuda.get_device_name(0), "N": N}

mb = {}
names = {0: "rmw_f32_random", 1: "rmw_f32_consecutive", 2: "atoms_u32_random", 3: "atoms_u32_consecutive",
         4: "atomicAdd_f32_cas_random", 5: "ffma_f2i_atoms_random"}
for warps in ((8, 16, 32) if "--mb" in sys.argv else ()):
    for mode in range(6):
        v = native.microbench_accumulators(mode, warps=warps, iters=4000)
        mb["%s_w%d" % (names[mode], warps)] = v
        print("microbench %-28s warps=%2d  %.3e updates/s  (%.2f /clk/SM @1.9GHz)" % (names[mode], warps, v, v / 148 / 1.9e9), flush=True)
res["microbench"] = mb

cfg = synth.CONFIGS["C3"]
D, t, B = cfg["D"], cfg["threshold"], cfg["batch"]
t0 = time.time()
data = synth.generate(N + 2 * B, D, cfg["nnz_mean"], seed=cfg["seed"], device="cuda")
torch.cuda.synchronize()
print("generated %d vectors, %d nnz in %.1fs" % (data.n, data.nnz, time.time() - t0), flush=True)
res["gen_s"] = time.time() - t0


def dev_rows(lo, hi):
    b = data.rows(lo, hi)
    return b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous()


sweep = []
SWEEP = [(1, 3584, 16, 8, 0), (2, 0, 16, 0, 16), (2, 0, 32, 0, 16), (2, 0, 8, 0, 16), (2, 0, 16, 0, 32), (2, 0, 32, 0, 32), (2, 0, 16, 0, 8), (2, 0, 32, 0, 8)]
if len(sys.argv) > 3:
    SWEEP = [tuple(int(x) for x in v.split(",")) for v in sys.argv[3].split(";")]
for algo, tile, warps, unroll, qb in SWEEP:
    g = native.Index(D, t, tile_vectors=tile, kernel_variant=(qb << 24) | (algo << 16) | (warps << 8) | unroll, reserve_vectors=N + 3 * B, reserve_nnz=int(data.nnz * 1.05))
    g2 = g
    torch.cuda.synchronize()
    # load the index by inserting (scores too); time separately
    t0 = time.time()
    loaded = 0
    for lo in range(0, N, B):
        hi = min(N, lo + B)
        torch.cuda.synchronize()
        r = g2.insert_batch(*dev_rows(lo, hi), index_only=(lo + B < N - 2 * B))
        loaded += r.postings_visited
    load_s = time.time() - t0
    st = g2.stats()
    # timed: 2 query-only batches against the full index
    recs = []
    for k in range(2):
        rows = dev_rows(N + k * B, N + (k + 1) * B)
        torch.cuda.synchronize()
        r = g2.insert_batch(*rows, query_only=True)
        recs.append(r)
    r = recs[-1]
    rate = r.postings_visited / (r.score_ms * 1e-3)
    tile = st['tile_vectors']
    row = dict(algo=algo, qb=qb, tile=tile, warps=warps, unroll=unroll, load_s=load_s, load_score_ms=st["tot_score_ms"], load_postings=loaded,
               load_rate=loaded / (st["tot_score_ms"] * 1e-3) if st["tot_score_ms"] else 0,
               q_postings=r.postings_visited, q_cands=r.candidates_unique, q_pairs=r.n_pairs, q_prefilter=r.n_prefilter,
               q_score_ms=r.score_ms, q_device_ms=r.device_ms, postings_per_s=rate, alg_GBs=rate * 8 / 1e9,
               frac_of_6548=rate * 8 / 1e9 / 6548.2, dir_bytes=st["bytes_directory"], post_bytes=st["bytes_postings"])
    sweep.append(row)
    print(json.dumps(row), flush=True)
    g.close()
res["sweep"] = sweep
json.dump(res, open(out_path, "w"), indent=1)
print("done")
