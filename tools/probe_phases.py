"""Per-phase cycle breakdown of the dense-head scoring kernel on a C3-shaped index."""
import json
import sys

import torch

sys.path.insert(0, ".")
import apss_b200
from apss_b200 import native, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
variants = sys.argv[2] if len(sys.argv) > 2 else "3,0,16,2,32"
PRUNE = len(sys.argv) > 3 and sys.argv[3] == "prune"
cfg = synth.CONFIGS["C3"]
D, t, B = cfg["D"], cfg["threshold"], cfg["batch"]
data = synth.generate(N + B, D, cfg["nnz_mean"], seed=cfg["seed"], device="cuda")
torch.cuda.synchronize()


def rows(lo, hi):
    b = data.rows(lo, hi)
    r = (b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous())
    torch.cuda.synchronize()
    return r


for v in variants.split(";"):
    algo, tile, warps, unroll, qb = [int(x) for x in v.split(",")]
    g = native.Index(D, t, tile_vectors=tile, kernel_variant=(qb << 24) | (algo << 16) | (warps << 8) | unroll, pruning=PRUNE)
    for lo in range(0, N, B):
        g.insert_batch(*rows(lo, min(N, lo + B)), index_only=True)
    r = g.insert_batch(*rows(N, N + B), query_only=True)
    st = g.stats()
    ph = st["phase_cycles"]
    rare = ph[6]; ph = ph[:6] + [ph[7]]
    tot = sum(ph) or 1
    names = ["setup", "phase1 lookup+short", "Wq+dense FFMA", "segments", "epilogue loop", "barrier+fetch", "selfclear+sync"]
    print(json.dumps(dict(variant=v, pruning=PRUNE, unindexed=st['n_unindexed'], postings=st['n_postings'], tile=st["tile_vectors"], score_ms=r.score_ms, postings_per_s=r.postings_visited / (r.score_ms * 1e-3), prefilter=r.n_prefilter, pairs=r.n_pairs,
                          items=r.work_items, rare_path_thread0=rare, cycles_per_item=tot / max(r.work_items, 1),
                          phases={n: round(100.0 * c / tot, 1) for n, c in zip(names, ph)})), flush=True)
    g.close()
