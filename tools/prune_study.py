"""CPU study for SURVEY 8(f)-3: how much posting traffic does exact index reduction remove, and how many
candidates survive the partial-score filter?  (numpy/scipy only; no reference code, no GPU.)

For every vector c the features are ranked by document frequency (descending) and the longest prefix U_c with
sum_{d in U_c} c[d]^2 <= (t / qnorm_max)^2 stays out of the index (Cauchy-Schwarz: dot(q, c_U) < t)."""
import argparse
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, ".")
import apss_b200
from apss_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
ap.add_argument("--n", type=int, default=200_000)
ap.add_argument("--queries", type=int, default=512)
args = ap.parse_args()
cfg = synth.CONFIGS[args.config]
t = cfg["threshold"]
t0 = time.time()
ip, ix, v = synth.generate_config(args.config, N=args.n).numpy()
N, D = len(ip) - 1, cfg["D"]
print("generated", N, "vectors,", len(ix), "features in %.1fs" % (time.time() - t0))
rows = np.repeat(np.arange(N), np.diff(ip))
df = np.bincount(ix, minlength=D)
# rank features inside each vector by df descending (ties: dim ascending)
order = np.lexsort((ix, -df[ix], rows))
r_s, ix_s, v_s = rows[order], ix[order], v[order]
sq = v_s * v_s
cs = np.cumsum(sq)
start = cs[ip[:-1]] - sq[ip[:-1]]
run = cs - np.repeat(start, np.diff(ip))                      # inclusive running sum of squares in rank order
for frac in (0.5, 0.8, 0.9, 1.0):
    lim = frac * (t * t) * (1 - 2.0 ** -20)
    print("---- alpha", frac)
    unindexed = run <= lim
    print("threshold %.2f: %.1f%% of features unindexed" % (t, 100 * unindexed.mean()))
    df_idx = np.bincount(ix_s[~unindexed], minlength=D)
    full = float(np.sum(df.astype(np.float64) ** 2))
    red = float(np.sum(df.astype(np.float64) * df_idx))
    print("postings visited (all-pairs, full batch semantics): full %.3e  pruned-index %.3e  ratio %.1fx" % (full, red, full / red))
    cu = np.sqrt(np.bincount(r_s[unindexed], weights=sq[unindexed], minlength=N))   # ||c_U||
    A = sp.csr_matrix((v, ix, ip), shape=(N, D))
    I = sp.csr_matrix((v_s[~unindexed], (r_s[~unindexed], ix_s[~unindexed])), shape=(N, D))
    rng = np.random.default_rng(7)
    qs = rng.choice(N, args.queries, replace=False)
    Q = A[qs]
    S = (Q @ I.T).tocsr()                                      # suffix dots
    Sfull = (Q @ A.T).tocsr()
    qn = np.sqrt(np.asarray(Q.multiply(Q).sum(axis=1)).ravel())
    touched = S.nnz
    touched_full = Sfull.nnz
    coo = S.tocoo()
    ub = cu[coo.col] * qn[coo.row]
    surv = int(np.sum(coo.data + ub >= t)); surv2 = int(np.sum(coo.data + cu[coo.col] >= t))
    # tighter: ||q restricted to dims of high df|| is unknown per candidate; try bound with q's own prefix norm
    true = int(np.sum(Sfull.data >= t)) - args.queries
    print("per query: touched full %.0f, touched pruned %.0f, survivors(acc+|c_U||q| >= t) %.1f (with qn=1: %.1f), true pairs %.2f" % (
        touched_full / args.queries, touched / args.queries, surv / args.queries, surv2 / args.queries, true / args.queries))
    for thr in (0.05, 0.1, 0.2, 0.3):
        print("   acc >= %.2f: %.1f per query" % (thr, np.sum(coo.data >= thr) / args.queries))
    # recall check: every true pair must be touched
    F = Sfull.tocoo()
    m = F.data >= t
    touched_set = set(zip(coo.row.tolist(), coo.col.tolist()))
    miss = sum(1 for r, c in zip(F.row[m].tolist(), F.col[m].tolist()) if (r, c) not in touched_set and qs[r] != c)
    print("true pairs not touched through the reduced index:", miss)
