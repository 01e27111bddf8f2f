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

# ---- would a position-based admission filter (L2AP's "remaining score" bound) shrink the candidate tables?
# A pair (q, c) first met at c's indexed component k (components taken in descending weight) can only reach t if
# |q| * sqrt(|c_U|^2 + sum_{indexed j with w_j <= w_k} w_j^2) >= t.  Count the touched pairs that pass.
if True:
    frac = 0.8
    lim = frac * (t * t) * (1 - 2.0 ** -20)
    unindexed = run <= lim
    keep = ~unindexed
    cu2 = np.bincount(r_s[unindexed], weights=sq[unindexed], minlength=N)
    # per vector: indexed weights sorted descending and the mass of the components not heavier than each
    ri, wi, di = r_s[keep], v_s[keep], ix_s[keep]
    o2 = np.lexsort((-wi, ri))
    ri, wi, di = ri[o2], wi[o2], di[o2]
    cnt = np.bincount(ri, minlength=N); ptr = np.concatenate([[0], np.cumsum(cnt)])
    csq = np.cumsum(wi * wi)
    tot_row = csq[ptr[1:] - 1] - (csq[ptr[:-1]] - (wi * wi)[ptr[:-1]]) if len(wi) else np.zeros(N)
    # suffix mass from position p (inclusive) in descending order
    start_c = np.repeat(csq[ptr[:-1]] - (wi * wi)[ptr[:-1]], cnt)
    suf = np.repeat(tot_row, cnt) - (csq - wi * wi - start_c)
    # inverted (reduced) index with the suffix mass attached to every posting
    oi = np.argsort(di, kind="stable")
    pd, pr, ps = di[oi], ri[oi], suf[oi]
    pptr = np.concatenate([[0], np.cumsum(np.bincount(pd, minlength=D))])
    rng = np.random.default_rng(7)
    qs = rng.choice(N, min(args.queries, 256), replace=False)
    touched_n = admitted_n = 0
    best = np.zeros(N)
    for qi_ in qs:
        a, b = ip[qi_], ip[qi_ + 1]
        qn_ = float(np.sqrt(np.sum(v[a:b] ** 2)))
        cands = []
        for d in ix[a:b]:
            s0, s1 = pptr[d], pptr[d + 1]
            if s1 > s0:
                np.maximum.at(best, pr[s0:s1], ps[s0:s1])        # the heaviest shared component has the largest suffix mass
                cands.append(pr[s0:s1])
        if not cands:
            continue
        cs_ = np.unique(np.concatenate(cands)); cs_ = cs_[cs_ != qi_]
        touched_n += len(cs_)
        admitted_n += int(np.sum(qn_ * np.sqrt(cu2[cs_] + best[cs_]) >= t))
        best[cs_] = 0.0; best[qi_] = 0.0
    print("---- position-based admission (alpha 0.8): touched %.0f per query, admitted by the remaining-score bound %.0f per query (%.1f%%)" % (
        touched_n / len(qs), admitted_n / len(qs), 100.0 * admitted_n / max(touched_n, 1)))

# ---- the reference's own plan: per-dimension max weights (EPA:51-57, HBaseUpLoader.scala:113-129).  How much could
# stay out of the index under  sum_{d in U} c[d] * maxw[d] < t  compared with the Cauchy-Schwarz bound (alpha -> 1)?
if True:
    maxw = np.zeros(D); np.maximum.at(maxw, ix, v)
    mw_s = v_s * maxw[ix_s]
    cmw = np.cumsum(mw_s)
    start_mw = cmw[ip[:-1]] - mw_s[ip[:-1]]
    run_mw = cmw - np.repeat(start_mw, np.diff(ip))
    un_mw = run_mw < t
    un_l2 = run <= (t * t) * (1 - 2.0 ** -20)
    either = un_mw | un_l2
    for name, m in (("max-weight bound", un_mw), ("Cauchy-Schwarz bound (alpha = 1)", un_l2), ("either", either)):
        dfi = np.bincount(ix_s[~m], minlength=D)
        print("---- %-34s un-indexed %.1f%% of components, postings visited %.3e" % (name, 100 * m.mean(), float(np.sum(df.astype(np.float64) * dfi))))
