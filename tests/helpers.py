"""Shared test helpers (CSR construction from {dim: value} dicts, comparisons)."""
import numpy as np


def csr_from_dicts(vecs):
    indptr = [0]
    idx, val = [], []
    for v in vecs:
        for d in sorted(v):
            idx.append(d)
            val.append(v[d])
        indptr.append(len(idx))
    return (np.asarray(indptr, np.int64), np.asarray(idx, np.int32), np.asarray(val, np.float64))


def csr_rows(csr, rows):
    indptr, idx, val = csr
    out_ptr = [0]
    oi, ov = [], []
    for r in rows:
        sl = slice(indptr[r], indptr[r + 1])
        oi.append(idx[sl]); ov.append(val[sl])
        out_ptr.append(out_ptr[-1] + (indptr[r + 1] - indptr[r]))
    return (np.asarray(out_ptr, np.int64),
            np.concatenate(oi).astype(np.int32) if oi else np.zeros(0, np.int32),
            np.concatenate(ov).astype(np.float64) if ov else np.zeros(0, np.float64))


def csr_slice(csr, lo, hi):
    indptr, idx, val = csr
    a, b = int(indptr[lo]), int(indptr[hi])
    return (indptr[lo:hi + 1] - indptr[lo], idx[a:b], val[a:b])


def assert_pairs_equal(got, want, rel=0.0):
    """got / want: {(q, c): sim}.  rel=0 -> bit-exact similarities."""
    gk, wk = set(got), set(want)
    assert gk == wk, "pair sets differ: missing %s extra %s" % (sorted(wk - gk)[:5], sorted(gk - wk)[:5])
    for k in wk:
        if rel == 0.0:
            assert got[k] == want[k], (k, got[k], want[k])
        else:
            assert abs(got[k] - want[k]) <= rel * abs(want[k]), (k, got[k], want[k])
