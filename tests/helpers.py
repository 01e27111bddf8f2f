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


class OracleEngine:
    """Test double with the interface of apss_b200.native.Index, backed by the CPU oracle.  Lives in
    tests/ on purpose: the product never routes through the oracle; this only lets the host-side logic
    (message mirror, shard dispatcher) be exercised on a box without a GPU."""

    def __init__(self, dim, similarity_threshold, index_threshold=0.0, as_built=False):
        from oracle import oracle as orc
        self._orc = orc
        self.o = orc.Oracle(dim, similarity_threshold, index_threshold, semantics=orc.R0 if as_built else orc.R1,
                            algo=orc.ALGO_FAITHFUL if as_built else orc.ALGO_FAST, threads=2)
        self.gid = []              # oracle ordinal -> global id
        self.next_id = 0
        self.last = None

    def set_next_id(self, next_id):
        self.next_id = int(next_id)

    def freeze(self):
        self.o.freeze()

    def insert_batch(self, indptr, indices, values, ext_keys=None, first_dim=None, query_only=False, skip_admit=False,
                     index_only=False, n=None):
        import types
        a = [x.numpy() if hasattr(x, "numpy") else np.asarray(x) for x in (indptr, indices, values)]
        nvec = len(a[0]) - 1
        keys = ext_keys
        if keys is None:           # default key = global id
            keys = np.arange(self.next_id, self.next_id + nvec, dtype=np.int64)
        r = self.o.insert_batch(a[0], a[1], a[2], keys=keys, query_only=query_only, index_only=index_only, skip_admit=skip_admit)
        id_base = self.next_id
        if not query_only:
            self.gid.extend(range(self.next_id, self.next_id + nvec))
            self.next_id += nvec
        self.last = r
        return types.SimpleNamespace(id_base=id_base, n_vectors=nvec, n_pairs=len(r.sim), n_pairs_r1=len(r.sim), n_prefilter=len(r.sim),
                                     postings_visited=r.postings_visited, candidates_unique=max(r.candidates_unique, 0),
                                     work_items=0, score_ms=0.0, device_ms=0.0,
                                     n_rejected=int((r.status == 0).sum()), n_empty=int((r.status == 1).sum()), n_active=int((r.status == 2).sum()))

    def fetch_pairs(self):
        r = self.last
        c = np.array([self.gid[int(x)] for x in r.c], np.int32)
        return r.q.astype(np.int32), c, r.sim.copy()

    def fetch_status(self, n):
        return self.last.status[:n]

    def stats(self):
        return {"n_vectors": self.o.n_vectors}
