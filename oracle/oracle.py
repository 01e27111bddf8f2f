"""ctypes front-end of the CPU oracle (oracle/apss_oracle.c).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (the reference ships no tests / golden vectors and
cannot be run here: no JVM).  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs only; never from all-pairs-similarity_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libapss_oracle.so")

R1, R0 = 0, 1
ALGO_FAITHFUL, ALGO_FAST = 0, 1
FLAG_QUERY_ONLY = 1
FLAG_INDEX_ONLY = 2
FLAG_SKIP_ADMIT = 4
ST_REJECTED, ST_EMPTY, ST_ACTIVE = 0, 1, 2
SET_ORDER_SCALA, SET_ORDER_ASCENDING = 0, 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "apss_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_void_p]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_freeze.argtypes = [C.c_void_p]
        L.oracle_set_threads.argtypes = [C.c_void_p, C.c_int32]
        L.oracle_max_threads.restype = C.c_int32
        L.oracle_set_pruning.restype = C.c_int32
        L.oracle_set_pruning.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.oracle_n_unindexed.restype = C.c_int64
        L.oracle_n_unindexed.argtypes = [C.c_void_p]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_last_error.argtypes = [C.c_void_p]
        L.oracle_insert_batch.restype = C.c_int32
        L.oracle_insert_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.oracle_n_pairs.restype = C.c_int64
        L.oracle_n_pairs.argtypes = [C.c_void_p]
        L.oracle_id_base.restype = C.c_int64
        L.oracle_id_base.argtypes = [C.c_void_p]
        L.oracle_n_vectors.restype = C.c_int64
        L.oracle_n_vectors.argtypes = [C.c_void_p]
        L.oracle_fetch_pairs.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.oracle_fetch_status.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_set_iteration_order.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        L.oracle_bruteforce.restype = C.c_int64
        L.oracle_bruteforce.argtypes = [C.c_int32] + [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 4 + [
            C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class BatchResult:
    q: np.ndarray          # query index within the batch
    c: np.ndarray          # candidate ordinal (global insertion order)
    qkey: np.ndarray
    ckey: np.ndarray
    sim: np.ndarray
    status: np.ndarray     # per input vector: ST_*
    id_base: int
    postings_visited: int
    candidates_unique: int  # -1 when the algorithm does not define it (faithful)
    dot_calls_ref: int      # -1 for ALGO_FAST

    def pair_set(self):
        """{(query ordinal, candidate ordinal): sim}"""
        return {(int(self.id_base + q), int(c)): float(s) for q, c, s in zip(self.q, self.c, self.sim)}

    def key_pair_set(self):
        """{(qkey, ckey): sim} -- the shape of SimilarityOutput.output (Message.scala:20-21)"""
        return {(int(a), int(b)): float(s) for a, b, s in zip(self.qkey, self.ckey, self.sim)}


class Oracle:
    def __init__(self, dim, similarity_threshold, index_threshold=0.0, semantics=R1, algo=ALGO_FAITHFUL,
                 max_shard_num=1, max_entry_num=1, max_index_entry_actor_num=1, set_order=SET_ORDER_SCALA,
                 max_weight=None, threads=1, pruning=False, prune_alpha=0.0, max_query_norm=0.0):
        self._L = lib()
        mw = None if max_weight is None else np.ascontiguousarray(max_weight, dtype=np.float64)
        if algo == ALGO_FAST and semantics != R1:
            raise ValueError("ALGO_FAST implements R1 only")
        self._h = self._L.oracle_create(int(dim), float(similarity_threshold), float(index_threshold), int(semantics),
                                        int(algo), int(max_shard_num), int(max_entry_num),
                                        int(max_index_entry_actor_num), int(set_order), _p(mw))
        self.dim = dim
        self._L.oracle_set_threads(self._h, int(threads))
        if pruning and self._L.oracle_set_pruning(self._h, float(prune_alpha), float(max_query_norm)) != 0:
            raise ValueError("pruning needs ALGO_FAST, 0 < alpha < 1 and max_query_norm > 0")

    def close(self):
        if self._h:
            self._L.oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def freeze(self):
        self._L.oracle_freeze(self._h)

    @property
    def n_unindexed(self):
        """components kept out of the index by exact index reduction so far"""
        return int(self._L.oracle_n_unindexed(self._h))

    @property
    def n_vectors(self):
        return int(self._L.oracle_n_vectors(self._h))

    def insert_batch(self, indptr, indices, values, keys=None, query_only=False, index_only=False, skip_admit=False) -> BatchResult:
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.float64)
        n = len(indptr) - 1
        k = None if keys is None else np.ascontiguousarray(keys, dtype=np.int64)
        rc = self._L.oracle_insert_batch(self._h, n, _p(indptr), _p(indices), _p(values), _p(k),
                                         (FLAG_QUERY_ONLY if query_only else 0) | (FLAG_INDEX_ONLY if index_only else 0) |
                                         (FLAG_SKIP_ADMIT if skip_admit else 0))
        if rc != 0:
            raise ValueError("oracle: %s (rc=%d)" % (self._L.oracle_last_error(self._h).decode(), rc))
        m = int(self._L.oracle_n_pairs(self._h))
        q = np.empty(m, np.int32); c = np.empty(m, np.int32)
        qk = np.empty(m, np.int64); ck = np.empty(m, np.int64); s = np.empty(m, np.float64)
        self._L.oracle_fetch_pairs(self._h, _p(q), _p(c), _p(qk), _p(ck), _p(s))
        st = np.empty(max(n, 1), np.uint8)
        self._L.oracle_fetch_status(self._h, _p(st))
        cnt = np.zeros(7, np.int64)
        self._L.oracle_counters(self._h, _p(cnt))
        return BatchResult(q, c, qk, ck, s, st[:n], int(self._L.oracle_id_base(self._h)), int(cnt[0]), int(cnt[1]), int(cnt[2]))

    def totals(self):
        cnt = np.zeros(7, np.int64)
        self._L.oracle_counters(self._h, _p(cnt))
        return {"postings_visited": int(cnt[3]), "candidates_unique": int(cnt[4]), "dot_calls_ref": int(cnt[5]), "pairs": int(cnt[6])}


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def set_iteration_order(dims, n_total=None, mode=SET_ORDER_SCALA):
    """Iteration order of the Scala Set holding `dims` (ascending ints), see apss_oracle.c."""
    d = np.ascontiguousarray(dims, dtype=np.int32)
    out = np.empty_like(d)
    lib().oracle_set_iteration_order(int(mode), int(len(d) if n_total is None else n_total), _p(d), len(d), _p(out))
    return out


def first_dims(indptr, indices, values, index_threshold=0.0, mode=SET_ORDER_SCALA):
    """first(q) for the single-worker parity configuration P0: the first element, in Set iteration
    order, of the dims of the PRUNED vector (WWA:192-193 then WWA:172 / EPA:43 / IWA:102).  -1 if empty."""
    n = len(indptr) - 1
    out = np.full(n, -1, np.int32)
    for v in range(n):
        sl = slice(indptr[v], indptr[v + 1])
        d = np.asarray(indices[sl])[np.asarray(values[sl]) > index_threshold]
        if len(d):
            out[v] = set_iteration_order(d, mode=mode)[0]
    return out


def bruteforce(q_csr, c_csr, thr, qkeys=None, ckeys=None, first_dim=None):
    """O(nq*nc) dense check.  Returns ({(q, c): sim}, n_candidates)."""
    qp, qi, qv = [np.ascontiguousarray(a, dtype=t) for a, t in zip(q_csr, (np.int64, np.int32, np.float64))]
    cp, ci, cv = [np.ascontiguousarray(a, dtype=t) for a, t in zip(c_csr, (np.int64, np.int32, np.float64))]
    nq, nc = len(qp) - 1, len(cp) - 1
    qk = np.arange(nq, dtype=np.int64) if qkeys is None else np.ascontiguousarray(qkeys, dtype=np.int64)
    ck = np.arange(nc, dtype=np.int64) if ckeys is None else np.ascontiguousarray(ckeys, dtype=np.int64)
    fd = None if first_dim is None else np.ascontiguousarray(first_dim, dtype=np.int32)
    cap = max(1024, nq * 8)
    while True:
        oq = np.empty(cap, np.int32); oc = np.empty(cap, np.int32); os_ = np.empty(cap, np.float64)
        ncand = C.c_int64(0)
        m = lib().oracle_bruteforce(nq, _p(qp), _p(qi), _p(qv), _p(qk), nc, _p(cp), _p(ci), _p(cv), _p(ck),
                                    float(thr), _p(fd), _p(oq), _p(oc), _p(os_), cap, C.byref(ncand))
        if m <= cap:
            break
        cap = int(m)
    return {(int(a), int(b)): float(s) for a, b, s in zip(oq[:m], oc[:m], os_[:m])}, int(ncand.value)
