"""Pure-Python, line-by-line restatement of the reference's index worker for TINY cases.

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED.  A second, independent restatement (dicts and sets
standing in for the Scala collections) used to cross-check oracle/apss_oracle.c on the known-answer
tests of SURVEY.md 8(c) and on small random inputs.  Paths below are relative to
/root/reference/core/src/main/scala/cpslab/.
"""
from __future__ import annotations

import time
from typing import Dict, List, Sequence, Tuple

M32 = 0xFFFFFFFF


def scala_improve(hcode: int) -> int:
    """scala.collection.immutable.HashSet.improve, Scala 2.10.4 (UNVERIFIED recall)."""
    h = (hcode + (~(hcode << 9) & M32)) & M32
    h ^= h >> 14
    h = (h + (h << 4)) & M32
    return h ^ (h >> 10)


def scala_set_order(dims: Sequence[int], n_total: int | None = None, ascending: bool = False) -> List[int]:
    """Iteration order of `indices.toSet.filter(..)`: Set1..Set4 keep insertion (= ascending) order;
    >= 5 elements make a HashSet (32-ary trie, 5-bit digits of the improved hash from the low end)."""
    n_total = len(dims) if n_total is None else n_total
    if ascending or n_total <= 4:
        return list(dims)
    return sorted(dims, key=lambda x: [(scala_improve(int(x) & M32) >> (5 * lvl)) & 31 for lvl in range(7)])


class SparseVector:
    """vector/SparseVector.scala:198-223 (size, ascending indices, values)"""

    def __init__(self, size: int, indices: Sequence[int], values: Sequence[float]):
        prev = -1
        for i in indices:                       # SparseVector.scala:100-104
            if not prev < i:
                raise ValueError("Found duplicate indices: %d." % i)
            prev = i
        if not prev < size:                     # SparseVector.scala:105
            raise ValueError("index out of range")
        self.size, self.indices, self.values = size, list(indices), list(values)


def calculate_similarity(v1: SparseVector, v2: SparseVector) -> float:
    """CommonUtils.scala:98-117 (iteration of vector1Map in ascending index order, see oracle header)"""
    if v1.size != v2.size:                      # CU:99
        raise ValueError("vector1 size: %d, vector2 size: %d" % (v1.size, v2.size))
    similarity = 0.0
    m1 = {i: x for i, x in zip(v1.indices, v1.values)}   # CU:102-105
    m2 = {i: x for i, x in zip(v2.indices, v2.values)}   # CU:106-109
    for idx in sorted(m1):                               # CU:110
        if idx in m2:
            similarity += m1[idx] * m2[idx]              # CU:111-114
    return similarity


class IndexingWorker:
    """deploy/server/IndexingWorkerActor.scala:21-149 without Akka."""

    def __init__(self, similarity_threshold: float, as_built: bool = True):
        self.vectors_store: List[Tuple[List[int], Tuple[str, SparseVector]]] = []   # IWA:22
        self.similarity_threshold = similarity_threshold                            # IWA:23
        self.inverted_index: Dict[int, List[int]] = {}                              # IWA:25
        self.stop_update_index = False                                              # IWA:35
        self.as_built = as_built
        self.dot_calls = 0
        self.postings_walked = 0

    def build_inverted_index(self, wrappers):             # IWA:61-71
        for w in wrappers:
            self.vectors_store.append(w)
            current_idx = len(self.vectors_store) - 1
            for d in w[0]:
                self.inverted_index.setdefault(d, []).append(current_idx)

    def query_similar_items(self, wrappers):              # IWA:74-111
        output_sim_set: Dict[str, Dict[str, float]] = {}

        def query_similar_vectors(query, candidate_list):  # IWA:80-99
            qid = query[1][0]
            sim_map: Dict[str, float] = {}
            for cidx in candidate_list:                    # IWA:86
                cand = self.vectors_store[cidx]            # IWA:87
                cid = cand[1][0]
                if qid in output_sim_set and cid not in output_sim_set[qid] and qid != cid:   # IWA:89-91
                    sim = calculate_similarity(cand[1][1], query[1][1])                       # IWA:92
                    self.dot_calls += 1
                    if sim >= self.similarity_threshold:   # IWA:93
                        sim_map[cid] = sim
            return sim_map

        for w in wrappers:                                 # IWA:101
            if not self.as_built:                          # R1: the entry exists before the first list
                output_sim_set.setdefault(w[1][0], {})
            for d in w[0]:                                 # IWA:102
                lst = self.inverted_index.get(d, [])       # IWA:104 (Q15: missing -> empty)
                self.postings_walked += len(lst)
                similar = query_similar_vectors(w, lst)
                output_sim_set.setdefault(w[1][0], {}).update(similar)   # IWA:106-107
        return output_sim_set

    def receive_index_data(self, wrappers):                # IWA:123-137 (immediate mode)
        if not self.stop_update_index:
            self.build_inverted_index(wrappers)
        return self.query_similar_items(wrappers), int(time.time() * 1000)

    def receive_timeout(self):                             # IWA:143-144
        self.stop_update_index = True


class Pipeline:
    """EPA admission -> WWA prune/batch -> routing -> IndexingWorker(s), parity configuration P0 style:
    one insert_batch call = one IOTrigger tick = one IndexData per worker."""

    def __init__(self, vector_dim, similarity_threshold, index_threshold=0.0, as_built=True,
                 max_shard_num=1, max_index_entry_actor_num=1, ascending_set_order=False):
        self.vector_dim, self.t, self.index_threshold = vector_dim, similarity_threshold, index_threshold
        self.max_shard_num, self.max_index = max_shard_num, max_index_entry_actor_num
        self.ascending = ascending_set_order
        self.workers: Dict[Tuple[int, int], IndexingWorker] = {}
        self.as_built = as_built

    def admit(self, v: SparseVector) -> bool:              # EPA:81-93 with the stub of EPA:51-57
        s = 0.0
        for i, x in zip(v.indices, v.values):
            if 0 <= i < self.vector_dim:
                s += 1.0 * x
        return s >= self.t

    def prune(self, v: SparseVector) -> SparseVector:      # WWA:188-193
        keep = [(i, x) for i, x in zip(v.indices, v.values) if x > self.index_threshold]
        return SparseVector(v.size, [i for i, _ in keep], [x for _, x in keep])

    def insert_batch(self, vectors: Sequence[Tuple[str, SparseVector]]):
        out: Dict[str, Dict[str, float]] = {}
        per_worker: Dict[Tuple[int, int], list] = {}
        for vid, vec in vectors:
            if not self.admit(vec):                        # EPA:97
                continue
            pv = self.prune(vec)                           # WWA:192-194
            n_total = len(pv.indices)
            for s in range(self.max_shard_num):            # WWA:172
                sub = [d for d in pv.indices if d % self.max_shard_num == s]
                if not sub:
                    continue
                for ch in range(self.max_index):           # EPA:41-46
                    dims = scala_set_order([d for d in sub if d % self.max_index == ch], n_total, self.ascending)
                    per_worker.setdefault((s, ch), []).append((dims, (vid, pv)))
        for key, wrappers in per_worker.items():           # EPA:113-122
            wk = self.workers.setdefault(key, IndexingWorker(self.t, self.as_built))
            res, _ = wk.receive_index_data(wrappers)
            for q, m in res.items():
                out.setdefault(q, {}).update(m)
        return out

    def freeze(self):
        for wk in self.workers.values():
            wk.receive_timeout()

    @property
    def dot_calls(self):
        return sum(w.dot_calls for w in self.workers.values())

    @property
    def postings_walked(self):
        return sum(w.postings_walked for w in self.workers.values())


def similarity_output_to_string(output: Dict[str, Dict[str, float]]) -> str:
    """message/Message.scala:23-34 (Double.toString differs from repr() only in corner cases)"""
    sb = []
    for q, m in output.items():
        sb.append("---------------------------------")
        sb.append(q + ":")
        for c, s in m.items():
            sb.append(c + "," + repr(float(s)) + ";")
        sb.append("\n")
    return "".join(sb)


# ---------------------------------------------------------------------------------------------------
# Exact index reduction -- NOT reference behaviour: a second, pure-Python restatement of the
# rule the GPU library applies with `pruning` on (include/apss.h) and oracle/apss_oracle.c restates in C
# (oracle_set_pruning), so that the two can be cross-checked on small inputs.

def index_reduction_limit(similarity_threshold: float, alpha: float = 0.8, max_query_norm: float = 1.0) -> float:
    t = similarity_threshold
    return alpha * t * t / (max_query_norm * max_query_norm) * (1.0 - 2.0 ** -20) if t > 0 else 0.0


def index_reduction_select(indices: Sequence[int], values: Sequence[float], df: Dict[int, int], lim: float) -> List[bool]:
    """skip[i] = True when component i stays OUT of the index: the longest prefix, in (document frequency
    descending, dim ascending) order, whose squared weights sum (left to right, fp64) to <= lim."""
    order = sorted(range(len(indices)), key=lambda i: (-df[int(indices[i])], int(indices[i])))
    skip = [False] * len(indices)
    s = 0.0
    for i in order:
        s2 = s + float(values[i]) * float(values[i])
        if not s2 <= lim:
            break
        s = s2
        skip[i] = True
    return skip


class ReducedIndexPipeline:
    """R1 scoring over the reduced index, batch by batch: returns pairs plus the two counters the library reports
    with pruning on (postings visited / candidates touched through INDEXED components only)."""

    def __init__(self, similarity_threshold: float, alpha: float = 0.8, max_query_norm: float = 1.0):
        self.t = similarity_threshold
        self.lim = index_reduction_limit(similarity_threshold, alpha, max_query_norm)
        self.df: Dict[int, int] = {}
        self.vectors: List[Tuple[List[int], List[float]]] = []
        self.postings: Dict[int, List[int]] = {}            # dim -> ids of the vectors that INDEX it
        self.n_unindexed = 0

    def insert_batch(self, batch: Sequence[Tuple[Sequence[int], Sequence[float]]]):
        base = len(self.vectors)
        for idx, _ in batch:                                # document frequencies include the whole batch
            for d in idx:
                self.df[int(d)] = self.df.get(int(d), 0) + 1
        for idx, val in batch:
            skip = index_reduction_select(idx, val, self.df, self.lim)
            self.n_unindexed += sum(skip)
            vid = len(self.vectors)
            self.vectors.append(([int(d) for d in idx], [float(x) for x in val]))
            for d, sk in zip(idx, skip):
                if not sk:
                    self.postings.setdefault(int(d), []).append(vid)
        pairs, postings, cands = {}, 0, 0
        for k, (idx, val) in enumerate(batch):
            q = base + k
            touched = set()
            for d in idx:
                lst = self.postings.get(int(d), [])
                postings += len(lst)
                touched.update(lst)
            touched.discard(q)
            cands += len(touched)
            qv = SparseVector(1 << 30, [int(d) for d in idx], [float(x) for x in val])
            for c in touched:
                cv = SparseVector(1 << 30, *self.vectors[c])
                sim = calculate_similarity(cv, qv)
                if sim >= self.t:
                    pairs[(q, c)] = sim
        return pairs, postings, cands
