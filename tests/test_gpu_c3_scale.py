"""Parity at the scale that is benchmarked (1 M vectors x 2^18 dims, nnz ~100, t = 0.7):
  (1) FULL parity on the first 100 K vectors, batch by batch (index, then query: IWA:122-134) -- pair sets and fp64
      similarities bit for bit for the parity-mode engine AND for exact index reduction (pruning = 3), plus the
      reference's counters (postings_visited, candidates_unique) in parity mode;
  (2) the whole 1 M-vector index, then a fixed sample of 2 000 of its vectors (seed 7) scored query-only against it:
      pairs, similarities and counters against the CPU oracle.
Everything goes through the C ABI; the data is generator g2 (the same bits on the oracle side and on the GPU)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import assert_pairs_equal

pytestmark = pytest.mark.gpu

N, D, NNZ, T, B, SEED = 1_000_000, 1 << 18, 100, 0.7, 16384, 20260103
N_FULL, N_SAMPLE = 100_000, 2000


@pytest.fixture(scope="module")
def c3():
    import torch
    from apss_b200 import synth
    data = synth.generate(N, D, NNZ, seed=SEED, device="cuda")
    torch.cuda.synchronize()
    return data


def _rows(data, lo, hi):
    b = data.rows(lo, hi)
    return b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous()


def _pairs(g, r):
    q, c, s = g.fetch_pairs()
    return {(int(r.id_base + a), int(b)): float(x) for a, b, x in zip(q, c, s)}


def test_c3_first_100k_full_parity_both_modes(c3):
    from apss_b200 import native as n
    import os
    threads = max(1, len(os.sched_getaffinity(0)))
    o = orc.Oracle(D, T, algo=orc.ALGO_FAST, threads=threads)
    g0 = n.Index(D, T, reserve_vectors=N_FULL + B, reserve_nnz=N_FULL * 110)
    g3 = n.Index(D, T, pruning=3, reserve_vectors=N_FULL + B, reserve_nnz=N_FULL * 110)
    tot = 0
    for lo in range(0, N_FULL, B):
        hi = min(N_FULL, lo + B)
        rows = _rows(c3, lo, hi)
        ro = o.insert_batch(*[x.cpu().numpy() for x in rows])
        want = ro.pair_set()
        r0 = g0.insert_batch(*rows)
        assert_pairs_equal(_pairs(g0, r0), want)
        assert (r0.postings_visited, r0.candidates_unique) == (ro.postings_visited, ro.candidates_unique)
        r3 = g3.insert_batch(*rows)
        assert_pairs_equal(_pairs(g3, r3), want)
        assert r3.postings_visited * 20 < max(r0.postings_visited, 1) or lo == 0
        tot += len(want)
    assert tot > 5000
    assert g3.stats()["n_tiles"] <= 8          # LSM: a handful of posting segments, not one per batch


def test_c3_sampled_queries_against_the_full_1m_index(c3):
    from apss_b200 import native as n
    import os
    threads = max(1, len(os.sched_getaffinity(0)))
    ip, ix, v = c3.numpy()
    o = orc.Oracle(D, T, algo=orc.ALGO_FAST, threads=threads)
    nnz = int(ip[-1])
    g0 = n.Index(D, T, reserve_vectors=N + B, reserve_nnz=int(nnz * 1.02))
    g3 = n.Index(D, T, pruning=3, reserve_vectors=N + B, reserve_nnz=int(nnz * 1.02))
    for lo in range(0, N, B):
        hi = min(N, lo + B)
        o.insert_batch(ip[lo:hi + 1] - ip[lo], ix[ip[lo]:ip[hi]], v[ip[lo]:ip[hi]], index_only=True)
        rows = _rows(c3, lo, hi)
        g0.insert_batch(*rows, index_only=True)
        g3.insert_batch(*rows, index_only=True)
    sample = np.sort(np.random.RandomState(7).choice(N, N_SAMPLE, replace=False))
    qp = np.zeros(N_SAMPLE + 1, np.int64)
    qp[1:] = np.cumsum(ip[sample + 1] - ip[sample])
    take = np.concatenate([np.arange(ip[s], ip[s + 1]) for s in sample])
    qi, qv = ix[take], v[take]
    ro = o.insert_batch(qp, qi, qv, query_only=True)
    want = ro.pair_set()
    assert len(want) >= N_SAMPLE                      # every sampled vector finds at least its own stored copy
    r0 = g0.insert_batch(qp, qi, qv, query_only=True)
    assert_pairs_equal(_pairs(g0, r0), want)
    assert (r0.postings_visited, r0.candidates_unique) == (ro.postings_visited, ro.candidates_unique)
    r3 = g3.insert_batch(qp, qi, qv, query_only=True)
    assert_pairs_equal(_pairs(g3, r3), want)
    assert r3.postings_visited * 100 < r0.postings_visited
    st = g3.stats()
    assert st["n_vectors"] == N and st["n_tiles"] <= 12
