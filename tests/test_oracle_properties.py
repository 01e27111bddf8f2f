"""Cross-checks of the oracle beyond the KATs (SURVEY.md 8(c) "other oracle cross-checks"): brute-force
O(N^2) dense check, the two C algorithms against each other, the C restatement against the independent
pure-Python one (including dimension routing to several emulated index workers), permutation invariance
of the R1 pair set, R0 subset of R1."""
import numpy as np
import pytest

import apss_b200
from oracle import oracle as orc
from oracle import pyref
from tests.helpers import csr_slice


def synth(N, D, nnz, seed, **kw):
    return apss_b200.synth.generate(N, D, nnz, seed=seed, **kw).numpy()


def run(data, B, t, **kw):
    o = orc.Oracle(data_dim(data), t, **kw)
    pairs, tot = {}, [0, 0, 0]
    N = len(data[0]) - 1
    for lo in range(0, N, B):
        r = o.insert_batch(*csr_slice(data, lo, min(N, lo + B)))
        pairs.update(r.pair_set())
        tot[0] += r.postings_visited; tot[1] += max(r.candidates_unique, 0); tot[2] += max(r.dot_calls_ref, 0)
    return pairs, tot


def data_dim(data):
    return int(data[1].max()) + 1 if len(data[1]) else 1


def test_fast_and_faithful_r1_agree_with_bruteforce():
    N, D, t, B = 900, 256, 0.35, 300
    data = synth(N, D, 12, seed=1)
    fast, tf = run(data, B, t, algo=orc.ALGO_FAST, threads=4)
    faith, tq = run(data, B, t, algo=orc.ALGO_FAITHFUL, semantics=orc.R1, threads=4)
    assert fast == faith and len(fast) > 20
    assert tf[0] == tq[0]                                   # postings walked are the same thing in both
    # brute force, batch by batch: queries of batch b against everything in batches <= b
    want, ncand = {}, 0
    for lo in range(0, N, B):
        got, c = orc.bruteforce(csr_slice(data, lo, lo + B), csr_slice(data, 0, lo + B), t,
                                qkeys=np.arange(lo, lo + B), ckeys=np.arange(0, lo + B))
        want.update({(lo + q, cc): s for (q, cc), s in got.items()})
        ncand += c
    assert fast == want and tf[1] == ncand


def test_r0_is_a_subset_of_r1_and_matches_bruteforce_rule():
    N, D, t, B = 600, 128, 0.3, 200
    data = synth(N, D, 7, seed=2)
    r1, _ = run(data, B, t, algo=orc.ALGO_FAITHFUL, semantics=orc.R1)
    r0, _ = run(data, B, t, algo=orc.ALGO_FAITHFUL, semantics=orc.R0)
    assert set(r0) < set(r1) and all(r0[k] == r1[k] for k in r0)
    # single worker: R0 = R1 minus pairs whose shared dims are all first(q)  
    want = {}
    for lo in range(0, N, B):
        q = csr_slice(data, lo, lo + B)
        got, _ = orc.bruteforce(q, csr_slice(data, 0, lo + B), t, qkeys=np.arange(lo, lo + B), ckeys=np.arange(0, lo + B),
                                first_dim=orc.first_dims(*q))
        want.update({(lo + a, b): s for (a, b), s in got.items()})
    assert r0 == want


def test_permuting_the_arrival_order_keeps_the_unordered_r1_pair_set():
    N, D, t, B = 500, 128, 0.3, 125
    ip, ix, v = synth(N, D, 8, seed=3)
    perm = np.random.RandomState(0).permutation(N)
    rows = [(ix[ip[i]:ip[i + 1]], v[ip[i]:ip[i + 1]]) for i in perm]
    ip2 = np.concatenate([[0], np.cumsum([len(r[0]) for r in rows])]).astype(np.int64)
    data2 = (ip2, np.concatenate([r[0] for r in rows]), np.concatenate([r[1] for r in rows]))
    a, _ = run((ip, ix, v), B, t, algo=orc.ALGO_FAST)
    b, _ = run(data2, B, t, algo=orc.ALGO_FAST)
    unordered = lambda pairs, m: {frozenset((m(q), m(c))) for q, c in pairs}
    assert unordered(a, lambda i: i) == unordered(b, lambda i: int(perm[i]))


@pytest.mark.parametrize("as_built", [True, False])
@pytest.mark.parametrize("shards,children", [(1, 1), (3, 2), (5, 1)])
def test_c_oracle_matches_python_restatement_with_dimension_routing(as_built, shards, children):
    """Several emulated index workers (dim % maxShardNum, dim % maxIndexEntryActorNum): the as-built result
    depends on the routing (each worker skips ITS first dim); both restatements must agree on it."""
    N, D, t, B = 160, 64, 0.3, 40
    ip, ix, v = synth(N, D, 6, seed=4 + shards)
    o = orc.Oracle(D, t, semantics=orc.R0 if as_built else orc.R1, algo=orc.ALGO_FAITHFUL,
                   max_shard_num=shards, max_index_entry_actor_num=children)
    p = pyref.Pipeline(D, t, as_built=as_built, max_shard_num=shards, max_index_entry_actor_num=children)
    for lo in range(0, N, B):
        r = o.insert_batch(*csr_slice((ip, ix, v), lo, lo + B))
        vecs = [(str(i), pyref.SparseVector(D, [int(d) for d in ix[ip[i]:ip[i + 1]]], [float(x) for x in v[ip[i]:ip[i + 1]]])) for i in range(lo, lo + B)]
        out = p.insert_batch(vecs)
        want = {(int(q), int(c)): s for q, m in out.items() for c, s in m.items()}
        assert r.pair_set() == want
    assert o.totals()["dot_calls_ref"] == p.dot_calls and o.totals()["postings_visited"] == p.postings_walked


def test_more_workers_can_only_lose_pairs_that_r1_has():
    N, D, t, B = 300, 64, 0.3, 100
    data = synth(N, D, 6, seed=9)
    r1, _ = run(data, B, t, algo=orc.ALGO_FAITHFUL, semantics=orc.R1, max_shard_num=4, max_index_entry_actor_num=3)
    r1_single, _ = run(data, B, t, algo=orc.ALGO_FAITHFUL, semantics=orc.R1)
    r0_multi, _ = run(data, B, t, algo=orc.ALGO_FAITHFUL, semantics=orc.R0, max_shard_num=4, max_index_entry_actor_num=3)
    assert r1 == r1_single                      # R1 does not depend on the sharding
    assert set(r0_multi) <= set(r1)


@pytest.mark.parametrize("t,alpha", [(0.6, 0.0), (0.4, 0.95), (0.9, 0.3)])
def test_index_reduction_keeps_the_pair_set(t, alpha):
    """oracle_set_pruning (the GPU library's exact index reduction) must not change pairs or sims"""
    import apss_b200
    N, D = 5000, 1 << 12
    ip, ix, v = apss_b200.synth.generate(N, D, 30, seed=11).numpy()
    full = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8)
    red = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8, pruning=True, prune_alpha=alpha)
    pf = pr = 0
    for lo in range(0, N, 1024):
        hi = min(N, lo + 1024)
        csr = (ip[lo:hi + 1] - ip[lo], ix[ip[lo]:ip[hi]], v[ip[lo]:ip[hi]])
        a = full.insert_batch(*csr); b = red.insert_batch(*csr)
        assert a.pair_set() == b.pair_set()
        pf += a.postings_visited; pr += b.postings_visited
    assert full.totals()["pairs"] > 0 and 0 < pr < pf and red.n_unindexed > 0
    with pytest.raises(ValueError):
        red.insert_batch(np.array([0, 1]), np.array([0], np.int32), np.array([2.0]))      # norm promise broken
    with pytest.raises(ValueError):
        orc.Oracle(D, t, algo=orc.ALGO_FAITHFUL, pruning=True)


@pytest.mark.parametrize("t,alpha", [(0.5, 0.8), (0.7, 0.5)])
def test_index_reduction_c_oracle_against_python_restatement(t, alpha):
    """two independent restatements of the reduction rule (C and pure Python) agree on pairs, similarities,
    postings visited, candidates touched and the number of un-indexed components"""
    import apss_b200
    N, D = 400, 256
    ip, ix, v = apss_b200.synth.generate(N, D, 12, seed=21).numpy()
    c_or = orc.Oracle(D, t, algo=orc.ALGO_FAST, pruning=True, prune_alpha=alpha)
    py = pyref.ReducedIndexPipeline(t, alpha)
    for lo in range(0, N, 100):
        hi = lo + 100
        rc = c_or.insert_batch(ip[lo:hi + 1] - ip[lo], ix[ip[lo]:ip[hi]], v[ip[lo]:ip[hi]])
        pairs, postings, cands = py.insert_batch([(ix[ip[i]:ip[i + 1]], v[ip[i]:ip[i + 1]]) for i in range(lo, hi)])
        assert rc.pair_set() == pairs
        assert (rc.postings_visited, rc.candidates_unique) == (postings, cands)
    assert c_or.n_unindexed == py.n_unindexed > 0


GOLDEN_G2_SHA256 = "a600feff811843cfa216ed6c3ff691564ff43903dbbff0c26d312e71ca9cd885"      # synth.generate(3000, 1 << 12, 30, seed=7), CPU == CUDA


def test_generator_is_counter_based_and_reproducible():
    """G(N, D, z, s, seed) is keyed by (seed, vector, draw): independent of the chunking, any row range can
    be regenerated alone, and a pinned digest guards the stream against silent changes (generator version g2)."""
    import hashlib
    import torch
    from apss_b200 import synth
    a = synth.generate(3000, 1 << 12, 30, seed=7)
    b = synth.generate(3000, 1 << 12, 30, seed=7, chunk=257)
    assert torch.equal(a.indptr, b.indptr) and torch.equal(a.indices, b.indices) and torch.equal(a.values, b.values)
    g = synth.Generator(3000, 1 << 12, 30, seed=7)
    ptr, dims, tf, dup = g.structure(1000, 1500)
    lo, hi = int(a.indptr[1000]), int(a.indptr[1500])
    assert torch.equal(dims, a.indices[lo:hi]) and int(tf.min()) >= 1
    ip, ix, v = a.numpy()
    nn = np.diff(ip)
    assert 25 < nn.mean() < 35 and nn.min() >= 4                      # max(4, Poisson(30))
    sq = np.add.reduceat(v * v, ip[:-1])
    assert np.allclose(sq, 1.0, atol=1e-12)                            # L2-normalised (LoadGenerator.scala:35-37)
    assert 0.05 < float(dup.float().mean()) < 0.16                     # ~10 % planted near-duplicates
    h = hashlib.sha256(ip.tobytes() + ix.tobytes() + v.tobytes()).hexdigest()
    assert synth.GEN_VERSION == "g2-splitmix64" and h == GOLDEN_G2_SHA256, h
