"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on identical
inputs.  Bit-exact: pair sets, integer counters and fp64 similarities (same summation order)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import csr_from_dicts, csr_slice, assert_pairs_equal

pytestmark = pytest.mark.gpu


def native():
    import apss_b200
    return apss_b200.native


def gpu_pairs(idx, res):
    q, c, s = idx.fetch_pairs()
    return {(int(res.id_base + a), int(b)): float(x) for a, b, x in zip(q, c, s)}


def run_both(batches, dim, t, idx_thr=0.0, sem="R1", keys=None, freeze_after=None, query_only=None, **kw):
    n = native()
    gsem = n.SEM_R0 if sem == "R0" else n.SEM_R1
    o = orc.Oracle(dim, t, idx_thr, semantics=orc.R0 if sem == "R0" else orc.R1,
                   algo=orc.ALGO_FAITHFUL if sem == "R0" else orc.ALGO_FAST, threads=8)
    g = n.Index(dim, t, idx_thr, semantics=gsem, **kw)
    out = []
    for i, csr in enumerate(batches):
        k = None if keys is None else keys[i]
        qo = bool(query_only and query_only[i])
        fd = orc.first_dims(*csr, index_threshold=idx_thr) if sem == "R0" else None
        ro = o.insert_batch(*csr, keys=k, query_only=qo)
        rg = g.insert_batch(*csr, ext_keys=k, first_dim=fd, query_only=qo)
        out.append((ro, rg, gpu_pairs(g, rg), g.fetch_status(len(csr[0]) - 1)))
        if freeze_after is not None and i == freeze_after:
            o.freeze(); g.freeze()
    return o, g, out


A = {0: .6, 1: .8}
B2 = {1: .8, 2: .6}


@pytest.mark.parametrize("algo", [2, 3])
def test_kats_other_kernels(algo):
    kv = dict(kernel_variant=algo << 16)
    _, _, out = run_both([csr_from_dicts([A]), csr_from_dicts([dict(A)])], 64, 0.5, **kv)
    assert out[1][2] == {(1, 0): .6 * .6 + .8 * .8} and out[1][1].postings_visited == 4 and out[1][1].candidates_unique == 1
    _, _, out = run_both([csr_from_dicts([A, B2])], 64, 0.5, **kv)
    assert out[0][2] == {(0, 1): .8 * .8, (1, 0): .8 * .8} and out[0][1].candidates_unique == 2
    a = {0: .5, 1: .5, 2: .5, 3: .5}; b = {0: .7, 1: .1, 2: .7, 3: .1}
    _, _, out = run_both([csr_from_dicts([a]), csr_from_dicts([b])], 64, 0.9, **kv)
    assert out[1][2] == {} and out[1][1].candidates_unique == 1 and out[1][1].postings_visited == 8
    _, g, out = run_both([csr_from_dicts([A]), csr_from_dicts([dict(A), dict(A)])], 64, 0.5, freeze_after=0, **kv)
    assert set(out[1][2]) == {(1, 0), (2, 0)}


def test_kat_a_b_c():
    _, _, out = run_both([csr_from_dicts([A]), csr_from_dicts([dict(A)])], 64, 0.5)
    ro, rg, gp, _ = out[1]
    assert gp == ro.pair_set() == {(1, 0): .6 * .6 + .8 * .8}
    assert rg.postings_visited == 4 and rg.candidates_unique == 1
    _, _, out = run_both([csr_from_dicts([A]), csr_from_dicts([B2])], 64, 0.5, sem="R0")
    assert out[1][2] == {} and out[1][1].n_pairs_r1 == 1
    _, _, out = run_both([csr_from_dicts([B2]), csr_from_dicts([A])], 64, 0.5, sem="R0")
    assert out[1][2] == {(1, 0): .8 * .8}
    _, _, out = run_both([csr_from_dicts([A, B2])], 64, 0.5)
    assert out[0][2] == {(0, 1): .8 * .8, (1, 0): .8 * .8} and out[0][1].candidates_unique == 2
    _, _, out = run_both([csr_from_dicts([A, B2])], 64, 0.5, sem="R0")
    assert out[0][2] == {(0, 1): .8 * .8}


def test_kat_d_e_i_counters_and_filters():
    a = {0: .5, 1: .5, 2: .5, 3: .5}
    b = {0: .7, 1: .1, 2: .7, 3: .1}
    _, _, out = run_both([csr_from_dicts([a]), csr_from_dicts([b])], 64, 0.9)
    assert out[1][2] == {} and out[1][1].candidates_unique == 1 and out[1][1].postings_visited == 8
    _, _, out = run_both([csr_from_dicts([{0: .1, 1: .9}]), csr_from_dicts([{1: .9, 5: .2}])], 64, 0.5, idx_thr=0.2)
    assert out[1][2] == {(1, 0): .9 * .9} and out[1][1].postings_visited == 2
    _, _, out = run_both([csr_from_dicts([{0: .5}]), csr_from_dicts([{0: 1.0}]), csr_from_dicts([{0: .95}])], 64, 0.9)
    n = native()
    assert list(out[0][3]) == [n.ST_REJECTED] and out[0][1].n_rejected == 1
    assert out[2][2] == {(2, 1): .95}
    _, _, out = run_both([csr_from_dicts([{0: .2, 1: .9}]), csr_from_dicts([{0: .15, 3: .9}])], 64, 0.1, idx_thr=0.9)
    assert list(out[1][3]) == [n.ST_EMPTY] and out[1][2] == {}


def test_kat_g_frozen_and_query_only():
    _, g, out = run_both([csr_from_dicts([A]), csr_from_dicts([dict(A), dict(A)])], 64, 0.5, freeze_after=0)
    assert set(out[1][2]) == {(1, 0), (2, 0)}
    assert g.stats()["n_vectors"] == 1 and g.stats()["frozen"] == 1
    _, g, out = run_both([csr_from_dicts([A]), csr_from_dicts([dict(A)]), csr_from_dicts([dict(A)])], 64, 0.5,
                         query_only=[False, True, False])
    assert set(out[1][2]) == {(1, 0)} and set(out[2][2]) == {(1, 0)}   # ids continue from the indexed ones


@pytest.mark.parametrize("algo", [1, 2, 3])
def test_kat_h_same_external_id_never_paired(algo):
    keys = [np.array([42]), np.array([42]), np.array([43])]
    _, _, out = run_both([csr_from_dicts([A]), csr_from_dicts([dict(A)]), csr_from_dicts([dict(A)])], 64, 0.5, keys=keys,
                         kernel_variant=algo << 16)
    assert out[1][2] == {} and out[1][1].candidates_unique == 0
    assert set(out[2][2]) == {(2, 0), (2, 1)} and out[2][1].candidates_unique == 2


def test_validation_all_or_nothing():
    n = native()
    g = n.Index(16, 0.5)
    g.insert_batch(*csr_from_dicts([{1: 1.0}]))
    with pytest.raises(n.ApssError) as e:
        g.insert_batch(np.array([0, 2]), np.array([3, 3]), np.array([.5, .5]))
    assert e.value.code == -4
    with pytest.raises(n.ApssError):
        g.insert_batch(np.array([0, 1]), np.array([16]), np.array([1.0]))
    assert g.stats()["n_vectors"] == 1
    r = g.insert_batch(*csr_from_dicts([{1: 1.0}]))
    assert r.id_base == 1 and gpu_pairs(g, r) == {(1, 0): 1.0}


def test_empty_and_ragged_inputs():
    n = native()
    g = n.Index(128, 0.3)
    r = g.insert_batch(np.array([0]), np.zeros(0, np.int32), np.zeros(0))
    assert r.n_pairs == 0
    # empty vectors in the middle of a batch; 1-component vectors; a vector spanning many dims
    vecs = [{}, {5: 1.0}, {}, {i: 1.0 / np.sqrt(100) for i in range(100)}, {5: .7, 6: .7}]
    csr = csr_from_dicts(vecs)
    o = orc.Oracle(128, 0.3, algo=orc.ALGO_FAST)
    ro = o.insert_batch(*csr)
    rg = g.insert_batch(*csr)
    assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
    assert rg.candidates_unique == ro.candidates_unique and rg.postings_visited == ro.postings_visited


def _synth(N, D, nnz, seed, **kw):
    import apss_b200
    return apss_b200.synth.generate(N, D, nnz, seed=seed, **kw).numpy()


@pytest.mark.parametrize("algo", [1, 2, 3])
@pytest.mark.parametrize("tile,batch", [(128, 100), (256, 333), (3584, 1000), (1024, 4096), (0, 2500), (256, 7)])
def test_synthetic_parity_small(tile, batch, algo):
    """Zipf data, several tile sizes / batch sizes (ragged last tile, tiles spanning batches)."""
    N, D, t = 6000, 1 << 12, 0.6
    data = _synth(N, D, 30, seed=11)
    n = native()
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8)
    g = n.Index(D, t, tile_vectors=tile, kernel_variant=algo << 16)
    for lo in range(0, N, batch):
        hi = min(N, lo + batch)
        csr = csr_slice(data, lo, hi)
        ro = o.insert_batch(*csr)
        rg = g.insert_batch(*csr)
        assert rg.id_base == lo
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())          # bit-exact sims
        assert rg.postings_visited == ro.postings_visited
        assert rg.candidates_unique == ro.candidates_unique
    assert g.stats()["tot_pairs"] == o.totals()["pairs"] > 0


def variant(algo=0, warps=0, unroll=0, qb=0):
    return (qb << 24) | (algo << 16) | (warps << 8) | unroll


@pytest.mark.parametrize("algo,warps,unroll,qb", [(1, 8, 4, 0), (1, 16, 2, 0), (1, 16, 8, 0), (1, 32, 4, 0),
                                                  (2, 16, 0, 16), (2, 8, 0, 4), (2, 32, 0, 32), (2, 16, 0, 1), (2, 16, 0, 7),
                                                  (3, 16, 2, 32), (3, 12, 2, 32), (3, 8, 2, 32), (3, 16, 2, 16), (3, 16, 4, 16),
                                                  (3, 8, 4, 16), (3, 16, 4, 8)])
def test_kernel_variants_agree(algo, warps, unroll, qb):
    """row kernel (fp32 plain RMW) and query-block kernel (fixed-point atomics) give identical results"""
    N, D, t = 3000, 1 << 11, 0.5
    data = _synth(N, D, 25, seed=5)
    n = native()
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8)
    g = n.Index(D, t, tile_vectors=1024 if algo != 3 else 512, kernel_variant=variant(algo, warps, unroll, qb))
    assert g.stats()["warps_per_cta"] == warps
    for lo in range(0, N, 1000):
        csr = csr_slice(data, lo, lo + 1000)
        ro = o.insert_batch(*csr); rg = g.insert_batch(*csr)
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
        assert (rg.postings_visited, rg.candidates_unique) == (ro.postings_visited, ro.candidates_unique)


@pytest.mark.parametrize("algo", [1, 2, 3])
def test_r0_parity_against_faithful_oracle(algo):
    """As-built semantics (first posting list skipped) through the R0 post-filter; vectors with >= 5
    components exercise the Scala HashSet iteration order."""
    N, D, t = 1500, 1 << 9, 0.4
    data = _synth(N, D, 8, seed=3)
    n = native()
    o = orc.Oracle(D, t, semantics=orc.R0, algo=orc.ALGO_FAITHFUL, threads=8)
    g = n.Index(D, t, semantics=n.SEM_R0, tile_vectors=256, kernel_variant=algo << 16)
    dropped = 0
    for lo in range(0, N, 500):
        csr = csr_slice(data, lo, lo + 500)
        ro = o.insert_batch(*csr)
        rg = g.insert_batch(*csr, first_dim=orc.first_dims(*csr))
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
        dropped += rg.n_pairs_r1 - rg.n_pairs
    assert dropped > 0


@pytest.mark.parametrize("algo,scale", [(1, 1.7), (2, 1.7), (3, 1.7), (2, 1e6), (3, 1e-6)])
def test_index_threshold_and_unnormalised_values(algo, scale):
    """Value prune changes what is scored (Q6); inputs need not be unit-norm (Q8): the fixed-point
    scale follows the largest squared norm seen."""
    N, D = 2000, 1 << 10
    ip, ix, v = _synth(N, D, 20, seed=9)
    v = v * scale
    t, ith = 0.3 * scale * scale, 0.12 * scale
    n = native()
    o = orc.Oracle(D, t, ith, algo=orc.ALGO_FAST, threads=8)
    g = n.Index(D, t, ith, tile_vectors=512, kernel_variant=algo << 16)
    for lo in range(0, N, 700):
        csr = csr_slice((ip, ix, v), lo, min(N, lo + 700))
        ro = o.insert_batch(*csr); rg = g.insert_batch(*csr)
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
        assert list(g.fetch_status(len(csr[0]) - 1)) == list(ro.status)
        assert (rg.postings_visited, rg.candidates_unique) == (ro.postings_visited, ro.candidates_unique)


@pytest.mark.parametrize("algo", [1, 3])
def test_pair_buffer_overflow_grows_and_replays(algo):
    """Threshold 0 makes every candidate a pair: far more than the initial pair buffer."""
    N, D = 1500, 1 << 8
    data = _synth(N, D, 10, seed=2, dup_frac=0.0)
    n = native()
    o = orc.Oracle(D, 0.0, algo=orc.ALGO_FAST, threads=8)
    g = n.Index(D, 0.0, tile_vectors=256, reserve_pairs=1024, kernel_variant=algo << 16)
    ro = o.insert_batch(*data); rg = g.insert_batch(*data)
    assert rg.n_pairs == len(ro.sim) == ro.candidates_unique > 1024
    assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_pruned_pair_buffer_overflow_and_zero_threshold(mode):
    """a low threshold with a tiny pair buffer: the reduced-index kernels replay after growing it; t = 0 keeps
    everything indexed (nothing may be left out when any shared dim makes a pair)"""
    N, D = 1500, 1 << 8
    data = _synth(N, D, 10, seed=2, dup_frac=0.0)
    n = native()
    for t in (0.05, 0.0):
        o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8)
        g = n.Index(D, t, tile_vectors=256, reserve_pairs=1024, pruning=mode)
        ro = o.insert_batch(*data); rg = g.insert_batch(*data)
        assert rg.n_pairs == len(ro.sim) > 1024
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
        if t == 0.0:
            assert g.stats()["n_unindexed"] == 0 and rg.candidates_unique == ro.candidates_unique


def test_device_pointer_entry_matches_host_entry():
    import torch
    N, D, t = 3000, 1 << 11, 0.5
    data = _synth(N, D, 25, seed=21)
    n = native()
    g1 = n.Index(D, t, tile_vectors=512); g2 = n.Index(D, t, tile_vectors=512)
    for lo in range(0, N, 1000):
        csr = csr_slice(data, lo, lo + 1000)
        r1 = g1.insert_batch(*csr)
        dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in csr]
        torch.cuda.synchronize()
        r2 = g2.insert_batch(dev[0], dev[1], dev[2])
        assert gpu_pairs(g1, r1) == gpu_pairs(g2, r2)
        assert (r1.postings_visited, r1.candidates_unique) == (r2.postings_visited, r2.candidates_unique)


@pytest.mark.parametrize("algo", [1, 3])
def test_c2_scale_sampled_parity(algo):
    """Config C2 shape (2^16 dims, nnz ~50, t=0.8) at 30K vectors: full parity + invariants."""
    import apss_b200
    N, D, t, B = 30000, 1 << 16, 0.8, 4096
    data = apss_b200.synth.generate(N, D, 50, seed=20260102).numpy()
    n = native()
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=orc.max_threads())
    g = n.Index(D, t, kernel_variant=algo << 16)
    seen = {}
    for lo in range(0, N, B):
        hi = min(N, lo + B)
        csr = csr_slice(data, lo, hi)
        ro = o.insert_batch(*csr); rg = g.insert_batch(*csr)
        gp = gpu_pairs(g, rg)
        assert_pairs_equal(gp, ro.pair_set())
        assert (rg.postings_visited, rg.candidates_unique) == (ro.postings_visited, ro.candidates_unique)
        seen.update(gp)
    # size-independent properties: symmetry inside the data set (every unordered pair reported once
    # across batches or twice inside a batch, with identical similarity)
    for (q, c), s in seen.items():
        assert q != c and s >= t
        if (c, q) in seen:
            assert seen[(c, q)] == s and q // B == c // B
        else:
            assert q // B > c // B


def test_dispatcher_single_rank_device_path():
    """ShardDispatcher with one rank (no process group): device tensors in, global ids out."""
    import torch
    import apss_b200
    from apss_b200.dispatcher import ShardDispatcher
    N, D, t, B = 4000, 1 << 11, 0.5, 1000
    data = _synth(N, D, 25, seed=31)
    n = native()
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8)
    disp = ShardDispatcher(n.Index(D, t, tile_vectors=512), device="cuda:0")
    disp.preload(*[torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in csr_slice(data, 0, B)])
    o.insert_batch(*csr_slice(data, 0, B), index_only=True)
    for lo in range(B, N, B):
        csr = csr_slice(data, lo, lo + B)
        r = disp.insert_batch(*[torch.from_numpy(np.ascontiguousarray(a)) for a in csr])
        ro = o.insert_batch(*csr)
        got = {(int(r.id_base + q), int(c)): float(s) for q, c, s in zip(r.q, r.c, r.sim)}
        assert_pairs_equal(got, ro.pair_set())
        assert (r.postings_visited, r.candidates_unique) == (ro.postings_visited, ro.candidates_unique)


@pytest.mark.parametrize("sem", ["R1", "R0"])
def test_c1_maildir_fixture_parity(sem):
    """Config C1: TF-IDF vectors of the reference's own corpus (a 1536-mail sample of data/maildir_small,
    fixture pinned by the reference's CRC side files, see tests/test_etl.py), L2-normalised as
    LoadGenerator.scala:35-37 does, D = 2^20, cosine >= 0.9, batches of 256."""
    import os
    import apss_b200
    from apss_b200 import etl
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "maildir_small_tfidf_sample.npz"))
    ip, ix = g["indptr"], g["indices"]
    v = etl.l2_normalise(ip, g["values"])
    D, t, B = etl.NUM_FEATURES, 0.9, 256
    n = native()
    N = len(ip) - 1 if sem == "R1" else 512
    o = orc.Oracle(D, t, semantics=orc.R0 if sem == "R0" else orc.R1, algo=orc.ALGO_FAITHFUL if sem == "R0" else orc.ALGO_FAST,
                   threads=orc.max_threads())
    gi = n.Index(D, t, semantics=n.SEM_R0 if sem == "R0" else n.SEM_R1)
    pairs = 0
    for lo in range(0, N, B):
        csr = csr_slice((ip, ix, v), lo, min(N, lo + B))
        fd = orc.first_dims(*csr) if sem == "R0" else None
        ro = o.insert_batch(*csr)
        rg = gi.insert_batch(*csr, first_dim=fd)
        assert_pairs_equal(gpu_pairs(gi, rg), ro.pair_set())
        assert list(gi.fetch_status(len(csr[0]) - 1)) == list(ro.status)
        if sem == "R1":
            assert (rg.postings_visited, rg.candidates_unique) == (ro.postings_visited, ro.candidates_unique)
        pairs += rg.n_pairs
    assert pairs == (595 if sem == "R1" else 92)      # the corpus is full of duplicated mails (sent vs sent_items ...)


def test_c5_shape_heavy_tail_parity():
    """Config C5's shape at small scale: 2^20 dims, heavy-tailed nnz (lognormal, up to thousands of
    components per vector), cosine >= 0.5."""
    import apss_b200
    from apss_b200 import synth
    N, D, t, B = 6000, 1 << 20, 0.5, 1500
    fs = synth.generate_flat(N, D, 200, seed=20260105, heavy_tail=True)
    ip, ix, v = fs.finalize(synth.idf_from_df(fs.df(), N)).numpy()
    assert np.diff(ip).max() > 1000
    n = native()
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=orc.max_threads())
    g = n.Index(D, t)
    pairs = 0
    for lo in range(0, N, B):
        csr = csr_slice((ip, ix, v), lo, lo + B)
        ro = o.insert_batch(*csr); rg = g.insert_batch(*csr)
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
        assert (rg.postings_visited, rg.candidates_unique) == (ro.postings_visited, ro.candidates_unique)
        pairs += rg.n_pairs
    assert pairs > 50


@pytest.mark.parametrize("hi", [1.0, 0.3])
def test_admission_with_real_max_weights(hi):
    """EPA:81-93 with real per-dimension max weights (the reference stubs them to 1.0, EPA:51-57): the
    admission predicate sum_d maxw(d) * v(d) >= t must agree with the oracle vector by vector."""
    N, D, t = 3000, 1 << 10, 0.45
    ip, ix, v = _synth(N, D, 12, seed=17)
    mw = np.random.RandomState(3).uniform(0.02, hi, D)
    n = native()
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST, max_weight=mw, threads=8)
    g = n.Index(D, t, max_weight=mw, tile_vectors=512)
    rejected = 0
    for lo in range(0, N, 1000):
        csr = csr_slice((ip, ix, v), lo, lo + 1000)
        ro = o.insert_batch(*csr); rg = g.insert_batch(*csr)
        assert list(g.fetch_status(1000)) == list(ro.status)
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
        assert (rg.postings_visited, rg.candidates_unique) == (ro.postings_visited, ro.candidates_unique)
        rejected += rg.n_rejected
    assert 0 < rejected < N


def test_two_handles_interleaved_and_threaded():
    """Distinct handles (different tile sizes / kernels) on the same device, used alternately and from two
    threads: a handle is single-caller but not thread-affine, distinct handles may run concurrently."""
    import threading
    N, D, t, B = 3000, 1 << 11, 0.5, 500
    data = _synth(N, D, 25, seed=77)
    n = native()
    want = {}
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8)
    for lo in range(0, N, B):
        want.update(o.insert_batch(*csr_slice(data, lo, lo + B)).pair_set())
    handles = [n.Index(D, t, tile_vectors=256), n.Index(D, t), n.Index(D, t, tile_vectors=512, kernel_variant=1 << 16)]
    got = [dict() for _ in handles]
    for lo in range(0, N, B):                       # interleaved on one thread
        for k, g in enumerate(handles):
            r = g.insert_batch(*csr_slice(data, lo, lo + B))
            got[k].update(gpu_pairs(g, r))
    assert all(x == want for x in got)
    # two threads, one handle each, concurrently
    res = [dict(), dict()]
    hs = [n.Index(D, t, tile_vectors=256), n.Index(D, t)]

    def work(k):
        for lo in range(0, N, B):
            r = hs[k].insert_batch(*csr_slice(data, lo, lo + B))
            res[k].update(gpu_pairs(hs[k], r))
    ths = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [th.start() for th in ths]; [th.join() for th in ths]
    assert res[0] == want and res[1] == want


def test_keys_only_on_some_batches():
    """ext_keys supplied for one batch only: the others fall back to the default key (= internal id)."""
    n = native()
    g = n.Index(64, 0.5)
    r0 = g.insert_batch(*csr_from_dicts([A]))                                   # id 0, default key 0
    r1 = g.insert_batch(*csr_from_dicts([dict(A)]), ext_keys=np.array([0]))      # same key as vector 0: never paired
    assert gpu_pairs(g, r1) == {} and r1.candidates_unique == 0
    r2 = g.insert_batch(*csr_from_dicts([dict(A)]))                              # default key 2: pairs with both
    assert set(gpu_pairs(g, r2)) == {(2, 0), (2, 1)}


# ---------------------------------------------------------------- exact index reduction

@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("tile,batch,t,alpha", [(0, 2500, 0.6, 0.0), (256, 333, 0.6, 0.5), (1024, 4096, 0.4, 0.95),
                                                (128, 7, 0.7, 0.0), (512, 1000, 0.9, 0.3)])
def test_pruned_index_same_pairs_fewer_postings(tile, batch, t, alpha, mode):
    """With `pruning` on the pair set and the fp64 similarities are those of the un-pruned run (bit-exact),
    while the counters are those of the oracle's restatement of the same reduction rule."""
    N, D = 6000, 1 << 12
    data = _synth(N, D, 30, seed=11)
    n = native()
    o_full = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8)
    o_pr = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8, pruning=True, prune_alpha=alpha)
    g = n.Index(D, t, tile_vectors=tile, pruning=mode, prune_alpha=alpha)
    tot_full = tot_pr = 0
    for lo in range(0, N, batch):
        csr = csr_slice(data, lo, min(N, lo + batch))
        rf = o_full.insert_batch(*csr); rp = o_pr.insert_batch(*csr); rg = g.insert_batch(*csr)
        assert_pairs_equal(gpu_pairs(g, rg), rf.pair_set())
        assert_pairs_equal(rp.pair_set(), rf.pair_set())
        assert (rg.postings_visited, rg.candidates_unique) == (rp.postings_visited, rp.candidates_unique)
        tot_full += rf.postings_visited; tot_pr += rg.postings_visited
    st = g.stats()
    assert st["n_unindexed"] == o_pr.n_unindexed > 0
    assert st["n_postings"] == len(data[1]) - st["n_unindexed"]
    assert st["tot_pairs"] == o_full.totals()["pairs"] > 0
    assert tot_pr * 2 < tot_full


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_pruned_index_query_only_keys_and_r0(mode):
    """frozen / query-only batches, duplicate external ids and the R0 post-filter on top of the reduced index"""
    N, D, t = 3000, 1 << 10, 0.5
    data = _synth(N, D, 12, seed=9)
    n = native()
    keys = np.arange(N, dtype=np.int64); keys[1::7] = keys[0:-1:7]          # some vectors share an id with their neighbour
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8)
    g = n.Index(D, t, tile_vectors=256, pruning=mode)
    for lo in range(0, 2000, 500):
        csr = csr_slice(data, lo, lo + 500)
        ro = o.insert_batch(*csr, keys=keys[lo:lo + 500]); rg = g.insert_batch(*csr, ext_keys=keys[lo:lo + 500])
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
    o.freeze(); g.freeze()
    csr = csr_slice(data, 2000, 3000)
    ro = o.insert_batch(*csr, keys=keys[2000:3000]); rg = g.insert_batch(*csr, ext_keys=keys[2000:3000])
    assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
    assert g.stats()["n_vectors"] == 2000 and rg.n_pairs > 0
    # R0 on the reduced index against the faithful as-built oracle
    o0 = orc.Oracle(D, t, semantics=orc.R0, algo=orc.ALGO_FAITHFUL, threads=8)
    g0 = n.Index(D, t, semantics=n.SEM_R0, tile_vectors=256, pruning=mode)
    for lo in range(0, 1500, 500):
        csr = csr_slice(data, lo, lo + 500)
        ro = o0.insert_batch(*csr); rg = g0.insert_batch(*csr, first_dim=orc.first_dims(*csr))
        assert_pairs_equal(gpu_pairs(g0, rg), ro.pair_set())


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_pruned_index_refuses_vectors_over_the_norm_promise(mode):
    n = native()
    g = n.Index(64, 0.5, pruning=mode)                       # max_query_norm defaults to 1
    g.insert_batch(*csr_from_dicts([A]))
    with pytest.raises(n.ApssError) as e:
        g.insert_batch(*csr_from_dicts([{0: .6, 1: .8}, {0: 3.0, 1: 4.0}]))
    assert e.value.code == -4 and "max_query_norm" in str(e.value)
    assert g.stats()["n_vectors"] == 1                       # all-or-nothing
    rg = g.insert_batch(*csr_from_dicts([dict(A)]))
    assert gpu_pairs(g, rg) == {(1, 0): .6 * .6 + .8 * .8}
    g5 = n.Index(64, 6.0, pruning=mode, max_query_norm=5.0)   # un-normalised data with a declared bound
    g5.insert_batch(*csr_from_dicts([{0: 3.0, 1: 4.0}]))
    rg = g5.insert_batch(*csr_from_dicts([{0: 3.0, 1: 4.0}, {0: 1.0}]))
    assert gpu_pairs(g5, rg) == {(1, 0): 25.0}
    for kv in (1 << 16, 2 << 16):                            # only the default scoring kernel applies the bound
        with pytest.raises(n.ApssError):
            n.Index(64, 0.5, pruning=mode, kernel_variant=kv)
    with pytest.raises(n.ApssError):
        n.Index(64, 0.5, pruning=True, prune_alpha=1.0)


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_pruned_c2_shape(mode):
    """C2's shape (100 K x 2^16, t = 0.8) on a 30 K prefix: full pair-set parity, large work reduction"""
    import apss_b200
    N, D, t = 30_000, 1 << 16, 0.8
    data = apss_b200.synth.generate(N, D, 50, seed=20260102).numpy()
    n = native()
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=16)
    g = n.Index(D, t, pruning=mode)
    tf = tp = 0
    for lo in range(0, N, 4096):
        csr = csr_slice(data, lo, min(N, lo + 4096))
        ro = o.insert_batch(*csr); rg = g.insert_batch(*csr)
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
        tf += ro.postings_visited; tp += rg.postings_visited
    assert o.totals()["pairs"] > 0 and tp * 20 < tf


@pytest.mark.parametrize("cap,items_cap,hot_cap", [(0, 0, 0), (64, 0, 0), (700, 50, 0), (8, 1, 0), (0, 0, 70000)])
def test_query_major_ranged_passes_merges_and_piece_replay(cap, items_cap, hot_cap, monkeypatch):
    """pruning = 3: a dimension shared by every vector (too heavy to stay out of the index) gives every query a list as
    long as the index: with the per-pass capacity lowered (APSS_QM_CAP) the query is scored in candidate-id ranges sized
    by the counting walk (halved until they fit); a tiny piece buffer (APSS_QM_ITEMS_CAP) forces the grow-and-replay
    path; small batches force segment merges; duplicate ids and in-batch pairs included"""
    rng = np.random.default_rng(5)
    N, D, t = 4000, 512, 0.5
    rows = []
    for i in range(N):
        dims = rng.choice(np.arange(1, D), size=6, replace=False)
        v = {int(d): float(x) for d, x in zip(dims, rng.uniform(0.05, 0.3, size=6))}
        v[0] = 0.9
        nrm = np.sqrt(sum(x * x for x in v.values()))
        rows.append({d: x / nrm for d, x in v.items()})
    keys = np.arange(N, dtype=np.int64); keys[5::11] = keys[4:-1:11]
    if cap:
        monkeypatch.setenv("APSS_QM_CAP", str(cap))
    if items_cap:
        monkeypatch.setenv("APSS_QM_ITEMS_CAP", str(items_cap))
    if hot_cap:      # two chunks of the hot-candidate buffer for 148 CTAs: the first call overflows, grows and replays
        monkeypatch.setenv("APSS_QM_HOT_CAP", str(hot_cap))
    n = native()
    for use_keys in (False, True):
        o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8, pruning=True)
        g = n.Index(D, t, pruning=3)
        for lo, hi in [(0, 1500), (1500, 1507), (1507, 1600), (1600, 3100), (3100, 3101), (3101, 4000)]:
            csr = csr_from_dicts(rows[lo:hi])
            kw_o = dict(keys=keys[lo:hi]) if use_keys else {}
            kw_g = dict(ext_keys=keys[lo:hi]) if use_keys else {}
            ro = o.insert_batch(*csr, **kw_o); rg = g.insert_batch(*csr, **kw_g)
            assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
            assert (rg.postings_visited, rg.candidates_unique) == (ro.postings_visited, ro.candidates_unique)
        st = g.stats()
        assert st["n_postings"] == sum(len(r) for r in rows) - st["n_unindexed"] and 1 <= st["n_tiles"] <= 6
        assert o.totals()["pairs"] > 1000
        g.close()


@pytest.mark.parametrize("slices", [0, 3, 16])
def test_candidate_major_heavy_vectors_and_many_queries(slices, monkeypatch):
    """stored vectors whose query lists overflow the per-warp table take the heavy pass: a dimension shared by every
    vector with a weight large enough to stay indexed, batches larger than one heavy-pass query chunk is not needed
    (chunking is covered by nq > 0 only), duplicate ids and in-batch pairs included"""
    rng = np.random.default_rng(3)
    N, D, t = 3000, 512, 0.5
    rows = []
    for i in range(N):
        dims = rng.choice(np.arange(1, D), size=6, replace=False)
        v = {int(d): float(x) for d, x in zip(dims, rng.uniform(0.05, 0.3, size=6))}
        v[0] = 0.9                                  # shared by all, too heavy to stay out of the index
        nrm = np.sqrt(sum(x * x for x in v.values()))
        rows.append({d: x / nrm for d, x in v.items()})
    n = native()
    if slices:
        monkeypatch.setenv("APSS_CAND_SLICES", str(slices))      # force the query-slice path (normally sized from the previous batch)
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8, pruning=True)
    g = n.Index(D, t, pruning=2)
    for lo in range(0, N, 1000):
        csr = csr_from_dicts(rows[lo:lo + 1000])
        ro = o.insert_batch(*csr); rg = g.insert_batch(*csr)
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
        assert (rg.postings_visited, rg.candidates_unique) == (ro.postings_visited, ro.candidates_unique)
    assert o.totals()["pairs"] > 1000


def test_generator_gives_the_same_bits_on_cpu_and_cuda():
    """generator g2: the oracle box and the GPU box regenerate the same data from the seed"""
    import torch
    from apss_b200 import synth
    for kw in (dict(N=6000, D=1 << 14, nnz_mean=40, seed=20260103), dict(N=1500, D=1 << 16, nnz_mean=200, seed=5, heavy_tail=True)):
        a = synth.generate(device="cpu", **kw)
        b = synth.generate(device="cuda", **kw)
        assert torch.equal(a.indptr, b.indptr.cpu()) and torch.equal(a.indices, b.indices.cpu())
        assert torch.equal(a.values, b.values.cpu())          # bit for bit, fp64


# ---------------------------------------------------------------- engine assumptions made explicit (VERDICT r1 "weak 8")

def test_very_long_vectors_u16_scale_and_limit():
    """The packed-u16 tile kernel bounds a sum by max_sq * 2^F + one quantum per shared dimension: the scale follows the
    longest vector seen (pairs sharing > 32 K dimensions used to wrap a half-word), and a vector beyond the limit is
    refused up front instead of being mis-scored."""
    rng = np.random.default_rng(1)
    D, t = 1 << 16, 0.5
    n = native()
    rows = []
    base = rng.choice(D, size=40000, replace=False)
    for i in range(6):                                   # six near-identical vectors sharing ~40 K dimensions
        w = rng.uniform(0.5, 1.0, size=base.size) if i == 0 else w0 * rng.uniform(0.98, 1.02, size=base.size)
        if i == 0:
            w0 = w
        v = dict(zip(base.tolist(), (w / np.sqrt((w * w).sum())).tolist()))
        rows.append(v)
    for i in range(300):                                 # and ordinary short ones
        dims = rng.choice(D, size=20, replace=False)
        w = rng.uniform(0.1, 1.0, size=20)
        rows.append(dict(zip(dims.tolist(), (w / np.sqrt((w * w).sum())).tolist())))
    csr = csr_from_dicts(rows)
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8)
    ro = o.insert_batch(*csr)
    for kw in (dict(), dict(pruning=3), dict(kernel_variant=2 << 16)):
        g = n.Index(D, t, **kw)
        rg = g.insert_batch(*csr)
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
        assert len(ro.sim) >= 30
        if not kw.get("pruning"):
            assert (rg.postings_visited, rg.candidates_unique) == (ro.postings_visited, ro.candidates_unique)
        g.close()
    big = np.arange(61000, dtype=np.int32)
    csr_big = (np.array([0, big.size], np.int64), big, np.full(big.size, 1.0 / np.sqrt(big.size)))
    g = n.Index(D, t)
    g.insert_batch(*csr_from_dicts([A]))
    with pytest.raises(n.ApssError) as e:
        g.insert_batch(*csr_big)
    assert e.value.code == -4 and "u16" in str(e.value) and g.stats()["n_vectors"] == 1
    g3 = n.Index(D, t, pruning=3)                        # the hash-table kernel has no such limit
    g3.insert_batch(*csr_big)
    assert g3.stats()["n_vectors"] == 1


@pytest.mark.parametrize("mode", [0, 2, 3])
def test_failure_after_the_index_append_is_rolled_back_or_retires_the_handle(mode, monkeypatch):
    """A batch that fails AFTER the index was touched must not leave the shard advanced behind the caller's back:
    append-only layouts (pruning 2 / 3) are cut back -- ids, document frequencies and posting segments as before, later
    batches give the oracle's pairs -- and the tile index, which rewrites its open tile in place, retires the handle."""
    N, D, t = 3000, 1 << 10, 0.5
    data = _synth(N, D, 12, seed=9)
    n = native()
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8, pruning=bool(mode))
    g = n.Index(D, t, tile_vectors=256, pruning=mode)
    b = [csr_slice(data, lo, lo + 500) for lo in range(0, N, 500)]
    for k in (0, 1):
        ro = o.insert_batch(*b[k]); rg = g.insert_batch(*b[k])
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
    st0 = g.stats()
    monkeypatch.setenv("APSS_TEST_FAIL_AFTER_APPEND", "1")
    with pytest.raises(n.ApssError) as e:
        g.insert_batch(*b[2])
    monkeypatch.delenv("APSS_TEST_FAIL_AFTER_APPEND")
    assert e.value.code == -3
    if mode == 0:
        assert "retired" in str(e.value)
        with pytest.raises(n.ApssError) as e2:
            g.insert_batch(*b[2])
        assert e2.value.code == -5                        # APSS_E_STATE: only apss_destroy is valid now
        return
    assert "rolled back" in str(e.value)
    st1 = g.stats()
    assert (st1["n_vectors"], st1["n_postings"], st1["n_unindexed"]) == (st0["n_vectors"], st0["n_postings"], st0["n_unindexed"])
    with pytest.raises(n.ApssError):
        g.fetch_pairs()                                   # no completed batch to fetch from
    for k in (2, 3, 4, 5):                                # the same batch again, then the rest: counters prove df was restored
        ro = o.insert_batch(*b[k]); rg = g.insert_batch(*b[k])
        assert rg.id_base == ro.id_base
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
        assert (rg.postings_visited, rg.candidates_unique) == (ro.postings_visited, ro.candidates_unique)
    assert g.stats()["n_unindexed"] == o.n_unindexed


def test_int32_bounds_of_ids_and_component_counts():
    """internal ids are int32: a batch that would run the id space over is refused before anything is touched"""
    n = native()
    g = n.Index(64, 0.5)
    g.set_next_id((1 << 31) - 3)
    with pytest.raises(n.ApssError) as e:
        g.insert_batch(*csr_from_dicts([A, dict(A), dict(A), dict(A)]))
    assert e.value.code == -1 and "id space" in str(e.value)
    rg = g.insert_batch(*csr_from_dicts([A, dict(A)]))     # two still fit
    assert rg.id_base == (1 << 31) - 3 and g.stats()["n_vectors"] == 2


# ---------------------------------------------------------------- shard dispatch below the C ABI

def _device_sets():
    import torch
    sets = [("dup", [0, 0, 0])]                       # three shards on GPU 0 (test hook): the dispatch logic on a one-GPU box
    if torch.cuda.device_count() >= 2:
        sets.append(("real", list(range(min(torch.cuda.device_count(), 4)))))
    return sets


@pytest.mark.parametrize("pruning", [0, 3])
def test_multi_device_handle_matches_the_oracle(pruning, monkeypatch):
    """apss_config.n_devices > 1: ONE handle, the index sharded by id range (block-cyclic by batch) over several engines.
    Pairs, similarities and the summed counters are those of a single index worker; ids stay global; keys, query-only
    batches, freeze, bulk load and device-resident batches (NVLink peer fan-out) go through the same entry points."""
    import torch
    N, D, t = 6000, 1 << 12, 0.6
    data = _synth(N, D, 30, seed=11)
    keys = np.arange(N, dtype=np.int64); keys[3::9] = keys[2:-1:9]
    n = native()
    monkeypatch.setenv("APSS_TEST_ALLOW_DUP_DEVICES", "1")
    for name, devs in _device_sets():
        o = orc.Oracle(D, t, algo=orc.ALGO_FAST, threads=8, pruning=bool(pruning))
        g = n.Index(D, t, pruning=pruning, devices=devs)
        assert g.stats()["n_devices"] == len(devs)
        # bulk load of the first 1000 vectors (one shard indexes, the others have nothing to do), then live batches
        o.insert_batch(*csr_slice(data, 0, 1000), index_only=True); g.insert_batch(*csr_slice(data, 0, 1000), index_only=True)
        for lo, hi, use_keys, on_dev in [(1000, 2000, False, False), (2000, 2007, False, True), (2007, 3500, True, False),
                                         (3500, 3501, True, True), (3501, 5000, False, False)]:
            csr = csr_slice(data, lo, hi)
            ro = o.insert_batch(*csr, keys=keys[lo:hi] if use_keys else None)
            args = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in csr] if on_dev else list(csr)
            k = None if not use_keys else (torch.from_numpy(keys[lo:hi].copy()).cuda() if on_dev else keys[lo:hi])
            rg = g.insert_batch(*args, ext_keys=k)
            assert rg.id_base == ro.id_base == lo
            assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
            if not pruning:
                assert (rg.postings_visited, rg.candidates_unique) == (ro.postings_visited, ro.candidates_unique)
            assert list(g.fetch_status(hi - lo)) == list(ro.status)
        o.freeze(); g.freeze()
        csr = csr_slice(data, 5000, 6000)
        ro = o.insert_batch(*csr); rg = g.insert_batch(*csr)
        assert_pairs_equal(gpu_pairs(g, rg), ro.pair_set())
        st = g.stats()
        assert st["n_vectors"] == 5000 and st["tot_pairs"] == o.totals()["pairs"] > 0
        with pytest.raises(n.ApssError):                       # a malformed batch is refused by every shard, nothing changes
            g.insert_batch(np.array([0, 2], np.int64), np.array([5, 5], np.int32), np.array([.5, .5]))
        assert g.stats()["n_vectors"] == 5000
        g.close()
