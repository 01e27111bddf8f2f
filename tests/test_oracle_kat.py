"""Known-answer tests of SURVEY.md 8(c) against BOTH oracle restatements (C and pure Python).

The reference has no tests of its own: these hand-derived KATs are what pins the oracle."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import pyref
from tests.helpers import csr_from_dicts

A = {0: .6, 1: .8}
B2 = {1: .8, 2: .6}


def run_c(batches, t, sem, idx_thr=0.0, keys=None, algo=orc.ALGO_FAITHFUL, freeze_after=None, **kw):
    o = orc.Oracle(2 ** 16, t, idx_thr, semantics=sem, algo=algo, **kw)
    res = []
    for i, b in enumerate(batches):
        k = None if keys is None else keys[i]
        res.append(o.insert_batch(*csr_from_dicts(b), keys=k))
        if freeze_after is not None and i == freeze_after:
            o.freeze()
    return o, res


def run_py(batches, t, as_built, idx_thr=0.0, ids=None, freeze_after=None, **kw):
    p = pyref.Pipeline(2 ** 16, t, idx_thr, as_built=as_built, **kw)
    outs, n = [], 0
    for i, b in enumerate(batches):
        vecs = []
        for j, v in enumerate(b):
            vid = str(n + j) if ids is None else ids[i][j]
            d = sorted(v)
            vecs.append((vid, pyref.SparseVector(2 ** 16, d, [v[x] for x in d])))
        n += len(b)
        outs.append(p.insert_batch(vecs))
        if freeze_after is not None and i == freeze_after:
            p.freeze()
    return p, outs


def py_pairs(out):
    return {(int(q), int(c)): s for q, m in out.items() for c, s in m.items()}


@pytest.mark.parametrize("sem", [orc.R0, orc.R1])
def test_kat_a(sem):
    o, res = run_c([[A], [dict(A)]], 0.5, sem)
    sim = .6 * .6 + .8 * .8
    assert res[1].pair_set() == {(1, 0): sim}
    assert res[1].postings_visited == 4
    if sem == orc.R0:
        assert res[1].dot_calls_ref == 1
    _, outs = run_py([[A], [dict(A)]], 0.5, sem == orc.R0)
    assert py_pairs(outs[1]) == {(1, 0): sim}
    of, rf = run_c([[A], [dict(A)]], 0.5, orc.R1, algo=orc.ALGO_FAST)
    assert rf[1].pair_set() == {(1, 0): sim} and rf[1].candidates_unique == 1 and rf[1].postings_visited == 4


def test_kat_b_first_dim_skip():
    _, r0 = run_c([[A], [B2]], 0.5, orc.R0)
    _, r1 = run_c([[A], [B2]], 0.5, orc.R1)
    assert r0[1].pair_set() == {}
    assert r1[1].pair_set() == {(1, 0): .8 * .8}
    # reversed arrival order: the shared dim 1 is no longer the query's first dim
    _, r0r = run_c([[B2], [A]], 0.5, orc.R0)
    assert r0r[1].pair_set() == {(1, 0): .8 * .8}
    _, p0 = run_py([[A], [B2]], 0.5, True)
    _, p1 = run_py([[A], [B2]], 0.5, False)
    assert py_pairs(p0[1]) == {} and py_pairs(p1[1]) == {(1, 0): .8 * .8}


def test_kat_c_same_batch_both_orders():
    _, r1 = run_c([[A, B2]], 0.5, orc.R1)
    _, r0 = run_c([[A, B2]], 0.5, orc.R0)
    assert r1[0].pair_set() == {(0, 1): .8 * .8, (1, 0): .8 * .8}
    assert r0[0].pair_set() == {(0, 1): .8 * .8}
    _, p0 = run_py([[A, B2]], 0.5, True)
    assert py_pairs(p0[0]) == {(0, 1): .8 * .8}
    _, rf = run_c([[A, B2]], 0.5, orc.R1, algo=orc.ALGO_FAST)
    assert rf[0].pair_set() == r1[0].pair_set() and rf[0].candidates_unique == 2


def test_kat_d_rescoring_of_failing_candidates():
    a = {0: .5, 1: .5, 2: .5, 3: .5}
    b = {0: .7, 1: .1, 2: .7, 3: .1}
    _, r0 = run_c([[a], [b]], 0.9, orc.R0)
    assert r0[1].pair_set() == {} and r0[1].dot_calls_ref == 3      # dims 1,2,3; dim 0 skipped (IWA:89)
    _, r1 = run_c([[a], [b]], 0.9, orc.R1)
    assert r1[1].dot_calls_ref == 4
    _, rf = run_c([[a], [b]], 0.9, orc.R1, algo=orc.ALGO_FAST)
    assert rf[1].candidates_unique == 1 and rf[1].pair_set() == {}
    p, _ = run_py([[a], [b]], 0.9, True)
    assert p.dot_calls == 3


def test_kat_e_value_prune_is_strict():
    v1 = {0: .1, 1: .9}
    v2 = {1: .9, 5: .2}          # .2 == indexThreshold -> dropped (strict >, WWA:192)
    _, r = run_c([[v1], [v2]], 0.5, orc.R1, idx_thr=0.2)
    assert r[1].pair_set() == {(1, 0): .9 * .9}
    assert r[1].postings_visited == 2            # only dim 1 survives in both
    _, p = run_py([[v1], [v2]], 0.5, False, idx_thr=0.2)
    assert py_pairs(p[1]) == {(1, 0): .9 * .9}
    # a vector pruned to empty is stored but never indexed / queried (WWA:195)
    _, r2 = run_c([[{0: .2, 1: .9}], [{0: .15, 3: .9}]], 0.1, orc.R1, idx_thr=0.9)
    assert list(r2[1].status) == [orc.ST_EMPTY] and r2[1].pair_set() == {}


def test_kat_f_single_dim_query_never_matches_as_built():
    _, r0 = run_c([[{7: 1.0}], [{7: 1.0}]], 0.5, orc.R0)
    _, r1 = run_c([[{7: 1.0}], [{7: 1.0}]], 0.5, orc.R1)
    assert r0[1].pair_set() == {} and r1[1].pair_set() == {(1, 0): 1.0}


@pytest.mark.parametrize("algo", [orc.ALGO_FAITHFUL, orc.ALGO_FAST])
def test_kat_g_frozen_index(algo):
    o, r = run_c([[A], [dict(A), dict(A)]], 0.5, orc.R1, algo=algo, freeze_after=0)
    # the frozen batch is queried against the old index only; its members do not see each other
    assert r[1].pair_set() == {(1, 0): .6 * .6 + .8 * .8, (2, 0): .6 * .6 + .8 * .8}
    assert o.n_vectors == 1
    _, p = run_py([[A], [dict(A), dict(A)]], 0.5, False, freeze_after=0)
    assert set(py_pairs(p[1])) == {(1, 0), (2, 0)}


def test_kat_i_admission():
    # t = 0.9: sum(v) = 0.5 < t  -> rejected by EPA:81-93: never indexed, queried or a candidate
    _, r = run_c([[{0: .5}], [{0: 1.0}], [{0: .95}]], 0.9, orc.R1)
    assert list(r[0].status) == [orc.ST_REJECTED]
    assert r[1].pair_set() == {} and r[1].postings_visited == 1
    assert r[2].pair_set() == {(2, 1): .95}
    _, p = run_py([[{0: .5}], [{0: 1.0}], [{0: .95}]], 0.9, False)
    assert py_pairs(p[2]) == {(2, 1): .95}


@pytest.mark.parametrize("algo", [orc.ALGO_FAITHFUL, orc.ALGO_FAST])
def test_kat_h_same_external_id_never_paired(algo):
    keys = [np.array([42]), np.array([42]), np.array([43])]
    _, r = run_c([[A], [dict(A)], [dict(A)]], 0.5, orc.R1, keys=keys, algo=algo)
    assert r[1].key_pair_set() == {}
    assert set(r[2].key_pair_set()) == {(43, 42)}
    _, p = run_py([[A], [dict(A)]], 0.5, False, ids=[["x"], ["x"]])
    assert p[1] == {"x": {}}


def test_scala_set_iteration_order_examples():
    # derived in SURVEY 8(a) from the (recalled, unverified) Scala 2.10.4 HashSet rule
    assert list(orc.set_iteration_order([3, 17, 100, 1000, 65537, 200000])) == [200000, 65537, 17, 1000, 3, 100]
    assert list(orc.set_iteration_order([5, 10, 15, 20, 25, 30, 35])) == [5, 10, 25, 20, 35, 30, 15]
    assert list(orc.set_iteration_order([9, 4, 1][::-1])) == [1, 4, 9]          # Set3: insertion order
    assert pyref.scala_set_order([3, 17, 100, 1000, 65537, 200000]) == [200000, 65537, 17, 1000, 3, 100]
    assert pyref.scala_set_order([5, 10, 15, 20, 25, 30, 35]) == [5, 10, 25, 20, 35, 30, 15]


def test_validation_is_all_or_nothing():
    o = orc.Oracle(16, 0.5)
    with pytest.raises(ValueError):
        o.insert_batch(np.array([0, 2]), np.array([3, 3]), np.array([.5, .5]))     # duplicate index
    with pytest.raises(ValueError):
        o.insert_batch(np.array([0, 1]), np.array([16]), np.array([1.0]))           # index >= size
    assert o.n_vectors == 0


def test_similarity_output_format():
    s = pyref.similarity_output_to_string({"7": {"3": 0.75}})
    assert s == "---------------------------------7:3,0.75;\n"
