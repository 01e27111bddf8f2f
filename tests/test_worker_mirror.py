"""The host-side mirror of the reference's message protocol (apss_b200.messages / .worker), driven
like the actor is driven.  On CPU the engine is the oracle-backed test double; the same scenarios run
against the CUDA engine in the gpu-marked test at the bottom."""
import pytest

from tests.helpers import OracleEngine

CONF = {"cpslab.allpair.similarityThreshold": 0.5, "cpslab.allpair.outputIODuration": 0,
        "cpslab.allpair.benchmark.expDuration": 0, "cpslab.allpair.vectorDim": 64, "cpslab.allpair.indexThreshold": 0.0}


def mk(conf=CONF, gpu=False, **over):
    import apss_b200
    from apss_b200.worker import GpuIndexingWorkerActor
    c = dict(conf); c.update(over)
    out = []
    as_built = str(c.get("cpslab.allpair.gpu.semantics", "R1")).upper() == "R0"
    eng = None if gpu else OracleEngine(c["cpslab.allpair.vectorDim"], c["cpslab.allpair.similarityThreshold"],
                                        c["cpslab.allpair.indexThreshold"], as_built=as_built)
    return GpuIndexingWorkerActor(c, replyTo=out.append, engine=eng), out, apss_b200


def scenario(gpu, **over):
    w, out, pkg = mk(gpu=gpu, **over)
    M = pkg.messages
    V = lambda d: M.SparkSparseVector.sparse(64, list(d.items()))
    w.receive(M.VectorIOMsg({("a", V({0: .6, 1: .8}))}))
    w.receive(M.VectorIOMsg([("b", V({1: .8, 2: .6})), ("c", V({0: .6, 1: .8})), ("tiny", V({5: .1}))]))
    assert isinstance(out[0], M.SimilarityOutput) and out[0].output == {"a": {}}
    o = out[1].output
    assert set(o) == {"b", "c"}                      # "tiny" fails the admission filter (sum < t): no entry
    assert o["c"]["a"] == .6 * .6 + .8 * .8 and o["b"]["a"] == .8 * .8
    assert o["b"]["c"] == o["c"]["b"] == .8 * .8     # in-batch pairs, both orders (IWA:125-132)
    assert str(out[0]) == "---------------------------------a:\n"
    # same external id is never paired with itself (IWA:91)
    w.receive(M.VectorIOMsg([("a", V({0: .6, 1: .8}))]))
    assert "a" not in out[2].output["a"] and set(out[2].output["a"]) == {"b", "c"}
    # Test echo (IWA:145-147) and ReceiveTimeout freeze (IWA:143-144)
    w.receive(M.Test("ping"))
    assert out[3] == M.Test("ping")
    w.receive(M.ReceiveTimeout())
    w.receive(M.VectorIOMsg([("z1", V({0: .6, 1: .8})), ("z2", V({0: .6, 1: .8}))]))
    assert "z2" not in out[4].output["z1"] and "a" in out[4].output["z1"]
    # a batch with a wrong-size vector is dropped whole, like the swallowed exception at IWA:135-137
    n_before = len(out)
    w.receive(M.VectorIOMsg([("bad", M.SparkSparseVector(32, [1], [1.0]))]))
    assert len(out) == n_before


def test_worker_messages_cpu():
    scenario(gpu=False)


def index_data_scenario(gpu, **over):
    """IndexData / DataPacket (Message.scala:16-18) through the worker and the router: no second admission filter
    (EPA:97 ran upstream), first(q) from the wrapper's own Set as built, and a refused batch leaves no trace."""
    w, out, pkg = mk(gpu=gpu, **over)
    M = pkg.messages
    from apss_b200.worker import RegionRouter
    V = lambda d: M.SparkSparseVector.sparse(64, list(d.items()))
    wrap = lambda vid, v: M.SparseVectorWrapper(frozenset(int(i) for i in v.indices), (vid, v))
    w.receive(M.IndexData({wrap("p", V({3: .3}))}))                     # sum 0.3 < t: would be rejected as VectorIOMsg
    assert out[0].output == {"p": {}}
    RegionRouter(CONF, w).tell(M.DataPacket(0, [wrap("q", V({3: .9, 4: .1})), wrap("r", V({3: 1.0}))]))   # EPA:113-122
    assert out[1].output == {"q": {"r": .9}, "r": {"q": .9}}           # q.p = .27 and r.p = .3 stay below t
    n = len(out)
    w.receive(M.VectorIOMsg([("dup", M.SparkSparseVector(32, [1], [1.0]))]))     # refused whole (IWA:135-137)
    assert len(out) == n
    w.receive(M.VectorIOMsg([("dup", V({3: 1.0}))]))
    assert "r" in out[n].output["dup"] and "dup" not in out[n].output["dup"]
    assert "1.0E-5" in str(M.SimilarityOutput({"x": {"y": 1e-5}}, 0))   # Double.toString, not Python's repr
    w0, out0, _ = mk(gpu=gpu, **dict(over, **{"cpslab.allpair.gpu.semantics": "R0"}))
    w0.receive(M.IndexData({wrap("a", V({0: .6, 1: .8}))}))
    w0.receive(M.IndexData({wrap("b", V({1: .8, 2: .6}))}))             # only shared dim is b's first: dropped as built
    assert out0[1].output == {"b": {}}


def test_index_data_and_data_packet_cpu():
    index_data_scenario(gpu=False)


def test_index_data_skips_admission():
    w, out, pkg = mk()
    M = pkg.messages
    v = M.SparkSparseVector.sparse(64, [(3, .3)])          # sum 0.3 < t: would be rejected as VectorIOMsg
    w.receive(M.IndexData({M.SparseVectorWrapper(frozenset([3]), ("p", v))}))
    assert out[0].output == {"p": {}}


def test_buffered_output_and_ioticket():
    w, out, pkg = mk(**{"cpslab.allpair.outputIODuration": 50})
    M = pkg.messages
    V = lambda d: M.SparkSparseVector.sparse(64, list(d.items()))
    w.receive(M.VectorIOMsg([("a", V({0: 1.0}))]))
    w.receive(M.VectorIOMsg([("b", V({0: 1.0}))]))
    assert out == []                                         # buffered (IWA:131-132)
    w.receive(M.IOTicket())
    assert len(out) == 1 and out[0].output == {"b": {"a": 1.0}}   # only non-empty results (IWA:115-119)
    w.receive(M.IOTicket())
    assert len(out) == 1                                     # buffer cleared (IWA:141)


def test_as_built_semantics_r0():
    w, out, pkg = mk(**{"cpslab.allpair.gpu.semantics": "R0"})
    M = pkg.messages
    V = lambda d: M.SparkSparseVector.sparse(64, list(d.items()))
    w.receive(M.VectorIOMsg([("a", V({0: .6, 1: .8}))]))
    w.receive(M.VectorIOMsg([("b", V({1: .8, 2: .6}))]))     # only shared dim is b's first: dropped as built
    assert out[1].output == {"b": {}}


def test_missing_config_key_fails_like_the_reference():
    import apss_b200
    from apss_b200.worker import GpuIndexingWorkerActor
    c = dict(CONF); del c["cpslab.allpair.benchmark.expDuration"]          # Q10: the key must exist
    with pytest.raises(KeyError):
        GpuIndexingWorkerActor(c, engine=OracleEngine(64, .5))


def test_client_connection_and_region_router():
    import apss_b200
    from apss_b200.worker import ClientConnection, LocalActorSystem, RegionRouter
    w, out, pkg = mk()
    M = pkg.messages
    sysm = LocalActorSystem()
    sysm.register("10.0.0.1:2551", RegionRouter(CONF, w))
    cc = ClientConnection(["10.0.0.1:2551"], sysm)
    V = lambda d: M.SparkSparseVector.sparse(64, list(d.items()))
    cc.insertNewVector({("a", V({0: 1.0}))})
    cc.insertNewVector({("b", V({0: 1.0}))})
    assert out[1].output == {"b": {"a": 1.0}}
    # timer-driven batching (WWA:164-183): vectors wait for the IOTrigger tick and form ONE batch
    w2, out2, _ = mk()
    r2 = RegionRouter(dict(CONF, **{"cpslab.allpair.ioTriggerPeriod": 10}), w2)
    r2.tell(M.VectorIOMsg([("a", V({0: 1.0}))])); r2.tell(M.VectorIOMsg([("b", V({0: 1.0}))]))
    assert out2 == []
    r2.tell(M.IOTrigger())
    assert out2[0].output == {"a": {"b": 1.0}, "b": {"a": 1.0}}


def test_vector_text_format_roundtrip():
    import apss_b200
    M = apss_b200.messages
    v = M.parse_vector("(1048576,[3,17,100],[0.5,0.25,1.0])")
    assert v.size == 1048576 and list(v.indices) == [3, 17, 100] and repr(v) == "(1048576,[3,17,100],[0.5,0.25,1.0])"
    with pytest.raises(ValueError):
        M.SparkSparseVector(8, [3, 3], [1.0, 1.0])


@pytest.mark.gpu
@pytest.mark.parametrize("pruning", [0, 3])
def test_index_data_and_data_packet_gpu(pruning):
    index_data_scenario(gpu=True, **{"cpslab.allpair.gpu.pruning": pruning})


@pytest.mark.gpu
@pytest.mark.parametrize("pruning", [0, 2, 3])
def test_worker_messages_gpu(pruning):
    # same messages, same SimilarityOutput with exact index reduction switched on through the config
    scenario(gpu=True, **{"cpslab.allpair.gpu.pruning": pruning})
