"""The reference's latency driver (benchmark/LoadGenerator.scala:15-173) mirrored in apss_b200.loadgen: concurrent
runners, StartTest / StartTime, response time = outputMoment - StartTime, immediate and buffered worker output.
Virtual clock on CPU with the oracle-backed engine double; the gpu-marked test runs the same experiment on the CUDA engine."""
import numpy as np
import pytest

from tests.helpers import OracleEngine

D = 32


def videos(n=6, seed=3):
    import apss_b200
    M = apss_b200.messages
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        dims = [0, 1, 2, 3 + i % 3, 8 + i]                  # three shared dimensions: every pair of videos is similar
        vals = rng.uniform(0.5, 2.0, size=5)                # NOT normalised: the runner normalises (LoadGenerator.scala:30-41)
        out.append(("v%d" % i, M.SparkSparseVector(D, dims, [float(x) for x in vals])))
    return out


def conf(**over):
    c = {"cpslab.allpair.similarityThreshold": 0.3, "cpslab.allpair.outputIODuration": 0, "cpslab.allpair.vectorDim": D,
         "cpslab.allpair.indexThreshold": 0.0, "cpslab.allpair.benchmark.expDuration": 500,
         "cpslab.allpair.benchmark.writeBatchingDuration": 10, "cpslab.allpair.benchmark.totalMessageCount": 4,
         "cpslab.allpair.benchmark.childrenNum": 2}
    c.update(over)
    return c


def mk_worker(c, gpu=False):
    import apss_b200
    from apss_b200.worker import GpuIndexingWorkerActor
    eng = None if gpu else OracleEngine(D, c["cpslab.allpair.similarityThreshold"], c["cpslab.allpair.indexThreshold"])
    return GpuIndexingWorkerActor(c, replyTo=None, engine=eng)


def experiment(gpu, **over):
    import apss_b200
    from apss_b200 import loadgen
    M = apss_b200.messages
    c = conf(**over)
    vids = videos()
    w = mk_worker(c, gpu)
    seen = []
    orig = w.receive

    def spy(msg):
        if isinstance(msg, M.VectorIOMsg):
            seen.append((loop.now(), [vid for vid, _ in msg.vectors][0]))
        orig(msg)
    w.receive = spy
    loop = loadgen.EventLoop(virtual=True, start_ms=1_000_000)
    lines = []
    rep = loadgen.run_experiment(c, vids, w, loop=loop, log=lines.append)
    return rep, seen, lines, c, vids


def check(rep, seen, lines, c, vids):
    total, kids, period = 4, 2, 10
    # warm-up: runner i counts from i * totalMessageCount (:22), one vector per tick, until msgCount > videos.size (:63-66);
    # the tick that crosses the bound still sends (the cancel comes before the send, :63-73)
    warm = [s for s in seen if s[0] < 1_000_000 + 500]
    ids0 = [int(v) for t, v in warm]
    assert sorted(ids0) == sorted(list(range(1, len(vids) + 2)) + list(range(total + 1, len(vids) + 2)))
    assert [t for t, _ in warm][:4] == [1_000_000, 1_000_000, 1_000_010, 1_000_010]          # both runners tick together
    # test phase: EVERY runner restarts at 1 (:79) -- the same ids are sent by all children (as built)
    test = [s for s in seen if s not in warm]
    assert sorted(int(v) for _, v in test) == sorted(list(range(1, total + 2)) * kids)
    t_first = min(t for t, _ in test)
    assert t_first >= 1_000_000 + 500                                                       # after the parent's ReceiveTimeout
    assert [t - t_first for t, _ in test][:4] == [0, 0, period, period]
    # response times: endTime - startTime over the ids with both (:112-131)
    assert rep["messages"] >= 1 and rep["with_both_times"] == rep["messages"]
    assert rep["min_ms"] >= 0 and rep["max_ms"] >= rep["min_ms"] and rep["min_ms"] <= rep["average_ms"] <= rep["max_ms"]
    assert rep["line"].startswith("LoadGenerator stopped with %d messages, average response time" % rep["messages"])
    assert all(" lasting Time:" in ln and " -> " in ln for ln in lines) and lines


def test_loadgen_protocol_immediate_output_cpu():
    rep, seen, lines, c, vids = experiment(False)
    check(rep, seen, lines, c, vids)
    assert rep["max_ms"] == 0            # outputIODuration = 0: the worker answers inside the tick (virtual clock)


def test_loadgen_protocol_buffered_output_cpu():
    # the worker buffers its output and flushes on its own IOTicket every 25 ms (IWA:113-120, 138-142): the response time
    # is the wait for the next flush, 0 .. 25 ms on the virtual clock, and queries sent in the same window share a moment
    rep, seen, lines, c, vids = experiment(False, **{"cpslab.allpair.outputIODuration": 25})
    check(rep, seen, lines, c, vids)
    assert 0 < rep["max_ms"] <= 25 and 0 <= rep["min_ms"] <= rep["max_ms"]


def test_index_is_frozen_in_the_test_phase_cpu():
    # the worker's own ReceiveTimeout (same expDuration key, IWA:37-39,143-144) fires in the quiet period before StartTest:
    # test-phase vectors are queried against the warm-up index only, never against each other
    import apss_b200
    from apss_b200 import loadgen
    c = conf()
    w = mk_worker(c)
    outs = []
    loop = loadgen.EventLoop(virtual=True)
    rep = loadgen.run_experiment(c, videos(), w, loop=loop)
    assert w.stopUpdateIndex
    assert rep["ready"] == 0             # totalMessageCount * childrenNum - 1 similar vectors per query never happens here


def test_event_loop_fixed_rate_and_cancel():
    from apss_b200.loadgen import EventLoop
    lp = EventLoop(virtual=True, start_ms=100)
    hits = []
    ev = lp.schedule_every(0, 10, lambda: hits.append(lp.now()))
    lp.schedule(35, lambda: EventLoop.cancel(ev))
    lp.schedule(50, lambda: hits.append(-lp.now()))
    lp.run()
    assert hits == [100, 110, 120, 130, -150]


@pytest.mark.gpu
def test_loadgen_protocol_gpu():
    for over in ({}, {"cpslab.allpair.outputIODuration": 25}, {"cpslab.allpair.gpu.pruning": 3}):
        rep, seen, lines, c, vids = experiment(True, **over)
        check(rep, seen, lines, c, vids)
