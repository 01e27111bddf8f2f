"""Host-side mirror of the reference's latency driver, benchmark/LoadGenerator.scala:15-173, speaking the same
messages (messages.StartTest / StartTime / VectorIOMsg / SimilarityOutput / IOTicket / ReceiveTimeout) to the GPU-backed
worker (worker.GpuIndexingWorkerActor through a RegionRouter).

What the reference does, and what is kept here AS BUILT:
  * `childrenNum` LoadRunner actors (LoadGenerator.scala:105-109), each ticking every `writeBatchingDuration` ms
    (:44-46); a tick sends ONE vector, `videos(msgCount % videos.size)`, L2-normalised (:30-41), under the id
    `msgCount.toString`.  Before the test phase runner i counts from `i * totalMessageCount` (:22) and stops ticking once
    `msgCount > videos.size` (:63-66) -- the warm-up that fills the index.
  * The parent's first ReceiveTimeout (`expDuration`, :100-102, :159-168) starts the test phase: `StartTest` to every
    child, which resets `msgCount = 0` in EVERY runner (:79) -- so in the test phase the runners send the SAME ids
    1, 2, 3 ... concurrently (an as-built quirk: `startTime` keeps the LAST runner's moment per id, :157-158).  Each
    test tick first reports `StartTime(id, now)` to the parent (:68), stops after `totalMessageCount` (:69-72).
  * The parent turns every SimilarityOutput into response times (:135-156): for each (query, similar) pair, when the
    query's set of found pairs GROWS, `endTime(query) = outputMoment`; a query is "ready" once it has
    `totalMessageCount * childrenNum - 1` pairs, and the system shuts down when that many queries are ready (or at the
    second ReceiveTimeout).  postStop prints the message count and the average / max / min of endTime - startTime
    over the ids that have both (:112-131; integer division for the average, as in the Scala).
  * The worker's own ReceiveTimeout (IWA:143-144, same `expDuration` key) freezes the index; with
    `outputIODuration > 0` the worker buffers its output and flushes on its own IOTicket (IWA:113-120, 138-142).

Akka's dispatcher and timers are replaced by `EventLoop`, a deterministic single-threaded scheduler: every actor's
mailbox is drained in timestamp order (ties: scheduling order), which is one of the interleavings Akka may produce.
The clock is either virtual (tests: milliseconds advance only through scheduled events) or the wall clock (the latency
tool: a tick that is due waits for its time, and a message handled late is handled late -- queueing shows up in the
response times exactly as it would with real timers)."""
import heapq
import math
import time
from typing import Callable, Dict, List, Optional, Sequence, Set, Tuple

from .messages import (IOTicket, ReceiveTimeout, SimilarityOutput, SparkSparseVector, StartTest, StartTime, VectorIOMsg)


class EventLoop:
    """schedule(delay_ms, fn) / schedule_every(initial_ms, period_ms, fn) -> cancel handle; run() until idle or stop()."""

    def __init__(self, virtual: bool = True, start_ms: int = 0):
        self.virtual = virtual
        self._t0 = start_ms if virtual else int(time.time() * 1000)
        self._now = self._t0
        self._q: List[Tuple[int, int, dict]] = []
        self._seq = 0
        self._stopped = False

    def now(self) -> int:                                    # System.currentTimeMillis
        return self._now if self.virtual else int(time.time() * 1000)

    def _push(self, at: int, ev: dict):
        self._seq += 1
        heapq.heappush(self._q, (at, self._seq, ev))

    def schedule(self, delay_ms: int, fn: Callable[[], None]) -> dict:
        ev = {"fn": fn, "period": 0, "cancelled": False}
        self._push(self.now() + max(0, int(delay_ms)), ev)
        return ev

    def schedule_every(self, initial_ms: int, period_ms: int, fn: Callable[[], None]) -> dict:
        ev = {"fn": fn, "period": max(1, int(period_ms)), "cancelled": False}
        self._push(self.now() + max(0, int(initial_ms)), ev)
        return ev

    @staticmethod
    def cancel(ev: Optional[dict]):
        if ev is not None:
            ev["cancelled"] = True

    def stop(self):                                          # context.system.shutdown()
        self._stopped = True

    def run(self, max_events: int = 10_000_000):
        n = 0
        while self._q and not self._stopped and n < max_events:
            at, _, ev = heapq.heappop(self._q)
            if ev["cancelled"]:
                continue
            if self.virtual:
                self._now = max(self._now, at)
            else:
                wait = at - int(time.time() * 1000)
                if wait > 0:
                    time.sleep(wait / 1000.0)
            ev["fn"]()
            n += 1
            if ev["period"] and not ev["cancelled"]:
                self._push(at + ev["period"], ev)            # fixed-rate, like scheduler.schedule
        return n


def _conf(conf, key, default=None):
    if key in conf:
        return conf[key]
    if default is None:
        raise KeyError(key)                                   # com.typesafe.config.ConfigException.Missing
    return default


class LoadRunner:
    """LoadGenerator.scala:15-92.  `videos`: [(id, SparkSparseVector)] as CCWEBVideoLoadGenerator.generateVectors gives."""

    def __init__(self, rid: int, conf, videos: Sequence[Tuple[str, SparkSparseVector]], remote: Callable, parent: "LoadGenerator",
                 loop: EventLoop):
        self.writeBatching = int(_conf(conf, "cpslab.allpair.benchmark.writeBatchingDuration"))
        self.totalMessageCount = int(_conf(conf, "cpslab.allpair.benchmark.totalMessageCount"))
        self.vectorDim = int(_conf(conf, "cpslab.allpair.vectorDim"))
        self.msgCount = rid * self.totalMessageCount                                   # :22
        self.videos = videos
        self.remote, self.parent, self.loop = remote, parent, loop
        self.testPhaseStarted = False
        self.stopped = False
        self.ioTask = loop.schedule_every(0, self.writeBatching, self._tick)          # preStart, :44-46

    def generateVector(self):                                                          # :30-41
        _, v = self.videos[self.msgCount % len(self.videos)]
        squareSum = math.sqrt(sum(float(x) * float(x) for x in v.values))
        values = [float(x) / squareSum for x in v.values]
        return {(str(self.msgCount), SparkSparseVector(self.vectorDim, list(v.indices), values))}

    def _tick(self):
        if not self.stopped:
            self.receive(IOTicket)

    def receive(self, msg):
        if msg is IOTicket or isinstance(msg, IOTicket):                               # :59-74
            self.msgCount += 1
            if not self.testPhaseStarted:
                if self.msgCount > len(self.videos):
                    EventLoop.cancel(self.ioTask)
            else:
                self.parent.receive(StartTime(str(self.msgCount), self.loop.now()))
                if self.msgCount > self.totalMessageCount:
                    EventLoop.cancel(self.ioTask)
                    self.stopped = True                                                # context.stop(self) -- after this send
            self.remote(VectorIOMsg(self.generateVector()))
        elif msg is StartTest or isinstance(msg, StartTest):                           # :75-83
            self.msgCount = 0
            self.testPhaseStarted = True
            EventLoop.cancel(self.ioTask)                                              # (the reference leaks the old timer if it
            self.ioTask = self.loop.schedule_every(0, self.writeBatching, self._tick)  #  still ran; two timers would double the rate)


class LoadGenerator:
    """LoadGenerator.scala:94-172."""

    def __init__(self, conf, videos, remote: Callable, loop: EventLoop, log: Optional[Callable[[str], None]] = None):
        self.totalMessageCount = int(_conf(conf, "cpslab.allpair.benchmark.totalMessageCount"))
        self.childNum = int(_conf(conf, "cpslab.allpair.benchmark.childrenNum"))
        self.expDuration = int(_conf(conf, "cpslab.allpair.benchmark.expDuration"))
        self.startTime: Dict[str, int] = {}
        self.endTime: Dict[str, int] = {}
        self.findPair: Dict[str, Set[Tuple[str, float]]] = {}
        self.readyVectors: Set[str] = set()
        self.testPhaseStarted = False
        self.loop, self.log = loop, log
        self.children = [LoadRunner(i, conf, videos, remote, self, loop) for i in range(self.childNum)]     # preStart
        self._timeout = None
        self._arm_timeout()

    def _arm_timeout(self):
        # context.setReceiveTimeout: fires after expDuration WITHOUT a message; every received message re-arms it
        EventLoop.cancel(self._timeout)
        if self.expDuration > 0:
            self._timeout = self.loop.schedule(self.expDuration, lambda: self.receive(ReceiveTimeout))

    def receive(self, msg):
        if isinstance(msg, SimilarityOutput):                                          # :135-156
            self._arm_timeout()
            if self.testPhaseStarted:
                for q, sims in msg.output.items():
                    for c, s in sims.items():
                        old = len(self.findPair[q]) if q in self.findPair else -1
                        self.findPair.setdefault(q, set()).add((c, s))
                        new = len(self.findPair[q])
                        if new != old:
                            if self.log:
                                # (the reference throws NoSuchElementException here for an id it never saw a StartTime
                                # for; such ids come from the warm-up and are skipped)
                                if q in self.startTime:
                                    self.log("%s -> %d lasting Time:%d" % (q, new, msg.outputMoment - self.startTime[q]))
                            self.endTime[q] = msg.outputMoment
                        if len(self.findPair[q]) >= self.totalMessageCount * self.childNum - 1:
                            self.readyVectors.add(q)
                    if len(self.readyVectors) >= self.totalMessageCount * self.childNum:
                        self.loop.stop()
        elif isinstance(msg, StartTime):                                               # :157-158
            self._arm_timeout()
            self.startTime[msg.vectorId] = msg.moment
        elif msg is ReceiveTimeout or isinstance(msg, ReceiveTimeout):                 # :159-168
            if not self.testPhaseStarted:
                self.testPhaseStarted = True
                for w in self.children:
                    w.receive(StartTest)
                self._arm_timeout()
            else:
                self.loop.stop()

    def report(self) -> dict:                                                          # postStop, :112-131
        messageNum = len(self.endTime)
        total, mx, mn, n = 0, None, None, 0
        for vid, start in self.startTime.items():
            if vid in self.endTime:
                d = self.endTime[vid] - start
                total += d
                mx = d if mx is None or d > mx else mx
                mn = d if mn is None or d < mn else mn
                n += 1
        avg = int(total / messageNum) if messageNum > 0 else None                      # Long division (truncates toward zero)
        line = None
        if messageNum > 0:
            line = "LoadGenerator stopped with %d messages, average response time %s, max:%s min:%s" % (messageNum, avg, mx, mn)
        return {"messages": messageNum, "average_ms": avg, "max_ms": mx, "min_ms": mn, "with_both_times": n, "line": line}


def run_experiment(conf, videos, worker, loop: Optional[EventLoop] = None, log=None) -> dict:
    """Wire LoadGenerator -> worker -> LoadGenerator the way conf/app.conf does (remoteTarget = the entry actor,
    outputActor = the LoadGenerator) and run to shutdown.  `worker`: a worker.GpuIndexingWorkerActor (or anything with
    receive(msg) and a settable replyTo).  The worker's ReceiveTimeout (IWA:37-39, 143-144: index frozen after expDuration
    without a message) and its output IOTicket (IWA:48-50) are timers of the same loop."""
    loop = loop or EventLoop(virtual=True)
    state = {"gen": None, "wt": None}
    exp = int(_conf(conf, "cpslab.allpair.benchmark.expDuration"))
    out_every = int(_conf(conf, "cpslab.allpair.outputIODuration", 0))

    def arm_worker_timeout():
        EventLoop.cancel(state["wt"])
        if exp > 0:
            state["wt"] = loop.schedule(exp, lambda: worker.receive(ReceiveTimeout()))

    def remote(msg):
        arm_worker_timeout()
        worker.receive(msg)

    worker.replyTo = lambda m: state["gen"].receive(m)
    if loop.virtual and hasattr(worker, "now_ms"):
        worker.now_ms = loop.now                             # outputMoment on the same (virtual) clock as StartTime
    gen = LoadGenerator(conf, videos, remote, loop, log=log)
    state["gen"] = gen
    arm_worker_timeout()
    if out_every > 0:
        loop.schedule_every(0, out_every, lambda: worker.receive(IOTicket()))
    events = loop.run()
    rep = gen.report()
    rep["events"] = events
    rep["ready"] = len(gen.readyVectors)
    return rep
