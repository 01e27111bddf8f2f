"""GPU-backed replacement of the reference's index worker, speaking the reference's messages.

Swap point: `context.actorOf(Props(new IndexingWorkerActor(conf)))` at EntryProxyActor.scala:116 (and
:135).  `GpuIndexingWorkerActor` handles the same messages as IndexingWorkerActor.receive
(IndexingWorkerActor.scala:122-148) and replies with the same `SimilarityOutput`; the index and the
scoring loop live on the GPU behind the C ABI (include/apss.h).  String ids never cross the ABI: the
actor keeps the String <-> internal-id tables (SURVEY.md 8(b) "ids").

`ClientConnection`, `LocalActorSystem` and `RegionRouter` give the reference's client call
(ClientConnection.scala:10-34) an in-process transport so the whole path
insertNewVector -> VectorIOMsg -> worker -> SimilarityOutput can be exercised without Akka.
"""
from __future__ import annotations

import random
import sys
import time
import traceback
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import native
from .messages import (DataPacket, IndexData, IOTicket, IOTrigger, ReceiveTimeout, SimilarityOutput, SparkSparseVector, Test,
                       VectorIOMsg, to_csr)

_M32 = 0xFFFFFFFF


def conf_get(conf, key, default=None, required=False):
    """typesafe-config style lookup: dotted key in a flat or nested dict."""
    if key in conf:
        return conf[key]
    cur = conf
    for part in key.split("."):
        if isinstance(cur, dict) and part in cur:
            cur = cur[part]
        else:
            if required:
                raise KeyError("No configuration setting found for key '%s'" % key)
            return default
    return cur


def _scala_improve(h: int) -> int:
    h = (h + (~(h << 9) & _M32)) & _M32
    h ^= h >> 14
    h = (h + (h << 4)) & _M32
    return h ^ (h >> 10)


def scala_set_first(dims: Sequence[int]) -> int:
    """first element, in iteration order, of the Scala 2.10.4 immutable Set built from ascending
    `dims` (WWA:172, IWA:102): Set1..Set4 iterate in insertion order, a HashSet (>= 5 elements) in
    trie order of the improved hash, 5 bits per level from the low end.  Needed only for the
    as-built semantics R0 (the first posting list of every query is never scored, IWA:89/106-107).
    The Scala stdlib is not vendored in the reference: this rule is recalled, not verified."""
    if len(dims) == 0:
        return -1
    if len(dims) <= 4:
        return int(dims[0])
    return int(min(dims, key=lambda x: [(_scala_improve(int(x) & _M32) >> (5 * lvl)) & 31 for lvl in range(7)]))


class GpuIndexingWorkerActor:
    """IndexingWorkerActor (IWA:21-149) with the inverted index and scoring loop on one GPU.

    conf keys (same names as the reference; IWA:23,26,33,44 / WWA:31,35):
      cpslab.allpair.similarityThreshold, cpslab.allpair.outputIODuration,
      cpslab.allpair.benchmark.expDuration (must exist, Q10), cpslab.allpair.vectorDim,
      cpslab.allpair.indexThreshold (default 0), cpslab.allpair.gpu.semantics ("R1" | "R0").
    """

    def __init__(self, conf, replyTo: Optional[Callable] = None, engine=None, device: int = 0):
        self.similarityThreshold = float(conf_get(conf, "cpslab.allpair.similarityThreshold", required=True))
        self.outputWritingDuration = int(conf_get(conf, "cpslab.allpair.outputIODuration", required=True))
        self.expDuration = int(conf_get(conf, "cpslab.allpair.benchmark.expDuration", required=True))
        self.vectorDim = int(conf_get(conf, "cpslab.allpair.vectorDim", required=True))
        self.indexThreshold = float(conf_get(conf, "cpslab.allpair.indexThreshold", 0.0))
        sem = str(conf_get(conf, "cpslab.allpair.gpu.semantics", "R1")).upper()
        self.as_built = sem == "R0"
        self.replyTo = replyTo
        self.now_ms = lambda: int(time.time() * 1000)     # System.currentTimeMillis (a virtual clock in protocol tests)
        self.writeBuffer: Dict[str, Dict[str, float]] = {}
        self.stopUpdateIndex = False
        self._ids: List[str] = []                 # internal id -> caller's String id
        self._first_of: Dict[str, int] = {}       # String id -> internal id of its first occurrence
        self._dups = False
        if engine is None:                        # the product path: CUDA or nothing
            engine = native.Index(self.vectorDim, self.similarityThreshold, self.indexThreshold, device=device,
                                  semantics=native.SEM_R0 if self.as_built else native.SEM_R1,
                                  pruning=int(conf_get(conf, "cpslab.allpair.gpu.pruning", 0)),    # 0 parity counters, 3 exact index reduction
                                  devices=conf_get(conf, "cpslab.allpair.gpu.devices", None))     # GPUs sharing the index (shards below the C ABI)
        self.engine = engine

    # -- IWA:122-148
    def receive(self, msg):
        if isinstance(msg, IndexData):
            # wrappers carry admitted, pruned vectors (EPA:97, WWA:192-194): do not re-admit.  As built, the first
            # posting list skipped (IWA:89 + IWA:106-107) is the first element of the WRAPPER's Set (IWA:102)
            ws = list(msg.vectors)
            firsts = [scala_set_first(sorted(w.indices)) for w in ws] if self.as_built else None
            self._handle_batch([w.sparseVector for w in ws], skip_admit=True, firsts=firsts)
        elif isinstance(msg, VectorIOMsg):
            self._handle_batch(list(msg.vectors), skip_admit=False)
        elif isinstance(msg, IOTicket) or msg is IOTicket:
            if self.writeBuffer:                                               # IWA:139-142
                self._reply(SimilarityOutput(dict(self.writeBuffer), self.now_ms()))
                self.writeBuffer = {}
        elif isinstance(msg, ReceiveTimeout) or msg is ReceiveTimeout:
            self.stopUpdateIndex = True                                        # IWA:143-144
            self.engine.freeze()
        elif isinstance(msg, Test):
            self._reply(msg)                                                   # IWA:145-147

    def _reply(self, m):
        if self.replyTo is not None:
            self.replyTo(m)

    def _handle_batch(self, vectors: List[Tuple[str, SparkSparseVector]], skip_admit: bool, firsts=None):
        try:                                                                   # IWA:124
            out = self.query_and_index(vectors, skip_admit, firsts)
            if self.replyTo is not None:                                       # IWA:128
                if self.outputWritingDuration <= 0:                            # IWA:129-130
                    self._reply(SimilarityOutput(out, self.now_ms()))
                else:                                                          # IWA:131-132, 113-120
                    for q, sims in out.items():
                        for c, s in sims.items():
                            self.writeBuffer.setdefault(q, {})[c] = s
        except Exception:                                                      # IWA:135-137
            traceback.print_exc(file=sys.stderr)

    def query_and_index(self, vectors, skip_admit=False, firsts=None) -> Dict[str, Dict[str, float]]:
        """buildInvertedIndex + querySimilarItems (IWA:61-111) for one batch; returns outputSimSet."""
        if not vectors:
            return {}
        indptr, indices, values = to_csr(vectors, self.vectorDim)
        n = len(vectors)
        base = len(self._ids)
        # keys of this batch: a String id seen before keeps the key of its first occurrence (IWA:91 compares Strings).
        # Nothing is recorded in _first_of / _ids until the engine has accepted the batch (a refused batch leaves no trace).
        keys = np.empty(n, np.int64)
        fresh: Dict[str, int] = {}
        dups = self._dups
        for i, (vid, _) in enumerate(vectors):
            if vid in self._first_of or vid in fresh:
                dups = True
            keys[i] = self._first_of[vid] if vid in self._first_of else fresh.setdefault(vid, base + i)
        first_dim = None
        if self.as_built:
            if firsts is None:
                firsts = [scala_set_first(indices[indptr[i]:indptr[i + 1]][values[indptr[i]:indptr[i + 1]] > self.indexThreshold])
                          for i in range(n)]
            first_dim = np.asarray(firsts, np.int32)
        res = self.engine.insert_batch(indptr, indices, values, ext_keys=keys if dups else None, first_dim=first_dim,
                                       query_only=self.stopUpdateIndex, skip_admit=skip_admit)
        self._dups = dups
        if not self.stopUpdateIndex:
            self._first_of.update(fresh)
        status = self.engine.fetch_status(n)
        q, c, s = self.engine.fetch_pairs()
        if not self.stopUpdateIndex:
            self._ids.extend(v[0] for v in vectors)
        out: Dict[str, Dict[str, float]] = {}
        for i, (vid, _) in enumerate(vectors):
            if status[i] == native.ST_ACTIVE:          # a key for every q with >= 1 dim (IWA:106)
                out.setdefault(vid, {})
        for qi, ci, si in zip(q, c, s):
            out[vectors[int(qi)][0]][self._ids[int(ci)]] = float(si)
        self.last_result = res
        return out


class RegionRouter:
    """What `regionRouter` -> ShardRegion -> EntryProxyActor -> WriteWorkerActor amount to for one GPU
    worker (SimilaritySearchService.scala:28-32, EPA:95-111, WWA:164-202): vectors are buffered and
    every IOTrigger tick turns the buffer into ONE batch for the worker.  With ioTriggerPeriod <= 0
    each VectorIOMsg is its own batch (the explicit-batch parity configuration P0)."""

    def __init__(self, conf, worker: GpuIndexingWorkerActor):
        self.ioTriggerPeriod = int(conf_get(conf, "cpslab.allpair.ioTriggerPeriod", 0))
        self.worker = worker
        self._buffer: List[Tuple[str, SparkSparseVector]] = []

    def tell(self, msg):
        if isinstance(msg, VectorIOMsg):
            if self.ioTriggerPeriod <= 0:
                self.worker.receive(msg)
            else:
                self._buffer.extend(msg.vectors)
        elif isinstance(msg, DataPacket):
            # EPA:113-122 handleDataPacket: the reference splits the packet by dimension over its index workers
            # (spawnToIndexActor, EPA:37-49); an id-range shard holds all dimensions, so the whole packet is ONE IndexData
            self.worker.receive(IndexData(msg.vectors))
        elif isinstance(msg, IOTrigger) or msg is IOTrigger:
            if self._buffer:
                buf, self._buffer = self._buffer, []
                self.worker.receive(VectorIOMsg(buf))
        else:
            self.worker.receive(msg)


class LocalActorSystem:
    """In-process stand-in for the ActorSystem argument of ClientConnection: resolves
    "akka.tcp://ClusterSystem@host:port/user/regionRouter" to a registered router."""

    def __init__(self):
        self._routes: Dict[str, RegionRouter] = {}

    def register(self, address: str, router: RegionRouter):
        self._routes["akka.tcp://ClusterSystem@%s/user/regionRouter" % address] = router

    def actorSelection(self, path: str) -> RegionRouter:
        return self._routes[path]


class ClientConnection:
    """ClientConnection.scala:10-34: fire-and-forget VectorIOMsg to a randomly chosen regionRouter."""

    def __init__(self, remoteAddresses: List[str], localActorSystem: LocalActorSystem):
        if isinstance(remoteAddresses, str):      # README.md:8-10 documents a single address (Q13)
            remoteAddresses = [remoteAddresses]
        self.remoteRouters = [localActorSystem.actorSelection("akka.tcp://ClusterSystem@%s/user/regionRouter" % a)
                              for a in remoteAddresses]                        # ClientConnection.scala:12-21
        self._n = 0

    def insertNewVector(self, vectors):
        """vectors: Set[(String, SparkSparseVector)]; the README form Set[SparkSparseVector] gets
        generated ids."""
        vs = []
        for v in vectors:
            if isinstance(v, SparkSparseVector):
                vs.append(("auto-%d" % self._n, v))
                self._n += 1
            else:
                vs.append(v)
        router = self.remoteRouters[random.randrange(len(self.remoteRouters))]     # ClientConnection.scala:24-25
        router.tell(VectorIOMsg(vs))                                               # ClientConnection.scala:32
