"""Host-side mirror of the reference's message protocol and vector types, names kept verbatim.

Reference: core/src/main/scala/cpslab/message/Message.scala:8-43,
           core/src/main/scala/cpslab/vector/SparseVector.scala:96-108,198-223,
           core/src/main/scala/cpslab/vector/SparseVectorWrapper.scala:9.
These are the shapes a JVM host would marshal across the JNI shim; here they are
plain Python so that the parity tests can drive the GPU worker exactly like the actor is driven.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, FrozenSet, Iterable, List, Sequence, Set, Tuple

import numpy as np


class SparkSparseVector:
    """org.apache.spark.mllib.linalg.SparseVector as the reference uses it: a (size, indices, values)
    holder (CU:94-95).  Construction checks follow Vectors.sparse (SparseVector.scala:96-108)."""
    __slots__ = ("size", "indices", "values")

    def __init__(self, size: int, indices: Sequence[int], values: Sequence[float]):
        idx = np.asarray(indices, dtype=np.int32)
        val = np.asarray(values, dtype=np.float64)
        if size <= 0:
            raise ValueError("requirement failed: size > 0")
        if idx.shape != val.shape:
            raise ValueError("requirement failed: indices and values differ in length")
        if idx.size and (np.any(np.diff(idx) <= 0)):
            raise ValueError("requirement failed: Found duplicate indices")
        if idx.size and not (0 <= int(idx[0]) and int(idx[-1]) < size):
            raise ValueError("requirement failed: index out of [0, size)")
        self.size, self.indices, self.values = int(size), idx, val

    @staticmethod
    def sparse(size: int, elements: Iterable[Tuple[int, float]]) -> "SparkSparseVector":
        """Vectors.sparse(size, Seq[(Int, Double)]): sorts by index (SparseVector.scala:96-108)."""
        el = sorted(elements, key=lambda e: e[0])
        return SparkSparseVector(size, [e[0] for e in el], [e[1] for e in el])

    def __repr__(self):     # Vector.toString, "(size,[i..],[v..])" (SparseVector.scala:204-205)
        return "(%d,[%s],[%s])" % (self.size, ",".join(str(int(i)) for i in self.indices), ",".join(repr(float(v)) for v in self.values))

    def __eq__(self, other):
        return (isinstance(other, SparkSparseVector) and self.size == other.size and
                np.array_equal(self.indices, other.indices) and np.array_equal(self.values, other.values))

    def __hash__(self):     # never densify (Q11): hash the sparse content
        return hash((self.size, self.indices.tobytes(), self.values.tobytes()))


def parse_vector(text: str) -> SparkSparseVector:
    """Vectors.parseNumeric for the sparse text form "(size,[i,...],[v,...])" (SparseVector.scala:132-156)."""
    s = text.strip()
    if not (s.startswith("(") and s.endswith(")")):
        raise ValueError("Cannot parse %r" % text)
    body = s[1:-1]
    size_s, rest = body.split(",", 1)
    a, b = rest.split("],[")
    idx = [int(float(x)) for x in a.strip("[]").split(",") if x.strip() != ""]
    val = [float(x) for x in b.strip("[]").split(",") if x.strip() != ""]
    return SparkSparseVector(int(float(size_s)), idx, val)


@dataclass(frozen=True)
class SparseVectorWrapper:                                   # SparseVectorWrapper.scala:9
    indices: FrozenSet[int]
    sparseVector: Tuple[str, SparkSparseVector]

    def __str__(self):
        return self.sparseVector[0]


@dataclass(frozen=True)
class LoadData:                                              # Message.scala:10
    tableName: str
    startRow: bytes
    endRow: bytes


@dataclass
class VectorIOMsg:                                           # Message.scala:13
    vectors: Set[Tuple[str, SparkSparseVector]]


@dataclass
class DataPacket:                                            # Message.scala:16
    shardId: int
    vectors: Set[SparseVectorWrapper]


@dataclass
class IndexData:                                             # Message.scala:18
    vectors: Set[SparseVectorWrapper]


@dataclass
class SimilarityOutput:                                      # Message.scala:20-35
    output: Dict[str, Dict[str, float]]
    outputMoment: int

    def __str__(self):
        sb: List[str] = []
        for qid, sims in self.output.items():
            sb.append("---------------------------------")
            sb.append(qid + ":")
            for cid, sim in sims.items():
                sb.append(cid + "," + java_double_to_string(float(sim)) + ";")     # Scala string concatenation = Double.toString
            sb.append("\n")
        return "".join(sb)


@dataclass(frozen=True)
class Test:                                                  # Message.scala:37
    content: str
    __test__ = False


class IOTicket:                                              # Message.scala:39 (case object)
    pass


class ReceiveTimeout:                                        # akka.actor.ReceiveTimeout (IWA:143)
    pass


class IOTrigger:                                             # WriteWorkerActor.IOTrigger (WWA:215)
    pass


class StartTest:                                             # Message.scala:42
    pass


@dataclass(frozen=True)
class StartTime:                                             # Message.scala:43
    vectorId: str
    moment: int


def java_double_to_string(x: float) -> str:
    """java.lang.Double.toString (what `cid + "," + sim` produces at Message.scala:29): see etl.java_double_to_string."""
    from .etl import java_double_to_string as f
    return f(x)


def to_csr(vectors: Sequence[Tuple[str, SparkSparseVector]], dim: int):
    """Marshal (id, vector) pairs to the CSR arrays of apss_insert_batch.  Mixed sizes are rejected up
    front (Q9: CU:99 would throw inside the batch and lose its whole output)."""
    indptr = np.zeros(len(vectors) + 1, np.int64)
    for i, (_, v) in enumerate(vectors):
        if v.size != dim:
            raise ValueError("vector1 size: %d, vector2 size: %d" % (v.size, dim))      # CU:99 message
        indptr[i + 1] = indptr[i] + len(v.indices)
    indices = np.concatenate([v.indices for _, v in vectors]).astype(np.int32) if vectors else np.zeros(0, np.int32)
    values = np.concatenate([v.values for _, v in vectors]).astype(np.float64) if vectors else np.zeros(0, np.float64)
    return indptr, indices, values
