"""CPU restatement of the reference's TF-IDF preprocessing (config C1's input), SURVEY.md 8(f)-1.

Reference: etl/src/main/scala/cpslab/etl/PreprocessWithTFIDF.scala:21-52,67 and Utils.scala:10-23, with the
third-party semantics it relies on (spark-mllib 1.2.0 HashingTF / IDF, java.lang.String.hashCode and
String.split) restated from their published behaviour -- the Spark/Java sources are not vendored in the
reference, so this is "recalled, unverified" like the rest of the oracle material.
Harness code: it produces input vectors; it is not on the scoring path.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

NUM_FEATURES = 1 << 20          # HashingTF() default (PreprocessWithTFIDF.scala:47) = vectorDim of conf/app.conf:6


def java_string_hashcode(s: str) -> int:
    """java.lang.String.hashCode: h = 31*h + c over UTF-16 code units, int32 wrap-around."""
    h = 0
    data = s.encode("utf-16-be", "surrogatepass")
    for i in range(0, len(data), 2):
        h = (31 * h + ((data[i] << 8) | data[i + 1])) & 0xFFFFFFFF
    return h - (1 << 32) if h & 0x80000000 else h


def non_negative_mod(x: int, mod: int) -> int:
    """org.apache.spark.util.Utils.nonNegativeMod"""
    r = x % mod if x >= 0 else -((-x) % mod)      # Java % keeps the sign of the dividend
    return r + mod if r < 0 else r


def java_split_space(s: str) -> List[str]:
    """String.split(" "): interior and leading empty strings kept, trailing empty strings removed."""
    parts = s.split(" ")
    while len(parts) > 1 and parts[-1] == "":
        parts.pop()
    if len(parts) == 1 and parts[0] == "" and s != "":
        return []          # e.g. " ".split(" ") -> [] in Java
    return parts


def read_lines_like_bufferedreader(data: bytes) -> List[str]:
    """BufferedReader.readLine: a line ends at \\n, \\r or \\r\\n; no terminator is returned."""
    text = data.decode("utf-8", "replace")
    lines = text.replace("\r\n", "\n").replace("\r", "\n").split("\n")
    if lines and lines[-1] == "":
        lines.pop()        # no extra empty line after a trailing terminator
    return lines


def file_to_single_line(path: str) -> str:
    """PreprocessWithTFIDF.scala:34-40: every line + " ", and -- because the loop tests `line != null`
    before reading -- the literal "null " once at the end."""
    with open(path, "rb") as f:
        lines = read_lines_like_bufferedreader(f.read())
    return "".join(l + " " for l in lines) + "null "


def list_files(root: str) -> List[str]:
    """Utils.getAllFilePath (Utils.scala:10-23): recursive, skips paths containing .DS_Store.  The
    reference's listing order is file-system dependent; here: sorted."""
    out = []
    for dirpath, dirnames, filenames in os.walk(root):
        dirnames.sort()
        for fn in sorted(filenames):
            p = os.path.join(dirpath, fn)
            if ".DS_Store" not in p:
                out.append(p)
    return out


class HashingTF:
    """spark-mllib 1.2.0 HashingTF: index = nonNegativeMod(term.##, numFeatures), value = raw count."""

    def __init__(self, num_features: int = NUM_FEATURES):
        self.num_features = num_features
        self._cache: Dict[str, int] = {}

    def index_of(self, term: str) -> int:
        i = self._cache.get(term)
        if i is None:
            i = non_negative_mod(java_string_hashcode(term), self.num_features)
            self._cache[term] = i
        return i

    def transform(self, terms: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
        tf: Dict[int, float] = {}
        for t in terms:
            i = self.index_of(t)
            tf[i] = tf.get(i, 0.0) + 1.0
        idx = np.fromiter(sorted(tf), dtype=np.int32, count=len(tf))
        return idx, np.array([tf[int(i)] for i in idx], dtype=np.float64)


def idf_fit(tf_vectors: Iterable[Tuple[np.ndarray, np.ndarray]], num_features: int = NUM_FEATURES):
    """IDF().fit (minDocFreq 0): idf(j) = ln((m + 1) / (df(j) + 1)), df = #docs with a positive value at j."""
    df = np.zeros(num_features, np.int64)
    m = 0
    for idx, val in tf_vectors:
        df[idx[val > 0]] += 1
        m += 1
    # math.log (libm, correctly rounded), NOT numpy's vectorised log: the latter is off by one ulp for some
    # arguments and the reference's output bytes (data/output/.part-*.crc) are reproduced only with the former
    return np.array([math.log((m + 1.0) / (d + 1.0)) for d in df.tolist()], dtype=np.float64), m


def tfidf_corpus(paths: Sequence[str], num_features: int = NUM_FEATURES):
    """computeTFIDFVector (PreprocessWithTFIDF.scala:45-52) over `paths`.  Returns CSR (indptr, indices,
    values) with explicit zeros kept (IDFModel.transform multiplies value-wise), NOT normalised."""
    htf = HashingTF(num_features)
    tfs = [htf.transform(java_split_space(file_to_single_line(p))) for p in paths]
    idf, m = idf_fit(tfs, num_features)
    indptr = np.zeros(len(tfs) + 1, np.int64)
    for i, (idx, _) in enumerate(tfs):
        indptr[i + 1] = indptr[i] + len(idx)
    indices = np.concatenate([t[0] for t in tfs]) if tfs else np.zeros(0, np.int32)
    values = np.concatenate([t[1] * idf[t[0]] for t in tfs]) if tfs else np.zeros(0, np.float64)
    return indptr, indices.astype(np.int32), values.astype(np.float64), idf, m


def java_double_to_string(x: float) -> str:
    """java.lang.Double.toString: decimal notation for 1e-3 <= |x| < 1e7, otherwise "d.dddE[-]n"; at least
    one digit after the point; shortest digits that round-trip (what repr() gives; JDKs before 19 emit a
    longer digit string for a few values, which this does not reproduce)."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    r = repr(float(x))
    sign = "-" if r.startswith("-") else ""
    r = r.lstrip("-")
    if "e" in r:
        mant, exp = r.split("e")
        exp = int(exp)
    else:
        mant, exp = r, 0
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    digits = (ip + fp).lstrip("0")
    # decimal exponent of the first significant digit
    if ip.strip("0"):
        e10 = len(ip.lstrip("0")) - 1 + exp
    else:
        e10 = -(len(fp) - len(fp.lstrip("0")) + 1) + exp
    digits = digits.rstrip("0") or "0"
    if -3 <= e10 < 7:
        if e10 >= 0:
            ipart = digits[:e10 + 1].ljust(e10 + 1, "0")
            fpart = digits[e10 + 1:] or "0"
        else:
            ipart = "0"
            fpart = "0" * (-e10 - 1) + digits
        return sign + ipart + "." + fpart
    return sign + digits[0] + "." + (digits[1:] or "0") + "E" + str(e10)


def vector_to_text(size: int, idx: np.ndarray, val: np.ndarray) -> str:
    """Vector.toString as saveAsTextFile writes it, "(size,[i,..],[v,..])" (SparseVector.scala:204-205)."""
    return "(%d,[%s],[%s])" % (size, ",".join(str(int(i)) for i in idx), ",".join(java_double_to_string(float(v)) for v in val))


def l2_normalise(indptr, values):
    """LoadGenerator.scala:35-37"""
    out = values.copy()
    for i in range(len(indptr) - 1):
        a, b = indptr[i], indptr[i + 1]
        n = math.sqrt(float(np.sum(values[a:b] * values[a:b])))
        if n > 0:
            out[a:b] = values[a:b] / n
    return out


def vector_from_text(text: str) -> Tuple[int, np.ndarray, np.ndarray]:
    """Inverse of vector_to_text: SparseVector.fromString (vector/SparseVector.scala:132-141).
    Splits on ",[" into exactly three parts or raises like the reference ("cannot parse ...")."""
    parts = text.split(",[")
    if len(parts) != 3:
        raise ValueError("cannot parse " + text)
    size = int(parts[0].replace("(", ""))
    idx = np.array([int(x) for x in parts[1].replace("]", "").split(",")], np.int32)
    val = np.array([float(x) for x in parts[2].replace("])", "").split(",")], np.float64)
    return size, idx, val


def ccweb_line_parser(line: str) -> Tuple[str, int, np.ndarray, np.ndarray]:
    """CCWEBVideoLoadGenerator.lineParser (benchmark/CCWEBVideoLoadGenerator.scala:10-21): a line
    "(id,size,[v0,v1,...])" holds a DENSE feature vector; every bracket is stripped, the LAST `size`
    comma fields are the values (takeRight, :16) and the non-zero ones become the sparse vector."""
    for ch in "()[]":
        line = line.replace(ch, "")
    fields = line.split(",")
    while fields and fields[-1] == "":           # Java split(",") drops trailing empty strings
        fields.pop()
    video_id, size = fields[0], int(fields[1])
    dense = np.array([float(x) for x in fields[len(fields) - size:]] if size > 0 else [], np.float64)
    if dense.shape[0] != size:
        # allValues(_) for _ in 0 until size would throw ArrayIndexOutOfBounds in the reference
        raise IndexError("line holds %d values, header says %d" % (dense.shape[0], size))
    nz = np.nonzero(dense != 0)[0].astype(np.int32)
    return video_id, size, nz, dense[nz]


def ccweb_generate_vectors(path: str) -> List[Tuple[str, int, np.ndarray, np.ndarray]]:
    """CCWEBVideoLoadGenerator.generateVectors (:23-29): one vector per line of the file."""
    with open(path, "r", encoding="utf-8", errors="replace") as f:
        return [ccweb_line_parser(ln.rstrip("\r\n")) for ln in f.read().split("\n") if ln != ""]


def load_runner_vector(videos, msg_count: int, vector_dim: int):
    """LoadRunner.generateVector (benchmark/LoadGenerator.scala:30-41): the video msg_count % len,
    values divided by sqrt(left-to-right sum of squares), size replaced by vectorDim, id = msgCount."""
    _, _, idx, val = videos[msg_count % len(videos)]
    sq = 0.0
    for x in val:                                 # foldLeft(0.0)(sum + value * value)
        sq = sq + float(x) * float(x)
    norm = math.sqrt(sq)
    with np.errstate(divide="ignore", invalid="ignore"):
        return str(msg_count), vector_dim, idx, val / norm
