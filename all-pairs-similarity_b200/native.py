"""ctypes binding of include/apss.h (libapss_b200.so, built in-tree for sm_100a).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable, every
entry point raises.  Nothing in this package imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("APSS_LIB_NAME") or ("libapss_b200_dbg.so" if os.environ.get("APSS_DEBUG_LIB") else
                        ("libapss_b200_prof.so" if os.environ.get("APSS_PROF_LIB") else "libapss_b200.so")))

ABI_VERSION = 3
MAX_DEVICES = 16
SEM_R1, SEM_R0 = 0, 1
BATCH_QUERY_ONLY, BATCH_DEVICE_PTRS, BATCH_SKIP_ADMIT, BATCH_INDEX_ONLY = 1, 2, 4, 8
ST_REJECTED, ST_EMPTY, ST_ACTIVE = 0, 1, 2

_STATUS_NAMES = {0: "APSS_OK", -1: "APSS_E_INVALID", -2: "APSS_E_CUDA", -3: "APSS_E_NOMEM", -4: "APSS_E_INPUT",
                 -5: "APSS_E_STATE", -6: "APSS_E_NO_DEVICE"}

# every symbol include/apss.h declares (checked by tests/test_abi.py)
EXPORTS = ["apss_abi_version", "apss_create", "apss_destroy", "apss_insert_batch", "apss_fetch_pairs",
           "apss_pairs_device", "apss_fetch_status", "apss_freeze", "apss_set_next_id", "apss_get_stats",
           "apss_last_error", "apss_stream", "apss_microbench_accumulators"]


class ApssError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (_STATUS_NAMES.get(code, str(code)), msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("dim", C.c_int32), ("similarity_threshold", C.c_double),
                ("index_threshold", C.c_double), ("max_weight", C.c_void_p), ("device", C.c_int32),
                ("semantics", C.c_int32), ("tile_vectors", C.c_int32), ("kernel_variant", C.c_int32),
                ("reserve_vectors", C.c_int64), ("reserve_nnz", C.c_int64), ("reserve_pairs", C.c_int64),
                ("pruning", C.c_int32), ("reserved0", C.c_int32), ("prune_alpha", C.c_double), ("max_query_norm", C.c_double),
                ("n_devices", C.c_int32), ("device_ids", C.c_int32 * MAX_DEVICES), ("reserved1", C.c_int32)]


class BatchResultC(C.Structure):
    _fields_ = [("id_base", C.c_int64), ("n_vectors", C.c_int32), ("n_rejected", C.c_int32), ("n_empty", C.c_int32),
                ("n_active", C.c_int32), ("n_pairs", C.c_int64), ("n_pairs_r1", C.c_int64), ("n_prefilter", C.c_int64),
                ("postings_visited", C.c_int64), ("candidates_unique", C.c_int64), ("work_items", C.c_int64),
                ("score_ms", C.c_double), ("device_ms", C.c_double), ("dense_postings", C.c_int64), ("dense_fma", C.c_int64)]


class StatsC(C.Structure):
    _fields_ = [("n_vectors", C.c_int64), ("n_postings", C.c_int64), ("n_tiles", C.c_int64), ("bytes_postings", C.c_int64),
                ("bytes_directory", C.c_int64), ("bytes_forward", C.c_int64), ("tot_postings_visited", C.c_int64),
                ("tot_candidates_unique", C.c_int64), ("tot_pairs", C.c_int64), ("tot_prefilter", C.c_int64),
                ("score_launches", C.c_int64), ("kernel_launches", C.c_int64), ("tot_score_ms", C.c_double),
                ("phase_cycles", C.c_int64 * 8),
                ("frozen", C.c_int32), ("tile_vectors", C.c_int32), ("warps_per_cta", C.c_int32), ("sm_count", C.c_int32),
                ("n_unindexed", C.c_int64), ("segment_merges", C.c_int64), ("merged_postings", C.c_int64),
                ("n_devices", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def load_library():
    """dlopen the in-tree library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ApssError(-6, "CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                            "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.apss_abi_version.restype = C.c_int32
    L.apss_create.restype = C.c_int32
    L.apss_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.apss_destroy.restype = None
    L.apss_destroy.argtypes = [C.c_void_p]
    L.apss_insert_batch.restype = C.c_int32
    L.apss_insert_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_uint32, C.POINTER(BatchResultC)]
    L.apss_fetch_pairs.restype = C.c_int32
    L.apss_fetch_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.apss_pairs_device.restype = C.c_int32
    L.apss_pairs_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.apss_fetch_status.restype = C.c_int32
    L.apss_fetch_status.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.apss_freeze.restype = C.c_int32
    L.apss_freeze.argtypes = [C.c_void_p]
    L.apss_set_next_id.restype = C.c_int32
    L.apss_set_next_id.argtypes = [C.c_void_p, C.c_int64]
    L.apss_get_stats.restype = C.c_int32
    L.apss_get_stats.argtypes = [C.c_void_p, C.POINTER(StatsC)]
    L.apss_last_error.restype = C.c_char_p
    L.apss_last_error.argtypes = [C.c_void_p]
    L.apss_stream.restype = C.c_void_p
    L.apss_stream.argtypes = [C.c_void_p]
    L.apss_microbench_accumulators.restype = C.c_int32
    L.apss_microbench_accumulators.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    if L.apss_abi_version() != ABI_VERSION:
        raise ApssError(-1, "ABI version mismatch: library %d, binding %d" % (L.apss_abi_version(), ABI_VERSION))
    _lib = L
    return L


@dataclass
class BatchResult:
    id_base: int
    n_vectors: int
    n_rejected: int
    n_empty: int
    n_active: int
    n_pairs: int
    n_pairs_r1: int
    n_prefilter: int
    postings_visited: int
    candidates_unique: int
    work_items: int
    score_ms: float
    device_ms: float
    dense_postings: int = 0
    dense_fma: int = 0


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):         # torch tensor (host pinned or device)
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class Index:
    """One GPU-resident index worker (apss_handle).  Mirrors what IndexingWorkerActor holds: the
    vector store and the inverted index (IWA:22-25), configured by the same three keys."""

    def __init__(self, dim, similarity_threshold, index_threshold=0.0, max_weight=None, device=0, semantics=SEM_R1,
                 tile_vectors=0, kernel_variant=0, reserve_vectors=0, reserve_nnz=0, reserve_pairs=0,
                 pruning=False, prune_alpha=0.0, max_query_norm=0.0, devices=None):
        """devices: list of CUDA ordinals that share the index (id-range shards below the C ABI); None = `device` alone."""
        self._L = load_library()
        self._mw = None if max_weight is None else np.ascontiguousarray(max_weight, dtype=np.float64)
        if self._mw is not None and len(self._mw) != dim:
            raise ValueError("max_weight must have `dim` entries")
        cfg = Config(C.sizeof(Config), int(dim), float(similarity_threshold), float(index_threshold),
                     None if self._mw is None else self._mw.ctypes.data, int(device), int(semantics), int(tile_vectors),
                     int(kernel_variant), int(reserve_vectors), int(reserve_nnz), int(reserve_pairs),
                     int(pruning), 0, float(prune_alpha), float(max_query_norm))
        if devices is not None:
            if not 1 <= len(devices) <= MAX_DEVICES:
                raise ValueError("1..%d devices" % MAX_DEVICES)
            cfg.n_devices = len(devices)
            for i, d in enumerate(devices):
                cfg.device_ids[i] = int(d)
            cfg.device = int(devices[0])
        h = C.c_void_p()
        rc = self._L.apss_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise ApssError(rc, "apss_create failed (dim=%d, device=%d)" % (dim, device))
        self._h = h
        self.dim, self.device, self.semantics = dim, device, semantics

    def close(self):
        if getattr(self, "_h", None):
            self._L.apss_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise ApssError(rc, self._L.apss_last_error(self._h).decode())

    def insert_batch(self, indptr, indices, values, ext_keys=None, first_dim=None, query_only=False, n=None,
                     skip_admit=False, index_only=False) -> BatchResult:
        """Host arrays (numpy / pinned torch) or, when all of them are CUDA tensors, device pointers."""
        on_device = hasattr(indptr, "is_cuda") and indptr.is_cuda
        if isinstance(indptr, np.ndarray) or not hasattr(indptr, "data_ptr"):
            indptr = np.ascontiguousarray(indptr, dtype=np.int64)
            indices = np.ascontiguousarray(indices, dtype=np.int32)
            values = np.ascontiguousarray(values, dtype=np.float64)
            if ext_keys is not None:
                ext_keys = np.ascontiguousarray(ext_keys, dtype=np.int64)
            if first_dim is not None:
                first_dim = np.ascontiguousarray(first_dim, dtype=np.int32)
        nvec = (len(indptr) - 1) if n is None else n
        flags = ((BATCH_QUERY_ONLY if query_only else 0) | (BATCH_DEVICE_PTRS if on_device else 0) |
                 (BATCH_SKIP_ADMIT if skip_admit else 0) | (BATCH_INDEX_ONLY if index_only else 0))
        self._keep = (indptr, indices, values, ext_keys, first_dim)
        out = BatchResultC()
        self._check(self._L.apss_insert_batch(self._h, nvec, _ptr(indptr), _ptr(indices), _ptr(values), _ptr(ext_keys),
                                              _ptr(first_dim), flags, C.byref(out)))
        return BatchResult(*[getattr(out, f) for f, _ in BatchResultC._fields_])

    def fetch_pairs(self, out_q=None, out_c=None, out_sim=None):
        """(q index within the batch, candidate internal id, fp64 similarity) of the last batch."""
        n = C.c_int64()
        self._check(self._L.apss_fetch_pairs(self._h, None, None, None, 0, C.byref(n)))
        m = n.value
        q = np.empty(m, np.int32) if out_q is None else out_q
        c = np.empty(m, np.int32) if out_c is None else out_c
        s = np.empty(m, np.float64) if out_sim is None else out_sim
        if m:
            self._check(self._L.apss_fetch_pairs(self._h, _ptr(q), _ptr(c), _ptr(s), m, C.byref(n)))
        return q[:m], c[:m], s[:m]

    def pairs_device(self):
        q, c, s, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        self._check(self._L.apss_pairs_device(self._h, C.byref(q), C.byref(c), C.byref(s), C.byref(n)))
        return q.value, c.value, s.value, n.value

    def fetch_status(self, n):
        st = np.empty(max(n, 1), np.uint8)
        self._check(self._L.apss_fetch_status(self._h, _ptr(st), n))
        return st[:n]

    def freeze(self):
        self._check(self._L.apss_freeze(self._h))

    def set_next_id(self, next_id):
        self._check(self._L.apss_set_next_id(self._h, int(next_id)))

    def stats(self):
        s = StatsC()
        self._check(self._L.apss_get_stats(self._h, C.byref(s)))
        d = {f: getattr(s, f) for f, _ in StatsC._fields_}
        d["phase_cycles"] = list(d["phase_cycles"])
        d.pop("reserved", None)
        return d

    @property
    def stream_ptr(self):
        return self._L.apss_stream(self._h)


def microbench_accumulators(mode, warps=16, iters=20000, device=0):
    v = C.c_double()
    rc = load_library().apss_microbench_accumulators(device, mode, warps, iters, C.byref(v))
    if rc != 0:
        raise ApssError(rc, "microbench failed")
    return v.value
