"""Shard dispatcher: the index partitioned by vector-id range across the GPUs of one box.

Replaces the reference's "remote router" (EntryProxyActor.scala:37-49,113-122 dimension sharding and
the Akka cluster-sharding hop, CommonUtils.scala:28-46): instead of copying every vector to each
shard that owns one of its dimensions, each rank (one process per GPU) owns whole vectors for a
block-cyclic range of internal ids (block = one insert batch, owner = batch_no mod world).  Per batch:

  1. rank 0 broadcasts the query batch (CSR) to all ranks          -- torch.distributed.broadcast
  2. every rank scores the batch against its own shard; the owner also indexes it (it alone sees
     the in-batch pairs, IWA:125-132)                              -- include/apss.h
  3. per-shard pair lists are gathered back to rank 0              -- one all_gather (counters + pairs)

Every pair is scored exactly once (a vector lives on exactly one shard).  The dispatcher owns the
global id space (apss_set_next_id), so shards report global candidate ids.  Backend-agnostic: with
NCCL the buffers are CUDA tensors handed to the C ABI as device pointers; with gloo (CPU tests of the
host logic) they are host tensors.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


@dataclass
class DispatchResult:
    id_base: int
    owner: int
    n_pairs: int                 # whole job (valid on every rank)
    postings_visited: int
    candidates_unique: int
    q: Optional[np.ndarray]      # rank 0 only: query index within the batch
    c: Optional[np.ndarray]      #              global candidate id
    sim: Optional[np.ndarray]
    local: object                # this rank's native BatchResult


class ShardDispatcher:
    PAIR_SLOT = 8192          # pairs per rank that travel with the counters (16 B each)

    def __init__(self, engine, group=None, device=None):
        """engine: this rank's index worker (native.Index on the rank's GPU)."""
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.backend = dist.get_backend(group) if dist.is_initialized() else "none"
        self.device = torch.device(device) if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu"))
        self.next_id = 0
        self.batch_no = 0
        self.frozen = False
        self._payload = None     # broadcast buffer, grow-only
        # APSS_DISPATCH_TIMING=1: wall-clock per phase (broadcast, local scoring, gather), summed over the calls
        import os
        self.timing = {"bcast": 0.0, "score": 0.0, "gather": 0.0, "calls": 0} if os.environ.get("APSS_DISPATCH_TIMING") else None

    # ---- helpers
    def _bcast(self, t, src=0):
        if self.world > 1:
            dist.broadcast(t, src=src, group=self.group)
        return t

    def owner_of(self, batch_no):
        return batch_no % self.world

    def freeze(self):
        self.frozen = True
        self.engine.freeze()

    def preload(self, indptr, indices, values):
        """Index a batch every rank already holds (bulk load): only the owner touches its shard; no
        scoring, no collectives.  Ids advance on every rank."""
        n = int(indptr.numel() if hasattr(indptr, "numel") else len(indptr)) - 1
        if self.owner_of(self.batch_no) == self.rank:
            self.engine.set_next_id(self.next_id)
            self.engine.insert_batch(indptr, indices, values, index_only=True)
        self.next_id += n
        self.batch_no += 1

    def insert_batch(self, indptr=None, indices=None, values=None, query_only=False) -> DispatchResult:
        """One insertNewVector batch.  Rank 0 passes the batch (torch tensors on self.device, or
        anything torch.as_tensor accepts); other ranks pass nothing."""
        dev = self.device
        import time
        t0 = time.perf_counter()
        # 1. broadcast: sizes, then ONE payload holding the three CSR arrays (indptr | values | indices: every
        #    part starts 8-byte aligned), so a batch costs two collectives instead of four
        if self.rank == 0:
            indptr = torch.as_tensor(indptr, dtype=torch.int64).to(dev).contiguous()
            indices = torch.as_tensor(indices, dtype=torch.int32).to(dev).contiguous()
            values = torch.as_tensor(values, dtype=torch.float64).to(dev).contiguous()
        if self.world > 1:
            hdr = torch.tensor([indptr.numel() - 1, indices.numel()] if self.rank == 0 else [0, 0], dtype=torch.int64, device=dev)
            self._bcast(hdr)
            n, nnz = (int(x) for x in hdr.tolist())
            b0, b1 = 8 * (n + 1), 8 * (n + 1) + 8 * nnz
            # the payload lives in ONE grow-only buffer per rank (25 % head-room): a fresh tensor per batch means a
            # cudaMalloc whenever a batch is a little larger than every one before, and with the shard's large
            # reservations mapped that call takes tens of milliseconds -- in the middle of a stream of 16 ms steps
            nbytes = b1 + 4 * nnz
            if self._payload is None or self._payload.numel() < nbytes:
                self._payload = torch.empty(nbytes + nbytes // 4 + 4096, dtype=torch.uint8, device=dev)
            payload = self._payload[:nbytes]
            if self.rank == 0:
                payload[:b0].copy_(indptr.view(torch.uint8)); payload[b0:b1].copy_(values.view(torch.uint8)); payload[b1:].copy_(indices.view(torch.uint8))
            self._bcast(payload)
            if self.rank != 0:
                indptr = payload[:b0].view(torch.int64)
                values = payload[b0:b1].view(torch.float64)
                indices = payload[b1:].view(torch.int32)
        else:
            n = int(indptr.numel()) - 1
        if dev.type == "cuda":
            torch.cuda.current_stream().synchronize()      # the engine runs on its own stream

        t1 = time.perf_counter()
        # 2. local scoring; the owner indexes
        query_only = query_only or self.frozen
        owner = self.owner_of(self.batch_no)
        mine = (owner == self.rank) and not query_only
        id_base = self.next_id
        if mine:
            self.engine.set_next_id(id_base)
        res = self.engine.insert_batch(indptr, indices, values, query_only=not mine)
        if not query_only:
            self.next_id += n
            self.batch_no += 1

        t2 = time.perf_counter()
        if self.world == 1:
            # single rank: nothing to exchange -- one device-to-host fetch through the C ABI, no torch temporaries
            # (candidate ids are global already: the engine's next_id was set above)
            qn, cn, sn = self.engine.fetch_pairs()
            if self.timing is not None:
                self.timing["bcast"] += t1 - t0; self.timing["score"] += t2 - t1; self.timing["gather"] += time.perf_counter() - t2; self.timing["calls"] += 1
            return DispatchResult(id_base, owner, res.n_pairs, res.postings_visited, res.candidates_unique, qn, cn, sn, res)
        # 3. pair lists to rank 0.  One all_gather of a fixed-size slot per rank carries the counters AND up to
        #    PAIR_SLOT pairs (sim fp64 | q int32 | c int32), so the usual batch needs a single collective; only when a
        #    rank reports more pairs than fit is a second, exactly sized gather issued.
        q, c, s = self._local_pairs(res.n_pairs)
        tot = torch.tensor([res.n_pairs, res.postings_visited, res.candidates_unique], dtype=torch.int64, device=dev)
        if self.world > 1:
            cap = self.PAIR_SLOT
            m = int(q.numel())
            k = min(m, cap)
            slot = torch.zeros(24 + 16 * cap, dtype=torch.uint8, device=dev)
            slot[:24] = tot.view(torch.uint8)
            slot[24:24 + 8 * cap].view(torch.float64)[:k] = s[:k]
            slot[24 + 8 * cap:24 + 12 * cap].view(torch.int32)[:k] = q[:k]
            slot[24 + 12 * cap:].view(torch.int32)[:k] = c[:k]
            slots = [torch.empty_like(slot) for _ in range(self.world)]
            dist.all_gather(slots, slot, group=self.group)
            allc = torch.stack([x[:24].view(torch.int64) for x in slots])
            tot = allc.sum(dim=0)
            counts = allc[:, 0].tolist()
            mx = max(counts)
            if mx <= cap:
                if self.rank == 0:
                    s = torch.cat([x[24:24 + 8 * cap].view(torch.float64)[:k] for x, k in zip(slots, counts)])
                    q = torch.cat([x[24 + 8 * cap:24 + 12 * cap].view(torch.int32)[:k] for x, k in zip(slots, counts)])
                    c = torch.cat([x[24 + 12 * cap:].view(torch.int32)[:k] for x, k in zip(slots, counts)])
            else:
                buf = torch.zeros(16 * mx, dtype=torch.uint8, device=dev)
                buf[:8 * mx].view(torch.float64)[:m] = s
                buf[8 * mx:12 * mx].view(torch.int32)[:m] = q
                buf[12 * mx:].view(torch.int32)[:m] = c
                gl = [torch.empty(16 * mx, dtype=torch.uint8, device=dev) for _ in range(self.world)] if self.rank == 0 else None
                dist.gather(buf, gl, dst=0, group=self.group)
                if self.rank == 0:
                    s = torch.cat([b[:8 * mx].view(torch.float64)[:k] for b, k in zip(gl, counts)])
                    q = torch.cat([b[8 * mx:12 * mx].view(torch.int32)[:k] for b, k in zip(gl, counts)])
                    c = torch.cat([b[12 * mx:].view(torch.int32)[:k] for b, k in zip(gl, counts)])
        tot = tot.tolist()
        if self.rank == 0:
            qn, cn, sn = q.cpu().numpy(), c.cpu().numpy(), s.cpu().numpy()
        else:
            qn = cn = sn = None
        if self.timing is not None:
            t3 = time.perf_counter()
            self.timing["bcast"] += t1 - t0; self.timing["score"] += t2 - t1; self.timing["gather"] += t3 - t2; self.timing["calls"] += 1
        return DispatchResult(id_base, owner, int(tot[0]), int(tot[1]), int(tot[2]), qn, cn, sn, res)

    def _local_pairs(self, n_pairs):
        dev = self.device
        if dev.type == "cuda" and hasattr(self.engine, "pairs_device"):
            qp, cp, sp, m = self.engine.pairs_device()
            q = torch.empty(m, dtype=torch.int32, device=dev)
            c = torch.empty(m, dtype=torch.int32, device=dev)
            s = torch.empty(m, dtype=torch.float64, device=dev)
            if m:
                _copy_from_ptr(q, qp); _copy_from_ptr(c, cp); _copy_from_ptr(s, sp)
            return q, c, s
        q, c, s = self.engine.fetch_pairs()
        return (torch.from_numpy(np.ascontiguousarray(q)).to(dev), torch.from_numpy(np.ascontiguousarray(c)).to(dev),
                torch.from_numpy(np.ascontiguousarray(s)).to(dev))


def _copy_from_ptr(dst: torch.Tensor, src_ptr: int):
    """device-to-device copy from a raw pointer owned by the C library into a torch tensor"""
    import ctypes
    rt = _cudart()
    nbytes = dst.numel() * dst.element_size()
    rc = rt.cudaMemcpy(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src_ptr), ctypes.c_size_t(nbytes), 3)   # cudaMemcpyDeviceToDevice
    if rc != 0:
        raise RuntimeError("cudaMemcpy D2D failed: %d" % rc)


_rt = None


def _cudart():
    global _rt
    if _rt is None:
        import ctypes
        import glob
        import os
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + \
            glob.glob("/usr/local/cuda/lib64/libcudart.so*")
        for cnd in cands + ["libcudart.so.12", "libcudart.so"]:
            try:
                _rt = ctypes.CDLL(cnd)
                break
            except OSError:
                continue
        if _rt is None:
            raise RuntimeError("libcudart not found")
        _rt.cudaMemcpy.restype = ctypes.c_int
    return _rt
