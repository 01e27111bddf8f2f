"""Shard dispatcher host logic with world_size 2 over gloo on CPU (oracle-backed engines injected from
tests/): broadcast of the batch, block-cyclic ownership, global ids, gather of pair lists."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import OracleEngine, csr_slice


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, N, D, t, B, ret, pair_slot=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import apss_b200
    from apss_b200.dispatcher import ShardDispatcher
    data = apss_b200.synth.generate(N, D, 20, seed=4).numpy()
    if pair_slot:
        ShardDispatcher.PAIR_SLOT = pair_slot        # force the second, exactly sized gather
    disp = ShardDispatcher(OracleEngine(D, t), device="cpu")
    # bulk-load the first two batches without scoring, then dispatch the rest
    for lo in (0, B):
        disp.preload(*[torch.from_numpy(np.ascontiguousarray(a)) for a in csr_slice(data, lo, lo + B)])
    got = {}
    tot = [0, 0]
    for lo in range(2 * B, N, B):
        csr = csr_slice(data, lo, min(N, lo + B)) if rank == 0 else (None, None, None)
        r = disp.insert_batch(*csr)
        assert r.id_base == lo and r.owner == (lo // B) % world
        tot[0] += r.postings_visited; tot[1] += r.candidates_unique
        if rank == 0:
            for q, c, s in zip(r.q, r.c, r.sim):
                got[(int(r.id_base + q), int(c))] = float(s)
            assert r.n_pairs == len(r.q)
    # frozen: a query-only batch is scored on every shard and indexed nowhere (2.5x the size of the earlier batches:
    # the grow-only broadcast buffer is re-allocated on every rank, then a small batch re-uses it)
    disp.freeze()
    r = disp.insert_batch(*(csr_slice(data, 0, 500) if rank == 0 else (None, None, None)))
    r2 = disp.insert_batch(*(csr_slice(data, 0, 50) if rank == 0 else (None, None, None)))
    if rank == 0:
        ret["pairs"] = got; ret["tot"] = tot
        ret["frozen_pairs"] = {(int(q), int(c)) for q, c in zip(r.q, r.c)}
        ret["frozen_pairs_small"] = {(int(q), int(c)) for q, c in zip(r2.q, r2.c)}
        ret["next_id"] = disp.next_id
    dist.destroy_process_group()


@pytest.mark.parametrize("pair_slot", [0, 2])
def test_dispatcher_world2_matches_single_oracle(pair_slot):
    from oracle import oracle as orc
    import apss_b200
    N, D, t, B = 1200, 512, 0.5, 200
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), N, D, t, B, ret, pair_slot), nprocs=2, join=True)
    data = apss_b200.synth.generate(N, D, 20, seed=4).numpy()
    o = orc.Oracle(D, t, algo=orc.ALGO_FAST)
    want = {}; tot = [0, 0]
    for lo in range(0, N, B):
        r = o.insert_batch(*csr_slice(data, lo, min(N, lo + B)), index_only=lo < 2 * B)
        if lo >= 2 * B:
            want.update(r.pair_set()); tot[0] += r.postings_visited; tot[1] += r.candidates_unique
    assert dict(ret["pairs"]) == want and len(want) > 0           # bit-exact, every pair exactly once
    assert list(ret["tot"]) == tot
    rq = o.insert_batch(*csr_slice(data, 0, 500), query_only=True)
    assert ret["frozen_pairs"] == {(int(q), int(c)) for q, c in zip(rq.q, rq.c)} and len(rq.q) > 0
    rq2 = o.insert_batch(*csr_slice(data, 0, 50), query_only=True)
    assert ret["frozen_pairs_small"] == {(int(q), int(c)) for q, c in zip(rq2.q, rq2.c)}
    assert ret["next_id"] == N
