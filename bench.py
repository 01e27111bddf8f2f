#!/usr/bin/env python
"""bench.py -- headline benchmark of the all-pairs similarity scoring path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3|C2]

Workload (config.workload): BASELINE.json configs[2] "synthetic 1M vectors x 2^18 dims, Zipf nnz ~100,
cosine >= 0.7" (C3 of SURVEY.md 8(d)).  The index is pre-loaded with the config's N vectors (sharded
block-cyclically over the ranks) and a "step" is one insertNewVector batch of 16384 fresh vectors:
admission + prune, index append, scoring against every indexed vector, fp64 verify, pair output.
Metric: candidate dot-products/s (candidates_unique / time); similar pairs/s alongside.

One JSON line on rank 0.  `value`: inputs resident in HBM, timed with CUDA events on the library's
stream, max over ranks.  `e2e`: the same call with pinned HOST buffers (H2D + pair fetch D2H inside
the timed region).  `roofline`: 8 B x postings_visited / CUDA-event time of the scoring kernel against
the measured HBM copy peak.  `cpu_baseline`: the oracle's restatement of the reference algorithm on
a bounded sample, on this box's host cores.  --impl reference times that CPU path as its own arm.
`roofline.onchip`: accumulator updates/s against the measured shared-memory atomic peak (what binds the
kernel).  `pruned`: the same batches of the value phase through a second engine with exact index reduction
(include/apss.h `pruning` = 2) -- identical pairs, counters of the reduced work, reported beside
the parity-mode headline; --prune N makes it the main arm instead, --no-pruned-leg skips it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

# stdout must carry exactly ONE JSON line, but native libraries (e.g. NCCL's version banner) write to fd 1
# directly: keep the real stdout aside, point fd 1 at stderr for everything else, emit the line at the end.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

import numpy as np
import torch

METRIC = "candidate_dot_products_per_sec"
UNIT = "candidates/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--n-index", type=int, default=0, help="override the number of pre-loaded vectors")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prune", type=int, default=0, nargs="?", const=2,
                    help="main arm with exact index reduction on (1 tile kernel, 2 candidate-major kernel); counters then count the reduced work")
    ap.add_argument("--prune-alpha", type=float, default=0.0)
    ap.add_argument("--no-pruned-leg", action="store_true", help="skip the extra 'pruned' measurement of the default run")
    ap.add_argument("--pruned-mode", type=int, default=3, help="pruning mode of the 'pruned' leg (3 query-major posting-list kernel, 2 candidate-major)")
    ap.add_argument("--shard-gen", action="store_true", help="generate per-rank shards (default for heavy-tail configs)")
    ap.add_argument("--verbose", action="store_true", help="per-step timings on stderr")
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket the timed value region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--cpu-index", type=int, default=50_000, help="index sample size for the CPU arms")
    ap.add_argument("--cpu-queries", type=int, default=0, help="query sample size for the CPU arms (0 = 2 per core)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, windows):
        sm, mx, reasons, pw = [], 0.0, set(), []
        for ts, line in self.rows:
            if not any(a <= ts <= b + 0.15 for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def known_traffic():
    """DRAM bytes per scoring-kernel launch from the committed `ncu --set full` capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def rows_np(data_np, lo, hi):
    ip, ix, v = data_np
    return ip[lo:hi + 1] - ip[lo], ix[ip[lo]:ip[hi]], v[ip[lo]:ip[hi]]


def cpu_sample(data_np, cfg, n_index, q_lo, n_queries, threads, algo_name):
    """The oracle's restatement of the reference path on a bounded sample: `n_queries` query vectors
    scored (query-only) against an index of the first `n_index` vectors.  algo "faithful" = what the
    reference computes (id-only postings, per-candidate hash-join dot, failing candidates re-scored per
    shared dim, IWA:74-111 + CU:98-117); "fast" = weighted postings + dense accumulator."""
    from oracle import oracle as orc
    D, t = cfg["D"], cfg["threshold"]
    o = orc.Oracle(D, t, algo=orc.ALGO_FAITHFUL if algo_name == "faithful" else orc.ALGO_FAST, semantics=orc.R1, threads=threads)
    for lo in range(0, n_index, 16384):
        o.insert_batch(*rows_np(data_np, lo, min(n_index, lo + 16384)), index_only=True)
    q = rows_np(data_np, q_lo, q_lo + n_queries)
    t0 = time.perf_counter()
    r = o.insert_batch(*q, query_only=True)
    dt = time.perf_counter() - t0
    o.close()
    return r, dt


def host_threads():
    """All host cores this process may use.  (torchrun exports OMP_NUM_THREADS=1; the oracle takes its thread
    count explicitly, so that default does not throttle the CPU arms.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference_arm(args, cfg, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the Scala original cannot run:
    no JVM) on this box's host cores.  Rank 0 alone works; other ranks exit 0."""
    if rank != 0:
        return
    from apss_b200 import synth
    threads = host_threads()
    n_index = args.cpu_index
    nq = args.cpu_queries or 2 * threads
    total_q = nq * (args.steps + args.warmup)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    data = synth.generate(n_index + total_q, cfg["D"], cfg["nnz_mean"], s=cfg["s"], seed=cfg["seed"], device=dev)
    data_np = data.numpy()
    from oracle import oracle as orc
    o = orc.Oracle(cfg["D"], cfg["threshold"], algo=orc.ALGO_FAITHFUL, semantics=orc.R1, threads=threads)
    ofast = orc.Oracle(cfg["D"], cfg["threshold"], algo=orc.ALGO_FAST, threads=threads)
    for lo in range(0, n_index, 16384):
        b = rows_np(data_np, lo, min(n_index, lo + 16384))
        o.insert_batch(*b, index_only=True); ofast.insert_batch(*b, index_only=True)
    cands = pairs = 0
    tt = 0.0
    for step in range(args.warmup + args.steps):
        q = rows_np(data_np, n_index + step * nq, n_index + (step + 1) * nq)
        t0 = time.perf_counter()
        r = o.insert_batch(*q, query_only=True)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            rf = ofast.insert_batch(*q, query_only=True)      # untimed: the candidate count of the same sample
            assert rf.pair_set() == r.pair_set()
            cands += rf.candidates_unique; pairs += len(r.sim); tt += dt
    val = cands / tt
    sample = "%d query vectors/step scored (query-only) against the first %d vectors of %s" % (nq, n_index, cfg["name"])
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"], "sample": sample},
            "pairs_per_sec": pairs / tt,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    args = parse_args()
    from apss_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict(synth.CONFIGS[args.config]); cfg["name"] = args.config
    if args.n_index:
        cfg["N"] = args.n_index
    if args.batch:
        cfg["batch"] = args.batch
    cfg["workload"] = "%s: synthetic %d vectors x 2^%d dims, Zipf(s=1) nnz~%d, cosine>=%.2f, batches of %d" % (
        args.config, cfg["N"], int(np.log2(cfg["D"])), cfg["nnz_mean"], cfg["threshold"], cfg["batch"])
    if args.impl == "reference":
        return run_reference_arm(args, cfg, rank, world)

    import torch.distributed as dist
    from apss_b200 import native
    from apss_b200.dispatcher import ShardDispatcher
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W, B, N, D, t = args.steps, args.warmup, cfg["batch"], cfg["N"], cfg["D"], cfg["threshold"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                            # early: see the wait before the timed region
    n_fresh = (W + K) + (1 + K)                    # value phase + e2e phase (1 warm-up)
    t_gen = time.time()
    shard_gen = bool(cfg.get("heavy_tail")) or args.shard_gen
    if not shard_gen:
        # every rank generates the whole data set (identical on all ranks); owners index their batches
        data = synth.generate(N + n_fresh * B, D, cfg["nnz_mean"], s=cfg["s"], seed=cfg["seed"], device=dev)
        torch.cuda.synchronize()
        t_gen = time.time() - t_gen

        def fresh_rows(i, pin=False):
            b = data.rows(N + i * B, N + (i + 1) * B)
            if pin:
                b = b.pin()
                return b.indptr, b.indices, b.values
            return b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous()

        def dev_rows(lo, hi):
            b = data.rows(lo, hi)
            return b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous()

        total_nnz = data.nnz
        eng = native.Index(D, t, device=local_rank, tile_vectors=args.tile, kernel_variant=args.variant, pruning=args.prune, prune_alpha=args.prune_alpha,
                           reserve_vectors=int((N + n_fresh * B) / world * 1.1) + 2 * B, reserve_nnz=int(total_nnz / world * 1.15) + (1 << 20))
        disp = ShardDispatcher(eng, device=dev)
        t_load = time.time()
        for lo in range(0, N, B):
            disp.preload(*dev_rows(lo, min(N, lo + B)))
        torch.cuda.synchronize()
        t_load = time.time() - t_load
    else:
        # per-rank shards (config C5: too large to generate everywhere): rank r generates and indexes the
        # batches r, r + world, ... itself; document frequencies are all-reduced so that every rank uses the
        # same IDF; rank 0 also generates the fresh query batches
        per_rank = max(1, (N // B) // world)
        N = per_rank * world * B
        fs = synth.generate_flat(per_rank * B, D, cfg["nnz_mean"], s=cfg["s"], seed=cfg["seed"] + 1009 * rank, device=dev,
                                 heavy_tail=bool(cfg.get("heavy_tail")))
        df = fs.df()
        if world > 1:
            dist.all_reduce(df)
        idf = synth.idf_from_df(df, N)
        shard = fs.finalize(idf)
        del fs
        fresh = None
        if rank == 0:
            fresh = synth.generate_flat(n_fresh * B, D, cfg["nnz_mean"], s=cfg["s"], seed=cfg["seed"] + 777, device=dev,
                                        heavy_tail=bool(cfg.get("heavy_tail"))).finalize(idf)
        torch.cuda.synchronize()
        t_gen = time.time() - t_gen

        def fresh_rows(i, pin=False):
            b = fresh.rows(i * B, (i + 1) * B)
            if pin:
                b = b.pin()
                return b.indptr, b.indices, b.values
            return b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous()

        total_nnz = shard.nnz * world
        eng = native.Index(D, t, device=local_rank, tile_vectors=args.tile, kernel_variant=args.variant, pruning=args.prune, prune_alpha=args.prune_alpha,
                           reserve_vectors=int((N + n_fresh * B) / world * 1.1) + 2 * B, reserve_nnz=int(shard.nnz * 1.15) + (1 << 22))
        disp = ShardDispatcher(eng, device=dev)
        t_load = time.time()
        for j in range(per_rank):
            b = shard.rows(j * B, (j + 1) * B)
            rows = (b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous())
            torch.cuda.synchronize()
            eng.set_next_id((j * world + rank) * B)
            eng.insert_batch(*rows, index_only=True)
        disp.next_id = N
        disp.batch_no = N // B
        torch.cuda.synchronize()
        t_load = time.time() - t_load
        cfg["workload"] = "%s: synthetic %d vectors x 2^%d dims, Zipf(s=1) %s nnz~%d, cosine>=%.2f, batches of %d" % (
            args.config, N, int(np.log2(D)), "heavy-tail" if cfg.get("heavy_tail") else "", cfg["nnz_mean"], t, B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        tns = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns[0])

    lib_stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    if rank == 0:          # nvidia-smi's start-up stalls the driver for ~1 s: it must be over before anything is timed
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 10.0:
            time.sleep(0.05)
    cursor = N

    # ------------------------------------------------------------ value: inputs resident in HBM
    fresh_i = [0]

    def step_device(_lo):
        i = fresh_i[0]; fresh_i[0] += 1
        return disp.insert_batch(*(fresh_rows(i) if rank == 0 else (None, None, None)))

    for _ in range(W):
        step_device(cursor); cursor += B
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.stats()["kernel_launches"]
    barrier()
    if args.profile_range:
        torch.cuda.profiler.start()
    w0 = time.time()
    ev0.record(lib_stream)
    tot = dict(cands=0, pairs=0, postings=0, score_ms=0.0, local_postings=0, items=0)
    step_wall = []
    for _ in range(K):
        ts_ = time.time()
        r = step_device(cursor); cursor += B
        step_wall.append((time.time() - ts_) * 1e3)
        if args.verbose and rank == 0:
            print("value step: wall %.1f ms score %.1f ms device %.1f ms pairs %d prefilter %d items %d" % (
                (time.time() - ts_) * 1e3, r.local.score_ms, r.local.device_ms, r.n_pairs, r.local.n_prefilter, r.local.work_items), file=sys.stderr, flush=True)
        tot["cands"] += r.candidates_unique; tot["pairs"] += r.n_pairs; tot["postings"] += r.postings_visited
        tot["score_ms"] += r.local.score_ms; tot["local_postings"] += r.local.postings_visited; tot["items"] += r.local.work_items
    ev1.record(lib_stream)
    barrier()
    w1 = time.time()
    if args.profile_range:
        torch.cuda.profiler.stop()
    st = eng.stats()
    launches = st["kernel_launches"] - launches0
    if disp.timing is not None and rank == 0:
        tm = disp.timing
        print("dispatcher timing per call (ms): bcast %.3f score %.3f gather %.3f | kernel %.3f" % (
            1e3 * tm["bcast"] / tm["calls"], 1e3 * tm["score"] / tm["calls"], 1e3 * tm["gather"] / tm["calls"], tot["score_ms"] / K), file=sys.stderr, flush=True)
    score_launches = K
    dt_value = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    wall_value = w1 - w0

    # ------------------------------------------------------------ e2e: pinned host buffers in, pairs out to host
    host_batches = []
    if rank == 0:
        for k in range(1 + K):
            host_batches.append(fresh_rows(W + K + k, pin=True))
    out_q = torch.empty(1 << 22, dtype=torch.int32).pin_memory()
    out_c = torch.empty(1 << 22, dtype=torch.int32).pin_memory()
    out_s = torch.empty(1 << 22, dtype=torch.float64).pin_memory()

    def step_host(k):
        if world == 1:
            ip, ix, v = host_batches[k]
            ts_ = time.time()
            r = eng.insert_batch(ip.numpy(), ix.numpy(), v.numpy())      # H2D inside the call
            tm_ = time.time()
            eng.fetch_pairs(out_q.numpy(), out_c.numpy(), out_s.numpy())  # D2H of the result
            if args.verbose:
                print("e2e step: insert %.1f ms fetch %.1f ms score %.1f ms device %.1f ms pairs %d prefilter %d" % (
                    (tm_ - ts_) * 1e3, (time.time() - tm_) * 1e3, r.score_ms, r.device_ms, r.n_pairs, r.n_prefilter), file=sys.stderr, flush=True)
            return r.candidates_unique, r.n_pairs, ip.numel() * 8 + ix.numel() * 4 + v.numel() * 8, r.n_pairs * 16
        ip, ix, v = host_batches[k] if rank == 0 else (None, None, None)
        r = disp.insert_batch(ip, ix, v)                                  # rank 0: pinned host -> device -> broadcast; pairs -> host
        h2d = (ip.numel() * 8 + ix.numel() * 4 + v.numel() * 8) if rank == 0 else 0
        return r.candidates_unique, r.n_pairs, h2d, r.n_pairs * 16

    step_host(0)
    barrier()
    w2 = time.time()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(lib_stream)
    e_c = e_p = e_h2d = e_d2h = 0
    for k in range(1, 1 + K):
        c_, p_, a_, b_ = step_host(k)
        e_c += c_; e_p += p_; e_h2d += a_; e_d2h += b_
    ev3.record(lib_stream)
    barrier()
    w3 = time.time()
    dt_e2e = max_over_ranks(max(ev2.elapsed_time(ev3) * 1e-3, 0.0))
    wall_e2e = w3 - w2

    # ------------------------------------------------------------ pruned leg: the same W + K batches of the value phase
    # through a second engine with exact index reduction on.  Same pairs, far less work; its
    # counters count the reduced work, so it is reported beside the parity-mode headline, not instead of it.
    pruned = None
    if not args.prune and not args.no_pruned_leg and not shard_gen:
        eng.close()
        del disp
        eng2 = native.Index(D, t, device=local_rank, pruning=args.pruned_mode, prune_alpha=args.prune_alpha,
                            reserve_vectors=int((N + n_fresh * B) / world * 1.1) + 2 * B, reserve_nnz=int(total_nnz / world * 1.15) + (1 << 20))
        disp2 = ShardDispatcher(eng2, device=dev)
        for lo in range(0, N, B):
            disp2.preload(*dev_rows(lo, min(N, lo + B)))
        lib_stream2 = torch.cuda.ExternalStream(eng2.stream_ptr, device=dev)
        p_tot = dict(cands=0, pairs=0, postings=0, score_ms=0.0, prefilter=0)
        for i in range(W):
            disp2.insert_batch(*(fresh_rows(i) if rank == 0 else (None, None, None)))
        ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w4 = time.time()
        ev4.record(lib_stream2)
        for i in range(W, W + K):
            r = disp2.insert_batch(*(fresh_rows(i) if rank == 0 else (None, None, None)))
            p_tot["cands"] += r.candidates_unique; p_tot["pairs"] += r.n_pairs; p_tot["postings"] += r.postings_visited
            p_tot["score_ms"] += r.local.score_ms; p_tot["prefilter"] += r.local.n_prefilter
        ev5.record(lib_stream2)
        barrier()
        w5 = time.time()
        dt_pr = max_over_ranks(ev4.elapsed_time(ev5) * 1e-3)
        st2 = eng2.stats()
        # the same K host batches as the e2e phase above, through host buffers (H2D and result fetch inside the call)
        pe = None
        if world == 1:
            ip, ix, v = host_batches[0]
            eng2.insert_batch(ip.numpy(), ix.numpy(), v.numpy()); eng2.fetch_pairs(out_q.numpy(), out_c.numpy(), out_s.numpy())
            torch.cuda.synchronize()
            ev6, ev7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w6 = time.time()
            ev6.record(lib_stream2)
            pe_pairs = 0
            for k in range(1, 1 + K):
                ip, ix, v = host_batches[k]
                r = eng2.insert_batch(ip.numpy(), ix.numpy(), v.numpy())
                eng2.fetch_pairs(out_q.numpy(), out_c.numpy(), out_s.numpy())
                pe_pairs += r.n_pairs
            ev7.record(lib_stream2)
            torch.cuda.synchronize()
            w7 = time.time()
            pe = {"ms_per_step": ev6.elapsed_time(ev7) / K, "wall_ms_per_step": (w7 - w6) * 1e3 / K,
                  "pairs_per_sec": pe_pairs / (ev6.elapsed_time(ev7) * 1e-3), "pairs_identical_to_parity_run": pe_pairs == e_p}
        pruned = {"kernel": "apss::k_score_qm (query-major posting-list traversal, reduced index)" if args.pruned_mode == 3 else "apss::k_score_cand (candidate-major, reduced index)", "ms_per_step": dt_pr / K * 1e3,
                  "pairs_per_sec": p_tot["pairs"] / dt_pr, "pairs_identical_to_parity_run": p_tot["pairs"] == tot["pairs"],
                  "speedup_vs_parity_run": dt_value / dt_pr,
                  "equivalent_candidates_per_sec": tot["cands"] / dt_pr,
                  "postings_visited_per_step": p_tot["postings"] / K, "candidates_touched_per_step": p_tot["cands"] / K,
                  "verify_records_per_step": p_tot["prefilter"] / K, "score_kernel_ms_per_step": p_tot["score_ms"] / K,
                  "unindexed_fraction": st2["n_unindexed"] / max(st2["n_unindexed"] + st2["n_postings"], 1),
                  "work_reduction_postings": tot["postings"] / max(p_tot["postings"], 1),
                  "e2e": pe,
                  "note": "exact index reduction (include/apss.h `pruning` = 2): same batches as the value phase, same pair set; "
                          "equivalent_candidates_per_sec = candidates of the parity run / time of this run"}
        windows_extra = [(w4, w5)]
    else:
        windows_extra = []
    sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    achieved = 8.0 * tot["local_postings"] / (tot["score_ms"] * 1e-3) / 1e9 if tot["score_ms"] > 0 else 0.0
    traffic = known_traffic()
    line = {
        "metric": METRIC, "value": tot["cands"] / dt_value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dt_value / K * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 accumulate + f64 verify", "data": "synthetic",
        "config": {"workload": cfg["workload"], "index_vectors": N, "batch": B, "sharding": "id-range block-cyclic x%d" % world,
                   "l2": "inputs larger than L2 (index %.0f MB postings per GPU)" % (st["bytes_postings"] / 1e6),
                   "tile_vectors": st["tile_vectors"], "warps_per_cta": st["warps_per_cta"], "kernel_variant": args.variant,
                   "pruning": ("exact index reduction ON: %.1f%% of stored components un-indexed; counters count the reduced work"
                               % (100.0 * st["n_unindexed"] / max(st["n_unindexed"] + st["n_postings"], 1))) if args.prune else "off (parity counters)"},
        "pairs_per_sec": tot["pairs"] / dt_value,
        "postings_per_sec": tot["postings"] / dt_value,
        "step_latency_ms": {"min": float(np.min(step_wall)), "p50": float(np.percentile(step_wall, 50)), "max": float(np.max(step_wall))},
        "wall_s_value": wall_value, "wall_s_e2e": wall_e2e, "gen_s": t_gen, "preload_s": t_load,
        "e2e": {"value": e_c / dt_e2e if dt_e2e > 0 else None, "unit": UNIT, "h2d_bytes_per_step": e_h2d // K,
                "d2h_bytes_per_step": e_d2h // K, "pairs_per_sec": e_p / dt_e2e if dt_e2e > 0 else None},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None if not traffic else traffic.get("dram_bytes_per_launch"),
                     "traffic_source": None if not traffic else traffic.get("kernel"),
                     "ncu": None if not traffic else traffic.get("ncu"),
                     "peak_source": peak_src,
                     "kernel": {1: "apss::k_score", 2: "apss::k_score_blk"}.get((args.variant >> 16) & 0xff, "apss::k_score_dense"),
                     "launches_timed": score_launches,
                     "algorithmic_bytes_per_launch": 8.0 * tot["local_postings"] / max(score_launches, 1),
                     "avg_launch_ms": tot["score_ms"] / max(score_launches, 1),
                     "kernel_share_of_step": tot["score_ms"] * 1e-3 / dt_value,
                     "note": "algorithmic bytes = 8 B per posting visited per QUERY TERM (SURVEY 8d); lists are re-read "
                             "from L2/shared memory across the queries of a batch, so DRAM traffic << algorithmic bytes "
                             "and achieved may exceed the HBM copy peak; the binding resource is on-chip (instruction issue / "
                             "shared-memory atomics), see the ncu figures"},
        "clocks": sampler.summary([(w0, w1), (w2, w3)] + windows_extra),
    }
    if pruned is not None:
        line["pruned"] = pruned
    try:
        # the binding resource is on-chip: one shared-memory accumulator update per posting visited, measured against
        # the micro-benchmarked peak of native u32 shared-memory atomic adds (random addresses, 16 warps per SM)
        peak_upd = 2.547e12
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_probe_microbench_and_v1_sweep.json")) as f:
                peak_upd = float(json.load(f)["microbench"]["atoms_u32_random_w16"])
        except Exception:
            pass
        upd = tot["local_postings"] / (tot["score_ms"] * 1e-3) if tot["score_ms"] > 0 else 0.0
        line["roofline"]["onchip"] = {"what": "accumulator updates/s per GPU (one per posting visited; dense dims go through FFMA instead)",
                                      "achieved": upd, "peak": peak_upd, "frac": upd / peak_upd,
                                      "peak_source": "measured ATOMS.ADD.U32 throughput, profiles/r01_probe_microbench_and_v1_sweep.json"}
    except Exception:
        pass
    if not args.no_cpu_baseline and not shard_gen:
        threads = host_threads()
        nq = args.cpu_queries or 2 * threads
        n_index = min(args.cpu_index, N)
        ip = data.indptr[: n_index + 1].cpu().numpy()
        nnz_i = int(ip[-1])
        idx_np = (ip, data.indices[:nnz_i].cpu().numpy(), data.values[:nnz_i].cpu().numpy())
        qb = data.rows(N, N + nq)
        q_np = qb.numpy()
        data_np = (np.concatenate([ip, ip[-1] + q_np[0][1:]]), np.concatenate([idx_np[1], q_np[1]]), np.concatenate([idx_np[2], q_np[2]]))
        rf, dt_f = cpu_sample(data_np, cfg, n_index, n_index, nq, threads, "faithful")
        ro, dt_o = cpu_sample(data_np, cfg, n_index, n_index, nq, threads, "fast")
        assert rf.pair_set() == ro.pair_set()
        line["cpu_baseline"] = {"value": ro.candidates_unique / dt_f, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d query vectors scored (query-only) against the first %d vectors of %s; reference "
                                          "algorithm (id-only postings + per-candidate hash-join dot, IWA:74-111/CU:98-117)" % (nq, n_index, args.config),
                                "seconds": dt_f,
                                "opt_value": ro.candidates_unique / dt_o, "opt_seconds": dt_o,
                                "opt_note": "same sample, weighted postings + dense accumulator (the GPU algorithm on CPU)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
