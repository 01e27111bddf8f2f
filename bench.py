#!/usr/bin/env python
"""bench.py -- headline benchmark of the all-pairs similarity scoring path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3|C2]

Workload (config.workload): BASELINE.json configs[2] "synthetic 1M vectors x 2^18 dims, Zipf nnz ~100,
cosine >= 0.7" (C3 of SURVEY.md 8(d)), generator g2 (counter-based: the same bits on every device).  The index is
pre-loaded with the config's N vectors (sharded block-cyclically over the ranks) and a "step" is one
insertNewVector batch of 16384 fresh vectors: admission + prune, index append, scoring against every indexed
vector, fp64 verify, pair output.

One JSON line on rank 0:
  value / e2e   parity mode (every posting visited, the reference's own counters): candidate dot-products/s with the
                inputs resident in HBM, and the same call with pinned HOST buffers (H2D + pair fetch D2H inside the
                timed region).  CUDA events on the library's stream, max over ranks.
  pruned        the same batches through exact index reduction (include/apss.h pruning = 3: query-major posting-list
                traversal, bulk-async producer/consumer kernel): identical pair set (hash compared), far less work.
  roofline      per kernel, each against the resource that binds it, peaks measured in this run:
                parity kernel  time at speed of light (sparse updates / ATOMS peak + dense FMAs / FP32 peak) / time;
                pruned kernel  8 B x postings visited / kernel time / measured HBM copy peak (every visit is a
                               physical 8-byte read, streamed by cp.async.bulk: no on-chip reuse across queries).
  parity        after the timed phases a sample of never-indexed query vectors is scored (query-only) by BOTH engines
                against the full sharded index and compared with the CPU oracle on rank 0: pair set, fp64
                similarities bit for bit, and (parity mode) candidates_unique / postings_visited.
  allpairs      the north-star job: the config's N vectors from an EMPTY index, batch by batch (index, then query:
                IWA:122-134), in both modes; total seconds, pairs, pairs/s and an order-independent hash of the
                (q, c, sim) set, equal across modes and across --gpus N.
  cpu_baseline  the oracle's restatement of the reference algorithm (cpu_ref, `value`) and the accumulator algorithm
                (cpu_opt) on this box's host cores, same index, bounded query samples.
--impl reference times the CPU path as its own arm on the same config (rank 0 only).
Timed regions: the step's input batches are cut out of the generated matrix BEFORE the clock starts (resident in HBM);
every step fetches its pairs to the host inside the region; the order-independent pair hash is computed after it.
`clocks`: SM clock / power / throttle reasons every 100 ms during the timed regions, read through in-process NVML.
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
    ap.add_argument("--prune", type=int, default=0, nargs="?", const=3,
                    help="main arm with exact index reduction on (3 query-major, 2 candidate-major, 1 tile kernel); counters then count the reduced work")
    ap.add_argument("--prune-alpha", type=float, default=0.0)
    ap.add_argument("--no-pruned-leg", action="store_true", help="skip the extra 'pruned' measurement of the default run")
    ap.add_argument("--pruned-mode", type=int, default=3, help="pruning mode of the 'pruned' leg")
    ap.add_argument("--no-allpairs", action="store_true", help="skip the all-pairs job from an empty index")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled oracle check")
    ap.add_argument("--parity-queries", type=int, default=256)
    ap.add_argument("--shard-gen", action="store_true", help="generate per-rank shards (default for heavy-tail configs)")
    ap.add_argument("--verbose", action="store_true", help="per-step timings on stderr")
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket the timed value region of one leg with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--profile-leg", default="main", choices=["main", "pruned"])
    ap.add_argument("--cpu-opt-queries", type=int, default=4096, help="query sample of the accumulator CPU arm")
    ap.add_argument("--cpu-queries", type=int, default=0, help="query sample of the reference-algorithm CPU arm (0 = cores / 4, >= 4)")
    return ap.parse_args()


class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region (the B200_PROFILING.md clocks line), every 100 ms.
    Read through NVML in-process (what nvidia-smi itself reads): a polling `nvidia-smi -lms` child stalls the driver for
    ~1 s at start-up and now and then for tens of ms per poll on a multi-GPU box -- inside a stream of 16 ms steps.
    Falls back to the nvidia-smi loop when the NVML binding is missing.  Rows: (time, "sm,max,power,hw,hwth,swth,swcap")."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc, self._stop = gpu_index, [], None, False

    def _nvml_loop(self, nv, hdl):
        bits = [getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8), getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = nv.nvmlDeviceGetMaxClockInfo(hdl, nv.NVML_CLOCK_SM)
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(hdl, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(hdl) / 1000.0
                r = int(get_reasons(hdl))
                self.rows.append((time.time(), "%d, %d, %.2f, %s" % (sm, mx, pw, ", ".join("Active" if r & b else "Not Active" for b in bits))))
            except Exception:      # noqa: BLE001
                pass
            time.sleep(0.1)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists ordinals
            vis = [x for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip().isdigit()]
            idx = int(vis[self.gpu]) if self.gpu < len(vis) else self.gpu
            hdl = nv.nvmlDeviceGetHandleByIndex(idx)
            threading.Thread(target=self._nvml_loop, args=(nv, hdl), daemon=True).start()
            return
        except Exception:      # noqa: BLE001
            pass
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
        self._stop = True
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


def committed_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of THIS configuration at 1 GPU
    (profiles/traffic.json: {kernel: {dram_bytes_per_launch, ncu}}), labelled with the capture it came from; None if
    there is none.  Only attached to a line that runs that very configuration (C3, 1 GPU, default sizes)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        rec = json.load(open(p)).get(kernel)
        return rec if isinstance(rec, dict) else None
    except Exception:
        return None


def host_threads():
    """All host cores this process may use.  (torchrun exports OMP_NUM_THREADS=1; the oracle takes its thread
    count explicitly, so that default does not throttle the CPU arms.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def workload_name(name, cfg, N):
    return "%s: synthetic %d vectors x 2^%d dims, Zipf(s=1) %snnz~%d, cosine>=%.2f, batches of %d" % (
        name, N, int(np.log2(cfg["D"])), "heavy-tail " if cfg.get("heavy_tail") else "", cfg["nnz_mean"], cfg["threshold"], cfg["batch"])


_U = np.uint64


def pairset_hash(q_global, c, sim):
    """order-independent 64-bit digest of a multiset of (query id, candidate id, fp64 similarity): wrap-around sum of a
    splitmix64 finaliser over each pair -- equal sets give equal digests whatever the shard / arrival order"""
    if len(q_global) == 0:
        return 0
    with np.errstate(over="ignore"):
        x = (np.asarray(q_global).astype(np.int64).astype(_U) << _U(32)) ^ np.asarray(c).astype(np.int64).astype(_U)
        x = x * _U(0x9E3779B97F4A7C15) ^ np.ascontiguousarray(sim, np.float64).view(_U) * _U(0xD1B54A32D192ED03)
        x = (x ^ (x >> _U(30))) * _U(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> _U(27))) * _U(0x94D049BB133111EB)
        x = x ^ (x >> _U(31))
        return int(np.sum(x, dtype=_U))


def rows_np(data_np, lo, hi):
    ip, ix, v = data_np
    return ip[lo:hi + 1] - ip[lo], ix[ip[lo]:ip[hi]], v[ip[lo]:ip[hi]]


def build_oracle(data_np, cfg, n_index, threads, algo_name):
    from oracle import oracle as orc
    o = orc.Oracle(cfg["D"], cfg["threshold"], algo=orc.ALGO_FAITHFUL if algo_name == "faithful" else orc.ALGO_FAST,
                   semantics=orc.R1, threads=threads)
    t0 = time.perf_counter()
    for lo in range(0, n_index, 16384):
        o.insert_batch(*rows_np(data_np, lo, min(n_index, lo + 16384)), index_only=True)
    return o, time.perf_counter() - t0


REF_ALGO = ("reference algorithm: id-only postings in hash sets + per-candidate hash-join dot, failing candidates re-scored per shared "
            "dimension (IWA:74-111, CU:98-117); all cores (a query's candidates are striped over the threads by key)")


# ---------------------------------------------------------------------------------------------- reference arm

def run_reference_arm(args, cfg, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the Scala original cannot run: no JVM) on this
    box's host cores, on the SAME configuration as the GPU arm: the config's N vectors indexed, each step a bounded
    sample of query vectors of the next insert batch scored against the whole index.  Rank 0 alone works."""
    if rank != 0:
        return
    from apss_b200 import synth
    threads = host_threads()
    N, B = cfg["N"], cfg["batch"]
    nq = args.cpu_queries or max(2, threads // 8)
    total_q = nq * (args.steps + args.warmup)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    data = synth.generate(N + total_q, cfg["D"], cfg["nnz_mean"], s=cfg["s"], seed=cfg["seed"], device=dev, heavy_tail=bool(cfg.get("heavy_tail")))
    data_np = data.numpy()
    del data
    o, t_build = build_oracle(data_np, cfg, N, threads, "faithful")
    ofast, _ = build_oracle(data_np, cfg, N, threads, "fast")
    cands = pairs = 0
    tt = 0.0
    for step in range(args.warmup + args.steps):
        q = rows_np(data_np, N + step * nq, N + (step + 1) * nq)
        t0 = time.perf_counter()
        r = o.insert_batch(*q, query_only=True)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            rf = ofast.insert_batch(*q, query_only=True)      # untimed: the candidate count of the same sample
            assert rf.pair_set() == r.pair_set()
            cands += rf.candidates_unique; pairs += len(r.sim); tt += dt
    val = cands / tt
    sample = "%d query vectors per step (of a %d-vector insert batch) scored query-only against the %d indexed vectors of %s; %s" % (
        nq, B, N, cfg["name"], REF_ALGO)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"], "index_vectors": N, "batch": B, "generator": synth.GEN_VERSION, "sample": sample},
            "pairs_per_sec": pairs / tt, "index_build_s": t_build,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------- GPU arm

class Ctx:
    pass


def run_leg(cx, mode, profile=False):
    """One engine mode through the whole measurement: pre-load, value phase (device-resident inputs), e2e phase (pinned
    host inputs, pairs to the host), then the untimed parity sample (query-only) against the final index."""
    from apss_b200 import native
    from apss_b200.dispatcher import ShardDispatcher
    args, cfg, dev, rank, world = cx.args, cx.cfg, cx.dev, cx.rank, cx.world
    K, W, B, N, D, t = args.steps, args.warmup, cfg["batch"], cx.N, cfg["D"], cfg["threshold"]
    eng = native.Index(D, t, device=cx.local_rank, tile_vectors=args.tile, kernel_variant=args.variant if not mode else 0, pruning=mode,
                       prune_alpha=args.prune_alpha, reserve_vectors=int((N + cx.n_fresh * B) / world * 1.1) + 2 * B,
                       reserve_nnz=int(cx.total_nnz / world * 1.15) + (1 << 20))
    disp = ShardDispatcher(eng, device=dev)
    t_load = time.time()
    cx.preload(disp, eng)
    torch.cuda.synchronize()
    t_load = time.time() - t_load
    lib_stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    leg = {"mode": mode, "preload_s": t_load}

    # the step's inputs are resident in HBM before the timed region starts: the batches are cut out of the generated
    # matrix (re-based indptr, contiguous arrays) up front, not inside the loop
    staged = [cx.fresh_rows(i) if rank == 0 else (None, None, None) for i in range(W + K)]
    torch.cuda.synchronize()

    def step_device(i):
        return disp.insert_batch(*staged[i])

    for i in range(W):
        step_device(i)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.stats()["kernel_launches"]
    cx.barrier()
    if profile:
        torch.cuda.profiler.start()
    w0 = time.time()
    ev0.record(lib_stream)
    tot = dict(cands=0, pairs=0, postings=0, score_ms=0.0, local_postings=0, prefilter=0, dense_post=0, dense_fma=0, hash=0)
    step_wall = []
    kept = []
    for i in range(W, W + K):
        ts_ = time.time()
        r = step_device(i)
        step_wall.append((time.time() - ts_) * 1e3)
        if args.verbose and rank == 0:
            print("mode %d value step: wall %.2f ms score %.2f ms device %.2f ms pairs %d prefilter %d" % (
                mode, step_wall[-1], r.local.score_ms, r.local.device_ms, r.n_pairs, r.local.n_prefilter), file=sys.stderr, flush=True)
        tot["cands"] += r.candidates_unique; tot["pairs"] += r.n_pairs; tot["postings"] += r.postings_visited
        tot["score_ms"] += r.local.score_ms; tot["local_postings"] += r.local.postings_visited; tot["prefilter"] += r.local.n_prefilter
        tot["dense_post"] += r.local.dense_postings; tot["dense_fma"] += r.local.dense_fma
        if rank == 0:
            kept.append((r.q, r.id_base, r.c, r.sim))      # the step's pairs are on the host; hashed after the timed region
    ev1.record(lib_stream)
    cx.barrier()
    w1 = time.time()
    if profile:
        torch.cuda.profiler.stop()
    for q_, b_, c_, s_ in kept:
        tot["hash"] = (tot["hash"] + pairset_hash(q_.astype(np.int64) + b_, c_, s_)) & 0xFFFFFFFFFFFFFFFF
    leg.update(tot=tot, step_wall=step_wall, window=(w0, w1), dt_value=cx.max_over_ranks(ev0.elapsed_time(ev1) * 1e-3),
               launches=eng.stats()["kernel_launches"] - launches0)
    if disp.timing is not None and rank == 0:
        tm = disp.timing
        print("mode %d dispatcher timing per call (ms): bcast %.3f score %.3f gather %.3f | kernel %.3f" % (
            mode, 1e3 * tm["bcast"] / tm["calls"], 1e3 * tm["score"] / tm["calls"], 1e3 * tm["gather"] / tm["calls"], tot["score_ms"] / K), file=sys.stderr, flush=True)

    # ---- e2e: pinned host buffers in, pairs out to the host
    def step_host(k):
        if world == 1:
            ip, ix, v = cx.host_batches[k]
            r = eng.insert_batch(ip.numpy(), ix.numpy(), v.numpy())                 # H2D inside the call
            eng.fetch_pairs(cx.out_q.numpy(), cx.out_c.numpy(), cx.out_s.numpy())   # D2H of the result
            return r.candidates_unique, r.n_pairs, ip.numel() * 8 + ix.numel() * 4 + v.numel() * 8, r.n_pairs * 16
        ip, ix, v = cx.host_batches[k] if rank == 0 else (None, None, None)
        r = disp.insert_batch(ip, ix, v)                    # rank 0: pinned host -> device -> broadcast; pairs -> host
        h2d = (ip.numel() * 8 + ix.numel() * 4 + v.numel() * 8) if rank == 0 else 0
        return r.candidates_unique, r.n_pairs, h2d, r.n_pairs * 16

    step_host(0)
    cx.barrier()
    w2 = time.time()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(lib_stream)
    e = dict(cands=0, pairs=0, h2d=0, d2h=0)
    for k in range(1, 1 + K):
        c_, p_, a_, b_ = step_host(k)
        e["cands"] += c_; e["pairs"] += p_; e["h2d"] += a_; e["d2h"] += b_
    ev3.record(lib_stream)
    cx.barrier()
    w3 = time.time()
    leg.update(e2e=e, dt_e2e=cx.max_over_ranks(max(ev2.elapsed_time(ev3) * 1e-3, 0.0)), window_e2e=(w2, w3), wall_e2e=w3 - w2)

    # ---- parity sample: never-indexed query vectors, query-only, against the final (sharded) index
    if not args.no_parity and cx.sample_rows is not None:
        r = disp.insert_batch(*(cx.sample_rows if rank == 0 else (None, None, None)), query_only=True)
        leg["sample"] = {"cands": r.candidates_unique, "postings": r.postings_visited, "n_pairs": r.n_pairs,
                         "pairs": None if rank != 0 else {(int(a), int(b)): float(s) for a, b, s in zip(r.q, r.c, r.sim)}}
    leg["stats"] = eng.stats()
    eng.close()
    return leg


def run_allpairs(cx, mode):
    """The north-star job: the config's N vectors from an EMPTY index, one insertNewVector batch after the other (each is
    indexed, then queried: IWA:122-134), sharded over the ranks.  Wall clock and device time of the whole job."""
    from apss_b200 import native
    from apss_b200.dispatcher import ShardDispatcher
    args, cfg, dev, rank, world = cx.args, cx.cfg, cx.dev, cx.rank, cx.world
    B, N, D, t = cfg["batch"], cx.N, cfg["D"], cfg["threshold"]
    eng = native.Index(D, t, device=cx.local_rank, pruning=mode, prune_alpha=args.prune_alpha,
                       reserve_vectors=int(N / world * 1.1) + 2 * B, reserve_nnz=int(cx.total_nnz / world * 1.15) + (1 << 20))
    disp = ShardDispatcher(eng, device=dev)
    lib_stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the job's input batches, cut out of the generated matrix before the clock starts (inputs resident in HBM)
    staged = [cx.dev_rows(lo, min(N, lo + B)) if rank == 0 else (None, None, None) for lo in range(0, N, B)]
    torch.cuda.synchronize()
    cx.barrier()
    w0 = time.time()
    ev0.record(lib_stream)
    tot = dict(pairs=0, cands=0, postings=0, score_ms=0.0, hash=0, batches=0)
    kept = []
    for bi, lo in enumerate(range(0, N, B)):
        r = disp.insert_batch(*staged[bi])
        tot["pairs"] += r.n_pairs; tot["cands"] += r.candidates_unique; tot["postings"] += r.postings_visited
        tot["score_ms"] += r.local.score_ms; tot["batches"] += 1
        if rank == 0:
            kept.append((r.q, r.id_base, r.c, r.sim))      # fetched to the host inside the timed job; hashed after it
    ev1.record(lib_stream)
    cx.barrier()
    w1 = time.time()
    for q_, b_, c_, s_ in kept:
        tot["hash"] = (tot["hash"] + pairset_hash(q_.astype(np.int64) + b_, c_, s_)) & 0xFFFFFFFFFFFFFFFF
    dt = cx.max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    st = eng.stats()
    eng.close()
    return {"mode": "parity (every posting visited)" if mode == 0 else "pruning = %d (exact index reduction)" % mode,
            "vectors": N, "batches": tot["batches"], "seconds": dt, "wall_seconds": w1 - w0, "pairs": tot["pairs"],
            "pairs_per_sec": tot["pairs"] / dt, "candidates": tot["cands"], "candidates_per_sec": tot["cands"] / dt,
            "postings_visited": tot["postings"], "score_kernel_seconds_rank0": tot["score_ms"] * 1e-3,
            "pairset_hash": "%016x" % tot["hash"], "index_postings_rank0": st["n_postings"], "window": (w0, w1)}


def main():
    args = parse_args()
    if os.environ.get("APSS_BENCH_WATCHDOG"):        # debugging aid: dump every thread's stack and exit after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["APSS_BENCH_WATCHDOG"]), exit=True)
    from apss_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict(synth.CONFIGS[args.config]); cfg["name"] = args.config
    if args.n_index:
        cfg["N"] = args.n_index
    if args.batch:
        cfg["batch"] = args.batch
    cfg["workload"] = workload_name(args.config, cfg, cfg["N"])
    if args.impl == "reference":
        return run_reference_arm(args, cfg, rank, world)

    import torch.distributed as dist
    from apss_b200 import native
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W, B, N, D = args.steps, args.warmup, cfg["batch"], cfg["N"], cfg["D"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                            # early: see the wait before the timed region
    cx = Ctx()
    cx.args, cx.cfg, cx.dev, cx.rank, cx.world, cx.local_rank = args, cfg, dev, rank, world, local_rank
    cx.n_fresh = n_fresh = (W + K) + (1 + K)       # value phase + e2e phase (1 warm-up)
    n_cpu_q = 0 if (args.no_cpu_baseline or world > 1) else args.cpu_opt_queries
    n_extra = (0 if args.no_parity else args.parity_queries) + n_cpu_q      # never-indexed rows behind the fresh batches
    t_gen = time.time()
    shard_gen = bool(cfg.get("heavy_tail")) or args.shard_gen
    cx.sample_rows = None
    data = None
    if not shard_gen:
        # every rank generates the whole data set (identical on all ranks and devices: generator g2); owners index their batches
        data = synth.generate(N + n_fresh * B + n_extra, D, cfg["nnz_mean"], s=cfg["s"], seed=cfg["seed"], device=dev)
        torch.cuda.synchronize()
        t_gen = time.time() - t_gen
        x0 = N + n_fresh * B                       # first never-indexed row

        def dev_rows(lo, hi):
            b = data.rows(lo, hi)
            return b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous()

        def fresh_rows(i, pin=False):
            b = data.rows(N + i * B, N + (i + 1) * B)
            if pin:
                b = b.pin()
                return b.indptr, b.indices, b.values
            return b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous()

        def preload(disp, eng):
            for lo in range(0, N, B):
                disp.preload(*dev_rows(lo, min(N, lo + B)))

        cx.total_nnz = data.nnz
        if not args.no_parity:
            cx.sample_rows = dev_rows(x0 + n_cpu_q, x0 + n_cpu_q + args.parity_queries)
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

        def preload(disp, eng):
            for j in range(per_rank):
                b = shard.rows(j * B, (j + 1) * B)
                rows = (b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous())
                torch.cuda.synchronize()
                eng.set_next_id((j * world + rank) * B)
                eng.insert_batch(*rows, index_only=True)
            disp.next_id = N
            disp.batch_no = N // B

        dev_rows = None
        cx.total_nnz = shard.nnz * world
        cfg["workload"] = workload_name(args.config, cfg, N)
    cx.N, cx.fresh_rows, cx.dev_rows, cx.preload = N, fresh_rows, dev_rows, preload

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

    cx.barrier, cx.max_over_ranks = barrier, max_over_ranks
    cx.host_batches = [fresh_rows(W + K + k, pin=True) for k in range(1 + K)] if rank == 0 else []
    cx.out_q = torch.empty(1 << 22, dtype=torch.int32).pin_memory()
    cx.out_c = torch.empty(1 << 22, dtype=torch.int32).pin_memory()
    cx.out_s = torch.empty(1 << 22, dtype=torch.float64).pin_memory()
    if rank == 0:          # nvidia-smi's start-up stalls the driver for ~1 s: it must be over before anything is timed
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 10.0:
            time.sleep(0.05)

    # measured on-chip peaks of this GPU, this run (shared-memory u32 atomics at random addresses; register FP32 FMA)
    onchip = {}
    try:
        onchip["atoms_per_s"] = native.microbench_accumulators(2, warps=16, iters=20000, device=local_rank)
        onchip["ffma_per_s"] = native.microbench_accumulators(6, warps=32, iters=20000, device=local_rank)
    except Exception as ex:      # noqa: BLE001
        print("microbench failed: %r" % (ex,), file=sys.stderr)

    main_mode = args.prune
    main_leg = run_leg(cx, main_mode, profile=args.profile_range and args.profile_leg == "main")
    pr = None
    if not args.prune and not args.no_pruned_leg:
        pr = run_leg(cx, args.pruned_mode, profile=args.profile_range and args.profile_leg == "pruned")
    jobs = []
    if not args.no_allpairs and not shard_gen:
        # the job's data set is G(N, ...) exactly -- independent of --steps / --warmup, so its pair total and hash are
        # comparable across runs, modes and --gpus N (the IDF depends on how many rows are generated)
        job_data = synth.generate(N, D, cfg["nnz_mean"], s=cfg["s"], seed=cfg["seed"], device=dev)

        def job_rows(lo, hi):
            b = job_data.rows(lo, hi)
            return b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous()

        cx.dev_rows, cx.total_nnz = job_rows, job_data.nnz
        for m in ([main_mode] if (args.no_pruned_leg or args.prune) else [0, args.pruned_mode]):
            jobs.append(run_allpairs(cx, m))
        del job_data
    sampler.stop()
    if world > 1:        # every collective is done: rank 0 goes on alone with the CPU-side checks
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    tot, dt_value, dt_e2e, e, st = main_leg["tot"], main_leg["dt_value"], main_leg["dt_e2e"], main_leg["e2e"], main_leg["stats"]
    peak, peak_src = measured_peak()
    windows = [main_leg["window"], main_leg["window_e2e"]] + ([pr["window"], pr["window_e2e"]] if pr else []) + [j["window"] for j in jobs]
    kname = {1: "apss::k_score", 2: "apss::k_score_blk"}.get((args.variant >> 16) & 0xff, "apss::k_score_dense")
    if main_mode:
        kname = {3: "apss::k_score_qm_flat", 2: "apss::k_score_cand", 1: "apss::k_score_dense<pruned>"}[main_mode]
    own_config = world == 1 and args.config == "C3" and not args.n_index and not args.batch
    line = {
        "metric": METRIC, "value": tot["cands"] / dt_value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dt_value / K * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u16/u32 fixed-point accumulate (f32 products) + f64 verify", "data": "synthetic",
        "config": {"workload": cfg["workload"], "index_vectors": N, "batch": B, "generator": synth.GEN_VERSION,
                   "sharding": "id-range block-cyclic x%d" % world,
                   "l2": "inputs larger than L2 (index %.0f MB postings per GPU)" % (st["bytes_postings"] / 1e6),
                   "tile_vectors": st["tile_vectors"], "warps_per_cta": st["warps_per_cta"], "kernel_variant": args.variant,
                   "pruning": ("exact index reduction ON (mode %d): %.1f%% of stored components un-indexed; counters count the reduced work"
                               % (main_mode, 100.0 * st["n_unindexed"] / max(st["n_unindexed"] + st["n_postings"], 1))) if main_mode else "off (parity counters)"},
        "pairs_per_sec": tot["pairs"] / dt_value,
        "postings_per_sec": tot["postings"] / dt_value,
        "pairset_hash": "%016x" % tot["hash"],
        "step_latency_ms": {"min": float(np.min(main_leg["step_wall"])), "p50": float(np.percentile(main_leg["step_wall"], 50)),
                            "max": float(np.max(main_leg["step_wall"]))},
        "gen_s": t_gen, "preload_s": main_leg["preload_s"],
        "e2e": {"value": e["cands"] / dt_e2e if dt_e2e > 0 else None, "unit": UNIT, "h2d_bytes_per_step": e["h2d"] // K,
                "d2h_bytes_per_step": e["d2h"] // K, "pairs_per_sec": e["pairs"] / dt_e2e if dt_e2e > 0 else None,
                "ms_per_step": dt_e2e / K * 1e3},
        "gpu_launches": int(main_leg["launches"]),
    }

    # ---- rooflines: each kernel against the resource that binds it
    score_s = tot["score_ms"] * 1e-3
    if main_mode == 0 and score_s > 0:
        sparse = tot["local_postings"] - tot["dense_post"]
        rl = {"kernel": kname, "launches_timed": K, "avg_launch_ms": tot["score_ms"] / K, "kernel_share_of_step": score_s / dt_value,
              "sparse_updates_per_launch": sparse / K, "dense_updates_per_launch": tot["dense_post"] / K,
              "dense_fma_executed_per_launch": tot["dense_fma"] / K}
        if onchip.get("atoms_per_s") and onchip.get("ffma_per_s"):
            t_sparse, t_dense = sparse / onchip["atoms_per_s"], tot["dense_fma"] / onchip["ffma_per_s"]
            t_min = t_sparse + t_dense
            rl.update({"bound": "on-chip: shared-memory atomics (sparse postings) + FP32 FMA (dense rows); HBM does not bind this kernel",
                       "achieved": (sparse + tot["dense_post"]) / score_s, "unit": "accumulator updates/s",
                       "peak": (sparse + tot["dense_post"]) / t_min if t_min > 0 else None, "frac": t_min / score_s,
                       "peak_source": "measured in this run: apss_microbench_accumulators mode 2 (ATOMS.ADD.U32, random addresses, 16 warps/SM) = %.3e /s, "
                                      "mode 6 (register FFMA) = %.3e /s; speed-of-light time = sparse updates / ATOMS peak + executed dense FMAs / FFMA peak; "
                                      "frac = that time / measured kernel time" % (onchip["atoms_per_s"], onchip["ffma_per_s"]),
                       "sparse": {"peak_updates_per_s": onchip["atoms_per_s"], "sol_ms_per_launch": 1e3 * t_sparse / K},
                       "dense": {"peak_fma_per_s": onchip["ffma_per_s"], "sol_ms_per_launch": 1e3 * t_dense / K,
                                 "useful_share_of_executed_fma": tot["dense_post"] / max(tot["dense_fma"], 1)}})
        tr = committed_traffic(kname) if own_config else None
        rl["hbm"] = {"algorithmic_bytes_per_launch": 8.0 * tot["local_postings"] / K, "algorithmic_GBps": 8.0 * tot["local_postings"] / score_s / 1e9,
                     "peak_GBps": peak, "peak_source": peak_src,
                     "note": "8 B per posting visited per QUERY TERM (SURVEY 8d); the lists are re-used on chip across the 16 queries of a block "
                             "and across blocks from L2, so HBM is NOT the bound of this kernel and this figure may exceed the copy peak",
                     "traffic_source": None if not tr else tr.get("ncu")}
        rl["traffic"] = None if not tr else tr.get("dram_bytes_per_launch")
        line["roofline"] = rl
    elif score_s > 0:
        line["roofline"] = {"bound": "hbm", "kernel": kname, "achieved": 8.0 * tot["local_postings"] / score_s / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": 8.0 * tot["local_postings"] / score_s / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                            "launches_timed": K, "avg_launch_ms": tot["score_ms"] / K, "kernel_share_of_step": score_s / dt_value}

    if pr is not None:
        p_tot, dt_pr, st2 = pr["tot"], pr["dt_value"], pr["stats"]
        ps = p_tot["score_ms"] * 1e-3
        tr = committed_traffic("apss::k_score_qm_flat") if own_config else None
        line["pruned"] = {
            "kernel": "apss::k_score_qm_flat (query-major posting-list traversal of the reduced index; bulk-async producer + 31 consumer warps)"
                      if args.pruned_mode == 3 else "apss::k_score_cand (candidate-major, reduced index)",
            "ms_per_step": dt_pr / K * 1e3, "pairs_per_sec": p_tot["pairs"] / dt_pr,
            "pairset_hash": "%016x" % p_tot["hash"],
            "pair_set_identical_to_parity_run": p_tot["hash"] == tot["hash"] and p_tot["pairs"] == tot["pairs"],
            "speedup_vs_parity_run": dt_value / dt_pr, "equivalent_candidates_per_sec": tot["cands"] / dt_pr,
            "postings_visited_per_step": p_tot["postings"] / K, "candidates_touched_per_step": p_tot["cands"] / K,
            "verify_records_per_step": p_tot["prefilter"] / K, "score_kernel_ms_per_step": p_tot["score_ms"] / K,
            "unindexed_fraction": st2["n_unindexed"] / max(st2["n_unindexed"] + st2["n_postings"], 1),
            "posting_segments": st2["n_tiles"], "segment_merges": st2.get("segment_merges"),
            "work_reduction_postings": tot["postings"] / max(p_tot["postings"], 1),
            "gpu_launches": int(pr["launches"]),
            "e2e": {"ms_per_step": pr["dt_e2e"] / K * 1e3, "pairs_per_sec": pr["e2e"]["pairs"] / pr["dt_e2e"] if pr["dt_e2e"] > 0 else None,
                    "h2d_bytes_per_step": pr["e2e"]["h2d"] // K, "d2h_bytes_per_step": pr["e2e"]["d2h"] // K,
                    "pairs_identical_to_parity_e2e": pr["e2e"]["pairs"] == e["pairs"]},
            "roofline": None if ps <= 0 else {
                "bound": "hbm", "achieved": 8.0 * p_tot["local_postings"] / ps / 1e9, "peak": peak, "unit": "GB/s",
                "frac": 8.0 * p_tot["local_postings"] / ps / 1e9 / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": 8.0 * p_tot["local_postings"] / K, "avg_launch_ms": p_tot["score_ms"] / K,
                "kernel_share_of_step": ps / dt_pr,
                "traffic": None if not tr else tr.get("dram_bytes_per_launch"), "traffic_source": None if not tr else tr.get("ncu"),
                "note": "8 B per posting visited (SURVEY 8d with pruning on: postings actually visited); every visit is one physical 8-byte read "
                        "(cp.async.bulk into shared memory), nothing is re-used on chip across queries: this is the north-star 'HBM roofline of "
                        "bytes of postings touched'; the timed interval also holds the piece-cutting kernels and the ranged kernel for deferred queries"},
            "note": "exact index reduction (include/apss.h `pruning`): same batches as the value phase, same pair set; "
                    "equivalent_candidates_per_sec = candidates of the parity run / time of this run"}
        line["pair_set_identical_across_modes"] = line["pruned"]["pair_set_identical_to_parity_run"]
    if jobs:
        for j in jobs:
            j.pop("window", None)
        line["allpairs"] = {"jobs": jobs, "pair_totals_equal": len({j["pairs"] for j in jobs}) == 1,
                            "pairset_hashes_equal": len({j["pairset_hash"] for j in jobs}) == 1,
                            "note": "north-star job: all N vectors from an empty index, every batch indexed then queried (IWA:122-134)"}
    line["clocks"] = sampler.summary(windows)

    # ---- CPU side on rank 0: sampled oracle parity, then the CPU baselines (N = 1 only)
    need_oracle = (not args.no_parity and cx.sample_rows is not None) or n_cpu_q
    if need_oracle and data is not None:
        threads = host_threads()
        data_np = data.numpy()
        del data
        cx.sample_rows = None
        torch.cuda.empty_cache()
        n_index = N + n_fresh * B
        ofast, t_build = build_oracle(data_np, cfg, n_index, threads, "fast")
        if not args.no_parity:
            s_lo = n_index + n_cpu_q
            ro = ofast.insert_batch(*rows_np(data_np, s_lo, s_lo + args.parity_queries), query_only=True)
            want = {(int(a), int(b)): float(s) for a, b, s in zip(ro.q, ro.c, ro.sim)}
            legs = {}
            ok = True
            named = [("parity_mode" if not main_mode else "pruning_%d" % main_mode, main_leg)] + ([("pruning_%d" % args.pruned_mode, pr)] if pr else [])
            for name, lg in named:
                got = lg["sample"]["pairs"]
                same_set = set(got) == set(want)
                same_sim = same_set and all(got[k] == want[k] for k in want)
                rec = {"pair_set_equal": same_set, "similarities_bit_exact": same_sim, "pairs": len(got)}
                if lg["mode"] == 0:
                    rec["candidates_unique_equal"] = lg["sample"]["cands"] == ro.candidates_unique
                    rec["postings_visited_equal"] = lg["sample"]["postings"] == ro.postings_visited
                    rec["candidates_unique"] = lg["sample"]["cands"]
                ok = ok and all(v for k, v in rec.items() if k.endswith("equal") or k.endswith("exact"))
                legs[name] = rec
            line["parity"] = {"ok": bool(ok), "queries": args.parity_queries, "index_vectors": n_index, "n_gpus": world,
                              "oracle": "oracle/apss_oracle.c ALGO_FAST (fp64, ascending-dimension order), %d threads, index built in %.1f s" % (threads, t_build),
                              "oracle_pairs": len(want), "legs": legs,
                              "what": "never-indexed query vectors scored query-only against the full sharded index after the timed phases"}
        if n_cpu_q:
            t0 = time.perf_counter()
            r_opt = ofast.insert_batch(*rows_np(data_np, n_index, n_index + n_cpu_q), query_only=True)
            dt_o = time.perf_counter() - t0
            nq = min(n_cpu_q, args.cpu_queries or max(4, threads // 4))
            q_ref = rows_np(data_np, n_index, n_index + nq)
            r_cnt = ofast.insert_batch(*q_ref, query_only=True)         # untimed: candidates and pairs of the reference sample
            ofast.close()
            oref, t_build_ref = build_oracle(data_np, cfg, n_index, threads, "faithful")
            t0 = time.perf_counter()
            r_ref = oref.insert_batch(*q_ref, query_only=True)
            dt_f = time.perf_counter() - t0
            oref.close()
            assert r_ref.pair_set() == r_cnt.pair_set()
            line["cpu_baseline"] = {
                "value": r_cnt.candidates_unique / dt_f, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "cpu_ref: %d query vectors scored query-only against the same %d indexed vectors of %s; %s" % (nq, n_index, args.config, REF_ALGO),
                "seconds": dt_f, "index_build_s": t_build_ref, "dot_calls": r_ref.dot_calls_ref, "dot_calls_per_sec": r_ref.dot_calls_ref / dt_f,
                "opt_value": r_opt.candidates_unique / dt_o, "opt_seconds": dt_o, "opt_queries": n_cpu_q,
                "opt_note": "cpu_opt: %d query vectors against the same index, weighted postings + dense fp64 accumulator, OpenMP over queries "
                            "(the GPU's parity-mode algorithm on the CPU)" % n_cpu_q}
        else:
            ofast.close()
    emit(line)


if __name__ == "__main__":
    main()
