#!/usr/bin/env python
"""Variant sweep of the reduced-index scoring kernels on one GPU: generate the workload once, then for every variant
(a comma-separated list of ENV=value settings read by apss_create, plus `mode=N` for the pruning mode) build the index,
run warm-up + timed batches and print the per-step kernel / device / wall times.

  python tools/qm_probe.py [--config C3] [--n-index N] [--steps K] -- "mode=3" "mode=3,APSS_QM_NT=512" "mode=2"
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from apss_b200 import native, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--n-index", type=int, default=0)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--profile-range", action="store_true", help="cudaProfilerStart/Stop around the timed batches (ncu --profile-from-start off)")
    ap.add_argument("variants", nargs="*", default=["mode=3"])
    args = ap.parse_args()
    cfg = dict(synth.CONFIGS[args.config])
    N = args.n_index or cfg["N"]
    B, D, t = cfg["batch"], cfg["D"], cfg["threshold"]
    nb = args.steps + args.warmup
    data = synth.generate(N + nb * B, D, cfg["nnz_mean"], s=cfg["s"], seed=cfg["seed"], device="cuda")
    torch.cuda.synchronize()

    def rows(lo, hi):
        b = data.rows(lo, hi)
        return b.indptr.contiguous(), b.indices.contiguous(), b.values.contiguous()

    out = []
    for var in args.variants:
        mode = 3
        env = {}
        for kv in var.split(","):
            k, v = kv.split("=")
            if k == "mode":
                mode = int(v)
            else:
                env[k] = v
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        eng = native.Index(D, t, pruning=mode, reserve_vectors=N + nb * B + B, reserve_nnz=int(data.nnz * 1.05))
        t0 = time.time()
        for lo in range(0, N, B):
            eng.insert_batch(*rows(lo, min(N, lo + B)), index_only=True)
        torch.cuda.synchronize()
        t_load = time.time() - t0
        sc, dv, wl, pv, cu, pr, pf = [], [], [], 0, 0, 0, 0
        for i in range(nb):
            r_in = rows(N + i * B, N + (i + 1) * B)
            torch.cuda.synchronize()
            if args.profile_range and i == args.warmup:
                torch.cuda.profiler.start()
            w0 = time.time()
            r = eng.insert_batch(*r_in)
            w1 = time.time()
            if i >= args.warmup:
                sc.append(r.score_ms); dv.append(r.device_ms); wl.append((w1 - w0) * 1e3)
                pv += r.postings_visited; cu += r.candidates_unique; pr += r.n_pairs; pf += r.n_prefilter
        if args.profile_range:
            torch.cuda.profiler.stop()
        st = eng.stats()
        rec = {"variant": var, "score_ms": sum(sc) / len(sc), "device_ms": sum(dv) / len(dv), "wall_ms": sum(wl) / len(wl),
               "postings_per_step": pv / len(sc), "cands_per_step": cu / len(sc), "pairs": pr, "prefilter_per_step": pf / len(sc), "preload_s": t_load,
               "segments_or_tiles": st["n_tiles"], "merges": st.get("segment_merges"), "alg_GBps": 8e-9 * pv / (sum(sc) * 1e-3) if sum(sc) else None}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        eng.close()
        for k, v in old.items():          # the variant's environment stays in force for its whole run (some knobs are read per call)
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return out


if __name__ == "__main__":
    main()
