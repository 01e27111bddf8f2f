"""Latency benchmark in the style of the reference's LoadGenerator (benchmark/LoadGenerator.scala:15-173):
warm-up pass (every vector sent once through ClientConnection.insertNewVector, one vector per message,
LoadGenerator.scala:59-73), ReceiveTimeout => the index freezes (IndexingWorkerActor.scala:143-144), then the
test phase replays `totalMessageCount` vectors and records the response time of each
(SimilarityOutput.outputMoment - StartTime, LoadGenerator.scala:135-157); prints what postStop prints
(:112-131): message count, average / max / min response time.  Here time is measured around the call
(the in-process transport is synchronous) with perf_counter, in milliseconds.

  python tools/latency_bench.py [--config C2] [--n-index 100000] [--messages 500] [--bulk] [--ccweb FILE]
--ccweb reads the reference's CC_WEB_VIDEO text format, "(id,size,[dense values])" per line
(CCWEBVideoLoadGenerator.scala:10-29), normalises like LoadRunner.generateVector (LoadGenerator.scala:30-41)
and uses those vectors instead of the synthetic ones (the dataset itself is not shipped with the reference).
--bulk pre-loads the index with large batches instead of one message per vector (the per-message
warm-up of the reference is itself reported when it is used)."""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import apss_b200
from apss_b200 import messages as M
from apss_b200 import synth
from apss_b200.worker import ClientConnection, GpuIndexingWorkerActor, LocalActorSystem, RegionRouter

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--n-index", type=int, default=0)
ap.add_argument("--messages", type=int, default=500)
ap.add_argument("--bulk", action="store_true")
ap.add_argument("--out", default="")
ap.add_argument("--ccweb", default="")
ap.add_argument("--prune", type=int, default=0, help="cpslab.allpair.gpu.pruning: 0 off, 2 exact index reduction")
ap.add_argument("--threshold", type=float, default=0.0)
args = ap.parse_args()
cfg = synth.CONFIGS[args.config]
if args.ccweb:
    from apss_b200 import etl
    videos = etl.ccweb_generate_vectors(args.ccweb)
    N = min(args.n_index, len(videos)) if args.n_index else len(videos)
    D = max(v[1] for v in videos)                                 # cpslab.allpair.vectorDim
    t = args.threshold or cfg["threshold"]

    def vec(i):
        _, dim, idx, val = etl.load_runner_vector(videos, i, D)
        return M.SparkSparseVector(dim, idx, val)
else:
    N = args.n_index or cfg["N"]
    D, t = cfg["D"], args.threshold or cfg["threshold"]
    data = synth.generate(N, D, cfg["nnz_mean"], seed=cfg["seed"], device="cuda" if torch.cuda.is_available() else "cpu").numpy()
    ip, ix, v = data
    vec = lambda i: M.SparkSparseVector(D, ix[ip[i]:ip[i + 1]], v[ip[i]:ip[i + 1]])

conf = {"cpslab.allpair.similarityThreshold": t, "cpslab.allpair.outputIODuration": 0, "cpslab.allpair.benchmark.expDuration": 30000,
        "cpslab.allpair.vectorDim": D, "cpslab.allpair.indexThreshold": 0.0, "cpslab.allpair.ioTriggerPeriod": 0,
        "cpslab.allpair.gpu.pruning": args.prune}
outputs = []
worker = GpuIndexingWorkerActor(conf, replyTo=outputs.append)
system = LocalActorSystem()
system.register("127.0.0.1:2551", RegionRouter(conf, worker))
client = ClientConnection(["127.0.0.1:2551"], system)

# ---- warm-up phase: index everything (LoadGenerator.scala:62-65)
t0 = time.perf_counter()
warm_lat = []
if args.bulk:
    B = 4096
    for lo in range(0, N, B):
        client.insertNewVector([(str(i), vec(i)) for i in range(lo, min(N, lo + B))])
else:
    for i in range(N):
        s = time.perf_counter()
        client.insertNewVector({(str(i), vec(i))})
        warm_lat.append((time.perf_counter() - s) * 1e3)
warm_s = time.perf_counter() - t0
outputs.clear()
# ---- idle for expDuration => ReceiveTimeout (LoadGenerator.scala:161-168, IWA:143-144)
worker.receive(M.ReceiveTimeout())
# ---- test phase (LoadGenerator.scala:66-67, 75-82, 135-157)
lat = []
neighbours = 0
for k in range(args.messages):
    i = k % N
    start = time.perf_counter()                                   # StartTime(vectorId, moment)
    client.insertNewVector({("t%d" % k, vec(i))})
    out = outputs[-1]                                             # SimilarityOutput for this query
    lat.append((time.perf_counter() - start) * 1e3)
    neighbours += sum(len(m) for m in out.output.values())
lat = np.array(lat)
res = {"config": args.ccweb or args.config, "pruning": args.prune, "index_vectors": N, "messages": args.messages, "warmup_s": warm_s,
       "warmup_mode": "bulk batches of 4096" if args.bulk else "one vector per message",
       "warmup_ms_per_message": (float(np.mean(warm_lat)) if warm_lat else None),
       "avg_ms": float(lat.mean()), "max_ms": float(lat.max()), "min_ms": float(lat.min()),
       "p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "neighbours_reported": neighbours}
print("total message number: %d\naverage response time: %.3f ms\nmax response time: %.3f ms\nmin response time: %.3f ms" % (
    args.messages, res["avg_ms"], res["max_ms"], res["min_ms"]))
print(json.dumps(res))
if args.out:
    json.dump(res, open(args.out, "w"), indent=1)
