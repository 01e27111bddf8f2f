#!/usr/bin/env python
"""The reference's latency experiment (benchmark/LoadGenerator.scala:15-173) against the GPU worker, on the WALL clock:
`childrenNum` runners tick every `writeBatchingDuration` ms, warm-up pass, the parent's ReceiveTimeout starts the test
phase, response time = SimilarityOutput.outputMoment - StartTime (apss_b200.loadgen mirrors the protocol as built).
Prints what LoadGenerator.postStop prints plus a JSON line.

  python tools/loadgen_bench.py [--config C2] [--videos 2000] [--children 4] [--messages 200] [--tick-ms 1]
                                [--exp-ms 300] [--output-io-ms 0] [--prune 0|3]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import apss_b200
from apss_b200 import loadgen, synth
from apss_b200 import messages as M
from apss_b200.worker import GpuIndexingWorkerActor

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--videos", type=int, default=2000)
ap.add_argument("--children", type=int, default=4)
ap.add_argument("--messages", type=int, default=200)
ap.add_argument("--tick-ms", type=int, default=1)
ap.add_argument("--exp-ms", type=int, default=300)
ap.add_argument("--output-io-ms", type=int, default=0)
ap.add_argument("--prune", type=int, default=0)
ap.add_argument("--out", default="")
args = ap.parse_args()
cfg = synth.CONFIGS[args.config]
D, t = cfg["D"], cfg["threshold"]
ip, ix, v = synth.generate(args.videos, D, cfg["nnz_mean"], seed=cfg["seed"], device="cuda" if torch.cuda.is_available() else "cpu").numpy()
videos = [("v%d" % i, M.SparkSparseVector(D, ix[ip[i]:ip[i + 1]], v[ip[i]:ip[i + 1]])) for i in range(args.videos)]
conf = {"cpslab.allpair.similarityThreshold": t, "cpslab.allpair.outputIODuration": args.output_io_ms, "cpslab.allpair.vectorDim": D,
        "cpslab.allpair.indexThreshold": 0.0, "cpslab.allpair.benchmark.expDuration": args.exp_ms,
        "cpslab.allpair.benchmark.writeBatchingDuration": args.tick_ms, "cpslab.allpair.benchmark.totalMessageCount": args.messages,
        "cpslab.allpair.benchmark.childrenNum": args.children, "cpslab.allpair.gpu.pruning": args.prune}
worker = GpuIndexingWorkerActor(conf)
rep = loadgen.run_experiment(conf, videos, worker, loop=loadgen.EventLoop(virtual=False))
print(rep["line"])
rep.update(config=args.config, videos=args.videos, children=args.children, messages_per_child=args.messages, tick_ms=args.tick_ms,
           output_io_ms=args.output_io_ms, pruning=args.prune)
print(json.dumps(rep))
if args.out:
    with open(args.out, "w") as f:
        json.dump(rep, f)
