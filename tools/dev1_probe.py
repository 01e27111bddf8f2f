#!/usr/bin/env python
"""Debug aid: the reduced-index engine on a device other than 0, in a process that sees several GPUs."""
import faulthandler
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(int(os.environ.get("WATCHDOG", "60")), exit=True)
import numpy as np
import torch

from apss_b200 import native, synth

dev = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 3
data = synth.generate(6000, 1 << 12, 30, seed=11).numpy()
ip, ix, v = data
print("devices visible:", torch.cuda.device_count(), "using", dev, "mode", mode, flush=True)
g = native.Index(1 << 12, 0.6, pruning=mode, device=dev)
tot = 0
for lo in range(0, 6000, 1000):
    hi = lo + 1000
    r = g.insert_batch(ip[lo:hi + 1] - ip[lo], ix[ip[lo]:ip[hi]], v[ip[lo]:ip[hi]])
    tot += r.n_pairs
    print("batch", lo, "pairs", r.n_pairs, "postings", r.postings_visited, flush=True)
print("ok", tot, g.stats()["n_tiles"], flush=True)
