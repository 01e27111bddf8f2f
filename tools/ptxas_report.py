#!/usr/bin/env python
"""Registers / spills / shared memory per kernel of csrc/apss_api.cu (ptxas -v), one line per entry point.
Usage: python tools/ptxas_report.py [regex]   (runs here, no GPU needed)"""
import re, subprocess, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(root, "all-pairs-similarity_b200", "csrc")
cmd = ["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
       "-I" + os.path.join(root, "include"), "-I" + src, "-Xptxas", "-v", "-cubin", "-o", "/tmp/apss_b200.cubin",
       os.path.join(src, "apss_api.cu")] + sys.argv[2:]
out = subprocess.run(cmd, capture_output=True, text=True).stderr
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else "k_score")
name = None
rows = []
for ln in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip(); cur = {"name": name}; rows.append(cur); continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
    if m and rows: rows[-1].update(stack=int(m.group(1)), st=int(m.group(2)), ld=int(m.group(3)))
    m = re.search(r"Used (\d+) registers", ln)
    if m and rows: rows[-1]["regs"] = int(m.group(1))
for r in rows:
    if pat.search(r["name"]):
        print("%-60s regs %3d  stack %3d  spill st/ld %3d/%3d" % (r["name"][:60], r.get("regs", -1), r.get("stack", 0), r.get("st", 0), r.get("ld", 0)))
