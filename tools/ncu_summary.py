#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU: `ncu -i`) into the JSON kept under profiles/.
  python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.json ["note"]"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def table(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def unit_bytes(v, u):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u)
    return float(v) * f if f else float(v)


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    rows = table(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    out = {"report": rep.split("/")[-1], "kernel": d.get("Kernel Name"), "note": sys.argv[3] if len(sys.argv) > 3 else "",
           "metrics": {k: {"value": d[k], "unit": u[k]} for k in KEYS if k in d}}
    if "dram__bytes_read.sum" in d:
        out["dram_bytes_per_launch"] = unit_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + unit_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
    out["stalls_per_issue_active"] = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(d[k])
                                      for k in hdr if "average_warps_issue_stalled" in k and "not_issued" not in k and d[k] not in ("", "n/a") and float(d[k]) > 0.05}
    src = table(rep, "source")
    if len(src) > 2:
        h = src[1]; ci = {x: i for i, x in enumerate(h)}
        data = [r for r in src[2:] if len(r) == len(h)]
        tot = sum(int(r[ci["# Samples"]]) for r in data) or 1
        out["warp_instructions_executed"] = sum(int(r[ci["Instructions Executed"]]) for r in data)
        top = sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:12]
        out["top_sass_by_stall_samples"] = [
            {"sass": r[ci["Source"]].strip(), "share_of_samples": round(int(r[ci["# Samples"]]) / tot, 4), "executed": int(r[ci["Instructions Executed"]]),
             "stalls": dict(sorted(((k, int(r[ci[k]])) for k in h if k.startswith("stall_") and "(" not in k and int(r[ci[k]]) > 0), key=lambda x: -x[1])[:2])}
            for r in top]
    json.dump(out, open(dst, "w"), indent=1)
    print(dst, out["metrics"].get("gpu__time_duration.sum"), out.get("dram_bytes_per_launch"))


if __name__ == "__main__":
    main()
