#!/usr/bin/env python
"""Summarise ncu outputs into the text files kept under profiles/.

    python tools/ncu_summary.py launches <launches.csv>          per-kernel totals and shares of a --metrics gpu__time_duration.sum run
    python tools/ncu_summary.py kernel <report.ncu-rep> [name]    key counters + hot SASS regions of a --set full capture
                                                                  (first kernel of the report, or the first whose name matches)
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ix = {h: k for k, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or not r[ix["Metric Value"]]:
            continue
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[ix["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r[ix["Kernel Name"]].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.3f} ms of device time (ncu-serialised, cold cache: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:100]:100s} n={v[0]:4d} total={v[1]:10.3f} ms  avg={v[1] / v[0]:9.4f} ms  share={100 * v[1] / tot:5.1f}%")


def kernel(path, name=None):
    sel = ["-k", "regex:" + name] if name else []
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(f"# {path}")
    print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:90s} {vals[i]:>18s} {units[i]}")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                v = float(vals[i])
            except ValueError:
                continue
            if v > 0.2:
                print(f"stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.3f} warps per issue")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = []
    for r in rows[2:]:  # first kernel of the report only
        if len(r) <= max(iE, iS) or not r[iE].isdigit():
            break
        data.append(r)
    tot = sum(int(r[iE]) for r in data)
    tots = max(sum(int(r[iS]) for r in data), 1)
    runs, cur = [], None
    for k, r in enumerate(data):
        e, s = int(r[iE]), int(r[iS])
        if cur and abs(e - cur["e"]) <= 0.02 * max(e, cur["e"], 1):
            cur["n"] += 1; cur["inst"] += e; cur["samp"] += s; cur["end"] = k
        else:
            cur = {"start": k, "end": k, "e": e, "n": 1, "inst": e, "samp": s}
            runs.append(cur)
    print(f"SASS regions with >= 1% of the {tot} executed warp instructions ({tots} stall samples):")
    for r in runs:
        if r["inst"] > 0.01 * tot or r["samp"] > 0.01 * tots:
            print(f"  sass {r['start']:5d}-{r['end']:5d}  {r['n']:4d} instr x {r['e']:>12d} executions  inst {100 * r['inst'] / tot:5.1f}%  samples {100 * r['samp'] / tots:5.1f}%")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](*sys.argv[2:])
