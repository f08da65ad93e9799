#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into profiles/ (tracked).

    python tools/ncu_summary.py <tag>            # e.g. r01a

Reads gpurun_out/launches.csv (the `--metrics gpu__time_duration.sum` launch list
of `bench.py`) and gpurun_out/sweep_full.ncu-rep (one `--set full` capture) and
writes
    profiles/<tag>_launches.csv         per-kernel totals and shares of the step
    profiles/<tag>_full_summary.csv     selected raw metrics per captured launch
    profiles/sweep_dram_bytes.json      dram bytes per sweep launch (bench.py's
                                        roofline.traffic reads this file)
Runs here (no GPU): `ncu -i` only parses the report.
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
    "derived__lts__lts2xbar_bytes.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(tag, src):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1.0)
        a = agg.setdefault(d["Kernel Name"], [0, 0.0, d["Grid Size"], d["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    dst = os.path.join(PROF, tag + "_launches.csv")
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_captured_time", "last_grid", "block"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, a[0], "%.1f" % a[1], "%.1f" % (a[1] / a[0]), "%.4f" % (a[1] / tot), a[2], a[3]])
    print("wrote", dst)


def full(tag, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in KEEP if c in idx]
    dst = os.path.join(PROF, tag + "_full_summary.csv")
    sweep_bytes = []
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(rows) - 2)])
        for c in cols:
            w.writerow([c, units[idx[c]]] + [r[idx[c]] for r in rows[2:]])
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for r in rows[2:]:
        if "sweep" in r[idx["Kernel Name"]] or "pass_kernel" in r[idx["Kernel Name"]]:
            b = 0.0
            for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                b += float(r[idx[c]]) * scale[units[idx[c]]]
            sweep_bytes.append(b)
    print("wrote", dst)
    if sweep_bytes:
        j = {"dram_bytes_per_launch": sum(sweep_bytes) / len(sweep_bytes), "launches": sweep_bytes,
             "source": "profiles/%s_full_summary.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % tag}
        json.dump(j, open(os.path.join(PROF, "sweep_dram_bytes.json"), "w"), indent=1)
        print("wrote profiles/sweep_dram_bytes.json", j["dram_bytes_per_launch"])


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    if os.path.exists(os.path.join(OUT, "launches.csv")):
        launches(tag, os.path.join(OUT, "launches.csv"))
    rep = sys.argv[2] if len(sys.argv) > 2 else os.path.join(OUT, "sweep_full.ncu-rep")
    if os.path.exists(rep):
        full(tag, rep)
