#!/usr/bin/env python
"""ncu_export.py REPORT.ncu-rep OUT_PREFIX -- selected raw metrics per launch of an ncu report as a small CSV
(OUT_PREFIX_summary.csv), so that the evidence travels back from the GPU box (gpurun_out/ is capped at 64 MiB)."""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "lts__throughput.avg.pct",
        "l1tex__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__lsu_writeback_active.avg.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__issue_active.avg.pct", "sm__warps_active.avg.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor",
        "smsp__average_warps_issue_stalled_long_scoreboard_per", "smsp__average_warps_issue_stalled_short_scoreboard_per",
        "smsp__average_warps_issue_stalled_wait_per", "smsp__average_warps_issue_stalled_barrier_per", "smsp__average_warps_issue_stalled_math_pipe",
        "smsp__average_warps_issue_stalled_lg_throttle", "smsp__average_warps_issue_stalled_mio_throttle", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__throughput.avg.pct", "sm__pipe_fma_cycles_active.avg.pct", "sm__inst_executed_pipe_xu.avg.pct",
        "launch__grid_size", "launch__block_size")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        print("ncu_export: no rows in", rep)
        return 1
    hdr = rows[0]
    pick = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name") or any(s in h for s in KEEP)]
    with open(out + "_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] if i < len(r) else "" for i in pick])
    print("ncu_export: wrote", out + "_summary.csv", len(rows) - 2, "launches")
    return 0


if __name__ == "__main__":
    sys.exit(main())
