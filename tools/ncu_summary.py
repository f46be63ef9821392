#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV kept under profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1/prof_summary.csv
"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sm__cycles_elapsed.avg", "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = []
    for w in WANT:
        hit = [i for i, h in enumerate(hdr) if h == w or h.endswith("." + w)]
        if hit:
            cols.append((w, hit[0]))
    with open(out, "w", newline="") as fh:
        wr = csv.writer(fh)
        wr.writerow([f"{w} [{units[i]}]" if units[i] else w for w, i in cols])
        for r in rows[2:]:
            wr.writerow([r[i] for _, i in cols])
    print(f"wrote {out}: {len(rows) - 2} kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
