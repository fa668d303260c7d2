#!/usr/bin/env python
"""Turn an `ncu --set full` report into the short per-kernel summary committed under profiles/.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
SCALE = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel:", r[idx["Kernel Name"]].split("(")[0], " grid", r[idx["Grid Size"]], " block", r[idx["Block Size"]])
        for k in KEYS:
            if k in idx:
                print(f"  {k:84s} {r[idx[k]]:>16s} {units[idx[k]]}")
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * SCALE[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * SCALE[units[idx["dram__bytes_write.sum"]]]
        print(f"  {'traffic = dram read + write per launch':84s} {(rd + wr) / 1e6:16.1f} MB")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
