#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python profiles/summarise.py launches gpurun_out/launches_r1.csv   > profiles/r1_launches.txt
    python profiles/summarise.py kernels  gpurun_out/prof_r1.ncu-rep   > profiles/r1_kernels.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "sm__cycles_elapsed.max"]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg.setdefault(r[ki].split("(")[0], []).append(v)
    ours = {k: v for k, v in agg.items() if k.startswith(("<unnamed>::", "void <unnamed>::"))}   # torch's have at:: in front
    per_eval = sum(sum(v) / len(v) for v in ours.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): compare SHARES")
    print("%-58s %6s %12s %8s" % ("kernel", "n", "mean us", "share"))
    for k, v in agg.items():
        mean = sum(v) / len(v)
        share = ("%7.2f%%" % (100 * mean / per_eval)) if k in ours else "   (torch)"
        print("%-58s %6d %12.1f %s" % (k[-58:], len(v), mean, share))
    print("# sum of our kernels per evaluation: %.1f us" % per_eval)


def kernels(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== " + r[hdr.index("Kernel Name")].split("(")[0])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("   %-82s %s %s" % (m, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2])
