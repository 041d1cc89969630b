#!/usr/bin/env python3
"""Summarise ncu outputs into small text files under profiles/ (the .ncu-rep itself stays in gpurun_out/).
  tools/ncu_summary.py launches <launches.csv> <out.txt>      per-kernel totals / shares from the --metrics gpu__time_duration pass
  tools/ncu_summary.py full <report.ncu-rep> <out.txt>        key counters of every kernel in a --set full capture
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, rows = r, rows[i + 1:]
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows:
        try:
            d.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES, not absolutes)\n")
        f.write("%-28s %5s %12s %12s %7s\n" % ("kernel", "n", "total_ms", "avg_us", "share"))
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-28s %5d %12.3f %12.3f %6.1f%%\n" % (k[:28], len(v), sum(v) / 1e6, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    print(open(out).read())


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none, one launch per kernel; source: %s\n" % rep)
        for r in rows[2:]:
            f.write("\n== %s\n" % r[hdr.index("Kernel Name")])
            for k in KEYS:
                if k in hdr:
                    f.write("  %-90s %s %s\n" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
