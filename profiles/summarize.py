#!/usr/bin/env python
"""Summarise ncu CSV exports kept under profiles/.

  python profiles/summarize.py launches profiles/r01_launches_bench_c5.csv
  python profiles/summarize.py raw      gpurun_out/prof.raw.csv

`launches` aggregates the `--metrics gpu__time_duration.sum` launch list per kernel (count,
mean, share of the summed GPU time); `raw` prints the roofline-relevant metrics of every
kernel in a `--set full --page raw --csv` export.
"""
import collections
import csv
import sys

KEYS = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__waves_per_multiprocessor",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
)


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    for n, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, rows = r, rows[n + 1:]
            break
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows:
        if r[mi] == "gpu__time_duration.sum":
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"| kernel | launches | mean us | total ms | share |\n|---|---|---|---|---|")
    for k, v in agg.items():
        print(f"| `{k[:90]}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / 1e6:.2f} | "
              f"{100 * sum(v) / tot:.1f}% |")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"### `{d['Kernel Name'][:90]}` grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                print(f"- {k}: {d[k]} {units[hdr.index(k)]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
