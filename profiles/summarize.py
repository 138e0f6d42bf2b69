#!/usr/bin/env python
"""Condenses gpurun_out/ artefacts into small tracked files under profiles/.

    python profiles/summarize.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof.ncu-rep]
                                       [--bench gpurun_out/bench.json]

  <tag>_launches.csv : per-kernel-name launch count / total / mean / share of GPU time, from the
                       `ncu --metrics gpu__time_duration.sum --clock-control none` launch list
                       (cold-cache, serialised: compare SHARES, not absolutes)
  <tag>_ncu_full.csv : selected `ncu --set full` metrics per captured launch (dram bytes, pipe use ...)
  <tag>_bench.json   : the bench.py line of the same build
"""
import argparse
import collections
import csv
import json
import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"at::native::|at::<unnamed>::|b200ssl::", "", name)
    return name[:90]


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, m, v, u = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[m] != "gpu__time_duration.sum":
            continue
        t = float(r[v].replace(",", ""))
        t_us = {"ns": t / 1e3, "us": t, "ms": t * 1e3, "s": t * 1e6}.get(r[u], t / 1e3)
        a = agg.setdefault(short(r[k]), [0, 0.0])
        a[0] += 1
        a[1] += t_us
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "mean_us", "share_of_gpu_time"])
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([name, n, round(t, 2), round(t / n, 2), round(t / total, 4)])
    return out


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [c for c in FULL_METRICS if c in hdr]
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{c} [{units[hdr.index(c)]}]" for c in cols])
        for r in rows[2:]:
            w.writerow([short(r[hdr.index("Kernel Name")])] + [r[hdr.index(c)] for c in cols])
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--bench")
    a = ap.parse_args()
    if a.launches:
        print(launches(a.launches, os.path.join(HERE, a.tag + "_launches.csv")))
    if a.rep:
        print(full(a.rep, os.path.join(HERE, a.tag + "_ncu_full.csv")))
    if a.bench:
        line = [l for l in open(a.bench) if l.startswith("{")][-1]
        json.dump(json.loads(line), open(os.path.join(HERE, a.tag + "_bench.json"), "w"), indent=1)
        print(os.path.join(HERE, a.tag + "_bench.json"))
