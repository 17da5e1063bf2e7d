#!/usr/bin/env python
"""Turns gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python profiles/summarize_ncu.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof.ncu-rep]
"""
import argparse
import collections
import csv
import io
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
       "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif row["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        agg.setdefault(row["Kernel Name"].split("(")[0][-60:], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"{'kernel':62s} {'n':>5s} {'mean_us':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:62s} {len(v):5d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / tot:7.1%}\n")
    print(open(out).read())


def raw(rep, out):
    """rep: an .ncu-rep, or the CSV of its raw page already exported on the GPU box (reports of a full-set capture are
    tens of MB and do not travel back)."""
    if str(rep).endswith(".csv"):
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    cols = [h.index("Kernel Name")] + [h.index(m) for m in RAW if m in h]
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow([h[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in rows[2:]:
            w.writerow([r[i].split("(")[0][-48:] if i == cols[0] else r[i] for i in cols])
    print(open(out).read())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    a = ap.parse_args()
    if a.launches:
        launches(a.launches, HERE / f"{a.tag}_launches.txt")
    if a.rep:
        raw(a.rep, HERE / f"{a.tag}_ncu_raw.csv")
