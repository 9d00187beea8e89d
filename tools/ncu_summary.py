#!/usr/bin/env python3
"""Summarise ncu output into text files for profiles/ (run where the .ncu-rep / launch csv files are).

  python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.txt
  python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep [top]      > profiles/rNN_kernel.txt
"""
import collections
import csv
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path}); times are cold-cache and serialised: compare shares")
    print(f"{'kernel':90s} {'launches':>8s} {'avg us':>10s} {'total us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:90]:90s} {len(v):8d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / 1e3:10.1f} {100 * sum(v) / total:6.1f}%")


def kernel(path, top=14):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units, vals = rows[0], rows[1], rows[2]
    print(f"# ncu --set full --clock-control none --import-source on  ({path})")
    print("kernel:", vals[h.index("Kernel Name")])
    for name in RAW:
        if name in h:
            i = h.index(name)
            print(f"  {name:70s} {vals[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, body = rows[1], rows[2:]
    si = hdr.index("# Samples")
    cols = [k for k, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(b[si]) for b in body)
    print(f"\nwarp-state samples: {tot}")
    agg = {hdr[k]: sum(int(b[k]) for b in body) for k in cols}
    for name, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {name:28s} {v:8d} {100 * v / max(1, tot):5.1f}%")
    print(f"\nhottest SASS instructions (of {len(body)}):")
    for b in sorted(body, key=lambda b: -int(b[si]))[:top]:
        st = sorted(((int(b[k]), hdr[k][6:]) for k in cols), reverse=True)[:2]
        print(f"  {int(b[si]):7d} {100 * int(b[si]) / max(1, tot):5.1f}%  {b[1].strip()[:70]:70s} {[x for x in st if x[0]]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 14)
