#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

    python profiles/summarize.py launches gpurun_out/r1_launches_fp32.csv  > profiles/r1_launches_fp32.txt
    python profiles/summarize.py kernel   gpurun_out/r1_decode_fp32.ncu-rep > profiles/r1_decode_fp32.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct",
        "sm__throughput.avg.pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "sm__inst_executed_pipe_uniform", "smsp__issue_active.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread ", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum ",
        "lts__t_bytes.sum ", "sm__cycles_elapsed.max", "smsp__cycles_active.avg ")


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki].split("(")[0][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; source {path}")
    print(f"# {'total us':>12} {'launches':>8} {'share':>7}  kernel   (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v[1]:14.1f} {v[0]:8d} {100 * v[1] / tot:6.2f}%  {k}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none ; source {path}")
    for rec in rows[2:]:
        print(f"## launch: {rec[H.index('Kernel Name')][:100]}  grid {rec[H.index('Grid Size')]} block {rec[H.index('Block Size')]}")
        for h, u, v in zip(H, U, rec):
            if any(h.startswith(k.strip()) if k.endswith(" ") else k in h for k in KEEP):
                print(f"{h:90s} {u:12s} {v}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
