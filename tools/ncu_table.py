#!/usr/bin/env python
"""Compact markdown table of an `ncu --page raw --csv` export: one row per profiled launch.
Usage: ncu_table.py <raw.csv> [pixels-per-launch ...]  (pixels: for the instr/px column)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
px = [float(x) for x in sys.argv[2:]]
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    try:
        v = float(r[col[name]].replace(",", ""))
    except Exception:
        return float("nan")
    u = units[col[name]]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "ns": 1e-3, "us": 1, "usecond": 1, "msecond": 1e3,
             "ms": 1e3, "nsecond": 1e-3}.get(u, 1)
    return v * scale


print("| kernel | µs | warp-instr | instr/px | issue % | warps active % | regs | DRAM rd MB | DRAM wr MB | L1 % | L2 % | FMA pipe % | ALU pipe % | FP64 % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for i, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]]
    inst = val(r, "smsp__inst_executed.sum")
    ipp = inst * 32 / px[i] if i < len(px) and px[i] else float("nan")
    print("| `%s` | %.1f | %.1fM | %s | %.0f | %.0f | %d | %.1f | %.1f | %.0f | %.0f | %.0f | %.0f | %.1f |" % (
        name[:60], val(r, "gpu__time_duration.sum"), inst / 1e6, ("%.0f" % ipp) if ipp == ipp else "-",
        val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        val(r, "launch__registers_per_thread"),
        val(r, "dram__bytes_read.sum") / 1e6, val(r, "dram__bytes_write.sum") / 1e6,
        val(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        val(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")))
