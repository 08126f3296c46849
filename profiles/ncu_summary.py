#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline counters, stall reasons, and the
hottest source lines.  Usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [n_warps]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("=== kernel:", name[:90])
    for k in KEYS:
        if k in hdr:
            print(f"  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    st = []
    for i, k in enumerate(hdr):
        if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                st.append((float(r[i]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    print("  stall reasons (pc sampling, all samples):")
    for v, k in sorted(st, reverse=True)[:9]:
        print(f"    {k:28s} {100 * v / tot:5.1f}%")
    break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, h, lines = None, None, []
seen_kernel = 0
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        h = r
        continue
    if h and len(r) == len(h) and r[0] != "":
        try:
            lines.append((cur, int(r[0]), r[1], int(r[7] or 0), int(r[4] or 0)))
        except ValueError:
            pass
n_kern = max(1, len(rows) - 2)
ti, ts = sum(l[3] for l in lines) or 1, sum(l[4] for l in lines) or 1
nw = float(sys.argv[2]) if len(sys.argv) > 2 else None
print(f"=== source page: {ti / n_kern:.0f} warp-instructions per launch" + (f" = {ti / n_kern / nw:.1f} per warp" if nw else ""))
byf = collections.Counter()
for f, n, s, i, sm in lines:
    byf[f] += i
for f, v in byf.most_common():
    print(f"  {f:28s} {100 * v / ti:5.1f}% of instructions")
print("  hottest lines by instructions:")
for f, n, s, i, sm in sorted(lines, key=lambda l: -l[3])[:28]:
    print(f"    {f}:{n:<4d} inst {100 * i / ti:5.1f}%  samples {100 * sm / ts:5.1f}%  {s.strip()[:88]}")
print("  hottest lines by stall samples:")
for f, n, s, i, sm in sorted(lines, key=lambda l: -l[4])[:12]:
    print(f"    {f}:{n:<4d} inst {100 * i / ti:5.1f}%  samples {100 * sm / ts:5.1f}%  {s.strip()[:88]}")
