#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed): key raw metrics + instructions / stall samples per source line.
usage: python scripts/ncu_summary.py gpurun_out/X.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

csv.field_size_limit(10**9)
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "sm__inst_executed_pipe_fp64.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:70s} {units[i]:14s} {[r[i] for r in rows[2:]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.OrderedDict()
cur = None
stall_cols = None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        stall_cols = {h: i for i, h in enumerate(r)}
        continue
    if len(r) > 7 and r[0].isdigit():
        try:
            inst, samp = int(r[7]), int(r[4])
        except ValueError:
            continue
        a = agg.setdefault((cur, int(r[0])), [0, 0, r[1]])
        a[0] += inst
        a[1] += samp
tot = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print(f"total warp-inst {tot} (all profiled launches), stall samples {ts}")
print("--- by instructions")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:4d} inst {a[0] / tot * 100:5.1f}% samp {a[1] / ts * 100:5.1f}%  {a[2][:110]}")
print("--- by stall samples")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]}:{k[1]:4d} inst {a[0] / tot * 100:5.1f}% samp {a[1] / ts * 100:5.1f}%  {a[2][:110]}")

# ---- instructions / stall samples per enclosing function of the kernel source
import re
src_path = None
for k in agg:
    if k[0] and k[0].endswith("b200aug_fused.cu"):
        src_path = k[0]
import os
cu = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neuralnet-tracker-traincode_b200", "csrc", "b200aug_fused.cu")
if os.path.exists(cu):
    lines = open(cu).read().split("\n")
    func_at = []
    cur_f = "(top)"
    pat = re.compile(r"^(?:template.*\n)?(?:__device__|__global__|static|extern)[^;{]*?\b([A-Za-z_0-9]+)\s*\(")
    for i, ln in enumerate(lines, 1):
        m = re.match(r"^(?:__device__|__global__)[^;]*?\b([A-Za-z_0-9]+)\s*\(", ln)
        if m and not ln.startswith(" "):
            cur_f = m.group(1)
        func_at.append(cur_f)
    per = collections.Counter()
    pers = collections.Counter()
    for (f, l), a in agg.items():
        name = func_at[l - 1] if f == "b200aug_fused.cu" and 0 < l <= len(func_at) else f
        per[name] += a[0]
        pers[name] += a[1]
    print("--- by function (lines of b200aug_fused.cu; inlined callees count where they are written)")
    for name, v in per.most_common(25):
        print(f"{name:32s} inst {v / tot * 100:5.1f}%  samples {pers[name] / ts * 100:5.1f}%")
