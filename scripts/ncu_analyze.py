#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, stall reasons, opcode mix (per tile-var if given).
usage: ncu_analyze.py report.ncu-rep [units_per_launch]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for h, u, v in zip(hdr, rows[1], vals):
    if h in want:
        print(f"{h:70s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for s in stalls:
        tot[s] += int(r[ix[s]] or 0)
T = sum(tot.values())
print("--- stall samples")
for s, v in tot.most_common(10):
    print(f"{s:28s} {v:8d} {100*v/T:5.1f}%")
op = collections.Counter(); smp = collections.Counter()
for r in data:
    parts = r[ix["Source"]].split()
    o = parts[0] if not parts[0].startswith("@") else parts[1]
    o = o.split(".")[0]
    op[o] += int(r[ix["Instructions Executed"]] or 0)
    smp[o] += int(r[ix["# Samples"]] or 0)
I = sum(op.values())
print("--- warp instructions", I, ("per unit %.1f" % (I/units)) if units else "")
for o, v in op.most_common(22):
    per = f"{v/units:9.1f}" if units else ""
    print(f"{o:10s} {per} {100*v/I:5.1f}%  samples {100*smp[o]/max(1,sum(smp.values())):5.1f}%")
