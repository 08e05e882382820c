#!/usr/bin/env python
"""Summarise an ncu --set full capture of one kernel: key counters, stall reasons, pipe utilisation and executed
instructions per code region (tools/ncu_summary.py report.ncu-rep [n_half_tiles] [bucket_bytes])."""
import csv, subprocess, sys, io
rep = sys.argv[1]
units_n = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
B = int(sys.argv[3], 0) if len(sys.argv) > 3 else 0x800
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "lts__t_bytes.sum"]:
    for h in d:
        if h.endswith(k):
            print(f"{h[-70:]:70s} {d[h][0]:>20s} {d[h][1]}")
print("instructions per unit:", float(d["smsp__inst_executed.sum"][0]) / units_n)
print("stalls per issue:", ", ".join(f"{h[34:-23]} {float(d[h][0]):.2f}" for h in sorted(d)
                                      if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and float(d[h][0]) > 0.15))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isrc, iex, ism = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(int(r[ia], 16), r[isrc].strip(), int(r[iex]), int(r[ism])) for r in rows[2:] if len(r) > iex]
tot, tots, base = sum(x[2] for x in data), sum(x[3] for x in data), data[0][0]
scale = float(d["smsp__inst_executed.sum"][0]) / tot
buckets = {}
for a, s, e, sm in data:
    x = buckets.setdefault((a - base) // B, [0, 0, None])
    x[0] += e; x[1] += sm
    if x[2] is None: x[2] = s
for b in sorted(buckets):
    e, sm, s = buckets[b]
    if e > 0.01 * tot or sm > 0.01 * tots:
        print(f"{b * B:6x}: {100.0 * e / tot:5.1f}% inst ({e * scale / units_n:7.1f}/unit) {100.0 * sm / tots:5.1f}% samples  {s[:60]}")
