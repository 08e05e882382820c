#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of one kernel from an ncu report (runs without a GPU).

    python tools/src_hot.py gpurun_out/x.ncu-rep score_tc2 [top]
"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None; agg = {}
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if not hdr or len(r) < len(hdr) or r[2] != "-": continue   # source-line summary rows have '-' as address
    ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
    try: n = int(r[ie]); s = int(r[isamp])
    except ValueError: continue
    agg[(cur, int(r[0]))] = (n, s, r[1].strip())
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f"total warp instructions {tot:.4g}, samples {tots}")
for (f, ln), (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:5d} {100*n/tot:5.1f}% inst {100*s/max(tots,1):5.1f}% samp  {src[:110]}")
