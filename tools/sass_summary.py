#!/usr/bin/env python
"""SASS mnemonic counts per kernel of librecsys_b200.so (cuobjdump -sass; runs without a GPU).

    python tools/sass_summary.py > profiles/r2_sass_summary.md

The counts show which hardware paths each kernel uses: UTCHMMA / UTCBAR (tcgen05.mma / commit), LDTM / STTM (tcgen05.ld / st),
UTMALDG (TMA tensor loads), UBLKCP (bulk copies), HMMA (warp-level mma.sync), LDGSTS (cp.async), SYNCS (mbarrier).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
so = os.path.join(ROOT, "recsys_lite_b200", "librecsys_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "HMMA", "LDGSTS", "SYNCS", "DFMA", "FFMA", "REDUX", "ATOM"]
kern, counts, total, arch = None, collections.OrderedDict(), {}, set()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1).split(".")[0]
        total[kern] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[kern][w] += 1
                break
print("# SASS summary of recsys_lite_b200/librecsys_b200.so (`cuobjdump -sass`, tools/sass_summary.py)\n")
print(f"arch: {', '.join(sorted(arch))}\n")
print("| kernel | SASS instr | " + " | ".join(WATCH) + " |")
print("|---|---|" + "---|" * len(WATCH))
for k, c in counts.items():
    name = demangle(k)
    name = re.sub(r"rl::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*$", "", name)
    if total[k] < 40:
        continue
    print(f"| `{name[:70]}` | {total[k]} | " + " | ".join(str(c.get(w, 0) or "") for w in WATCH) + " |")
