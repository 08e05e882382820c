#!/usr/bin/env python
"""Turn the ncu outputs of a profiling call into the tracked summaries under profiles/.

    python tools/summarize_profiles.py --launches gpurun_out/launches_r1.csv --rep gpurun_out/prof_r1_bench.ncu-rep --tag r1

Inputs (produced on the GPU box, see /opt/skills/guides/B200_PROFILING.md and profiles/README.md):
  --launches  CSV log of `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python bench.py ...`
  --rep       report of `ncu --set full --clock-control none --import-source on -k regex:... -o ... python bench.py ...`
Outputs: profiles/<tag>_launches_ncu.csv (copy), profiles/<tag>_launches.md (per-kernel shares),
         profiles/<tag>_ncu_full_summary.json (selected counters per captured launch),
         profiles/<tag>_traffic.json (dram bytes per launch and tensor-pipe activity: read by bench.py).
Runs here (no GPU needed): `ncu -i` only reads the report.
"""
import argparse
import csv
import io
import json
import os
import shutil
import subprocess

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]


def read_ncu_csv(text):
    lines = [ln for ln in text.splitlines() if ln.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def launches_table(path, tag):
    rows = read_ncu_csv(open(path).read())
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if len(r) != len(hdr) or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[r[iu]]
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    total = sum(a[1] for a in agg.values())
    out = [f"# {tag} launch list under ncu --metrics gpu__time_duration.sum --clock-control none",
           "# (cold-cache, serialised: compare SHARES, not absolutes; command in profiles/README.md)", "",
           "| kernel | launches | total ms | avg ms | share |", "|---|---|---|---|---|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:80]}` | {n} | {ms:.3f} | {ms / n:.3f} | {ms / total:.3f} |")
    return "\n".join(out) + "\n"


def full_summary(rep):
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = read_ncu_csv(text)
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"Kernel Name": r[hdr.index("Kernel Name")]}
        for m in KEEP:
            if m in hdr:
                d[m] = f"{r[hdr.index(m)]} {units[hdr.index(m)]}".strip()
        out.append(d)
    return out


def num(s):
    v, _, unit = s.partition(" ")
    v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(unit, 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--tag", default="r1")
    a = ap.parse_args()
    pdir = os.path.join(ROOT, "profiles")
    if a.launches:
        shutil.copyfile(a.launches, os.path.join(pdir, f"{a.tag}_launches_ncu.csv"))
        open(os.path.join(pdir, f"{a.tag}_launches.md"), "w").write(launches_table(a.launches, a.tag))
    if a.rep:
        summ = full_summary(a.rep)
        json.dump(summ, open(os.path.join(pdir, f"{a.tag}_ncu_full_summary.json"), "w"), indent=1)
        # bench.py order inside a step: user half-step, item half-step, scoring
        als = [d for d in summ if "als_k64" in d["Kernel Name"]]
        sc = [d for d in summ if "score_topn_tc" in d["Kernel Name"]]
        tr = {"source": f"profiles/{a.tag}_ncu_full_summary.json: ncu --set full --clock-control none on `python bench.py --steps 2 "
                        "--warmup 1 --no-cpu` (C4, 1 x B200), dram__bytes_read.sum + dram__bytes_write.sum per launch"}
        tp = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
        pct = {}
        if sc:
            tr["score_topn"] = int(num(sc[0]["dram__bytes_read.sum"]) + num(sc[0]["dram__bytes_write.sum"]))
            pct["score_topn"] = float(sc[0][tp].split()[0])
        if len(als) >= 2:
            for name, d in (("als_user_half_step", als[0]), ("als_item_half_step", als[1])):
                tr[name] = int(num(d["dram__bytes_read.sum"]) + num(d["dram__bytes_write.sum"]))
                pct[name] = float(d[tp].split()[0])
        tr["tensor_pipe_active_pct"] = pct
        json.dump(tr, open(os.path.join(pdir, f"{a.tag}_traffic.json"), "w"), indent=1)
    print("profiles updated")


if __name__ == "__main__":
    main()
