#!/usr/bin/env python
"""profiles/<tag>_ncu_full_summary.json + profiles/r2_traffic.json from an `ncu --set full` report of the three hot launches
(user half-step, item half-step, scoring) -- runs without a GPU.

    python tools/summarize_r2.py gpurun_out/r2b_full.ncu-rep r2 "<command line of the capture>" "<note>"
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
rep, tag, cmd, note = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
SC = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
BY = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def val(r, name):
    i = hdr.index(name)
    return float(r[i].replace(",", "")), units[i]


launches = []
names = ["als_k64_kernel (user half-step)", "als_k64_kernel (item half-step)", "score_tc2_kernel"]
for r, nm in zip(rows[2:], names):
    d = {"kernel": nm, "kernel_name": r[hdr.index("Kernel Name")]}
    v, u = val(r, "gpu__time_duration.sum"); d["duration_ms"] = v * SC[u]
    for key, m in (("dram_read_bytes", "dram__bytes_read.sum"), ("dram_write_bytes", "dram__bytes_write.sum")):
        v, u = val(r, m); d[key] = v * BY[u]
    for key, m in (("tensor_pipe_active_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                   ("alu_pipe_active_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                   ("fma_pipe_active_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                   ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   ("registers_per_thread", "launch__registers_per_thread"),
                   ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                   ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
                   ("inst_executed", "smsp__inst_executed.sum"),
                   ("sm_cycles", "sm__cycles_elapsed.avg"),
                   ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                   ("grid", "launch__grid_size"), ("block", "launch__block_size"),
                   ("smem_dyn_bytes", "launch__shared_mem_per_block_dynamic")):
        if m in hdr:
            d[key] = val(r, m)[0]
    d["stalls_per_issue"] = {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: round(float(r[i]), 2)
                             for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and
                             h.endswith("_per_issue_active.ratio") and float(r[i]) > 0.15}
    launches.append(d)
out = {"command": cmd, "note": note, "launches": launches}
json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.json"), "w"), indent=1)
tr = {"source": f"profiles/{tag}_ncu_full_summary.json ({note})",
      "als_user_half_step": int(launches[0]["dram_read_bytes"] + launches[0]["dram_write_bytes"]),
      "als_item_half_step": int(launches[1]["dram_read_bytes"] + launches[1]["dram_write_bytes"]),
      "score_topn": int(launches[2]["dram_read_bytes"] + launches[2]["dram_write_bytes"]),
      "tensor_pipe_active_pct": {"score_topn": launches[2]["tensor_pipe_active_pct"], "als_user_half_step": launches[0]["tensor_pipe_active_pct"],
                                 "als_item_half_step": launches[1]["tensor_pipe_active_pct"]}}
json.dump(tr, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
for d in launches:
    print(d["kernel"], round(d["duration_ms"], 3), "ms; dram", round((d["dram_read_bytes"] + d["dram_write_bytes"]) / 1e9, 3), "GB; tensor",
          round(d["tensor_pipe_active_pct"], 1), "% alu", round(d.get("alu_pipe_active_pct", 0), 1), "% issue", round(d["issue_active_pct"], 1), "%", d["stalls_per_issue"])
