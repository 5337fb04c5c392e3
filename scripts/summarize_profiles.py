#!/usr/bin/env python
"""Turn the ncu artefacts of one GPU visit (gpurun_out/) into tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag> [launches_csv] [ncu_rep]

Writes profiles/<tag>_launches.md (launch list: kernel, count, mean device time, share of the step),
profiles/<tag>_kernels.json (key ncu --set full metrics per captured kernel) and refreshes
profiles/traffic.json (DRAM bytes per launch of the forward / gradient kernel, read by bench.py).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
rep = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag)
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

if os.path.exists(launches):
    rows = list(csv.reader(open(launches)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[h]
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) < len(hdr):
            continue
        name = r[hdr.index("Kernel Name")]
        agg.setdefault(name, []).append(float(r[hdr.index("Metric Value")]))
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(out_dir, "%s_launches.md" % tag), "w") as f:
        f.write("# ncu launch list, `%s`\n\n" % tag)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv python bench.py --steps 5 --warmup 3`\n")
        f.write("(cold-cache, serialised: compare SHARES, not absolutes)\n\n")
        f.write("| kernel | launches | mean device time (us) | share of listed time |\n|---|---:|---:|---:|\n")
        for k, v in agg.items():
            f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (k[:110], len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / total))
    print("wrote", os.path.join(out_dir, "%s_launches.md" % tag))

raw_csv = os.path.join(ROOT, "gpurun_out", "prof_%s_raw.csv" % tag)     # exported on the GPU box (the .ncu-rep of two
if os.path.exists(rep) or os.path.exists(raw_csv):                       # kernels exceeds gpurun's 64 MiB return limit)
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        raw = open(raw_csv).read()
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_active",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sass__inst_executed_global_loads", "smsp__inst_executed_op_global_red.sum",
            "smsp__inst_executed_op_shared_atom.sum", "sass__inst_executed_shared_loads",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
            "sm__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
    out = []
    for r in rows[2:]:
        d = {}
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                d[k] = r[i] + (" " + units[i] if units[i] else "")
        out.append(d)
    json.dump(out, open(os.path.join(out_dir, "%s_kernels.json" % tag), "w"), indent=1)
    print("wrote", os.path.join(out_dir, "%s_kernels.json" % tag))

    def to_bytes(s):
        v, u = s.split()[0], (s.split() + [""])[1]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    traffic = {}
    for d in out:
        if "dram__bytes_read.sum" not in d:
            continue
        t = to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"])
        n = d["Kernel Name"]
        if "edf_" in n:
            if "_fwd_" in n:
                key = "fwd"
            elif "grad" in n:
                key = "grad"
            else:
                key = "grad" if ("true" in n or ", 1>" in n) else "fwd"
            traffic[key] = int(t)
    if traffic:
        traffic["source"] = "ncu --set full, %s" % tag
        json.dump(traffic, open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)
        print("traffic", traffic)
