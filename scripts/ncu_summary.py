#!/usr/bin/env python
"""Key metrics of an `ncu --page raw --csv` export: python scripts/ncu_summary.py gpurun_out/prof_<tag>_raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_atom.sum",
        "smsp__inst_executed_op_shared_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_tma.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in KEYS:
        if k in d and d[k] != "":
            print("  %-72s %s" % (k, d[k]))
    st = []
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and d.get(k, "") not in ("", "n/a"):
            st.append((k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(d[k].replace(",", ""))))
    print("  stalls per issue: " + ", ".join("%s %.2f" % kv for kv in sorted(st, key=lambda x: -x[1])[:9]))
    print()
