#!/usr/bin/env python
"""Summarises one kernel of an `ncu --page raw --csv` export: usage ncu_summary.py raw.csv [row]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
vals = rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_requests_srcunit_tex_op_red.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for k in keys:
    if k in d:
        print("%-70s %s %s" % (k, d[k][0], d[k][1]))
st = []
for h in hdr:
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        try:
            st.append((float(d[h][0].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
print("stall reasons (warps stalled per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:9]))
