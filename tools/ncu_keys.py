#!/usr/bin/env python
"""Key metrics and stall reasons of every launch in an .ncu-rep (prints only).
    python tools/ncu_keys.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct",
        "sm__cycles_elapsed.avg", "local_load_bytes", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
STALLS = "smsp__pcsamp_warps_issue_stalled_"
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print("== " + d.get("Kernel Name", "?")[:100])
    for k in KEYS:
        if k in d:
            print("   %s = %s %s" % (k, d[k], u.get(k, "")))
    st = {h[len(STALLS):]: float(v) for h, v in d.items()
          if h.startswith(STALLS) and "not_issued" not in h and v not in ("", "n/a")}
    tot = sum(st.values()) or 1.0
    print("   stall samples: " + ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]))
