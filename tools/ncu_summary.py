"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "local_load", "local_store", "smsp__inst_executed_op_local"]
STALL = "smsp__average_warps_issue_stalled_"

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=== kernel:", d.get("Kernel Name", "?")[:110], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for h, u in zip(hdr, units):
        if any(h == k or (k in h and not h.endswith("pct") and k in ("local_load", "local_store")) for k in KEYS) or h in KEYS:
            print("  %-86s %s %s" % (h, d[h], u))
    stalls = sorted(((float(d[h]), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h in hdr if h.startswith(STALL) and d[h]), reverse=True)
    print("  stalls per issue:", ", ".join("%s=%.2f" % (n, v) for v, n in stalls[:7]))
