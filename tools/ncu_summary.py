"""Prints the handful of ncu metrics we track from a .ncu-rep (raw page)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "sass__inst_executed_global_loads",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in KEYS:
        if k in d:
            print("  %-80s %s %s" % (k, d[k], units[hdr.index(k)]))
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            v = float(d[k] or 0)
            if v >= 0.05:
                print("  stall %-60s %.3f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
