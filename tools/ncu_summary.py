"""condense `ncu -i X.ncu-rep --page raw --csv` into the handful of counters DESIGN.md quotes, one line per kernel launch"""
import csv, sys, re
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_membar.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum", "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct",
        "derived__memory_l2_theoretical_sectors_global_excessive", "local_load_bytes", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    out = [name]
    for k in KEYS:
        if k in col and r[col[k]] != "":
            out.append("%s=%s%s" % (k, r[col[k]], (" " + units[col[k]]) if units[col[k]] else ""))
    print(" | ".join(out))
