"""Summarise `ncu --page raw --csv` exports: the roofline-relevant metrics per captured launch."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct"]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        print(path, "empty"); continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]][:70]
        print("== %s :: %s" % (path.split("/")[-1], name))
        for k in KEYS:
            if k in idx:
                print("   %-80s %s %s" % (k, r[idx[k]], units[idx[k]]))
        if "--all" in sys.argv:
            pass
