"""Condense an `ncu --page raw --csv` export to the columns quoted in profiles/README.md.
usage: python tools/ncu_summary.py gpurun_out/x_raw.csv profiles/x_ncu_full_summary.csv"""
import csv, sys
COLS = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__maximum_warps_per_active_cycle_pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = [hdr.index(c) for c in COLS if c in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i][:90] for i in idx])
