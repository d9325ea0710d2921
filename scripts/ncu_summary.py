"""Prints the headline metrics of every kernel in an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
idx = [(h, i) for i, h in enumerate(hdr) if h in want]
for r in rows[2:]:
    print("---")
    for h, i in idx:
        print(f"  {h} = {r[i][:100]} {rows[1][i]}")
