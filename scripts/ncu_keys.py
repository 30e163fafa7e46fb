import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__cycles_active.avg", "sm__cycles_active.avg", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    for k in keys:
        for i, h in enumerate(hdr):
            if h == k:
                print(f"{k:80s} {r[i]:>22s} {units[i]}")
    print("---")
