"""Extracts the judged metrics from an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into a small text file.
usage: python profiles/extract.py gpurun_out/X.ncu-rep profiles/X_summary.txt"""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], "w") as f:
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        f.write("kernel: %s\n" % name)
        for i, h in enumerate(hdr):
            if h in WANT or "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
                f.write("  %-80s %-12s %s\n" % (h, units[i], r[i]))
print(open(sys.argv[2]).read())
