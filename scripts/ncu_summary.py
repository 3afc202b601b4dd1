"""Developer aid (CPU box): condense an .ncu-rep into the handful of metrics the roofline discussion needs."""
import csv, subprocess, sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum ",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
print("kernel:", v[h.index("Kernel Name")] if "Kernel Name" in h else "?")
for i, n in enumerate(h):
    if any(n == w.strip() or n.startswith(w) for w in WANT):
        print(f"{n:90s} {v[i]:>16s} {u[i]}")
