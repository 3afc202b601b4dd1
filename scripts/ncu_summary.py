"""Developer aid (CPU box): condense an .ncu-rep into the handful of metrics the roofline discussion needs.
usage: python scripts/ncu_summary.py <file.ncu-rep> [out.txt]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct"]


def summarize(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    lines = []
    for v in rows[2:]:
        lines.append("kernel: " + v[h.index("Kernel Name")] + "   grid " + v[h.index("launch__grid_size")])
        for i, n in enumerate(h):
            if n in WANT:
                lines.append(f"  {n:88s} {v[i]:>18s} {u[i]}")
        lines.append("")
    return "\n".join(lines)


if __name__ == "__main__":
    s = summarize(sys.argv[1])
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(s)
    print(s)
