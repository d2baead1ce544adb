#!/usr/bin/env python
"""ncu_summary.py REPORT.ncu-rep OUT.csv -- the handful of `ncu --page raw` counters the design notes quote, as
metric,unit,value rows (run here on the CPU box: `ncu -i` needs no GPU)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, kernels = rows[0], rows[1], rows[2:]          # one row per captured launch
    names = [k[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "" for k in kernels]
    with open(out, "w") as f:
        f.write("metric,unit," + ",".join(f"value_{i}" for i in range(len(kernels))) + "\n" if len(kernels) > 1 else "metric,unit,value\n")
        f.write("kernel,," + ",".join(f'"{n}"' for n in names) + "\n")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"{w},{units[i]}," + ",".join(k[i] for k in kernels) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
