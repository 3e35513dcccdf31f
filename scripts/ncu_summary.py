#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small tracked text file under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_name.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg.per_second", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_drain_per_warp_active.pct",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    units, data = rows[1], rows[2:]
    ki = h.index("Kernel Name")
    lines = ["# ncu --set full --clock-control none summary of %s" % rep, ""]
    for n, r in enumerate(data):
        lines.append("## launch %d: %s" % (n, r[ki][:120]))
        for k in KEYS:
            if k in h:
                i = h.index(k)
                lines.append("%-78s %s %s" % (k, r[i], units[i]))
        for i, k in enumerate(h):   # every stall reason the capture holds (cycles per issued instruction)
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and k not in KEYS:
                lines.append("%-78s %s %s" % (k, r[i], units[i]))
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
