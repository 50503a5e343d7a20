#!/usr/bin/env python
"""Prints the metrics we track from an .ncu-rep (run here, no GPU needed):  python profiles/ncu_summary.py rep [more]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fmaheavy",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block", "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.avg.per_cycle_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled", "smsp__warp_issue_stalled",
        "smsp__pcsamp_warps_issue_stalled", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "smsp__cycles_active.avg", "sm__inst_executed_pipe_uniform", "smsp__thread_inst_executed_per_inst_executed"]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            print(f"== {rep} :: {name[:90]}")
            for h, u, v in zip(hdr, units, r):
                if any(h.startswith(w) for w in WANT):
                    print(f"{h:95s} {u:14s} {v}")


if __name__ == "__main__":
    main()
