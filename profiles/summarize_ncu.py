#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the handful of counters the design argues
from.  Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
Runs on the CPU box (ncu -i only reads the report)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "lts__t_bytes.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel:", r[col["Kernel Name"]])
        for k in KEYS:
            if k in col:
                print(f"  {k:75s} {r[col[k]]:>16s} {units[col[k]]}")
        stalls = [(float(r[i] or 0), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h, i in col.items()
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
        print("  warp stall reasons (avg warps stalled per issue-active cycle, top 8):")
        for v, name in sorted(stalls, reverse=True)[:8]:
            print(f"    {name:30s} {v:8.3f}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
