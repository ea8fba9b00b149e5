#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export (ncu -i X.ncu-rep --page raw --csv > X.csv) into one short block per launch:
duration, DRAM bytes, DRAM / L1TEX / L2 utilisation, occupancy, registers, FP64 / DMMA pipe activity, main stall reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
     ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex throughput %"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 throughput %"), ("lts__t_sector_hit_rate.pct", "l2 hit %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "registers"),
     ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__occupancy_limit_registers", "occ limit regs"),
     ("launch__occupancy_limit_shared_mem", "occ limit smem"),
     ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma inst % of peak"),
     ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
     ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
     ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
     ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
     ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
     ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
     ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait")]
for n, r in enumerate(rows[2:]):
    name = r[col["Kernel Name"]].replace("void unnamed>::", "").split("(")[0]
    print(f"[{n}] {name}  grid={r[col['Grid Size']]} block={r[col['Block Size']]}")
    for key, label in M:
        if key in col and r[col[key]] != "":
            print(f"      {label:26s} {r[col[key]]} {units[col[key]]}")
