"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of metrics the roofline discussion uses.
usage: python tools/ncu_summary.py raw.csv > profiles/<name>.md"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("smsp__warps_active.avg.per_cycle_active", "warps active / scheduler"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor pipe instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed (max)"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard (LDS)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (global)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction (I$)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
print(f"# ncu summary of {sys.argv[1]}\n")
for r in rows[2:]:
    print(f"## {r[ix['Kernel Name']].split('(')[0]}  (launch id {r[ix['ID']]})\n")
    print("| metric | value | unit |\n|---|---|---|")
    for key, label in WANT:
        if key in ix:
            print(f"| {label} (`{key}`) | {r[ix[key]]} | {units[ix[key]]} |")
    print()
