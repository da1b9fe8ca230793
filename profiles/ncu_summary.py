"""Summarises .ncu-rep captures into markdown:  python profiles/ncu_summary.py rep [rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(f"## {rep}: unreadable\n")
        continue
    hdr, units = rows[0], rows[1]
    print(f"## {rep}\n")
    print("| metric | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |")
    print("|---|" + "---|" * (len(rows) - 2))
    print("| kernel | " + " | ".join(r[hdr.index("Kernel Name")][:60] for r in rows[2:]) + " |")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} [{units[i]}] | " + " | ".join(r[i] for r in rows[2:]) + " |")
    print()
