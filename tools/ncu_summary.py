"""Summarise an .ncu-rep (first matching kernel) into a small text block for profiles/."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__inst_issued.avg.per_cycle_active",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:90s} {r[i]} {units[i]}")
print("-- warp stall reasons (pct of issue-slot samples, > 1 %) --")
for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
        try:
            v = float(r[i])
        except ValueError:
            continue
        if v > 0.02:
            print(f"{h:90s} {v:.3f}")
