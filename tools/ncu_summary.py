"""Print the handful of ncu metrics the roofline discussion needs from a .ncu-rep (run where ncu is installed)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_pipe_xu", "local", "l1tex__t_bytes_pipe_lsu_mem_local"]
for r in rows[2:]:
    print("-----")
    for h, u, v in zip(hdr, rows[1], r):
        if any(h == w or (w in h and ("stalled" in w)) for w in want) or "issue_stalled" in h and h.endswith("per_issue_active.ratio") or h.startswith("smsp__sass_thread_inst_executed_op_d") and h.endswith(".sum") or "lsu_mem_local" in h and h.endswith(".sum") or h.startswith("smsp__sass_thread_inst_executed_op_d") and h.endswith(".sum.per_cycle_elapsed") or h == "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained":
            print(f"{h} [{u}] = {v}")

# ncu-counted fp64 FLOP rate as a fraction of the DFMA peak: (2 dfma + dadd + dmul) / (2 * peak dfma/cycle)
for r in rows[2:]:
    d = dict(zip(hdr, r))
    try:
        f = lambda k: float(d[k].replace(",", ""))
        frac = (2 * f("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") + f("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed")
                + f("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed")) / (2 * f("sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained"))
        print(f"derived: counted fp64 FLOP rate (2 DFMA + DADD + DMUL) / DFMA peak = {frac:.3f}")
    except Exception:
        pass
