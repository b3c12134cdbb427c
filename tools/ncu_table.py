"""one line per captured kernel of an .ncu-rep (development helper)"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
def col(name):
    return hdr.index(name) if name in hdr else None
cols = [("t_us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("regs", "launch__registers_per_thread"),
        ("inst", "smsp__inst_executed.sum"), ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("dramR", "dram__bytes_read.sum"), ("dramW", "dram__bytes_write.sum"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), ("fp64%", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("LS", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("SS", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("bar", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
        ("lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
        ("mpt", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("l1hit%", "l1tex__t_sector_hit_rate.pct"), ("l2hit%", "lts__t_sector_hit_rate.pct")]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    out = [r[ki][:44].ljust(44)]
    for short, name in cols:
        c = col(name)
        if c is None:
            continue
        v = r[c].replace(",", "")
        try:
            f = float(v)
            v = "%.3g" % f
        except ValueError:
            pass
        out.append("%s=%s" % (short, v))
    print(" ".join(out))
