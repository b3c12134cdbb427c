#!/usr/bin/env python
"""Turn ncu outputs brought back from the GPU box (gpurun_out/) into the small text / json summaries kept here.

    python profiles/summarize.py launches <launches.csv> <out.txt> [title]
    python profiles/summarize.py full <report.ncu-rep> <out.txt> [title]

`full` needs the `ncu` CLI (present in the authoring image; no GPU needed to read a report)."""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    # L1 data pipe: the limit several round-2 kernels turned out to sit at (wavefronts = 128-byte passes of the LSU)
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    t, n = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        if "time" not in r[mi]:
            continue
        try:
            t[r[ki]] += float(r[vi].replace(",", ""))
            n[r[ki]] += 1
        except ValueError:
            pass
    tot = sum(t.values())
    with open(out, "w") as f:
        f.write("# %s\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n" % title)
        f.write("# total kernel time %.1f us over %d launches\n" % (tot / 1e3, sum(n.values())))
        for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
            f.write("%-78s n=%4d total=%9.1f us avg=%8.1f us share=%5.1f%%\n" % (k[:78], n[k], v / 1e3, v / n[k] / 1e3, 100 * v / tot))


def full(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# %s\n# ncu --set full --clock-control none --import-source on\n" % title)
        for r in rows[2:]:
            f.write("\n%s\n" % r[hdr.index("Kernel Name")])
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("  %-86s %s %s\n" % (m, r[i], units[i]))


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, dst, title)
