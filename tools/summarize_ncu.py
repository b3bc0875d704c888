#!/usr/bin/env python
"""Key counters of every kernel in an `ncu --set full` report, as a markdown table.

  python tools/summarize_ncu.py gpurun_out/prof_X.ncu-rep > profiles/r01_ncu_X.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "stall long_sb %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long_sb/issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "lg_throttle/issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "math_throttle/issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "short_sb/issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "barrier/issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "mio_throttle/issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "wait/issue"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sectors_op_atom.sum", "L2 atom sectors"),
    ("lts__t_sectors_op_red.sum", "L2 red sectors"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    print("Report: `%s`\n" % rep.split("/")[-1])
    for r in rows[2:]:
        print("### `%s`\n" % r[ki].split("(")[0].replace("void ", ""))
        print("| counter | value |")
        print("|---|---:|")
        for key, label in WANT:
            if key in h:
                i = h.index(key)
                print("| %s (`%s`) | %s %s |" % (label, key, r[i], units[i]))
        print()


if __name__ == "__main__":
    main()
