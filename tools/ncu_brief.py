"""Print the handful of ncu metrics used in profiles/ from a .ncu-rep (one line per captured launch)."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "us"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
    ("dram__bytes_read.sum", "rd"),
    ("dram__bytes_write.sum", "wr"),
    ("smsp__inst_executed.sum", "inst"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_nsel"),
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    parts = []
    for name, short in WANT:
        if name in hdr:
            v = r[hdr.index(name)]
            u = units[hdr.index(name)]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            parts.append(f"{short}={v}{u if short in ('rd', 'wr') else ''}")
    print(r[ik][:60], " ".join(parts))
