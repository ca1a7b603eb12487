"""Condense `ncu --set full` reports into the small CSV kept under profiles/.

usage: python tools/ncu_summary.py OUT.csv REPORT.ncu-rep [REPORT2.ncu-rep ...]
One column per profiled launch, one row per metric; durations are converted to us and byte counts to
Mbyte (ncu picks a unit per report), everything else is copied as printed.
"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "lts__t_sector_hit_rate.pct",
]


TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
BYTES = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def normalise(value, unit):
    try:
        x = float(value.replace(",", ""))
    except ValueError:
        return value, unit
    if unit in TIME:
        return f"{x * TIME[unit]:.3f}", "us"
    if unit in BYTES:
        return f"{x * BYTES[unit]:.3f}", "Mbyte"
    return value, unit


def load(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    cols = []
    for r in launches:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("unnamed>::", "").replace("hpd::", "")
        vals = {}
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                vals[m] = normalise(r[i], units[i])
        cols.append((name, vals))
    return cols


def main():
    out, reports = sys.argv[1], sys.argv[2:]
    cols = [c for rep in reports for c in load(rep)]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [c[0] for c in cols])
        for m in METRICS:
            unit = next((c[1][m][1] for c in cols if m in c[1]), "")
            w.writerow([m, unit] + [c[1].get(m, ("", ""))[0] for c in cols])


if __name__ == "__main__":
    main()
