#!/usr/bin/env python3
"""Text summaries of ncu output for profiles/:
    ncu_summary.py launches <launches.csv>      per-kernel totals of a `--metrics gpu__time_duration.sum` launch list
    ncu_summary.py full <report.ncu-rep>        key metrics per captured launch of a `--set full` report
"""
import collections
import csv
import subprocess
import sys

FULL = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[start]
    ix = {n: i for i, n in enumerate(h)}
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) < len(h):
            continue
        try:
            v = float(r[ix["Metric Value"]])
        except ValueError:
            continue
        unit = r[ix["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        key = (r[ix["Kernel Name"]].split("(")[0], r[ix["Grid Size"]], r[ix["Block Size"]])
        agg.setdefault(key, []).append(v)
    total = sum(sum(v) for v in agg.values())
    print("%-46s %-16s %-14s %5s %10s %10s %6s" % ("kernel", "grid", "block", "n", "avg us", "total us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-46s %-16s %-14s %5d %10.1f %10.1f %5.1f%%" % (k[0][:46], k[1], k[2], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / total))
    print("total %.1f us over %d launches" % (total, sum(len(v) for v in agg.values())))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print("%s  grid %s block %s" % (r[h.index("Kernel Name")].split("(")[0], r[h.index("Grid Size")], r[h.index("Block Size")]))
        for m in FULL:
            if m in h:
                print("    %-82s %14s %s" % (m, r[h.index(m)], units[h.index(m)]))


def units(path, out_json=None):
    """per-unit figures of the dominant kernel for bench.py's roofline: executed thread-instructions per luma
    pixel per reference and DRAM bytes per launch of luma_search_2step (1080p, averaged over the captures)"""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, un = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    inst, dram, launches_, refs = 0.0, 0.0, 0, 0
    for r in rows[2:]:
        if "k_luma_search_2step" not in r[h.index("Kernel Name")]:
            continue
        g = [int(x) for x in r[h.index("Grid Size")].strip("() ").split(",")]
        inst += float(r[h.index("smsp__inst_executed.sum")])
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            dram += float(r[h.index(m)]) * scale.get(un[h.index(m)], 1.0)
        launches_ += 1
        refs += g[1]
    res = {"kernel": "luma_search_2step", "captures": launches_, "reference_launch_units": refs,
           "thread_instr_per_pixel_per_ref": inst * 32 / refs / (1920 * 1088),
           "dram_bytes_per_ref": dram / refs, "source": path.split("/")[-1]}
    print(json.dumps(res, indent=1))
    if out_json:
        json.dump(res, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full, "units": units}[sys.argv[1]](*sys.argv[2:])
