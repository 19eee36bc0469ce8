#!/usr/bin/env python
"""Summarises an .ncu-rep (raw page) into the handful of counters DESIGN.md / bench.py quote.

    python tools/ncu_summary.py gpurun_out/r01_fine_c2_4k.ncu-rep [--json out.json]
"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_no_instructions",
    "smsp__pcsamp_warps_issue_stalled_dispatch_stall", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
    "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_barrier",
]


def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in WANT:
                try:
                    v = float(vals[i].replace(",", ""))
                except ValueError:
                    v = vals[i]
                d[h] = {"value": v, "unit": units[i]}
        out.append(d)
    return out


if __name__ == "__main__":
    res = summarise(sys.argv[1])
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as fp:
            json.dump(res, fp, indent=1)
    for d in res:
        print(d["kernel"][:110])
        for k in WANT:
            if k in d:
                print(f"  {k:72s} {d[k]['value']:>16} {d[k]['unit']}")
