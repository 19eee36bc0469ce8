#!/usr/bin/env python
"""Summarises an .ncu-rep (raw page) into the handful of counters DESIGN.md / bench.py quote.

    python tools/ncu_summary.py gpurun_out/r01_fine_c2_4k.ncu-rep [--json out.json] [--traffic WORKLOAD [--kernel REGEX]]

--traffic WORKLOAD writes the capture's DRAM traffic and issue counters into profiles/traffic.json under that workload
name, stamped with the hash of the kernel sources and the commit it was taken at: bench.py quotes the entry only while
the kernel sources still hash to the same value (a capture of older code is withheld, not reported stale).
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
    # the L1 / L2 side of the "HBM / L2 roofline", and the pipes the loop issues to
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "lts__t_bytes.sum", "lts__t_sectors.sum", "l1tex__t_sectors.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.avg.per_cycle_active", "sm__cycles_active.avg",
]
# plus anything about NVLink / peer / system-memory apertures (multi-GPU gather, zero-copy host frames)
WANT_SUBSTRINGS = ("nvlrx", "nvltx", "nvlink", "aperture_peer", "aperture_sysmem", "pcie")


def kernel_sass_hash(binary, pattern="finePassKernelILb1EjE"):
    """sha256 (16 hex digits) over the SASS of one kernel inside a cubin-carrying file (the library, an object file):
    ties an ncu capture to the MACHINE CODE it profiled -- edits elsewhere in the source files do not change it, any change
    of the kernel's code does. None when cuobjdump is unavailable or the kernel is not found."""
    import hashlib
    import re
    try:
        out = subprocess.run(["cuobjdump", "-sass", str(binary)], capture_output=True, text=True, timeout=300).stdout
    except Exception:  # noqa: BLE001
        return None
    keep, lines = False, []
    for ln in out.splitlines():
        if "Function :" in ln:
            keep = pattern in ln
        elif keep and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            lines.append(re.sub(r"/\* 0x[0-9a-f]* \*/", "", ln).rstrip())
    return hashlib.sha256("\n".join(lines).encode()).hexdigest()[:16] if lines else None


def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in WANT or any(x in h for x in WANT_SUBSTRINGS):
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
        for k in list(WANT) + sorted(k for k in d if k not in WANT and k != "kernel"):
            if k in d:
                print(f"  {k:72s} {d[k]['value']:>16} {d[k]['unit']}")
    if "--traffic" in sys.argv:
        import hashlib
        import re
        from pathlib import Path
        root = Path(__file__).resolve().parent.parent
        workload = sys.argv[sys.argv.index("--traffic") + 1]
        pattern = sys.argv[sys.argv.index("--kernel") + 1] if "--kernel" in sys.argv else "finePass"
        pick = [d for d in res if re.search(pattern, d["kernel"])]
        if not pick:
            raise SystemExit(f"no kernel matching {pattern!r} in {sys.argv[1]}")
        d = pick[0]
        g = lambda k: d[k]["value"] if k in d else None  # noqa: E731
        byte_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

        def gb(k):      # ncu prints byte counters in scaled units
            return None if k not in d else d[k]["value"] * byte_scale.get(d[k]["unit"], 1.0)
        to_ms = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
        h = hashlib.sha256()
        for name in ("svo_traverse.cuh", "svo_kernels.cu", "svo_kernels.cuh"):
            h.update((root / "sparse-voxel-octrees_b200" / "csrc" / name).read_bytes())
        commit = subprocess.run(["git", "-C", str(root), "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        entry = {
            "source": f"{sys.argv[1]} (ncu --set full --clock-control none, one launch)", "kernel": d["kernel"][:80],
            "kernel_source_hash": h.hexdigest()[:16], "commit": commit or None,
            "kernel_sass_hash": kernel_sass_hash(root / "sparse-voxel-octrees_b200" / "libsvo_b200.so"),
            "dram_bytes_per_launch": (gb("dram__bytes_read.sum") or 0) + (gb("dram__bytes_write.sum") or 0),
            "dram_bytes_read": gb("dram__bytes_read.sum"), "dram_bytes_write": gb("dram__bytes_write.sum"),
            "l1_hit_pct": g("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
            "l1_throughput_pct": g("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            "l2_throughput_pct": g("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "l2_bytes": gb("lts__t_bytes.sum"), "l1_bytes": gb("l1tex__t_bytes.sum"),
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "threads_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "warp_instructions": g("smsp__inst_executed.sum"),
            "pipe_alu_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "pipe_fma_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "pipe_lsu_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "ncu_duration_ms": (g("gpu__time_duration.sum") or 0) * to_ms.get(d.get("gpu__time_duration.sum", {}).get("unit", "ns"), 1e-6),
            "sm_cycles_elapsed": g("sm__cycles_elapsed.max"),
        }
        tp = root / "profiles" / "traffic.json"
        table = json.loads(tp.read_text()) if tp.exists() else {}
        table[workload] = entry
        tp.write_text(json.dumps(table, indent=1) + "\n")
        print(f"profiles/traffic.json[{workload}] <- {entry['kernel'][:50]} at {entry['commit']} ({entry['kernel_source_hash']})")
