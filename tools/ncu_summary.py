#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): per kernel duration, DRAM bytes, throughput, occupancy,
issue utilisation and the warp-stall mix.  Usage: python tools/ncu_summary.py report.ncu-rep [out.md]"""
import csv
import subprocess
import sys

WANT = [
    ("duration_us", "gpu__time_duration.sum"),
    ("dram_read_MB", "dram__bytes_read.sum"),
    ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct_of_peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs_per_thread", "launch__registers_per_thread"),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("l2_read_sectors_from_sm", "lts__t_sectors_srcunit_tex_op_read.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("dyn_smem", "launch__shared_mem_per_block_dynamic"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "not_selected", "lg_throttle", "mio_throttle",
          "math_pipe_throttle", "drain", "membar", "branch_resolving", "dispatch_stall", "no_instruction", "sleeping"]


def to_float(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(head)}
    out = ["# ncu summary of %s" % rep, ""]
    for r in rows[2:]:
        out.append("## %s" % r[col["Kernel Name"]][:110])
        for label, metric in WANT:
            if metric in col:
                v, u = r[col[metric]], units[col[metric]]
                f = to_float(v)
                if f is not None and u in ("byte", "Kbyte", "Mbyte", "Gbyte"):
                    f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
                    v, u = "%.3f" % f, "MB"
                out.append("- %s: %s %s" % (label, v, u))
        mix = []
        for s in STALLS:
            m = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            if m in col:
                f = to_float(r[col[m]])
                if f:
                    mix.append((f, s))
        out.append("- stalls (warps per issue-active cycle): " + ", ".join("%s %.2f" % (s, f) for f, s in sorted(mix, reverse=True)[:7]))
        out.append("")
    text = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
