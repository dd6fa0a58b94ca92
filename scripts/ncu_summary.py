#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/:
   launches.csv (gpu__time_duration per launch)  -> profiles/<tag>_launches.md
   prof_fused.ncu-rep (--set full, fused kernel) -> profiles/<tag>_fused_full.md + profiles/fused_kernel_traffic.json
Usage: python scripts/ncu_summary.py r01 [gpurun_out]"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

lp = os.path.join(src, "launches.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = OrderedDict()
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        us = float(r[vi].replace(",", "")) / 1000.0
        agg.setdefault(name, []).append(us)
        seq.append((name, us))
    total = sum(sum(v) for k, v in agg.items() if "synth_fill" not in k)
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py` "
                "(cold-cache, serialised: compare shares, not absolutes). synth_fill (input generation) excluded from shares.\n\n")
        f.write("| kernel | launches | mean us | min us | max us | share of GPU time |\n|---|---|---|---|---|---|\n")
        for k, v in agg.items():
            sh = "-" if "synth_fill" in k else f"{100 * sum(v) / total:.1f}%"
            f.write(f"| `{k}` | {len(v)} | {sum(v) / len(v):.2f} | {min(v):.2f} | {max(v):.2f} | {sh} |\n")
        big = [(n, u) for n, u in seq if "fused" in n and u > 200]
        if big:
            step = [u for n, u in big]
            tail = [u for n, u in seq if ("finalize" in n or "decide" in n)]
            f.write(f"\nPer 2048-frame step: fused kernel {sum(step) / len(step):.1f} us; finalize+decide mean "
                    f"{sum(tail) / max(1, len(tail)):.1f} us per launch (overlapped with the next fused kernel on the library stream).\n")
    print("wrote", f"{tag}_launches.md")

rp = os.path.join(src, "prof_fused.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
            "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.max",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]
    traffic = []
    with open(os.path.join(out_dir, f"{tag}_fused_full.md"), "w") as f:
        f.write(f"# ncu --set full, fused_score_kernel ({tag})\n\n`ncu --set full --clock-control none --import-source on -k regex:fused_score` "
                "over `bench.py`; one column per captured launch (2048 frames of 1080p, 3 397 386 240 algorithmic bytes).\n\n")
        data = rows[2:]
        f.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n|---|---|" + "---|" * len(data) + "\n")
        for w in want:
            if w in H:
                i = H.index(w)
                f.write(f"| `{w}` | {U[i]} | " + " | ".join(r[i] for r in data) + " |\n")
        for r in data:
            def val(name):
                i = H.index(name)
                v = float(r[i].replace(",", ""))
                u = U[i].lower()
                return v * (1e9 if u.startswith("gbyte") else 1e6 if u.startswith("mbyte") else 1e3 if u.startswith("kbyte") else 1)
            traffic.append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    if traffic:
        json.dump({"dram_bytes_per_launch": sum(traffic) / len(traffic), "launches": len(traffic), "frames_per_launch": 2048,
                   "algorithmic_bytes_per_launch": 2048 * 1658880, "source": f"profiles/{tag}_fused_full.md"},
                  open(os.path.join(out_dir, "fused_kernel_traffic.json"), "w"), indent=1)
    print("wrote", f"{tag}_fused_full.md")
