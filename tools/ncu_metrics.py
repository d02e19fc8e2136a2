#!/usr/bin/env python3
"""Summarise `ncu --set full` captures (gpurun_out/prof_<kernel>.ncu-rep) into profiles/r02_kernel_metrics.json: per kernel
(first captured launch; all launches listed) duration, registers, DRAM bytes, issue-active %, threads per instruction
(warp execution efficiency), pipe utilisation and the stall breakdown.  bench.py reads the file for `roofline.limiter`.
usage: ncu_metrics.py [tag] kernel1 kernel2 ...   (needs ncu on PATH; runs in the build container, no GPU)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
kernels = sys.argv[2:]
M = {
    "duration_us": ("gpu__time_duration.sum", None),
    "registers": ("launch__registers_per_thread", None),
    "grid": ("launch__grid_size", None),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", None),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", None),
    "warp_instructions": ("smsp__inst_executed.sum", None),
    "thread_inst_per_inst": ("smsp__thread_inst_executed_per_inst_executed.ratio", None),
    "fp64_pipe_pct": ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", None),
    "xu_pipe_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", None),
    "dram_read_bytes": ("dram__bytes_read.sum", None),
    "dram_write_bytes": ("dram__bytes_write.sum", None),
    "dram_throughput_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", None),
}
STALLS = ["no_instruction", "wait", "short_scoreboard", "long_scoreboard", "barrier", "branch_resolving", "math_pipe_throttle",
          "not_selected", "dispatch_stall", "lg_throttle", "mio_throttle", "imc_miss", "tex_throttle", "drain", "membar", "sleeping"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
out = {"source": "ncu --set full --clock-control none -k regex:<kernel> -c 2 python tools/profile_run.py 10000 1 (C2: 10 001 stereo frames); tag " + tag,
       "kernels": {}}
for k in kernels:
    rep = os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % k)
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        def val(name):
            if name not in hdr:
                return None
            i = hdr.index(name)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                return None
            return v * UNIT.get(units[i], 1) if units[i] in UNIT else v
        d = {key: val(m) for key, (m, _) in M.items()}
        st = {s: val("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s) for s in STALLS}
        st = {s: round(v, 3) for s, v in st.items() if v}
        d["stall_warps_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1]))
        d["top_stall"] = next(iter(d["stall_warps_per_issue"]), None)
        if d["dram_read_bytes"] is not None and d["dram_write_bytes"] is not None:
            d["dram_bytes_per_launch"] = d["dram_read_bytes"] + d["dram_write_bytes"]
        launches.append(d)
    if launches:
        e = dict(launches[0])
        e["kernel"] = k
        e["captured"] = tag
        e["other_launches"] = [{kk: l[kk] for kk in ("duration_us", "issue_active_pct", "thread_inst_per_inst", "warp_instructions", "dram_bytes_per_launch")} for l in launches[1:]]
        out["kernels"][k] = e
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_kernel_metrics.json"), "w"), indent=1)
for k, e in out["kernels"].items():
    print("%-20s %8.1f us  regs %3d  issue %5.1f%%  thr/inst %5.2f  fp64 %5.1f%%  xu %5.1f%%  dram %6.1f MB  top stall %s" % (
        k, e["duration_us"], e["registers"], e["issue_active_pct"], e["thread_inst_per_inst"], e["fp64_pipe_pct"], e["xu_pipe_pct"],
        (e.get("dram_bytes_per_launch") or 0) / 1e6, e["top_stall"]))
