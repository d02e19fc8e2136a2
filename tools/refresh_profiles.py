#!/usr/bin/env python3
"""Copy the round-end measurement artefacts from gpurun_out/ into profiles/ under a tag and refresh
profiles/r01_kernel_traffic.json from whatever --set full captures are present.  usage: refresh_profiles.py r01e"""
import collections, csv, json, os, shutil, subprocess, sys
tag = sys.argv[1]
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
for src, dst in (("bench.json", "bench.json"), ("bench_ref.json", "bench_reference.json"), ("launches.csv", "launches_bench.csv"), ("bench_n2.json", "bench_n2.json")):
    if os.path.exists(os.path.join(G, src)): shutil.copy(os.path.join(G, src), os.path.join(P, "%s_%s" % (tag, dst)))
tj = os.path.join(P, "r01_kernel_traffic.json")
out = json.load(open(tj))
names = {"k_quantize_pack": "quantize_pack", "k_psy_analysis": "psy", "k_subband_analysis": "subband_analysis", "k_mdct": "mdct", "k_psy_masking": "psy_masking"}
for k, n in names.items():
    rep = os.path.join(G, "prof_%s.ncu-rep" % k)
    if not os.path.exists(rep) or (len(sys.argv) > 2 and k not in sys.argv[2:]): continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = dict(zip(rows[0], rows[-1])); u = dict(zip(rows[0], rows[1]))
    mb = lambda key: float(d[key]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u[key]]
    rd, wr = mb("dram__bytes_read.sum"), mb("dram__bytes_write.sum")
    out["kernels"][n] = {"kernel": k, "captured": tag, "dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                         "duration_under_ncu": d["gpu__time_duration.sum"] + " " + u["gpu__time_duration.sum"],
                         "registers": int(d["launch__registers_per_thread"]),
                         "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                         "warps_active_pct": float(d["sm__warps_active.avg.pct_of_peak_sustained_active"]),
                         "xu_pipe_pct": float(d["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]),
                         "fp64_pipe_pct": float(d["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]),
                         "warp_instructions": int(float(d["smsp__inst_executed.sum"]))}
    print(n, out["kernels"][n])
out["source"] = "ncu --set full --clock-control none -k regex:<kernel> -c 1 python tools/profile_run.py 10000 1 (C2: 10 001 stereo frames); see each entry's `captured` tag"
json.dump(out, open(tj, "w"), indent=1)
lines = open(os.path.join(G, "launches.csv")).read().splitlines()
i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(lines[i:]):
    n = r["Kernel Name"].split("(")[0].replace("void ", "")[:48]
    agg[n][0] += 1; agg[n][1] += float(r["Metric Value"]) / 1000
tot = sum(v[1] for v in agg.values())
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print("%-50s n=%3d total %9.1f us share %5.1f%% avg %8.1f us" % (n, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
d = json.loads(open(os.path.join(G, "bench.json")).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches")}, {k: round(v["ms"], 3) for k, v in d["kernels"].items()})
