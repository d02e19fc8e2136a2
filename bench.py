#!/usr/bin/env python3
"""bench.py -- encoded audio seconds per second on BASELINE.json config #2 (stereo 44.1 kHz, 128 kbps CBR,
10 000 synthetic sine-sweep frames, one stream per GPU), measured on B200 through libmp3b200.so.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched under torch.distributed.run, one rank/GPU)
  python bench.py --impl reference ...                   (the CPU restatement of lamejs on the host cores)

One "step" = one pass of the whole hot path (psy analysis -> scans -> masking -> filterbank+MDCT -> quantize+pack)
over the batch.  `value` is timed with the PCM already resident in HBM (CUDA events, max over ranks); `e2e` is the
same work through the host-buffer C-ABI call (pinned host memory, H2D and D2H inside the timed region).
Multi-GPU: streams are sharded statically over ranks (weak scaling: one C2 stream per GPU); the only collective is
the final NCCL gather of the encoded bytes to rank 0, inside the timed region.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "encoded audio seconds/sec (44.1kHz stereo 128kbps CBR)"
UNIT = "audio_s/s"
CH, SR, KBPS = 2, 44100, 128
FRAMES = 10000
N_SAMPLES = FRAMES * 1152
# algorithmic bytes per frame x channel of each kernel
ALGO_BYTES = {"filterbank_mdct": 6912, "psy": 3304, "quantize_pack": 5800}


def make_input():
    from synth import make_signal
    return make_signal("sweep", N_SAMPLES, SR)


def cpu_reference_run(threads, streams_per_thread=1, frames=FRAMES):
    """Times the CPU oracle (port of lamejs) with one independent stream per host thread."""
    import oracle_lib as O
    O.lib()
    l, r = make_input()
    l, r = l[: frames * 1152], r[: frames * 1152]
    results = [0] * threads

    def work(i):
        for _ in range(streams_per_thread):
            data, _, _ = O.encode_stream(CH, SR, KBPS, l, r)
            results[i] += len(data)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    audio_s = threads * streams_per_thread * (frames + 1) * 1152 / SR
    return audio_s / dt, dt


def host_cores():
    """Host threads this process can really use: the CPU affinity mask, capped by the cgroup CPU quota if there is one."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        try:
            quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota > 0:
                n = max(1, min(n, int(quota / period + 0.5)))
        except Exception:
            pass
    return n


def kernel_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/r01_kernel_traffic.json)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")))
        return d["kernels"][kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.reasons, self.proc = index, [], set(), None
        self.max_mhz = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            try:
                self.samples.append(float(p[0])); self.max_mhz = float(p[1])
                for n, v in zip(names, p[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = host_cores()
    frames = 2000   # bounded sample: cores x 2000 frames of the C2 sweep per step (~0.6 s of CPU work per core)
    for _ in range(args.warmup):
        cpu_reference_run(cores, 1, 500)
    vals, t_tot = [], 0.0
    for _ in range(args.steps):
        v, dt = cpu_reference_run(cores, 1, frames)
        vals.append(v); t_tot += dt
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": "BASELINE config #2: stereo 44.1kHz 128kbps CBR sine sweep, prefix of %d frames per host thread" % frames},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d threads x %d-frame prefix of the C2 sweep per step; lamejs itself cannot run (no JS engine): C++ restatement -O2" % (cores, frames)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams-per-gpu", type=int, default=1)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import lamejs_b200 as M

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = M.lib()
    assert L.mp3b200_set_device(local_rank) == 0

    # ---- workload: S streams per GPU, each the C2 sweep (rank-specific channel swap keeps shards distinct) ----
    S = args.streams_per_gpu
    l, r = make_input()
    frames = M.stream_frames(N_SAMPLES)
    nbytes = M.stream_bytes(CH, SR, KBPS, N_SAMPLES)
    host_pcm = torch.empty(S * 2 * N_SAMPLES, dtype=torch.int16).pin_memory()
    for s in range(S):
        a, b = (l, r) if (rank + s) % 2 == 0 else (r, l)
        host_pcm[(2 * s) * N_SAMPLES:(2 * s + 1) * N_SAMPLES] = torch.from_numpy(a)
        host_pcm[(2 * s + 1) * N_SAMPLES:(2 * s + 2) * N_SAMPLES] = torch.from_numpy(b)
    host_out = torch.empty(S * nbytes, dtype=torch.uint8).pin_memory()
    d_pcm = host_pcm.to(dev)
    d_out = torch.zeros(S * nbytes + 64, dtype=torch.uint8, device=dev)
    pcm_off = np.array([2 * s * N_SAMPLES for s in range(S)], dtype=np.int64)
    nsamp = np.full(S, N_SAMPLES, dtype=np.int64)
    out_off = np.array([s * nbytes for s in range(S)], dtype=np.int64)
    gather = [torch.empty_like(d_out) for _ in range(world)] if (world > 1 and rank == 0) else None
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_device():
        tm = M.encode_streams_device(CH, SR, KBPS, d_pcm.data_ptr(), pcm_off, nsamp, d_out.data_ptr(), out_off)
        if world > 1:
            dist.gather(d_out, gather, dst=0)      # final NCCL byte-gather (north_star)
        return tm

    lp = (ctypes.c_void_p * S)(*[host_pcm.data_ptr() + 2 * (2 * s) * N_SAMPLES for s in range(S)])
    rp = (ctypes.c_void_p * S)(*[host_pcm.data_ptr() + 2 * (2 * s + 1) * N_SAMPLES for s in range(S)])
    op = (ctypes.c_void_p * S)(*[host_out.data_ptr() + s * nbytes for s in range(S)])
    caps = np.full(S, nbytes, dtype=np.int64)
    got = np.zeros(S, dtype=np.int64)

    def step_e2e():
        rc = L.mp3b200_encode_streams(CH, SR, KBPS, S, lp, rp, nsamp.ctypes.data, op, caps.ctypes.data, got.ctypes.data)
        assert rc == 0, L.mp3b200_last_error()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.mp3b200_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms, ktimes = 0.0, np.zeros(8)
    for _ in range(args.steps):
        flush_buf.fill_(1)                      # evict L2 between timed iterations
        barrier()
        ev0.record()
        tm = step_device()
        ev1.record()
        torch.cuda.synchronize()
        total_ms += ev0.elapsed_time(ev1)
        ktimes += tm
    launches = L.mp3b200_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    audio_s = world * S * frames * 1152 / SR
    value = audio_s / (ms_per_step / 1000.0)

    # ---- end to end through the host-buffer C-ABI ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = audio_s / float(e2e_t.item())

    if rank == 0:
        ktimes /= args.steps
        units = S * frames * CH                               # frame x channel units per launch
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, peak_src = 6650.0, "fallback"
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        kern = {
            "psy": {"ms": float(ktimes[0] + ktimes[1] + ktimes[2]), "bytes": ALGO_BYTES["psy"]},
            "filterbank_mdct": {"ms": float(ktimes[3]), "bytes": ALGO_BYTES["filterbank_mdct"]},
            "quantize_pack": {"ms": float(ktimes[4] + ktimes[5]), "bytes": ALGO_BYTES["quantize_pack"]},
        }
        for k in kern.values():
            k["gbps"] = units * k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
            k["frac"] = k["gbps"] / peak if k["gbps"] else None
        dom = max(kern, key=lambda k: kern[k]["ms"])
        ncores = host_cores()
        cpu_v, cpu_dt = cpu_reference_run(ncores, 1, 5000)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "BASELINE config #2: stereo 44.1kHz 128kbps CBR, %d-frame sine sweep, %d stream(s) per GPU" % (FRAMES, S),
                       "l2": "256 MiB buffer written between timed iterations (L2 flush)", "realtime_factor": value,
                       "quantizer_passes": float(ktimes[7]), "bit_exact_vs": "oracle (tests/test_gpu_parity.py)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(S * 2 * N_SAMPLES * 2), "d2h_bytes_per_step": int(S * nbytes)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbps"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": kern[dom]["frac"], "traffic": kernel_traffic(dom),
                         "note": "algorithmic bytes/launch = %d B x %d frame-channels; exact-double arithmetic keeps every kernel FP64/latency bound (DESIGN.md)" % (kern[dom]["bytes"], units)},
            "kernels": kern,
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": ncores, "kind": "port",
                             "sample": "one 5000-frame prefix of the C2 sweep per host thread, %.1f s wall; lamejs restatement (C++ -O2), lamejs itself needs a JS engine" % cpu_dt},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
