#!/usr/bin/env python3
"""bench.py -- encoded audio seconds per second of the lamejs Mp3Encoder hot path on B200 (libmp3b200.so).

  python bench.py [--config c2|c3|c4|c5] --gpus N --steps K --warmup W     (N>1: under torch.distributed.run, one rank/GPU)
  python bench.py --impl reference ...        (the CPU restatement of lamejs, oracle/, on the box's host cores)

Workloads (BASELINE.json configs); c2 is the headline the metric is quoted on:
  c2  stereo 44.1 kHz 128 kbps, one 10 000-frame sine sweep per GPU                       (weak scaling)
  c3  stereo 48 kHz 320 kbps, 100 streams x 1000 white-noise frames, sharded over ranks   (strong scaling)
  c4  mono 44.1 kHz 128 kbps, 1000 streams x 1000 frames octave noise, round-robin shards (strong scaling)
  c5  stereo 44.1 kHz 128 kbps CBR, 100 streams x 1000 transient-burst frames             (strong scaling)

One "step" = one pass of the whole hot path (psy analysis -> scans -> masking -> filterbank+MDCT -> quantizer kernels ->
bit packing) over the batch.  `value`: PCM already resident in HBM, CUDA events, max over ranks, NCCL byte gather to rank 0
inside the timed region when N>1.  `e2e`: the same work through the host-buffer C-ABI call, H2D and D2H (and the gather)
inside the timed region, from pinned and from pageable host memory.  `handle_api`: the lamejs call pattern.
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

UNIT = "audio_s/s"
# name -> (channels, samplerate, kbps, streams, frames per stream, signal kind, scaling, label)
CONFIGS = {
    "c2": (2, 44100, 128, 1, 10000, "sweep", "weak", "BASELINE config #2: stereo 44.1kHz 128kbps CBR, one 10000-frame sine sweep per GPU"),
    "c3": (2, 48000, 320, 100, 1000, "white", "strong", "BASELINE config #3: stereo 48kHz 320kbps CBR, 100 streams x 1000 white-noise frames"),
    "c4": (1, 44100, 128, 1000, 1000, "octave", "strong", "BASELINE config #4: mono 44.1kHz 128kbps, 1000 streams x 1000 frames (1M-frame batch) sharded round-robin + NCCL byte gather"),
    "c5": (2, 44100, 128, 100, 1000, "burst", "strong", "BASELINE config #5 input (transient bursts, block switching) under CBR 128k (VBR is dead code in lamejs): 100 streams x 1000 frames"),
}
DISTINCT = 64      # c3-c5: this many distinct seeded streams, cycled (generation time; every stream is still encoded)
# algorithmic bytes per frame x channel: groups of kernels, and the dominant single kernel
ALGO_BYTES = {"filterbank_mdct": 6912, "psy": 3304, "quantizer": 5800}
OUTER_BYTES_PER_GC = 2304 + 2304 + 1152 + 288 + 184 + 1152 + 288   # k_q_outer per granule-channel: xr, xrpow, lines, side info in; lines, side info out


def metric_name(cfg):
    ch, sr, kbps = CONFIGS[cfg][:3]
    return "encoded audio seconds/sec (%gkHz %s %dkbps CBR)" % (sr / 1000.0, "stereo" if ch == 2 else "mono", kbps)


def make_stream(cfg, j):
    """PCM of stream j of a config: (left, right) int16."""
    from synth import bursts, octave_hold, sweep, white
    ch, sr, kbps, S, frames, kind = CONFIGS[cfg][:6]
    n = frames * 1152
    if kind == "sweep":
        l, r = sweep(n, sr)
        return (l, r) if j % 2 == 0 else (r, l)          # rank-specific channel swap keeps shards distinct
    j = j % DISTINCT
    if kind == "white":
        return white(n, 0x5EED0003, offset=j << 32)
    if kind == "octave":
        x = octave_hold(n, 0x5EED0004 + 16 * j)
        return x, x
    return bursts(n, 0x5EED0005 + 64 * j)


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


def cpu_reference_run(cfg, threads, streams):
    """Times the CPU oracle (port of lamejs) on `streams` whole streams of the config spread over `threads` host threads."""
    import oracle_lib as O
    O.lib()
    ch, sr, kbps, S, frames = CONFIGS[cfg][:5]
    distinct = 1 if cfg == "c2" else min(streams, DISTINCT)
    sig = [make_stream(cfg, j) for j in range(distinct)]
    nxt, lock, done = [0], threading.Lock(), [0]

    def work():
        while True:
            with lock:
                j = nxt[0]
                nxt[0] += 1
            if j >= streams:
                return
            l, r = sig[j % distinct]
            data, _, _ = O.encode_stream(ch, sr, kbps, l, r if ch == 2 else None)
            with lock:
                done[0] += len(data)

    ths = [threading.Thread(target=work) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    audio_s = streams * (frames + 1) * 1152 / sr
    return audio_s / dt, dt


def kernel_metrics():
    """Limiter metrics of the committed ncu --set full captures (profiles/r02_kernel_metrics.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_kernel_metrics.json")))["kernels"]
    except Exception:
        return {}


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


def reference_sample(cfg, cores):
    """Bounded sample of the config for the CPU arm: whole streams, one per host thread (c2: the full 10 001-frame sweep)."""
    frames = CONFIGS[cfg][4]
    return cores, "%d host threads x one whole %d-frame stream of the config each (lamejs restatement oracle/, C++ -O2; real lamejs needs a JS " \
                  "engine: it runs in the build container under Qt's QJSEngine and pins the oracle, tools/jsrun/)" % (cores, frames + 1)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg = args.config
    cores = host_cores()
    streams, sample = reference_sample(cfg, cores)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(cfg, cores, max(1, cores // 4))
    vals, t_tot = [], 0.0
    for _ in range(args.steps):
        v, dt = cpu_reference_run(cfg, cores, streams)
        vals.append(v); t_tot += dt
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": metric_name(cfg), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000 * t_tot / args.steps, "higher_is_better": True, "scaling": CONFIGS[cfg][6], "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": CONFIGS[cfg][7], "cpu_sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def handle_api_numbers(M, L, dev_index):
    """The lamejs call pattern on the GPU: (a) ONE Mp3Encoder fed README-style 1152-sample encodeBuffer calls; (b) 256 live
    encoders advanced one 1152-sample call each per mp3b200_encode_batch launch."""
    from synth import make_signal
    out = {}
    l, r = make_signal("noise", 400 * 1152, 44100, 3)
    enc = M.Mp3Encoder(2, 44100, 128)
    for k in range(0, 20 * 1152, 1152):
        enc.encodeBuffer(l[k:k + 1152], r[k:k + 1152])
    t0 = time.perf_counter()
    calls = 0
    for k in range(20 * 1152, 320 * 1152, 1152):
        enc.encodeBuffer(l[k:k + 1152], r[k:k + 1152]); calls += 1
    dt = time.perf_counter() - t0
    enc.flush(); enc.close()
    out["single_encoder_1152_calls"] = {"value": calls * 1152 / 44100 / dt, "unit": UNIT, "us_per_call": 1e6 * dt / calls}
    N = 256
    encs = [M.Mp3Encoder(2, 44100, 128) for _ in range(N)]
    chunks = [(l[k:k + 1152], r[k:k + 1152]) for k in range(0, 60 * 1152, 1152)]
    for c in chunks[:10]:
        M.encode_batch(encs, [c[0]] * N, [c[1]] * N)
    t0 = time.perf_counter()
    for c in chunks[10:]:
        M.encode_batch(encs, [c[0]] * N, [c[1]] * N)
    dt = time.perf_counter() - t0
    M.flush_batch(encs)
    for e in encs:
        e.close()
    out["encode_batch_256_live_encoders"] = {"value": N * (len(chunks) - 10) * 1152 / 44100 / dt, "unit": UNIT, "ms_per_batch_call": 1e3 * dt / (len(chunks) - 10)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--streams-per-gpu", type=int, default=1, help="c2 only: sweep streams per GPU")
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
    from lamejs_b200.sharding import shard_streams

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = M.lib()
    assert L.mp3b200_set_device(local_rank) == 0

    cfg = args.config
    CH, SR, KBPS, S_total, FRAMES, kind, scaling, label = CONFIGS[cfg]
    N_SAMPLES = FRAMES * 1152
    # ---- this rank's streams: c2 = S sweeps per GPU (weak); others = round-robin shard of the global list (strong) ----
    if cfg == "c2":
        mine = [rank + s for s in range(args.streams_per_gpu)]
        S_global = world * args.streams_per_gpu
    else:
        mine = shard_streams(S_total, world, rank)
        S_global = S_total
    S = len(mine)
    frames = M.stream_frames(N_SAMPLES)
    nbytes = M.stream_bytes(CH, SR, KBPS, N_SAMPLES)
    per = CH * N_SAMPLES
    host_pcm = torch.empty(max(S, 1) * per, dtype=torch.int16).pin_memory()
    cache = {}
    for i, j in enumerate(mine):
        key = j if cfg == "c2" else j % DISTINCT
        if key not in cache:
            cache[key] = make_stream(cfg, j)
        a, b = cache[key]
        host_pcm[i * per:i * per + N_SAMPLES] = torch.from_numpy(a)
        if CH == 2:
            host_pcm[i * per + N_SAMPLES:(i + 1) * per] = torch.from_numpy(b)
    del cache
    host_out = torch.empty(max(S, 1) * nbytes, dtype=torch.uint8).pin_memory()
    d_pcm = host_pcm.to(dev)
    d_out = torch.zeros(max(S, 1) * nbytes + 64, dtype=torch.uint8, device=dev)
    pcm_off = np.array([s * per for s in range(S)], dtype=np.int64)
    nsamp = np.full(S, N_SAMPLES, dtype=np.int64)
    out_off = np.array([s * nbytes for s in range(S)], dtype=np.int64)
    # byte gather: every rank contributes a buffer of the largest shard's size (closed-form offsets inside)
    S_max = (S_global + world - 1) // world if cfg != "c2" else args.streams_per_gpu
    send = torch.zeros(S_max * nbytes, dtype=torch.uint8, device=dev) if world > 1 else None
    gather = [torch.empty_like(send) for _ in range(world)] if (world > 1 and rank == 0) else None
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    ev_c0, ev_c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def gather_bytes():
        """final NCCL byte gather (north_star): returns the device-timed collective ms of this rank"""
        if world == 1:
            return 0.0
        send[:S * nbytes].copy_(d_out[:S * nbytes])
        ev_c0.record()
        dist.gather(send, gather, dst=0)
        ev_c1.record()
        return None

    def step_device():
        tm = M.encode_streams_device(CH, SR, KBPS, d_pcm.data_ptr(), pcm_off, nsamp, d_out.data_ptr(), out_off)
        gather_bytes()
        return tm

    def host_ptrs(buf):
        lp = (ctypes.c_void_p * max(S, 1))(*[buf.data_ptr() + 2 * (s * per) for s in range(S)])
        rp = (ctypes.c_void_p * max(S, 1))(*[buf.data_ptr() + 2 * (s * per + (N_SAMPLES if CH == 2 else 0)) for s in range(S)])
        return lp, rp

    op = (ctypes.c_void_p * max(S, 1))(*[host_out.data_ptr() + s * nbytes for s in range(S)])
    caps = np.full(S, nbytes, dtype=np.int64)
    got = np.zeros(S, dtype=np.int64)

    def step_e2e(ptrs):
        rc = L.mp3b200_encode_streams(CH, SR, KBPS, S, ptrs[0], ptrs[1], nsamp.ctypes.data, op, caps.ctypes.data, got.ctypes.data)
        assert rc == 0, L.mp3b200_last_error()
        if world > 1:                              # the gather is part of the job at N>1: results travel from the host copy
            send[:S * nbytes].copy_(host_out[:S * nbytes], non_blocking=True)
            dist.gather(send, gather, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                         # nvidia-smi needs ~0.3 s to deliver its first sample: start before the warm-up
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    launches0 = L.mp3b200_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms, coll_ms, ktimes = 0.0, 0.0, np.zeros(16)
    for _ in range(args.steps):
        flush_buf.fill_(1)                      # evict L2 between timed iterations
        barrier()
        ev0.record()
        tm = step_device()
        ev1.record()
        torch.cuda.synchronize()
        total_ms += ev0.elapsed_time(ev1)
        if world > 1:
            coll_ms += ev_c0.elapsed_time(ev_c1)
        ktimes += tm
    launches = L.mp3b200_launch_count() - launches0
    own_ms = total_ms / args.steps
    t = torch.tensor([total_ms, -total_ms, coll_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t[0].item()) / args.steps
    # A C2 step lasts ~6 ms and nvidia-smi ticks every 100 ms: keep the same load running (untimed, the same count on every
    # rank because step_device contains the gather) so that the clock sampler sees several ticks under this load.
    for _ in range(min(400, int(900.0 / max(ms_per_step, 1e-3))) if ms_per_step < 300 else 0):
        step_device()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    skew_ms = (float(t[0].item()) + float(t[1].item())) / args.steps        # slowest rank minus fastest rank
    collective_ms = float(t[2].item()) / args.steps
    audio_s = S_global * frames * 1152 / SR
    value = audio_s / (ms_per_step / 1000.0)

    # ---- end to end through the host-buffer C-ABI: pinned and pageable caller memory ----
    def time_e2e(ptrs):
        for _ in range(2):
            step_e2e(ptrs)
        barrier()
        t0 = time.perf_counter()
        n = max(2, min(args.steps, 5))
        for _ in range(n):
            step_e2e(ptrs)
        torch.cuda.synchronize()
        tt = torch.tensor([(time.perf_counter() - t0) / n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return audio_s / float(tt.item())

    e2e_pinned = time_e2e(host_ptrs(host_pcm))
    pageable = torch.empty_like(host_pcm, pin_memory=False).copy_(host_pcm)
    e2e_pageable = time_e2e(host_ptrs(pageable))
    del pageable

    if rank == 0:
        ktimes /= args.steps
        units = S * frames * CH                               # frame x channel units per launch on this rank
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        km = kernel_metrics()
        kern = {
            # the subband analysis (k_subband_analysis) runs beside the per-stream scan inside ktimes[1]; ktimes[3] is k_mdct:
            # psy and filterbank are reported as one group
            "psy+filterbank_mdct": {"ms": float(ktimes[0] + ktimes[1] + ktimes[2] + ktimes[3]),
                                    "bytes_per_unit": ALGO_BYTES["psy"] + ALGO_BYTES["filterbank_mdct"]},
            "quantizer": {"ms": float(ktimes[4] + ktimes[5]), "bytes_per_unit": ALGO_BYTES["quantizer"],
                          "by_kernel_ms": {"k_q_prepare": float(ktimes[8]), "k_q_search": float(ktimes[9]), "k_q_outer": float(ktimes[10]),
                                           "k_q_finish": float(ktimes[11]), "k_q_pack": float(ktimes[12]), "revalidation_passes": float(ktimes[5])}},
        }
        for k in kern.values():
            k["gbps"] = units * k["bytes_per_unit"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
            k["frac"] = k["gbps"] / peak if k["gbps"] else None
        # dominant single kernel: the rate loop k_q_outer, launched once per granule (2 launches per step)
        gcs_per_launch = S * frames * CH
        outer_ms = float(ktimes[10]) / 2.0
        outer_gbps = gcs_per_launch * OUTER_BYTES_PER_GC / (outer_ms * 1e-3) / 1e9 if outer_ms > 0 else None
        mo = km.get("k_q_outer", {})
        ncores = host_cores()
        cpu_streams, cpu_sample = reference_sample(cfg, ncores)
        cpu_v, cpu_dt = cpu_reference_run(cfg, ncores, cpu_streams)
        try:
            handle_api = handle_api_numbers(M, L, local_rank) if world == 1 else None
        except Exception as e:   # noqa: BLE001
            handle_api = {"error": repr(e)}
        try:      # container step: CRC-16 of the bytes the packer left in HBM; outside the timed region
            if world == 1:
                best = None
                for _ in range(5):
                    _crc, ms = M.debug_music_crc(d_out.data_ptr(), out_off, np.full(S, nbytes, dtype=np.int64), timed=True)
                    best = ms if best is None else min(best, ms)
                tag = {"kernel": "k_music_crc", "ms": best, "bytes": int(S * nbytes), "gbps": S * nbytes / (best * 1e-3) / 1e9,
                       "frac": S * nbytes / (best * 1e-3) / 1e9 / peak,
                       "note": "range upload + clear + launch + 4 B/stream read-back, CUDA events; algorithmic bytes = 1 B read per output byte; "
                               "replaces lamejs's per-byte table CRC in copy_buffer (BitStream.js:924-928)"}
            else:
                tag = None
        except Exception as e:   # noqa: BLE001
            tag = {"error": repr(e)}
        line = {
            "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": label, "streams_total": S_global, "streams_on_rank0": S, "frames_per_stream": frames,
                       "l2": "256 MiB buffer written between timed iterations (L2 flush)", "realtime_factor": value,
                       "quantizer_passes": float(ktimes[7]),
                       "bit_exact_vs": "oracle/ on every GPU test; oracle/ == real lamejs (QJSEngine) on 306 committed fixtures (tests/test_lamejs_pin.py)"},
            "e2e": {"value": e2e_pinned, "unit": UNIT, "h2d_bytes_per_step": int(S * per * 2), "d2h_bytes_per_step": int(S * nbytes),
                    "host_memory": "pinned", "pageable_value": e2e_pageable, "includes_gather": world > 1},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_q_outer", "achieved": outer_gbps, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": outer_gbps / peak if outer_gbps else None, "traffic": mo.get("dram_bytes_per_launch"),
                         "launch_ms": outer_ms,
                         "note": "algorithmic bytes/launch = %d B x %d granule-channels (DESIGN.md 4); the rate loop is instruction-issue / latency bound, "
                                 "not bandwidth bound: see limiter" % (OUTER_BYTES_PER_GC, gcs_per_launch),
                         "limiter": {"issue_active_pct": mo.get("issue_active_pct"), "warp_exec_eff_threads_per_inst": mo.get("thread_inst_per_inst"),
                                     "fp64_pipe_pct": mo.get("fp64_pipe_pct"), "xu_pipe_pct": mo.get("xu_pipe_pct"),
                                     "warps_active_pct": mo.get("warps_active_pct"), "top_stall": mo.get("top_stall"),
                                     "source": "profiles/r02_kernel_metrics.json (ncu --set full)"}},
            "kernels": kern,
            "multi_gpu": {"collective_ms": collective_ms, "skew_ms": skew_ms, "rank0_step_ms": own_ms} if world > 1 else None,
            "handle_api": handle_api,
            "tag": tag,
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": ncores, "kind": "port", "sample": cpu_sample + ", %.1f s wall" % cpu_dt},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
