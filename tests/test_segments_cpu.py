"""Host logic of one-stream-over-several-ranks (lamejs_b200/sharding.py encode_stream_segments): boundaries, warm-up,
state hand-over and the re-encode path, with a toy encoder that has the real interface and FIFO geometry (a frame completes
after frame*1152 + 224 samples) and a sequential state that forgets its start after a few frames.  World-size-2 and -3 gloo
runs on CPU; the CUDA encoder goes through the same functions in tests/test_gpu_segments.py."""
import hashlib
import os
import struct
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lamejs_b200 import sharding  # noqa: E402

FS = 1152


class ToyEncoder:
    """state' = f(state >> shift, frame samples): the start state is forgotten after ceil(16 / shift) frames."""

    def __init__(self, shift=4):
        self.shift, self.state, self.frames, self.buf, self.base, self.flushed = shift, 0xBEEF, 0, np.zeros(0, np.int16), 0, False

    def _emit(self, upto_fed, pad=0):
        out = b""
        data = np.concatenate([self.buf, np.zeros(pad, np.int16)])
        while (self.frames + 1) * FS + 224 <= upto_fed:
            lo = self.frames * FS - self.base
            frame = data[max(lo, 0):lo + FS]
            self.state = ((self.state >> self.shift) + int(np.abs(frame.astype(np.int64)).sum())) & 0xFFFF
            out += struct.pack("<IH", self.frames, self.state) + hashlib.sha1(frame.tobytes()).digest()[:2]
            self.frames += 1
        keep = max(0, self.frames * FS - 1104 - self.base)
        self.buf, self.base = self.buf[keep:], self.base + keep
        return out

    def encodeBuffer(self, left, right=None):
        self.buf = np.concatenate([self.buf, np.asarray(left, np.int16)])
        return self._emit(self.base + len(self.buf))

    def flush(self):
        if self.flushed:
            return b""
        self.flushed = True
        fed = self.base + len(self.buf)
        target = -(-(fed + 1152) // FS) * FS + 224
        return self._emit(target, pad=target - fed)

    def export_state(self):
        return struct.pack("<qqH", self.frames, self.base, self.state) + self.buf.tobytes()

    def import_state(self, blob):
        self.frames, self.base, self.state = struct.unpack("<qqH", blob[:18])
        self.buf = np.frombuffer(blob[18:], np.int16).copy()

    def seek(self, frame, left_hist, right_hist=None):
        assert self.frames == 0 and len(self.buf) == 0 and frame >= 1
        assert len(left_hist) == frame * FS + 224 - max(0, frame * FS - 1104)
        self.frames, self.base, self.buf = frame, max(0, frame * FS - 1104), np.asarray(left_hist, np.int16).copy()

    def close(self):
        pass


def whole(pcm, shift):
    e = ToyEncoder(shift)
    return e.encodeBuffer(pcm) + e.flush()


def signal(n, seed=3):
    return np.random.default_rng(seed).integers(-3000, 3000, n).astype(np.int16)


@pytest.mark.parametrize("nseg,warmup,shift,expect_redone", [(1, 8, 4, 0), (2, 8, 4, 0), (5, 8, 4, 0), (4, 2, 1, None), (3, 0, 4, None), (7, 40, 4, 0)])
def test_segments_local_equal_whole_stream(nseg, warmup, shift, expect_redone):
    pcm = signal(60 * FS + 517)
    got, redone = sharding.encode_stream_segments_local(lambda: ToyEncoder(shift), pcm, None, FS, nseg, warmup)
    assert got == whole(pcm, shift)
    if expect_redone is not None:
        assert redone == expect_redone          # the warm-up converged everywhere: nothing was encoded twice
    else:
        assert redone >= 1                      # a state that does not converge in `warmup` frames is caught and repaired


def test_segment_bounds_cover_the_stream_once():
    for n in (0, 100, 1375, 1376, 5000, 123456):
        for world in (1, 2, 3, 8):
            b, body = sharding.segment_bounds(n, FS, world)
            assert b[0][0] == 0 and b[-1][1] is None
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert all(hi is None or hi == 0 or hi * FS + 224 <= n for _, hi in b)
            assert body == max(0, (n - 224) // FS)


def test_short_streams_and_more_ranks_than_frames():
    for n in (0, 300, 2000, 5 * FS):
        pcm = signal(n, seed=n + 1)
        got, _ = sharding.encode_stream_segments_local(lambda: ToyEncoder(4), pcm, None, FS, 6, 3)
        assert got == whole(pcm, 4)


def _worker(rank, world, port, shift, warmup, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pcm = signal(50 * FS + 99)
    got, redone = sharding.encode_stream_segments(lambda: ToyEncoder(shift), pcm, None, FS, warmup=warmup)
    if rank == 0:
        q.put((got == whole(pcm, shift), redone))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shift,warmup", [(2, 4, 8), (3, 4, 8), (3, 1, 2)])
def test_segments_over_gloo(world, shift, warmup):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + 7 * world + shift
    procs = [ctx.Process(target=_worker, args=(r, world, port, shift, warmup, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, redone = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok
    assert (redone == 0) if shift == 4 else (redone >= 1)
