"""ctypes binding of the CPU oracle (oracle/liblamejs_oracle.so).  Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg -- never by the product."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liblamejs_oracle.so")
TABLES_H = os.path.join(ORACLE_DIR, "lj_tables.h")

SBMAX_l, SBMAX_s, SFBMAX = 22, 13, 39

TRACE_DTYPE = np.dtype(
    [
        ("xr", np.float32, (2, 2, 576)),
        ("en_l", np.float32, (2, 2, SBMAX_l)),
        ("thm_l", np.float32, (2, 2, SBMAX_l)),
        ("en_s", np.float32, (2, 2, SBMAX_s, 3)),
        ("thm_s", np.float32, (2, 2, SBMAX_s, 3)),
        ("blocktype", np.int32, (2, 2)),
        ("ath_adjust", np.float64),
        ("l3_enc", np.int32, (2, 2, 576)),
        ("global_gain", np.int32, (2, 2)),
        ("part2_3_length", np.int32, (2, 2)),
        ("part2_length", np.int32, (2, 2)),
        ("big_values", np.int32, (2, 2)),
        ("count1", np.int32, (2, 2)),
        ("scalefac", np.int32, (2, 2, SFBMAX)),
        ("scalefac_compress", np.int32, (2, 2)),
        ("table_select", np.int32, (2, 2, 3)),
        ("region0", np.int32, (2, 2)),
        ("region1", np.int32, (2, 2)),
        ("preflag", np.int32, (2, 2)),
        ("scalefac_scale", np.int32, (2, 2)),
        ("count1table", np.int32, (2, 2)),
        ("subblock_gain", np.int32, (2, 2, 3)),
        ("scfsi", np.int32, (2, 4)),
        ("frame_bytes", np.int32),
        ("padding", np.int32),
        ("old_value_in", np.int32, (2,)),
        ("old_value_out", np.int32, (2,)),
        ("cur_step_in", np.int32, (2,)),
        ("cur_step_out", np.int32, (2,)),
    ],
    align=True,
)

_lib = None


def build():
    import fcntl
    with open(os.path.join(ORACLE_DIR, ".build.lock"), "w") as lk:   # pytest-xdist workers build once, in turn
        fcntl.flock(lk, fcntl.LOCK_EX)
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = ctypes.CDLL(ORACLE_SO)
        L.lj_create.restype = ctypes.c_void_p
        L.lj_create.argtypes = [ctypes.c_int] * 3
        L.lj_encode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.lj_flush.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.lj_destroy.argtypes = [ctypes.c_void_p]
        L.lj_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.lj_trace_count.argtypes = [ctypes.c_void_p]
        L.lj_query_out_samplerate.argtypes = [ctypes.c_int] * 3
        L.lj_enable_vbr_tag.argtypes = [ctypes.c_void_p]
        L.lj_enable_reservoir.argtypes = [ctypes.c_void_p]
        L.lj_enable_joint_stereo.argtypes = [ctypes.c_void_p]
        L.lj_enable_reservoir_integer_bytes.argtypes = [ctypes.c_void_p]
        L.lj_music_crc.argtypes = [ctypes.c_void_p]
        L.lj_bytes_written.argtypes = [ctypes.c_void_p]
        L.lj_bytes_written.restype = ctypes.c_longlong
        L.lj_vbr_frames.argtypes = [ctypes.c_void_p]
        L.lj_encoder_padding.argtypes = [ctypes.c_void_p]
        L.lj_get_lametag_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.lj_crc16.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]
        L.lj_wav_read_header.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]
        assert L.lj_trace_size() == TRACE_DTYPE.itemsize, (L.lj_trace_size(), TRACE_DTYPE.itemsize)
        _lib = L
    return _lib


def out_samplerate(channels, samplerate, kbps):
    """Output sample rate lame_init_params picks (Lame.js:285-364): != samplerate means lamejs resamples."""
    return lib().lj_query_out_samplerate(channels, samplerate, kbps)


class OracleEncoder:
    """Mirror of lamejs.Mp3Encoder (src/js/index.js:66-136) backed by the C++ restatement."""

    def __init__(self, channels, samplerate, kbps, trace_frames=0, write_vbr_tag=False, reservoir=False, joint_stereo=False):
        self.L = lib()
        self.h = self.L.lj_create(channels, samplerate, kbps)
        if not self.h:
            raise ValueError("unsupported configuration")
        self.channels = channels
        if joint_stereo:                   # gfp.mode = JOINT_STEREO (Mp3Encoder uses STEREO)
            assert self.L.lj_enable_joint_stereo(self.h) == 0
        if reservoir == "java":            # NOT lamejs: integer byte counts as in Reservoir.java (decodable streams)
            assert self.L.lj_enable_reservoir_integer_bytes(self.h) == 0
        elif reservoir:                    # gfp.disable_reservoir = false (Mp3Encoder never does this)
            assert self.L.lj_enable_reservoir(self.h) == 0
        self.tag_on = bool(write_vbr_tag) and self.L.lj_enable_vbr_tag(self.h) == 1   # gfp.bWriteVbrTag (InitVbrTag may refuse)
        self.trace = None
        if trace_frames:
            self.trace = np.zeros(trace_frames, dtype=TRACE_DTYPE)
            self.L.lj_set_trace(self.h, self.trace.ctypes.data, trace_frames)

    def encode_buffer(self, left, right=None):
        left = np.ascontiguousarray(left, dtype=np.int16)
        right = left if (right is None or self.channels == 1) else np.ascontiguousarray(right, dtype=np.int16)
        assert len(left) == len(right)
        cap = int(1.25 * len(left) + 7200)
        buf = np.empty(cap, dtype=np.uint8)
        k = self.L.lj_encode(self.h, left.ctypes.data, right.ctypes.data, len(left), buf.ctypes.data, cap)
        if k < 0:
            raise RuntimeError("lj_encode error %d" % k)
        return buf[:k].tobytes()

    def flush(self):
        cap = 7200 + 4 * 1440
        buf = np.empty(cap, dtype=np.uint8)
        k = self.L.lj_flush(self.h, buf.ctypes.data, cap)
        if k < 0:
            raise RuntimeError("lj_flush error %d" % k)
        return buf[:k].tobytes()

    def traces(self):
        return self.trace[: self.L.lj_trace_count(self.h)]

    # gfc.nMusicCRC / VBR_seek_table.nBytesWritten: kept by copy_buffer on every call (BitStream.js:924-935)
    def music_crc(self):
        return self.L.lj_music_crc(self.h)

    def bytes_written(self):
        return self.L.lj_bytes_written(self.h)

    def vbr_frames(self):
        return self.L.lj_vbr_frames(self.h)

    def encoder_padding(self):
        return self.L.lj_encoder_padding(self.h)

    def lametag_frame(self):
        """VBRTag.getLameTagFrame: b'' when the tag is off."""
        buf = np.zeros(2880, dtype=np.uint8)
        k = self.L.lj_get_lametag_frame(self.h, buf.ctypes.data, 2880)
        return buf[:k].tobytes()

    def close(self):
        if self.h:
            self.L.lj_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def encode_stream(channels, samplerate, kbps, left, right=None, chunk=None, trace_frames=0):
    """encodeBuffer(whole stream or chunks) + flush(); returns (bytes, per-call sizes, traces)."""
    enc = OracleEncoder(channels, samplerate, kbps, trace_frames)
    out, sizes = bytearray(), []
    n = len(left)
    step = chunk or max(n, 1)
    for i in range(0, n, step):
        b = enc.encode_buffer(left[i : i + step], None if right is None else right[i : i + step])
        sizes.append(len(b))
        out += b
    b = enc.flush()
    sizes.append(len(b))
    out += b
    tr = enc.traces().copy() if trace_frames else None
    enc.close()
    return bytes(out), sizes, tr


def crc16(data, crc=0):
    """CRC-16 of VBRTag.js:547-556 (reflected 0xA001, start value `crc`)."""
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    return lib().lj_crc16(a.ctypes.data if len(a) else None, len(a), crc)


def wav_read_header(data):
    """WavHeader.readHeader (index.js:154-193): dict, None (`return undefined`), or raises like the reference throws."""
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    out = (ctypes.c_longlong * 4)()
    rc = lib().lj_wav_read_header(a.ctypes.data if len(a) else None, len(a), out)
    if rc == 0:
        return None
    if rc == -1:
        raise ValueError("extended fmt chunk not implemented")
    if rc == -2:
        raise IndexError("read past the end of the buffer")
    return {"dataOffset": out[0], "dataLen": out[1], "channels": out[2], "sampleRate": out[3]}


def encode_stream_tagged(channels, samplerate, kbps, left, right=None, chunk=None):
    """Like encode_stream with gfp.bWriteVbrTag = true: the first call's bytes start with the placeholder frame
    (InitVbrTag); returns (bytes, per-call sizes, info) with info = tag frame, music CRC, byte / frame counts, padding."""
    enc = OracleEncoder(channels, samplerate, kbps, write_vbr_tag=True)
    out, sizes = bytearray(), []
    n = len(left)
    step = chunk or max(n, 1)
    for i in range(0, n, step):
        b = enc.encode_buffer(left[i : i + step], None if right is None else right[i : i + step])
        sizes.append(len(b))
        out += b
    b = enc.flush()
    sizes.append(len(b))
    out += b
    info = {"tag_on": enc.tag_on, "tag": enc.lametag_frame(), "music_crc": enc.music_crc(), "bytes_written": enc.bytes_written(),
            "frames": enc.vbr_frames(), "encoder_padding": enc.encoder_padding()}
    enc.close()
    return bytes(out), sizes, info
