"""lamejs_b200 -- host-side mirror of the lamejs `Mp3Encoder` API over the B200-native C-ABI library.

    from lamejs_b200 import Mp3Encoder
    enc = Mp3Encoder(2, 44100, 128)           # new lamejs.Mp3Encoder(channels, sampleRate, kbps)
    mp3 = enc.encodeBuffer(left, right)        # Int16 arrays -> bytes (frames completed by this call)
    mp3 += enc.flush()

(reference: zhuker/lamejs src/js/index.js:66-136).  All computation happens in libmp3b200.so on the GPU; the
module raises if the library or a CUDA device is missing -- there is no CPU fallback.
"""
from .encoder import Mp3Encoder, WavHeader, id3v1_tag, id3v2_tag, ID3_ADD_V2, ID3_V1_ONLY, ID3_V2_ONLY, ID3_SPACE_V1, ID3_PAD_V2, lametag_size, lametag_build, get_vbr_tag, crc16_combine, encode_streams_tagged, debug_music_crc, encode_batch, flush_batch, encode_streams, encode_streams_device, debug_stages, lib, stream_bytes, stream_frames, granules_per_frame, Mp3B200Error  # noqa: F401
