/* mp3b200.h -- C ABI of the B200-native batch MP3 encoder (libmp3b200.so).
 *
 * Drop-in boundary for the lamejs `Mp3Encoder` hot path.  Each entry point names the reference
 * interface it replaces (paths relative to zhuker/lamejs @ 582bbba).  Plain pointers and sizes only;
 * no torch / CUDA types cross this boundary.  All work runs on the CUDA device selected with
 * mp3b200_set_device() (default: device 0); there is NO CPU fallback -- every call fails with
 * MP3B200_ERR_CUDA when no usable sm_100 device is present.
 *
 * Error codes mirror lamejs/LAME where one exists (src/js/Lame.js:1045-1061,1494; BitStream.js:916-919).
 */
#ifndef MP3B200_H
#define MP3B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP3B200_OK 0
#define MP3B200_ERR_CONFIG (-1)   /* unsupported channels/samplerate/bitrate: lame_init_params returns -1 */
#define MP3B200_ERR_BUFFER (-1)   /* output buffer too small: copy_buffer returns -1 (BitStream.js:916-919) */
#define MP3B200_ERR_HANDLE (-3)   /* bad handle: lame_encode_buffer returns -3 (Lame.js:1494) */
#define MP3B200_ERR_CUDA (-100)   /* CUDA failure or no device (no reference equivalent) */

typedef struct mp3b200_encoder mp3b200_encoder;

/* Select the CUDA device used by subsequently created encoders / batch calls (one process per GPU). */
int mp3b200_set_device(int device);

/* Replaces `new lamejs.Mp3Encoder(channels, samplerate, kbps)` (src/js/index.js:66-115).
 * channels 1|2; samplerate 8000|11025|12000|16000|22050|24000 (MPEG-2 / 2.5 LSF: one granule, 576-sample frames) or
 * 32000|44100|48000 (MPEG-1); kbps snapped to the nearest legal rate of that MPEG version like FindNearestBitrate
 * (src/js/Lame.js:408-423).  Configurations for which lamejs would RESAMPLE (lame_init_params picks out_samplerate !=
 * in_samplerate from the bitrate's low-pass, Lame.js:285-364 -- e.g. 44.1 kHz stereo below 112 kbps) return
 * MP3B200_ERR_CONFIG: the CPU oracle models lamejs's resampler byte for byte (all 306 reference fixtures), but lamejs
 * reads its input with fractional / out-of-range typed-array indices there (whole-buffer calls and every flush of a
 * non-integer rate ratio put NaN samples into the stream, oracle/lj_init.cpp fill_buffer_resample), so that row is not
 * offered as a drop-in. */
int mp3b200_create(int channels, int samplerate, int kbps, mp3b200_encoder** out);

/* Replaces `encodeBuffer(left, right)` (src/js/index.js:117-130 -> Lame.js:1490-1667).  `right` may be NULL
 * for mono.  Writes the bytes of the frames completed by this call (possibly 0) into `out` and returns their
 * count; lamejs sizes its buffer as trunc(1.25*n + 7200).  Buffers are borrowed for the call only. */
int mp3b200_encode(mp3b200_encoder* h, const int16_t* left, const int16_t* right, int nsamples, uint8_t* out, int cap);

/* Replaces `flush()` (src/js/index.js:132-135 -> Lame.js:1381-1488): encodes the buffered tail padded with
 * zeros; a second flush returns 0; the encoder stays usable. */
int mp3b200_flush(mp3b200_encoder* h, uint8_t* out, int cap);

/* Replaces garbage collection of the Mp3Encoder object. */
void mp3b200_destroy(mp3b200_encoder* h);

/* Batched encodeBuffer / flush over N live encoders of one configuration (SURVEY.md 8(b) "Batch"): equivalent to
 * calling mp3b200_encode / mp3b200_flush on every handle in turn, but the frames all handles complete in this call
 * are encoded by ONE pipeline launch (handle i = stream i).  left[i]/right[i]: nsamples[i] Int16 (right or right[i]
 * NULL: mono / duplicate left); out[i] receives out_bytes[i] bytes (cap[i] >= trunc(1.25 * nsamples[i] + 7200) like
 * lamejs, 0 = unchecked); out_bytes[i] < 0 reports a per-handle error (NULL handle -3, buffer too small -1).
 * A JS caller loops over its encoders instead (worker-example/worker.js, worker-realtime.js:46). */
int mp3b200_encode_batch(mp3b200_encoder* const* handles, const int16_t* const* left, const int16_t* const* right,
                         const int* nsamples, uint8_t* const* out, const int* cap, int nstreams, int* out_bytes);
int mp3b200_flush_batch(mp3b200_encoder* const* handles, uint8_t* const* out, const int* cap, int nstreams, int* out_bytes);

/* ---- encoder state: checkpoint / resume, and one stream cut into segments ----------------------------
 * lamejs keeps an Mp3Encoder's state in JS objects (gfc.*: ATH adjust, block-type FSM, OldValue / CurrentStep, the previous
 * granule's masking, mfbuf); a JS caller checkpoints by keeping the object alive.  Here the state is a blob:
 *   export_state  writes everything the handle carries between calls (scalars, the previous unit's masking row from device
 *                 memory, the PCM tail later frames still read, FIFO accounting); buf == NULL returns the size.  Two handles
 *                 that encoded the same frames from the same samples export identical bytes.
 *   import_state  makes a handle of the same configuration continue exactly where the exporting one stood.
 *   seek          positions a FRESH handle at frame `frame` >= 1 with the sequential state of a stream start: the start of
 *                 a warm-up.  A segment encoder seeks W frames before its first frame, encodes them (discarding the bytes)
 *                 and compares its state with the predecessor segment's exported end state: equal blobs prove its frames are
 *                 the ones a single encoder would produce; otherwise it imports the predecessor's state and encodes again
 *                 (lamejs_b200/sharding.py encode_stream_segments).  `hist`: the stream samples
 *                 [max(0, frame*framesize - 1104), frame*framesize + 224), framesize = 576 * granules per frame; feeding
 *                 continues with sample frame*framesize + 224.
 * Return MP3B200_OK / bytes written, or a negative error (wrong configuration -2, buffer -1, handle -3). */
int mp3b200_export_state(mp3b200_encoder* h, void* buf, int cap);
int mp3b200_import_state(mp3b200_encoder* h, const void* buf, int len);
int mp3b200_seek(mp3b200_encoder* h, int64_t frame, const int16_t* left_hist, const int16_t* right_hist, int nhist);

/* ---- batch extension (same semantics, many independent streams per launch sequence) -------------------
 * Equivalent to, for each stream s: e = new Mp3Encoder(ch, sr, kbps); bytes = e.encodeBuffer(L_s, R_s) ++
 * e.flush().  This is the throughput path (a JS caller would loop over encoders, worker-example/worker.js). */

/* Number of bytes / frames that stream of `nsamples` per channel produces (closed form: CBR, no reservoir). */
int64_t mp3b200_stream_bytes(int channels, int samplerate, int kbps, int64_t nsamples);
int64_t mp3b200_stream_frames(int64_t nsamples);                 /* MPEG-1 configurations (1152-sample frames) */
/* any accepted configuration (MPEG-2 / 2.5 frames carry 576 samples); -1 if the configuration is rejected */
int64_t mp3b200_stream_frames_cfg(int channels, int samplerate, int kbps, int64_t nsamples);
/* granules per frame: 2 (MPEG-1: 32 / 44.1 / 48 kHz) or 1 (MPEG-2 / 2.5: 8 .. 24 kHz); -1 if rejected */
int mp3b200_granules_per_frame(int channels, int samplerate, int kbps);

/* Host buffers.  left[s]/right[s]: nsamples[s] Int16 each (right NULL or ignored for mono).  out[s] receives
 * out_bytes[s] = mp3b200_stream_bytes(...) bytes (cap[s] must be >= that).  Returns 0 or a negative error. */
int mp3b200_encode_streams(int channels, int samplerate, int kbps, int nstreams, const int16_t* const* left,
                           const int16_t* const* right, const int64_t* nsamples, uint8_t* const* out,
                           const int64_t* cap, int64_t* out_bytes);

/* Device-resident variant for benchmarking kernel throughput: d_pcm is ONE device allocation holding, per
 * stream s, nsamples[s] Int16 of the left channel at sample offset pcm_off[s] and (stereo) the right channel at
 * pcm_off[s] + nsamples[s].  d_out is a device buffer; stream s is written at out_off[s].  `timings_ms`
 * (optional, 16 floats) receives per-kernel CUDA-event times in ms: [0] psy analysis + loudness, [1] attack pre-pass + per-stream
 * scan with the subband analysis (polyphase filterbank) running beside it, [2] masking, [3] MDCT, [4] quantizer first pass (all of its kernels), [5] re-validation passes, [6] total, [7] number of
 * quantizer passes; first pass by kernel: [8] k_q_prepare, [9] k_q_search (gr0 + gr1), [10] k_q_outer (gr0 + gr1),
 * [11] k_q_finish (gr0 + gr1), [12] k_q_pack, [13] the re-validation folded into the first pass (verify + repaired
 * searches / rate loops of the few frames whose speculated start did not stand); [14..15] reserved (0).
 * The call runs on a stream of its own that first waits for work already queued on the legacy default stream (where torch /
 * plain CUDA callers produced d_pcm) and returns after that stream has drained. */
int mp3b200_encode_streams_device(int channels, int samplerate, int kbps, int nstreams, const int16_t* d_pcm,
                                  const int64_t* pcm_off, const int64_t* nsamples, uint8_t* d_out,
                                  const int64_t* out_off, float* timings_ms);

/* ---- container / metadata step after the path ------------------------------------------------------
 * lamejs carries LAME's Xing / Info / LAME tag writer (src/js/VBRTag.js) and keeps its two inputs up to date on every
 * Mp3Encoder call -- gfc.nMusicCRC and VBR_seek_table.nBytesWritten (BitStream.js:924-935) -- but switches the writer off
 * (index.js:107: gfp.bWriteVbrTag = false).  Here the writer can be switched on.  Semantics are LAME's (VBRTag.java, which
 * VBRTag.js transliterates): with the tag on, the first bytes an encoder hands out are an all-zero frame of the stream's own
 * bitrate (InitVbrTag); after flush() mp3b200_get_lametag_frame returns the finished frame, which the caller writes over
 * that placeholder (lame_get_lametag_frame / putVbrTag).  The CRC-16 over all audio bytes is computed on the GPU
 * (k_music_crc, lamejs_b200/csrc/k_tag.cuh) from the bytes where the packer left them.
 *   set_write_vbr_tag   gfp.bWriteVbrTag, before the first sample.  Returns 1 (on), 0 (off: asked to, or InitVbrTag refused
 *                       because the frame cannot hold side info + 156 bytes, VBRTag.js:508-513), negative on error.
 *   get_lametag_frame   VBRTag.getLameTagFrame (VBRTag.js:829-923): 0 when the tag is off or no frame has been encoded;
 *                       the size needed when `cap` is too small (or buf NULL); else writes that many bytes and returns it.
 *   music_crc / bytes_written   gfc.nMusicCRC / nBytesWritten so far (-1 with the tag off: the accumulators are idle then).
 *   lametag_size        the tag frame's size for a configuration (0: does not fit; negative: rejected configuration).
 *   lametag_build       the same frame from numbers instead of a handle (pure host arithmetic, no device needed): for callers
 *                       that encode one stream in segments and combine counts and CRCs themselves.
 *   encode_streams_tagged   mp3b200_encode_streams with the tag on: out[s] = finished tag frame ++ audio frames
 *                       (cap[s] >= mp3b200_stream_bytes + mp3b200_lametag_size); one k_music_crc launch for the batch. */
int mp3b200_set_write_vbr_tag(mp3b200_encoder* h, int on);
int mp3b200_get_lametag_frame(mp3b200_encoder* h, uint8_t* buf, int cap);
int mp3b200_music_crc(mp3b200_encoder* h);
int64_t mp3b200_bytes_written(mp3b200_encoder* h);
int mp3b200_lametag_size(int channels, int samplerate, int kbps);
int mp3b200_lametag_build(int channels, int samplerate, int kbps, int64_t nframes, int64_t music_bytes, int music_crc,
                          int encoder_padding, uint8_t* buf, int cap);
int mp3b200_encode_streams_tagged(int channels, int samplerate, int kbps, int nstreams, const int16_t* const* left,
                                  const int16_t* const* right, const int64_t* nsamples, uint8_t* const* out,
                                  const int64_t* cap, int64_t* out_bytes);

/* The rest of VBRTag.js's surface:
 *   put_vbr_tag     putVbrTag (VBRTag.js:937-965) on a stream held in memory: writes the finished frame over the placeholder,
 *                   behind an ID3v2 tag if the stream starts with one (skipId3v2; the port's inverted test is not reproduced).
 *                   0 ok / nothing to write, -1 like the reference (no frame counted yet, empty or too short stream).
 *   get_vbr_tag     getVbrTag (VBRTag.js:375-470): the reader side -- frame / byte counts, seek table, quality, encoder delay
 *                   and padding from the first frame of any Xing / Info tagged stream.  1 ok, 0 no tag (reference: null), -2
 *                   buffer too short.  Pure host code.
 *   crc16_combine   CRC-16 of A || B from crc(A), crc(B) and |B| (the rule k_music_crc and the handles use): lets callers that
 *                   encode one stream as segments on several GPUs put one tag on the joined stream. */
typedef struct mp3b200_vbr_tag_data {
  int32_t h_id, samprate, flags, frames, bytes, vbr_scale, headersize, enc_delay, enc_padding;
  uint8_t toc[100];
} mp3b200_vbr_tag_data;
int mp3b200_put_vbr_tag(mp3b200_encoder* h, uint8_t* stream, int64_t len);
int mp3b200_get_vbr_tag(const uint8_t* frame, int64_t len, mp3b200_vbr_tag_data* out);
int mp3b200_crc16_combine(int crc_a, int crc_b, int64_t len_b);

/* ID3 tags.  lamejs carries only a stub (index.js:56-64) and switches the automatic tags off (index.js:109);
 * the writer is the Java original's, src/main/java/mp3/ID3Tag.java: lame_get_id3v2_tag :961-1102 (ID3v2.3, ISO-8859-1 text
 * frames TSSE TIT2 TPE1 TALB TYER COMM TRCK TCON TLEN, optional padding), lame_get_id3v1_tag :1141-1189 (128 bytes, v1.1 when a
 * track is set).  Fields are Latin-1 strings, NULL or "" = not set (the id3tag_set_* calls, applied in the order of the struct);
 * `genre` is a number 0..147 or a name of ID3Tag.java:56-89.  Both return the tag size written, the size needed when `cap` is
 * smaller, 0 when the reference writes no such tag (v2: nothing asks for it and every field fits version 1; v1: nothing set, or
 * V2_ONLY), negative for a malformed year / track / genre.  The version 2 tag goes in front of the stream (and of the Info / LAME
 * tag frame), the version 1 tag behind it.  Pure host code. */
#define MP3B200_ID3_ADD_V2 2     /* id3tag_add_v2 */
#define MP3B200_ID3_V1_ONLY 4    /* id3tag_v1_only */
#define MP3B200_ID3_V2_ONLY 8    /* id3tag_v2_only */
#define MP3B200_ID3_SPACE_V1 16  /* id3tag_space_v1: pad version 1 fields with spaces */
#define MP3B200_ID3_PAD_V2 32    /* id3tag_set_pad(padding), 128 bytes if padding <= 0 */
typedef struct mp3b200_id3tag {
  const char *title, *artist, *album, *year, *comment, *track, *genre;
  int flags, padding;
  int64_t num_samples;           /* gfp.num_samples for the TLEN frame; -1 = unknown (no TLEN) */
  int samplerate;
} mp3b200_id3tag;
int mp3b200_id3v2_tag(const mp3b200_id3tag* t, uint8_t* buf, int cap);
int mp3b200_id3v1_tag(const mp3b200_id3tag* t, uint8_t* buf, int cap);
const char* mp3b200_id3_genre_name(int index);

/* Test / bench tap of k_music_crc: CRC-16 (VBRTag.js:547-556, start 0) of the ranges [off[i], off[i] + len[i]) of a DEVICE
 * buffer; `ms` (optional) receives the CUDA-event time of one launch sequence incl. its 4-byte-per-range read-back. */
int mp3b200_debug_music_crc(const uint8_t* d_buf, const int64_t* off, const int64_t* len, int nranges, uint32_t* crc, float* ms);

/* Replaces `lamejs.WavHeader.readHeader(dataView)` (src/js/index.js:154-193): the RIFF/WAVE front-end lamejs ships for its
 * examples.  Returns 1 and fills `out`; 0 where the reference returns undefined (not RIFF / not WAVE / "fmt " not first);
 * -1 where it throws 'extended fmt chunk not implemented' (fmt length other than 16 or 18); -2 where its DataView read runs
 * past the buffer (RangeError).  Pure host code, no device needed.  PCM starts at data + data_offset. */
typedef struct mp3b200_wav_header { int64_t data_offset, data_len; int32_t channels; uint32_t sample_rate; } mp3b200_wav_header;
int mp3b200_wav_read_header(const uint8_t* data, int64_t len, mp3b200_wav_header* out);

/* ---- stage taps for parity tests (one stream, whole-stream semantics) -----------------------------------
 * Run the pipeline for one stream given host PCM and copy intermediate results back.  Any output pointer may be
 * NULL.  Shapes ([F] = mp3b200_stream_frames(n)):
 *   xr          float [F][2 gr][nch][576]   MDCT spectrum (NewMDCT.js mdct_sub48 output)
 *   blocktype   int32 [F][2][nch]           final block type per granule (PsyModel.js block_type_set)
 *   en_l/thm_l  float [F][2][nch][22], en_s/thm_s float [F][2][nch][13][3]   masking handed to the quantizer
 *   ath_adjust  double[F]                   ATH.adjust after adjust_ATH (Encoder.js:166-243)
 *   l3_enc      int32 [F][2][nch][576], ginfo int32 [F][2][nch][16] (global_gain, part2_3_length, part2_length,
 *               big_values, count1, scalefac_compress, table_select[3], region0, region1, preflag, scalefac_scale,
 *               count1table, block_type, reserved)
 * If `force_blocktype` is non-NULL (int32 [F][2][nch]) it overrides the psy model's block decision for the
 * filterbank stage (used to test the MDCT in isolation). */
int mp3b200_debug_stages(int channels, int samplerate, int kbps, const int16_t* left, const int16_t* right,
                         int64_t nsamples, const int32_t* force_blocktype, float* xr, int32_t* blocktype,
                         float* en_l, float* thm_l, float* en_s, float* thm_s, double* ath_adjust,
                         int32_t* l3_enc, int32_t* ginfo, uint8_t* bytes_out, int64_t bytes_cap);

const char* mp3b200_last_error(void);
/* total number of kernel launches issued by this library since load (bench.py "gpu_launches") */
int64_t mp3b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
