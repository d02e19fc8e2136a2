/* lj_vbrtag.cpp -- Xing / Info / LAME tag of the oracle and the WAV front-end.  TEST INFRASTRUCTURE.
 *
 * Follows src/js/VBRTag.js (= src/main/java/mp3/VBRTag.java, which it was transliterated from):
 *   crcUpdateLookup / updateMusicCRC  VBRTag.js:547-556   (CRC-16, polynomial x^16+x^15+x^2+1, reflected, start 0)
 *   addVbr / addVbrFrame              VBRTag.js:149-167,192-197
 *   xingSeekTable                     VBRTag.js:169-185
 *   setLameTagFrameHeader             VBRTag.js:281-364
 *   InitVbrTag                        VBRTag.js:472-538
 *   putLameVBR                        VBRTag.js:558-802
 *   getLameTagFrame                   VBRTag.js:829-923
 * and src/js/index.js:154-193 (WavHeader.readHeader).
 *
 * What of this lamejs actually executes: copy_buffer(..., mp3data = 1) updates gfc.nMusicCRC and
 * VBR_seek_table.nBytesWritten on EVERY Mp3Encoder call (BitStream.js:924-935), so those two are part of the
 * hot path's state and are pinned against the engine (tests/test_lamejs_pin.py).  The tag writer itself is
 * switched off by Mp3Encoder (index.js:107) and VBRTag.js is not runnable as shipped -- `new int[400]`,
 * `case vbr_abr:`, `lame.BitrateIndex`, `Lame.LAME_ID`, `Tables` are unbound names, `0xff & version.charAt(j)` is 0 for
 * every character, `bag[i / 2]` and `TotalFrameSize` are fractional where Java divides integers.  With those names
 * bound by a test shim (tools/jsrun/tag_probe.py) the file does run; where its arithmetic is integer-exact (48 and 32 kHz
 * frame sizes) it agrees with this restatement in every byte the string bug does not touch.  Where JavaScript and Java
 * differ, Java's integer semantics are followed (they produce a decodable stream; the fractional ones do not):
 *   TotalFrameSize = integer quotient; bag halving by integer index; the magic and version strings as character codes.
 * xingSeekTable follows the JavaScript (double) arithmetic; Java computes `j`, `act`, `sum` in float.
 */
#include "lj_encoder.h"

static const int NUMTOCENTRIES = 100;
static const int MAXFRAMESIZE = 2880;
static const int VBRHEADERSIZE = NUMTOCENTRIES + 4 + 4 + 4 + 4 + 4;
static const int LAMEHEADERSIZE = VBRHEADERSIZE + 9 + 1 + 1 + 8 + 1 + 1 + 3 + 1 + 1 + 2 + 4 + 2 + 2;

/* crc16Lookup[i] (VBRTag.js:113-145) is the byte-wise table of the reflected polynomial 0xA001; generated, then
 * spot-checked against the listed constants in tests/test_tag_cpu.py */
static uint16_t crc_table[256];
static bool crc_table_ready = false;
static void crc_table_init(void) {
  for (int i = 0; i < 256; i++) {
    unsigned c = (unsigned)i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0xA001u : c >> 1;
    crc_table[i] = (uint16_t)c;
  }
  crc_table_ready = true;
}
static inline int crcUpdateLookup(int value, int crc) {
  int tmp = crc ^ value;
  crc = (crc >> 8) ^ crc_table[tmp & 0xff];
  return crc;
}

void lj_update_music_crc(LjEnc* e, const uint8_t* buf, int size) {
  if (!crc_table_ready) crc_table_init();
  int crc = e->nMusicCRC;
  for (int i = 0; i < size; ++i) crc = crcUpdateLookup(buf[i], crc);
  e->nMusicCRC = crc;
}

static void addVbr(LjEnc* e, int bitrate) {
  e->vbr_nframes++;
  e->vbr_sum += bitrate;
  e->vbr_seen++;
  if (e->vbr_seen < e->vbr_want) return;
  if (e->vbr_pos < 400) {
    e->vbr_bag[e->vbr_pos] = e->vbr_sum;
    e->vbr_pos++;
    e->vbr_seen = 0;
  }
  if (e->vbr_pos == 400) {
    for (int i = 1; i < 400; i += 2) e->vbr_bag[i / 2] = e->vbr_bag[i];
    e->vbr_want *= 2;
    e->vbr_pos /= 2;
  }
}

void lj_add_vbr_frame(LjEnc* e) {            /* Encoder.js:640-641 */
  if (e->bWriteVbrTag) addVbr(e, e->brate);  /* Tables.bitrate_table[version][bitrate_index] */
}

static void xingSeekTable(const LjEnc* e, uint8_t* t) {
  if (e->vbr_pos <= 0) return;
  for (int i = 1; i < NUMTOCENTRIES; ++i) {
    double j = (double)i / NUMTOCENTRIES;
    int indx = js_toint32(floor(j * e->vbr_pos));
    if (indx > e->vbr_pos - 1) indx = e->vbr_pos - 1;
    double act = e->vbr_bag[indx];
    double sum = e->vbr_sum;
    int seek_point = js_toint32(256. * act / sum);
    if (seek_point > 255) seek_point = 255;
    t[i] = (uint8_t)(0xff & seek_point);
  }
}

static void createInteger(uint8_t* buf, int pos, int value) {
  buf[pos + 0] = (uint8_t)((value >> 24) & 0xff);
  buf[pos + 1] = (uint8_t)((value >> 16) & 0xff);
  buf[pos + 2] = (uint8_t)((value >> 8) & 0xff);
  buf[pos + 3] = (uint8_t)(value & 0xff);
}
static void createShort(uint8_t* buf, int pos, int value) {
  buf[pos + 0] = (uint8_t)((value >> 8) & 0xff);
  buf[pos + 1] = (uint8_t)(value & 0xff);
}
static int shiftInBitsValue(int x, int n, int v) { return 0xff & ((x << n) | (v & ~(-1 << n))); }

static int BitrateIndex(int bRate, int version, int samplerate) {   /* Lame.js:431-443 */
  static const int bt[3][16] = {{0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, -1},
                                {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, -1},
                                {0, 8, 16, 24, 32, 40, 48, 56, 64, -1, -1, -1, -1, -1, -1, -1}};
  if (samplerate < 16000) version = 2;
  for (int i = 0; i <= 14; i++)
    if (bt[version][i] > 0 && bt[version][i] == bRate) return i;
  return -1;
}

static void setLameTagFrameHeader(const LjEnc* e, uint8_t* buffer) {
  buffer[0] = (uint8_t)shiftInBitsValue(buffer[0], 8, 0xff);
  buffer[1] = (uint8_t)shiftInBitsValue(buffer[1], 3, 7);
  buffer[1] = (uint8_t)shiftInBitsValue(buffer[1], 1, (e->out_samplerate < 16000) ? 0 : 1);
  buffer[1] = (uint8_t)shiftInBitsValue(buffer[1], 1, e->version);
  buffer[1] = (uint8_t)shiftInBitsValue(buffer[1], 2, 4 - 3);
  buffer[1] = (uint8_t)shiftInBitsValue(buffer[1], 1, 1);                 /* !error_protection */
  buffer[2] = (uint8_t)shiftInBitsValue(buffer[2], 4, e->bitrate_index);
  buffer[2] = (uint8_t)shiftInBitsValue(buffer[2], 2, e->samplerate_index);
  buffer[2] = (uint8_t)shiftInBitsValue(buffer[2], 1, 0);
  buffer[2] = (uint8_t)shiftInBitsValue(buffer[2], 1, 0);                 /* gfp.extension */
  buffer[3] = (uint8_t)shiftInBitsValue(buffer[3], 2, e->mode_mono ? 3 : (e->mode_joint ? 1 : 0));   /* gfp.mode.ordinal(): STEREO 0, JOINT_STEREO 1, MONO 3 */
  buffer[3] = (uint8_t)shiftInBitsValue(buffer[3], 2, e->mode_ext);
  buffer[3] = (uint8_t)shiftInBitsValue(buffer[3], 1, 0);                 /* copyright */
  buffer[3] = (uint8_t)shiftInBitsValue(buffer[3], 1, 1);                 /* original */
  buffer[3] = (uint8_t)shiftInBitsValue(buffer[3], 2, 0);                 /* emphasis */
  buffer[0] = 0xff;
  int abyte = 0xff & (buffer[1] & 0xf1);
  int bitrate = e->brate;                                                /* VBR == vbr_off */
  int bbyte = 0xff & (16 * BitrateIndex(bitrate, e->version, e->out_samplerate));
  if (e->version == 1) {
    buffer[1] = (uint8_t)(0xff & (abyte | 0x0a));
    abyte = 0xff & (buffer[2] & 0x0d);
    buffer[2] = (uint8_t)(0xff & (bbyte | abyte));
  } else {
    buffer[1] = (uint8_t)(0xff & (abyte | 0x02));
    abyte = 0xff & (buffer[2] & 0x0d);
    buffer[2] = (uint8_t)(0xff & (bbyte | abyte));
  }
}

/* putbits_noheaders(val, 8) for byte-aligned data (BitStream.js:140-163) + add_dummy_byte's shift of every header's
 * write_timing (BitStream.js:817-826) */
static void add_dummy_byte(LjEnc* e, int val) {
  e->bs_byteidx++;
  e->bs_buf[e->bs_byteidx] = (uint8_t)val;
  e->bs_totbit += 8;
  for (int i = 0; i < 256; ++i) e->header[i].write_timing += 8;      /* BitStream.js:823-824 */
}

extern "C" {

/* gfp.bWriteVbrTag = true before lame_init_params: lame_init_bitstream calls InitVbrTag (Lame.js:706-708).
 * Returns 1 when the tag is on, 0 when InitVbrTag switched it off (the frame is too small for it). */
int lj_enable_vbr_tag(LjEnc* e) {
  if (!e || e->frameNum != 0 || e->bs_byteidx != -1) return -1;
  if (!crc_table_ready) crc_table_init();
  const int kbps_header = e->brate;                                       /* VBR == vbr_off */
  const int totalFrameSize = ((e->version + 1) * 72000 * kbps_header) / e->out_samplerate;
  const int headerSize = e->sideinfo_len + LAMEHEADERSIZE;
  e->vbr_TotalFrameSize = totalFrameSize;
  if (totalFrameSize < headerSize || totalFrameSize > MAXFRAMESIZE) { e->bWriteVbrTag = 0; return 0; }
  e->bWriteVbrTag = 1;
  e->vbr_nframes = 0; e->nBytesWritten = 0; e->vbr_sum = 0; e->vbr_seen = 0; e->vbr_want = 1; e->vbr_pos = 0;
  uint8_t buffer[MAXFRAMESIZE];
  memset(buffer, 0, sizeof buffer);
  setLameTagFrameHeader(e, buffer);
  for (int i = 0; i < totalFrameSize; ++i) add_dummy_byte(e, buffer[i] & 0xff);
  return 1;
}

int lj_music_crc(const LjEnc* e) { return e->nMusicCRC; }
long long lj_bytes_written(const LjEnc* e) { return e->nBytesWritten; }
int lj_vbr_frames(const LjEnc* e) { return e->vbr_nframes; }
int lj_encoder_padding(const LjEnc* e) { return e->encoder_padding; }

static int putLameVBR(const LjEnc* e, int musicLength, uint8_t* streamBuffer, int streamBufferPos, int crc) {
  int bytesWritten = 0;
  const int encDelay = ENCDELAY;                                           /* gfp.encoder_delay (Lame.js:941) */
  const int encPadding = e->encoder_padding;
  int quality = 100 - 10 * 4 - e->quality;                                 /* VBR_q = 4 (Lame.js:153) */
  static const char version[] = "LAME3.98r";                               /* Version.js:56-59 */
  const int revision = 0x00;
  static const int vbrTypeTranslator[7] = {1, 5, 3, 2, 4, 0, 3};
  const double lp = e->lowpass_final / 100.0 + .5;
  const int lowpass = js_toint32(lp > 255 ? 255 : lp);
  const int peakSignalAmplitude = 0, radioReplayGain = 0, audiophileReplayGain = 0;   /* findReplayGain / findPeakSample off */
  const int noiseShaping = e->noise_shaping;
  int stereoMode, nonOptimal = 0, sourceFreq;
  const bool expNPsyTune = (e->exp_nspsytune & 1) != 0;
  const bool safeJoint = (e->exp_nspsytune & 2) != 0;
  const int athType = e->ATHtype;
  const int abrBitrate = e->brate;                                         /* vbr_off */
  const int vbr = vbrTypeTranslator[0];                                    /* VbrMode.vbr_off.ordinal() == 0 */
  const int revMethod = 0x10 * revision + vbr;
  /* nogap_total == nogap_current == 0 (LameInternalFlags.js:335-336): neither flag */
  const int flags = athType + ((expNPsyTune ? 1 : 0) << 4) + ((safeJoint ? 1 : 0) << 5);
  if (quality < 0) quality = 0;
  stereoMode = e->mode_mono ? 0 : (e->mode_joint ? 3 : 1);                 /* MONO 0, STEREO 1, JOINT_STEREO 3 (force_ms off) */
  if (e->in_samplerate <= 32000) sourceFreq = 0x00;
  else if (e->in_samplerate == 48000) sourceFreq = 0x02;
  else if (e->in_samplerate > 48000) sourceFreq = 0x03;
  else sourceFreq = 0x01;
  /* disable_reservoir (index.js:108) && brate < 320, or a source rate <= 32 kHz */
  if ((e->disable_reservoir && e->brate < 320) || athType == 0 || e->in_samplerate <= 32000) nonOptimal = 1;
  const int misc = noiseShaping + (stereoMode << 2) + (nonOptimal << 5) + (sourceFreq << 6);
  const int musicCRC = e->nMusicCRC;

  createInteger(streamBuffer, streamBufferPos + bytesWritten, quality);
  bytesWritten += 4;
  for (int j = 0; j < 9; j++) streamBuffer[streamBufferPos + bytesWritten + j] = (uint8_t)(0xff & version[j]);
  bytesWritten += 9;
  streamBuffer[streamBufferPos + bytesWritten] = (uint8_t)(0xff & revMethod);
  bytesWritten++;
  streamBuffer[streamBufferPos + bytesWritten] = (uint8_t)(0xff & lowpass);
  bytesWritten++;
  createInteger(streamBuffer, streamBufferPos + bytesWritten, peakSignalAmplitude);
  bytesWritten += 4;
  createShort(streamBuffer, streamBufferPos + bytesWritten, radioReplayGain);
  bytesWritten += 2;
  createShort(streamBuffer, streamBufferPos + bytesWritten, audiophileReplayGain);
  bytesWritten += 2;
  streamBuffer[streamBufferPos + bytesWritten] = (uint8_t)(0xff & flags);
  bytesWritten++;
  if (abrBitrate >= 255) streamBuffer[streamBufferPos + bytesWritten] = 0xFF;
  else streamBuffer[streamBufferPos + bytesWritten] = (uint8_t)(0xff & abrBitrate);
  bytesWritten++;
  streamBuffer[streamBufferPos + bytesWritten] = (uint8_t)(0xff & (encDelay >> 4));
  streamBuffer[streamBufferPos + bytesWritten + 1] = (uint8_t)(0xff & ((encDelay << 4) + (encPadding >> 8)));
  streamBuffer[streamBufferPos + bytesWritten + 2] = (uint8_t)(0xff & encPadding);
  bytesWritten += 3;
  streamBuffer[streamBufferPos + bytesWritten] = (uint8_t)(0xff & misc);
  bytesWritten++;
  streamBuffer[streamBufferPos + bytesWritten++] = 0;
  createShort(streamBuffer, streamBufferPos + bytesWritten, e->brate);     /* gfp.preset = the (snapped) bitrate, Presets.js:415 */
  bytesWritten += 2;
  createInteger(streamBuffer, streamBufferPos + bytesWritten, musicLength);
  bytesWritten += 4;
  createShort(streamBuffer, streamBufferPos + bytesWritten, musicCRC);
  bytesWritten += 2;
  for (int i = 0; i < bytesWritten; i++) crc = crcUpdateLookup(streamBuffer[streamBufferPos + i], crc);
  createShort(streamBuffer, streamBufferPos + bytesWritten, crc);
  bytesWritten += 2;
  return bytesWritten;
}

/* VBRTag.getLameTagFrame: 0 when the tag is off or no frame was counted; the needed size when `cap` is too small;
 * else writes TotalFrameSize bytes and returns that */
int lj_get_lametag_frame(LjEnc* e, uint8_t* buffer, int cap) {
  if (!e || !e->bWriteVbrTag) return 0;
  if (e->vbr_pos <= 0) return 0;
  if (cap < e->vbr_TotalFrameSize) return e->vbr_TotalFrameSize;
  memset(buffer, 0, e->vbr_TotalFrameSize);
  setLameTagFrameHeader(e, buffer);
  uint8_t toc[NUMTOCENTRIES];
  memset(toc, 0, sizeof toc);
  xingSeekTable(e, toc);
  int streamIndex = e->sideinfo_len;
  buffer[streamIndex++] = 'I'; buffer[streamIndex++] = 'n'; buffer[streamIndex++] = 'f'; buffer[streamIndex++] = 'o';   /* vbr_off */
  createInteger(buffer, streamIndex, 0x0001 + 0x0002 + 0x0004 + 0x0008);
  streamIndex += 4;
  createInteger(buffer, streamIndex, e->vbr_nframes);
  streamIndex += 4;
  const int streamSize = (int)(e->nBytesWritten + e->vbr_TotalFrameSize);
  createInteger(buffer, streamIndex, streamSize);
  streamIndex += 4;
  memcpy(buffer + streamIndex, toc, NUMTOCENTRIES);
  streamIndex += NUMTOCENTRIES;
  int crc = 0x00;
  for (int i = 0; i < streamIndex; i++) crc = crcUpdateLookup(buffer[i], crc);
  streamIndex += putLameVBR(e, streamSize, buffer, streamIndex, crc);
  return e->vbr_TotalFrameSize;
}

/* CRC-16 of a byte string continued from `crc` (exposed for the tests of the product's parallel CRC) */
int lj_crc16(const uint8_t* buf, long long n, int crc) {
  if (!crc_table_ready) crc_table_init();
  for (long long i = 0; i < n; i++) crc = crcUpdateLookup(buf[i], crc);
  return crc;
}
int lj_crc16_table(int i) { if (!crc_table_ready) crc_table_init(); return crc_table[i & 255]; }

/* WavHeader.readHeader (index.js:154-193).  Returns 1 and fills out[4] = {dataOffset, dataLen, channels, sampleRate};
 * 0 = `return undefined` (not RIFF / WAVE / "fmt "); -1 = throws 'extended fmt chunk not implemented';
 * -2 = the DataView read past the end (RangeError in JavaScript). */
static int rd32(const uint8_t* d, long long n, long long pos, int little, uint32_t* v) {
  if (pos < 0 || pos + 4 > n) return 0;
  *v = little ? ((uint32_t)d[pos] | (uint32_t)d[pos + 1] << 8 | (uint32_t)d[pos + 2] << 16 | (uint32_t)d[pos + 3] << 24)
              : ((uint32_t)d[pos] << 24 | (uint32_t)d[pos + 1] << 16 | (uint32_t)d[pos + 2] << 8 | (uint32_t)d[pos + 3]);
  return 1;
}
int lj_wav_read_header(const uint8_t* d, long long n, long long* out) {
  const uint32_t RIFF = 0x52494646u, WAVE = 0x57415645u, fmt_ = 0x666d7420u, data = 0x64617461u;
  uint32_t header, v, fmtLen, len = 0;
  out[0] = out[1] = out[2] = out[3] = 0;
  if (!rd32(d, n, 0, 0, &header)) return -2;
  if (RIFF != header) return 0;
  if (!rd32(d, n, 4, 1, &v)) return -2;                 /* fileLen, unused */
  if (!rd32(d, n, 8, 0, &v)) return -2;
  if (WAVE != v) return 0;
  if (!rd32(d, n, 12, 0, &v)) return -2;
  if (fmt_ != v) return 0;
  if (!rd32(d, n, 16, 1, &fmtLen)) return -2;
  long long pos = 16 + 4;
  switch (fmtLen) {
    case 16: case 18: {
      if (pos + 2 + 2 > n) return -2;
      out[2] = d[pos + 2] | d[pos + 3] << 8;
      if (!rd32(d, n, pos + 4, 1, &v)) return -2;
      out[3] = v;
      break;
    }
    default: return -1;
  }
  pos += fmtLen;
  while (data != header) {
    if (!rd32(d, n, pos, 0, &header)) return -2;
    if (!rd32(d, n, pos + 4, 1, &len)) return -2;
    if (data == header) break;
    pos += ((long long)len + 8);
  }
  out[1] = len;
  out[0] = pos + 8;
  return 1;
}

} /* extern "C" */
