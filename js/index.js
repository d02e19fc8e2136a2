'use strict';
/**
 * lamejs-compatible facade over libmp3b200.so (B200-native MP3 encoder).
 *
 *   const lamejs = require('mp3b200');            // instead of require('lamejs')
 *   const enc = new lamejs.Mp3Encoder(2, 44100, 128);
 *   const mp3 = enc.encodeBuffer(left, right);    // Int16Array in, Int8Array out (frames completed by this call)
 *   const tail = enc.flush();
 *
 * Same constructor / encodeBuffer / flush surface and return types as zhuker/lamejs src/js/index.js:66-136.
 * Binding: ffi-napi over the C ABI declared in include/mp3b200.h.  NOT EXECUTABLE in the build image (no node);
 * shipped as the reference-side binding a maintainer would use.
 */
const ffi = require('ffi-napi');
const ref = require('ref-napi');

const voidPtr = ref.refType(ref.types.void);
const voidPtrPtr = ref.refType(voidPtr);
const lib = ffi.Library(process.env.MP3B200_LIB || 'libmp3b200', {
  mp3b200_create: ['int', ['int', 'int', 'int', voidPtrPtr]],
  mp3b200_encode: ['int', [voidPtr, 'pointer', 'pointer', 'int', 'pointer', 'int']],
  mp3b200_flush: ['int', [voidPtr, 'pointer', 'int']],
  mp3b200_destroy: ['void', [voidPtr]],
  mp3b200_export_state: ['int', [voidPtr, 'pointer', 'int']],
  mp3b200_import_state: ['int', [voidPtr, 'pointer', 'int']],
  mp3b200_seek: ['int', [voidPtr, 'int64', 'pointer', 'pointer', 'int']],
  mp3b200_last_error: ['string', []],
});

function Mp3Encoder(channels, samplerate, kbps) {
  if (arguments.length !== 3) {              // index.js:67-72
    console.error('WARN: Mp3Encoder(channels, samplerate, kbps) not specified');
    channels = 1; samplerate = 44100; kbps = 128;
  }
  const hp = ref.alloc(voidPtr);
  const rc = lib.mp3b200_create(channels, samplerate, kbps, hp);
  if (rc !== 0) throw new Error('mp3b200_create failed (' + rc + '): ' + lib.mp3b200_last_error());
  const h = hp.deref();
  let maxSamples = 1152;
  let buf = Buffer.alloc(0 | (1.25 * maxSamples + 7200));   // index.js:113-114

  const asBuf = (a) => Buffer.from(a.buffer, a.byteOffset, a.byteLength);

  this.encodeBuffer = function (left, right) {
    if (channels === 1) right = left;
    if (left.length > maxSamples) {                          // index.js:122-126
      maxSamples = left.length;
      buf = Buffer.alloc(0 | (1.25 * maxSamples + 7200));
    }
    const n = lib.mp3b200_encode(h, asBuf(left), asBuf(right), left.length, buf, buf.length);
    if (n < 0) throw new Error('mp3b200_encode failed (' + n + '): ' + lib.mp3b200_last_error());
    return new Int8Array(buf.subarray(0, n));                // a fresh copy, like index.js:129
  };

  this.flush = function () {
    const n = lib.mp3b200_flush(h, buf, buf.length);
    if (n < 0) throw new Error('mp3b200_flush failed (' + n + '): ' + lib.mp3b200_last_error());
    return new Int8Array(buf.subarray(0, n));
  };

  // ---- beyond lamejs: the encoder state as a blob (checkpoint / resume; segment workers, see INTEGRATION.md) ----
  this.exportState = function () {
    const n = lib.mp3b200_export_state(h, ref.NULL, 0);
    if (n < 0) throw new Error('mp3b200_export_state failed (' + n + '): ' + lib.mp3b200_last_error());
    const blob = Buffer.alloc(n);
    const m = lib.mp3b200_export_state(h, blob, n);
    if (m < 0) throw new Error('mp3b200_export_state failed (' + m + '): ' + lib.mp3b200_last_error());
    return blob.subarray(0, m);
  };
  this.importState = function (blob) {
    const rc = lib.mp3b200_import_state(h, blob, blob.length);
    if (rc !== 0) throw new Error('mp3b200_import_state failed (' + rc + '): ' + lib.mp3b200_last_error());
  };
  this.seek = function (frame, leftHist, rightHist) {
    if (channels === 1 || !rightHist) rightHist = leftHist;
    const rc = lib.mp3b200_seek(h, frame, asBuf(leftHist), asBuf(rightHist), leftHist.length);
    if (rc !== 0) throw new Error('mp3b200_seek failed (' + rc + '): ' + lib.mp3b200_last_error());
  };

  this.close = function () { lib.mp3b200_destroy(h); };
}

/** WavHeader.readHeader stays in JS exactly as in lamejs (src/js/index.js:138-193); it is not on the hot path. */
function WavHeader() { this.dataOffset = 0; this.dataLen = 0; this.channels = 0; this.sampleRate = 0; }
function fourccToInt(f) { return f.charCodeAt(0) << 24 | f.charCodeAt(1) << 16 | f.charCodeAt(2) << 8 | f.charCodeAt(3); }
WavHeader.RIFF = fourccToInt('RIFF'); WavHeader.WAVE = fourccToInt('WAVE');
WavHeader.fmt_ = fourccToInt('fmt '); WavHeader.data = fourccToInt('data');
WavHeader.readHeader = function (dataView) {
  const w = new WavHeader();
  const header = dataView.getUint32(0, false);
  if (WavHeader.RIFF !== header) return undefined;
  if (WavHeader.WAVE !== dataView.getUint32(8, false)) return undefined;
  if (WavHeader.fmt_ !== dataView.getUint32(12, false)) return undefined;
  const fmtLen = dataView.getUint32(16, true);
  let pos = 16 + 4;
  if (fmtLen !== 16 && fmtLen !== 18) return undefined;
  w.channels = dataView.getUint16(pos + 2, true);
  w.sampleRate = dataView.getUint32(pos + 4, true);
  pos += fmtLen;
  let len = 0;
  for (let i = 0; i < 10 && WavHeader.data !== dataView.getUint32(pos, false); i++) {
    len = dataView.getUint32(pos + 4, true);
    pos += len + 8;
  }
  if (WavHeader.data !== dataView.getUint32(pos, false)) return undefined;
  w.dataLen = dataView.getUint32(pos + 4, true);
  w.dataOffset = pos + 8;
  return w;
};

module.exports = { Mp3Encoder, WavHeader };
