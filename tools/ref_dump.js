#!/usr/bin/env node
/* Out-of-band pin against the real reference: run with node where `lamejs` is installed (npm i lamejs@1.2.x or a
 * checkout of zhuker/lamejs: `node tools/ref_dump.js /path/to/lamejs/src/js/index.js`).  Re-implements the integer
 * generators of tests/synth.py (splitmix64 counter PRNG) and prints sha256 + length of lamejs's output for the cases
 * of tests/golden/golden.json that do not need libm (white / octave / burst / silence). */
const crypto = require('crypto');
const lamejs = require(process.argv[2] || 'lamejs');
const M1 = 0xBF58476D1CE4E5B9n, M2 = 0x94D049BB133111EBn, G = 0x9E3779B97F4A7C15n, MASK = (1n << 64n) - 1n;
const mix = (z) => { z = ((z ^ (z >> 30n)) * M1) & MASK; z = ((z ^ (z >> 27n)) * M2) & MASK; return z ^ (z >> 31n); };
const h64 = (seed, i) => mix((BigInt(seed) + (BigInt(i) + 1n) * G) & MASK);
const s16 = (seed, i) => { const v = Number(h64(seed, i) >> 48n); return v >= 32768 ? v - 65536 : v; };
function gen(kind, n, seed) {
  const l = new Int16Array(n), r = new Int16Array(n);
  for (let i = 0; i < n; i++) {
    if (kind === 'white') { l[i] = s16(0x5EED0003 + seed, i); r[i] = s16((0x5EED0003 + seed) ^ 0xFFFF0000, i); }
    else if (kind === 'octave') {
      const f = (sd) => (s16(sd, i) >> 1) + (s16(sd + 1, i >> 1) >> 2) + (s16(sd + 2, i >> 2) >> 3) + (s16(sd + 3, i >> 3) >> 4);
      l[i] = f(0x5EED0004 + seed); r[i] = f(((0x5EED0004 + seed) ^ 0xFFFF0000) >>> 0);
    } else if (kind === 'burst') {
      const on = (i % 4099) < 64;
      const c = (sd) => on ? s16(sd, i) : (Number((h64(sd, i) >> 63n) & 1n) * 2 - 1);
      l[i] = c(0x5EED0005 + seed); r[i] = c(((0x5EED0005 + seed) ^ 0xFFFF0000) >>> 0);
    }
  }
  return [l, r];
}
const cases = [['c1_silence_mono_128', 'silence', 1, 44100, 128, 44100, 0], ['c3_white_stereo_48k_320', 'white', 2, 48000, 320, 120 * 1152, 3],
  ['c4_octave_mono_128', 'octave', 1, 44100, 128, 150 * 1152, 4], ['c5_burst_stereo_128', 'burst', 2, 44100, 128, 150 * 1152, 5],
  ['white_mono_44k_320', 'white', 1, 44100, 320, 40 * 1152, 8]];
for (const [name, kind, ch, sr, kbps, n, seed] of cases) {
  const [l, r] = gen(kind, n, seed);
  const enc = new lamejs.Mp3Encoder(ch, sr, kbps);
  const parts = [Buffer.from(enc.encodeBuffer(l, r).buffer), Buffer.from(enc.flush().buffer)];
  const out = Buffer.concat(parts);
  console.log(name, out.length, crypto.createHash('sha256').update(out).digest('hex'));
}
