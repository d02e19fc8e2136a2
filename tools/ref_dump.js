#!/usr/bin/env node
/* Out-of-band check of the committed fixtures against lamejs under node / V8 (the fixtures themselves were produced by
 * the unmodified reference under Qt's QJSEngine, see tools/jsrun/ -- this script answers "does V8, whose Math.* is an
 * fdlibm port, give the same bytes?").
 *
 *   node tools/ref_dump.js /path/to/lamejs/src/js/index.js --check tests/golden/lamejs_golden.json
 *       -> PASS/FAIL per case (cases whose input needs libm -- kinds `sweep`, `sine` -- are skipped), exit code 1 on FAIL
 *   node tools/ref_dump.js /path/to/lamejs/src/js/index.js --check tests/golden/lamejs_golden.json --only c5_burst_stereo_128
 *   node tools/ref_dump.js ... --dump NAME            -> per-call byte counts and the first 64 bytes of every frame
 *
 * Inputs are regenerated with the integer-only generators of tests/synth.py (splitmix64 counter PRNG). */
const crypto = require('crypto');
const fs = require('fs');
const args = process.argv.slice(2);
const lamejs = require(args[0] || 'lamejs');
const opt = (k) => { const i = args.indexOf(k); return i >= 0 ? args[i + 1] : null; };
const M1 = 0xBF58476D1CE4E5B9n, M2 = 0x94D049BB133111EBn, G = 0x9E3779B97F4A7C15n, MASK = (1n << 64n) - 1n;
const mix = (z) => { z = ((z ^ (z >> 30n)) * M1) & MASK; z = ((z ^ (z >> 27n)) * M2) & MASK; return z ^ (z >> 31n); };
const h64 = (seed, i) => mix((BigInt(seed) + (BigInt(i) + 1n) * G) & MASK);
const s16 = (seed, i) => { const v = Number(h64(seed, i) >> 48n); return v >= 32768 ? v - 65536 : v; };
const u32 = (x) => x >>> 0;
function gen(kind, n, seed) {
  const l = new Int16Array(n), r = new Int16Array(n);
  for (let i = 0; i < n; i++) {
    if (kind === 'white') { l[i] = s16(0x5EED0003 + seed, i); r[i] = s16(u32((0x5EED0003 + seed) ^ 0xFFFF0000), i); }
    else if (kind === 'noise') {        /* tests/synth.py make_signal("noise"): (a + c) >> 4 of two white draws */
      const a = s16(0x5EED0100 + seed, i), b = s16(u32((0x5EED0100 + seed) ^ 0xFFFF0000), i);
      const c = s16(0x5EED0200 + seed, i), d = s16(u32((0x5EED0200 + seed) ^ 0xFFFF0000), i);
      l[i] = (a + c) >> 4; r[i] = (b + d) >> 4;
    } else if (kind === 'octave') {
      const f = (sd) => (s16(sd, i) >> 1) + (s16(sd + 1, i >> 1) >> 2) + (s16(sd + 2, i >> 2) >> 3) + (s16(sd + 3, i >> 3) >> 4);
      l[i] = f(0x5EED0004 + seed); r[i] = f(u32((0x5EED0004 + seed) ^ 0xFFFF0000));
    } else if (kind === 'burst') {
      const on = (i % 4099) < 64;
      const c = (sd) => on ? s16(sd, i) : (Number((h64(sd, i) >> 63n) & 1n) * 2 - 1);
      l[i] = c(0x5EED0005 + seed); r[i] = c(u32((0x5EED0005 + seed) ^ 0xFFFF0000));
    }                                   /* 'silence': zeros */
  }
  return [l, r];
}
function encode(c) {
  const [l, r] = gen(c.kind, c.samples, c.seed);
  const enc = new lamejs.Mp3Encoder(c.channels, c.samplerate, c.kbps);
  const parts = [], step = c.chunk > 0 ? c.chunk : Math.max(c.samples, 1);
  for (let i = 0; i < c.samples; i += step)
    parts.push(Buffer.from(c.channels === 2 ? enc.encodeBuffer(l.subarray(i, i + step), r.subarray(i, i + step)) : enc.encodeBuffer(l.subarray(i, i + step))));
  parts.push(Buffer.from(enc.flush()));
  return parts;
}
const fixtures = JSON.parse(fs.readFileSync(opt('--check') || 'tests/golden/lamejs_golden.json')).cases;
const only = opt('--only'), dump = opt('--dump');
let fail = 0, pass = 0, skipped = 0;
for (const name of Object.keys(fixtures).sort()) {
  const c = fixtures[name];
  if ((only && name !== only) || (dump && name !== dump)) continue;
  if (c.kind === 'sweep' || c.kind === 'sine') { skipped++; continue; }
  const parts = encode(c), out = Buffer.concat(parts);
  const sha = crypto.createHash('sha256').update(out).digest('hex');
  const sizes = crypto.createHash('sha256').update(JSON.stringify(parts.map((p) => p.length)).replace(/,/g, ', ')).digest('hex');
  const ok = sha === c.sha256 && out.length === c.bytes && sizes === c.sizes_sha256;
  console.log((ok ? 'PASS ' : 'FAIL ') + name + ' bytes ' + out.length + (ok ? '' : ' (fixture ' + c.bytes + ') sha ' + sha));
  if (ok) pass++; else fail++;
  if (dump) {
    console.log('per-call sizes:', parts.map((p) => p.length).join(' '));
    let off = 0, k = 0;
    while (off + 4 <= out.length) {   /* walk frames by header: enough to localise the first differing frame */
      const v1 = (out[off + 1] & 0x08) !== 0, br = out[off + 2] >> 4, sr = (out[off + 2] >> 2) & 3, pad = (out[off + 2] >> 1) & 1;
      const rates = v1 ? [44100, 48000, 32000] : ((out[off + 1] & 0x10) ? [22050, 24000, 16000] : [11025, 12000, 8000]);
      const kb = (v1 ? [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320] : [0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160])[br];
      const len = Math.floor((v1 ? 144000 : 72000) * kb / rates[sr]) + pad;
      console.log('frame', k++, 'off', off, 'len', len, crypto.createHash('sha256').update(out.subarray(off, off + len)).digest('hex').slice(0, 16), out.subarray(off, off + 24).toString('hex'));
      off += len;
    }
  }
}
console.log(pass + ' passed, ' + fail + ' failed, ' + skipped + ' skipped (input needs libm)');
process.exit(fail ? 1 : 0);
