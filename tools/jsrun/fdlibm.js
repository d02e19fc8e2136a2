/* fdlibm.js -- Sun's fdlibm 5.3 e_log.c / e_log10.c / e_exp.c / e_pow.c in JavaScript, installed over Math.log / log10 / exp / pow.
 *
 * Why: V8 (node, Chrome -- where lamejs normally runs) computes these four with ports of fdlibm (base/ieee754); Qt's QV4, the
 * engine in the build image, calls the C library.  Loading this file before lamejs makes the engine's Math behave like V8's for
 * the functions lamejs calls per frame and at init, so the fixtures can be re-made "as under V8" and compared: the bytes do not
 * change (tests/test_lamejs_pin.py), i.e. the float32 store points absorb the last-ulp differences between the two libms.
 * The same algorithm is what oracle/js_math.h and lamejs_b200/csrc/mp3_math.cuh restate in C; tests/test_js_math.py checks this
 * file against js_math.h bit for bit on random arguments (the engine evaluates it, the C side prints the same table).
 * TEST INFRASTRUCTURE.  Plain IEEE doubles, evaluated left to right; words are read through a shared Float64Array / Uint32Array.
 */
(function () {
  'use strict';
  var f64 = new Float64Array(1), u32 = new Uint32Array(f64.buffer);      // little endian: u32[0] low word, u32[1] high word
  function hi(x) { f64[0] = x; return u32[1] | 0; }
  function lo(x) { f64[0] = x; return u32[0] >>> 0; }
  function setHi(x, h) { f64[0] = x; u32[1] = h >>> 0; return f64[0]; }
  function setLo(x, l) { f64[0] = x; u32[0] = l >>> 0; return f64[0]; }
  function fromWords(h, l) { u32[1] = h >>> 0; u32[0] = l >>> 0; return f64[0]; }
  var sqrt = Math.sqrt, abs = Math.abs;                                  // correctly rounded / exact in every engine

  function log(x) {
    var ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10, two54 = 1.80143985094819840000e+16,
      Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
      Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    var hfsq, f, s, z, R, w, t1, t2, dk, k, hx, i, j, lx;
    x = +x;
    hx = hi(x); lx = lo(x);
    k = 0;
    if (hx < 0x00100000) {
      if (((hx & 0x7fffffff) | lx) === 0) return -Infinity;
      if (hx < 0) return NaN;
      k -= 54; x *= two54; hx = hi(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    x = setHi(x, hx | (i ^ 0x3ff00000));
    k += (i >> 20);
    f = x - 1.0;
    if ((0x000fffff & (2 + hx)) < 3) {
      if (f === 0.0) {
        if (k === 0) return 0.0;
        dk = k; return dk * ln2_hi + dk * ln2_lo;
      }
      R = f * f * (0.5 - 0.33333333333333333 * f);
      if (k === 0) return f - R;
      dk = k; return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    s = f / (2.0 + f);
    dk = k;
    z = s * s;
    i = (hx - 0x6147a) | 0;
    w = z * z;
    j = (0x6b851 - hx) | 0;
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
      hfsq = 0.5 * f * f;
      if (k === 0) return f - (hfsq - s * (hfsq + R));
      return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k === 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
  }

  function log10(x) {
    var two54 = 1.80143985094819840000e+16, ivln10 = 4.34294481903251816668e-01, log10_2hi = 3.01029995663611771306e-01,
      log10_2lo = 3.69423907715893078616e-13;
    var y, z, i, k, hx, lx;
    x = +x;
    hx = hi(x); lx = lo(x);
    k = 0;
    if (hx < 0x00100000) {
      if (((hx & 0x7fffffff) | lx) === 0) return -Infinity;
      if (hx < 0) return NaN;
      k -= 54; x *= two54; hx = hi(x); lx = lo(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    i = (k & 0x80000000) >>> 31;
    hx = (hx & 0x000fffff) | ((0x3ff - i) << 20);
    y = k + i;
    x = fromWords(hx, lx);
    z = y * log10_2lo + ivln10 * log(x);
    return z + y * log10_2hi;
  }

  function exp(x) {
    var halF = [0.5, -0.5], huge = 1.0e+300, twom1000 = 9.33263618503218878990e-302, o_threshold = 7.09782712893383973096e+02,
      u_threshold = -7.45133219101941108420e+02, ln2HI = [6.93147180369123816490e-01, -6.93147180369123816490e-01],
      ln2LO = [1.90821492927058770002e-10, -1.90821492927058770002e-10], invln2 = 1.44269504088896338700e+00,
      P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
      P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    var y, h = 0, l = 0, c, t, k = 0, xsb, hx;
    x = +x;
    hx = hi(x) >>> 0;
    xsb = (hx >>> 31) & 1;
    hx = (hx & 0x7fffffff) >>> 0;
    if (hx >= 0x40862E42) {
      if (hx >= 0x7ff00000) {
        if (((hx & 0xfffff) | lo(x)) !== 0) return x + x;
        return (xsb === 0) ? x : 0.0;
      }
      if (x > o_threshold) return huge * huge;
      if (x < u_threshold) return twom1000 * twom1000;
    }
    if (hx > 0x3fd62e42) {
      if (hx < 0x3FF0A2B2) {
        h = x - ln2HI[xsb]; l = ln2LO[xsb]; k = 1 - xsb - xsb;
      } else {
        k = (invln2 * x + halF[xsb]) | 0;                                // C: (int) of a value well inside the int range
        t = k;
        h = x - t * ln2HI[0];
        l = t * ln2LO[0];
      }
      x = h - l;
    } else if (hx < 0x3e300000) {
      if (huge + x > 1.0) return 1.0 + x;
    } else k = 0;
    t = x * x;
    c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k === 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    y = 1.0 - ((l - (x * c) / (2.0 - c)) - h);
    if (k >= -1021) return setHi(y, (hi(y) + (k << 20)) | 0);
    y = setHi(y, (hi(y) + ((k + 1000) << 20)) | 0);
    return y * twom1000;
  }

  function scalbn(z, n) {                                                // ldexp for the subnormal results of pow
    var k = (hi(z) & 0x7ff00000) >> 20;
    if (k === 0) { if (z === 0) return z; z *= 1.80143985094819840000e+16; k = ((hi(z) & 0x7ff00000) >> 20) - 54; }
    k = k + n;
    if (k > 0x7fe) return 1.0e300 * (z < 0 ? -1.0e300 : 1.0e300);
    if (k > 0) return setHi(z, (hi(z) & 0x800fffff) | (k << 20));
    if (k <= -54) return 1.0e-300 * (z < 0 ? -1.0e-300 : 1.0e-300);
    k += 54;
    return setHi(z, (hi(z) & 0x800fffff) | (k << 20)) * 5.55111512312578270212e-17;
  }

  function pow(x, y) {
    var bp = [1.0, 1.5], dp_h = [0.0, 5.84962487220764160156e-01], dp_l = [0.0, 1.35003920212974897128e-08],
      two53 = 9007199254740992.0, huge = 1.0e300, tiny = 1.0e-300,
      L1 = 5.99999999999994648725e-01, L2 = 4.28571428578550184252e-01, L3 = 3.33333329818377432918e-01,
      L4 = 2.72728123808534006489e-01, L5 = 2.30660745775561754067e-01, L6 = 2.06975017800338417784e-01,
      P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
      P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08,
      lg2 = 6.93147180559945286227e-01, lg2_h = 6.93147182464599609375e-01, lg2_l = -1.90465429995776804525e-09,
      ovt = 8.0085662595372944372e-0017, cp = 9.61796693925975554329e-01, cp_h = 9.61796700954437255859e-01,
      cp_l = -7.02846165095275826516e-09, ivln2 = 1.44269504088896338700e+00, ivln2_h = 1.44269502162933349609e+00,
      ivln2_l = 1.92596299112661746887e-08;
    var z, ax, z_h, z_l, p_h, p_l, y1, t1, t2, r, s, t, u, v, w, i, j, k, yisint, n, hx, hy, ix, iy, lx, ly;
    var ss, s2, s_h, s_l, t_h, t_l;
    x = +x; y = +y;
    hx = hi(x); lx = lo(x);
    hy = hi(y); ly = lo(y);
    ix = hx & 0x7fffffff; iy = hy & 0x7fffffff;
    if ((iy | ly) === 0) return 1.0;
    if (ix > 0x7ff00000 || ((ix === 0x7ff00000) && (lx !== 0)) || iy > 0x7ff00000 || ((iy === 0x7ff00000) && (ly !== 0))) return x + y;
    yisint = 0;
    if (hx < 0) {
      if (iy >= 0x43400000) yisint = 2;
      else if (iy >= 0x3ff00000) {
        k = (iy >> 20) - 0x3ff;
        if (k > 20) {
          j = ly >>> (52 - k);
          if (((j << (52 - k)) >>> 0) === ly) yisint = 2 - (j & 1);
        } else if (ly === 0) {
          j = iy >> (20 - k);
          if ((j << (20 - k)) === iy) yisint = 2 - (j & 1);
        }
      }
    }
    if (ly === 0) {
      if (iy === 0x7ff00000) {
        if (((ix - 0x3ff00000) | lx) === 0) return NaN;                   // ECMAScript: (+-1) ** +-Infinity is NaN
        else if (ix >= 0x3ff00000) return (hy >= 0) ? y : 0.0;
        else return (hy < 0) ? -y : 0.0;
      }
      if (iy === 0x3ff00000) { if (hy < 0) return 1.0 / x; return x; }
      if (hy === 0x40000000) return x * x;
      if (hy === 0x3fe00000) { if (hx >= 0) return sqrt(x); }
    }
    ax = abs(x);
    if (lx === 0) {
      if (ix === 0x7ff00000 || ix === 0 || ix === 0x3ff00000) {
        z = ax;
        if (hy < 0) z = 1.0 / z;
        if (hx < 0) {
          if (((ix - 0x3ff00000) | yisint) === 0) z = NaN;
          else if (yisint === 1) z = -z;
        }
        return z;
      }
    }
    n = (hx < 0) ? 0 : 1;
    if ((n | yisint) === 0) return NaN;
    s = 1.0;
    if ((n | (yisint - 1)) === 0) s = -1.0;
    if (iy > 0x41e00000) {
      if (iy > 0x43f00000) {
        if (ix <= 0x3fefffff) return (hy < 0) ? huge * huge : tiny * tiny;
        if (ix >= 0x3ff00000) return (hy > 0) ? huge * huge : tiny * tiny;
      }
      if (ix < 0x3fefffff) return (hy < 0) ? s * huge * huge : s * tiny * tiny;
      if (ix > 0x3ff00000) return (hy > 0) ? s * huge * huge : s * tiny * tiny;
      t = ax - 1.0;
      w = (t * t) * (0.5 - t * (0.3333333333333333333333 - t * 0.25));
      u = ivln2_h * t;
      v = t * ivln2_l - w * ivln2;
      t1 = u + v;
      t1 = setLo(t1, 0);
      t2 = v - (t1 - u);
    } else {
      n = 0;
      if (ix < 0x00100000) { ax *= two53; n -= 53; ix = hi(ax); }
      n += ((ix) >> 20) - 0x3ff;
      j = ix & 0x000fffff;
      ix = j | 0x3ff00000;
      if (j <= 0x3988E) k = 0;
      else if (j < 0xBB67A) k = 1;
      else { k = 0; n += 1; ix -= 0x00100000; }
      ax = setHi(ax, ix);
      u = ax - bp[k];
      v = 1.0 / (ax + bp[k]);
      ss = u * v;
      s_h = ss;
      s_h = setLo(s_h, 0);
      t_h = fromWords(((ix >> 1) | 0x20000000) + 0x00080000 + (k << 18), 0);
      t_l = ax - (t_h - bp[k]);
      s_l = v * ((u - s_h * t_h) - s_h * t_l);
      s2 = ss * ss;
      r = s2 * s2 * (L1 + s2 * (L2 + s2 * (L3 + s2 * (L4 + s2 * (L5 + s2 * L6)))));
      r += s_l * (s_h + ss);
      s2 = s_h * s_h;
      t_h = 3.0 + s2 + r;
      t_h = setLo(t_h, 0);
      t_l = r - ((t_h - 3.0) - s2);
      u = s_h * t_h;
      v = s_l * t_h + t_l * ss;
      p_h = u + v;
      p_h = setLo(p_h, 0);
      p_l = v - (p_h - u);
      z_h = cp_h * p_h;
      z_l = cp_l * p_h + p_l * cp + dp_l[k];
      t = n;
      t1 = (((z_h + z_l) + dp_h[k]) + t);
      t1 = setLo(t1, 0);
      t2 = z_l - (((t1 - t) - dp_h[k]) - z_h);
    }
    y1 = y;
    y1 = setLo(y1, 0);
    p_l = (y - y1) * t1 + y * t2;
    p_h = y1 * t1;
    z = p_l + p_h;
    j = hi(z);
    i = lo(z) | 0;
    if (j >= 0x40900000) {
      if (((j - 0x40900000) | i) !== 0) return s * huge * huge;
      if (p_l + ovt > z - p_h) return s * huge * huge;
    } else if ((j & 0x7fffffff) >= 0x4090cc00) {
      if (((j - (0xc090cc00 | 0)) | i) !== 0) return s * tiny * tiny;
      if (p_l <= z - p_h) return s * tiny * tiny;
    }
    i = j & 0x7fffffff;
    k = (i >> 20) - 0x3ff;
    n = 0;
    if (i > 0x3fe00000) {
      n = (j + (0x00100000 >> (k + 1))) | 0;
      k = ((n & 0x7fffffff) >> 20) - 0x3ff;
      t = fromWords(n & ~(0x000fffff >> k), 0);
      n = ((n & 0x000fffff) | 0x00100000) >> (20 - k);
      if (j < 0) n = -n;
      p_h -= t;
    }
    t = p_l + p_h;
    t = setLo(t, 0);
    u = t * lg2_h;
    v = (p_l - (t - p_h)) * lg2 + t * lg2_l;
    z = u + v;
    w = v - (z - u);
    t = z * z;
    t1 = z - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    r = (z * t1) / (t1 - 2.0) - (w + z * w);
    z = 1.0 - (r - z);
    j = hi(z);
    j = (j + (n << 20)) | 0;
    if ((j >> 20) <= 0) z = scalbn(z, n);
    else z = setHi(z, (hi(z) + (n << 20)) | 0);
    return s * z;
  }

  Math.__fdlibm = {log: log, log10: log10, exp: exp, pow: pow, native: {log: Math.log, log10: Math.log10, exp: Math.exp, pow: Math.pow}};
  Math.log = log; Math.log10 = log10; Math.exp = exp; Math.pow = pow;
})();
