/* lj_mdct.cpp -- polyphase analysis filterbank + MDCT of the oracle.  TEST INFRASTRUCTURE.
 * Follows src/js/NewMDCT.js: window_subband :534-914, mdct_short :927-979,
 * mdct_long :981-1051, mdct_sub48 :1053-1161.  Same statement order; doubles with
 * Float32Array store points (sb_sample, xr, work are F32).
 */
#include <string.h>
#include "lj_encoder.h"
#include "lj_tables.h"

#define EW(i) LJ_ENWINDOW[(i)]
#define WIN(t, i) LJ_MDCT_WIN[(t) * 36 + (i)]
/* tantab_l, cx, ca, cs alias win[SHORT_TYPE] (NewMDCT.js:507-510) */
#define WS(i) WIN(SHORT_TYPE, (i))

/* x1 + x1Pos indexing like the reference (x1Pos moves). */
static void window_subband(const F32* x1, int x1Pos, F32* a) {
  int wp = 10;
  int x2 = x1Pos + 238 - 14 - 286;
  for (int i = -15; i < 0; i++) {
    double w, s, t;
    w = EW(wp + -10); s = x1[x2 + -224] * w; t = x1[x1Pos + 224] * w;
    w = EW(wp + -9); s += x1[x2 + -160] * w; t += x1[x1Pos + 160] * w;
    w = EW(wp + -8); s += x1[x2 + -96] * w; t += x1[x1Pos + 96] * w;
    w = EW(wp + -7); s += x1[x2 + -32] * w; t += x1[x1Pos + 32] * w;
    w = EW(wp + -6); s += x1[x2 + 32] * w; t += x1[x1Pos + -32] * w;
    w = EW(wp + -5); s += x1[x2 + 96] * w; t += x1[x1Pos + -96] * w;
    w = EW(wp + -4); s += x1[x2 + 160] * w; t += x1[x1Pos + -160] * w;
    w = EW(wp + -3); s += x1[x2 + 224] * w; t += x1[x1Pos + -224] * w;

    w = EW(wp + -2); s += x1[x1Pos + -256] * w; t -= x1[x2 + 256] * w;
    w = EW(wp + -1); s += x1[x1Pos + -192] * w; t -= x1[x2 + 192] * w;
    w = EW(wp + 0); s += x1[x1Pos + -128] * w; t -= x1[x2 + 128] * w;
    w = EW(wp + 1); s += x1[x1Pos + -64] * w; t -= x1[x2 + 64] * w;
    w = EW(wp + 2); s += x1[x1Pos + 0] * w; t -= x1[x2 + 0] * w;
    w = EW(wp + 3); s += x1[x1Pos + 64] * w; t -= x1[x2 + -64] * w;
    w = EW(wp + 4); s += x1[x1Pos + 128] * w; t -= x1[x2 + -128] * w;
    w = EW(wp + 5); s += x1[x1Pos + 192] * w; t -= x1[x2 + -192] * w;

    s *= EW(wp + 6);
    w = t - s;
    a[30 + i * 2] = t + s;
    a[31 + i * 2] = EW(wp + 7) * w;
    wp += 18;
    x1Pos--;
    x2++;
  }
  {
    double s, t, u, v;
    t = x1[x1Pos + -16] * EW(wp + -10);
    s = x1[x1Pos + -32] * EW(wp + -2);
    t += (x1[x1Pos + -48] - x1[x1Pos + 16]) * EW(wp + -9);
    s += x1[x1Pos + -96] * EW(wp + -1);
    t += (x1[x1Pos + -80] + x1[x1Pos + 48]) * EW(wp + -8);
    s += x1[x1Pos + -160] * EW(wp + 0);
    t += (x1[x1Pos + -112] - x1[x1Pos + 80]) * EW(wp + -7);
    s += x1[x1Pos + -224] * EW(wp + 1);
    t += (x1[x1Pos + -144] + x1[x1Pos + 112]) * EW(wp + -6);
    s -= x1[x1Pos + 32] * EW(wp + 2);
    t += (x1[x1Pos + -176] - x1[x1Pos + 144]) * EW(wp + -5);
    s -= x1[x1Pos + 96] * EW(wp + 3);
    t += (x1[x1Pos + -208] + x1[x1Pos + 176]) * EW(wp + -4);
    s -= x1[x1Pos + 160] * EW(wp + 4);
    t += (x1[x1Pos + -240] - x1[x1Pos + 208]) * EW(wp + -3);
    s -= x1[x1Pos + 224];

    u = s - t;
    v = s + t;

    t = a[14];
    s = a[15] - t;

    a[31] = v + t;
    a[30] = u + s;
    a[15] = u - s;
    a[14] = v - t;
  }
  {
    double xr;
    xr = a[28] - a[0]; a[0] += a[28]; a[28] = xr * EW(wp + -2 * 18 + 7);
    xr = a[29] - a[1]; a[1] += a[29]; a[29] = xr * EW(wp + -2 * 18 + 7);

    xr = a[26] - a[2]; a[2] += a[26]; a[26] = xr * EW(wp + -4 * 18 + 7);
    xr = a[27] - a[3]; a[3] += a[27]; a[27] = xr * EW(wp + -4 * 18 + 7);

    xr = a[24] - a[4]; a[4] += a[24]; a[24] = xr * EW(wp + -6 * 18 + 7);
    xr = a[25] - a[5]; a[5] += a[25]; a[25] = xr * EW(wp + -6 * 18 + 7);

    xr = a[22] - a[6]; a[6] += a[22]; a[22] = xr * LJ_SQRT2;
    xr = a[23] - a[7]; a[7] += a[23]; a[23] = xr * LJ_SQRT2 - a[7];
    a[7] -= a[6];
    a[22] -= a[7];
    a[23] -= a[22];

    xr = a[6]; a[6] = a[31] - xr; a[31] = a[31] + xr;
    xr = a[7]; a[7] = a[30] - xr; a[30] = a[30] + xr;
    xr = a[22]; a[22] = a[15] - xr; a[15] = a[15] + xr;
    xr = a[23]; a[23] = a[14] - xr; a[14] = a[14] + xr;

    xr = a[20] - a[8]; a[8] += a[20]; a[20] = xr * EW(wp + -10 * 18 + 7);
    xr = a[21] - a[9]; a[9] += a[21]; a[21] = xr * EW(wp + -10 * 18 + 7);

    xr = a[18] - a[10]; a[10] += a[18]; a[18] = xr * EW(wp + -12 * 18 + 7);
    xr = a[19] - a[11]; a[11] += a[19]; a[19] = xr * EW(wp + -12 * 18 + 7);

    xr = a[16] - a[12]; a[12] += a[16]; a[16] = xr * EW(wp + -14 * 18 + 7);
    xr = a[17] - a[13]; a[13] += a[17]; a[17] = xr * EW(wp + -14 * 18 + 7);

    xr = -a[20] + a[24]; a[20] += a[24]; a[24] = xr * EW(wp + -12 * 18 + 7);
    xr = -a[21] + a[25]; a[21] += a[25]; a[25] = xr * EW(wp + -12 * 18 + 7);

    xr = a[4] - a[8]; a[4] += a[8]; a[8] = xr * EW(wp + -12 * 18 + 7);
    xr = a[5] - a[9]; a[5] += a[9]; a[9] = xr * EW(wp + -12 * 18 + 7);

    xr = a[0] - a[12]; a[0] += a[12]; a[12] = xr * EW(wp + -4 * 18 + 7);
    xr = a[1] - a[13]; a[1] += a[13]; a[13] = xr * EW(wp + -4 * 18 + 7);
    xr = a[16] - a[28]; a[16] += a[28]; a[28] = xr * EW(wp + -4 * 18 + 7);
    xr = -a[17] + a[29]; a[17] += a[29]; a[29] = xr * EW(wp + -4 * 18 + 7);

    xr = LJ_SQRT2 * (a[2] - a[10]); a[2] += a[10]; a[10] = xr;
    xr = LJ_SQRT2 * (a[3] - a[11]); a[3] += a[11]; a[11] = xr;
    xr = LJ_SQRT2 * (-a[18] + a[26]); a[18] += a[26]; a[26] = xr - a[18];
    xr = LJ_SQRT2 * (-a[19] + a[27]); a[19] += a[27]; a[27] = xr - a[19];

    xr = a[2]; a[19] -= a[3]; a[3] -= xr; a[2] = a[31] - xr; a[31] += xr;
    xr = a[3]; a[11] -= a[19]; a[18] -= xr; a[3] = a[30] - xr; a[30] += xr;
    xr = a[18]; a[27] -= a[11]; a[19] -= xr; a[18] = a[15] - xr; a[15] += xr;

    xr = a[19]; a[10] -= xr; a[19] = a[14] - xr; a[14] += xr;
    xr = a[10]; a[11] -= xr; a[10] = a[23] - xr; a[23] += xr;
    xr = a[11]; a[26] -= xr; a[11] = a[22] - xr; a[22] += xr;
    xr = a[26]; a[27] -= xr; a[26] = a[7] - xr; a[7] += xr;

    xr = a[27]; a[27] = a[6] - xr; a[6] += xr;

    xr = LJ_SQRT2 * (a[0] - a[4]); a[0] += a[4]; a[4] = xr;
    xr = LJ_SQRT2 * (a[1] - a[5]); a[1] += a[5]; a[5] = xr;
    xr = LJ_SQRT2 * (a[16] - a[20]); a[16] += a[20]; a[20] = xr;
    xr = LJ_SQRT2 * (a[17] - a[21]); a[17] += a[21]; a[21] = xr;

    xr = -LJ_SQRT2 * (a[8] - a[12]); a[8] += a[12]; a[12] = xr - a[8];
    xr = -LJ_SQRT2 * (a[9] - a[13]); a[9] += a[13]; a[13] = xr - a[9];
    xr = -LJ_SQRT2 * (a[25] - a[29]); a[25] += a[29]; a[29] = xr - a[25];
    xr = -LJ_SQRT2 * (a[24] + a[28]); a[24] -= a[28]; a[28] = xr - a[24];

    xr = a[24] - a[16]; a[24] = xr;
    xr = a[20] - xr; a[20] = xr;
    xr = a[28] - xr; a[28] = xr;

    xr = a[25] - a[17]; a[25] = xr;
    xr = a[21] - xr; a[21] = xr;
    xr = a[29] - xr; a[29] = xr;

    xr = a[17] - a[1]; a[17] = xr;
    xr = a[9] - xr; a[9] = xr;
    xr = a[25] - xr; a[25] = xr;
    xr = a[5] - xr; a[5] = xr;
    xr = a[21] - xr; a[21] = xr;
    xr = a[13] - xr; a[13] = xr;
    xr = a[29] - xr; a[29] = xr;

    xr = a[1] - a[0]; a[1] = xr;
    xr = a[16] - xr; a[16] = xr;
    xr = a[17] - xr; a[17] = xr;
    xr = a[8] - xr; a[8] = xr;
    xr = a[9] - xr; a[9] = xr;
    xr = a[24] - xr; a[24] = xr;
    xr = a[25] - xr; a[25] = xr;
    xr = a[4] - xr; a[4] = xr;
    xr = a[5] - xr; a[5] = xr;
    xr = a[20] - xr; a[20] = xr;
    xr = a[21] - xr; a[21] = xr;
    xr = a[12] - xr; a[12] = xr;
    xr = a[13] - xr; a[13] = xr;
    xr = a[28] - xr; a[28] = xr;
    xr = a[29] - xr; a[29] = xr;

    xr = a[0]; a[0] += a[31]; a[31] -= xr;
    xr = a[1]; a[1] += a[30]; a[30] -= xr;
    xr = a[16]; a[16] += a[15]; a[15] -= xr;
    xr = a[17]; a[17] += a[14]; a[14] -= xr;
    xr = a[8]; a[8] += a[23]; a[23] -= xr;
    xr = a[9]; a[9] += a[22]; a[22] -= xr;
    xr = a[24]; a[24] += a[7]; a[7] -= xr;
    xr = a[25]; a[25] += a[6]; a[6] -= xr;
    xr = a[4]; a[4] += a[27]; a[27] -= xr;
    xr = a[5]; a[5] += a[26]; a[26] -= xr;
    xr = a[20]; a[20] += a[11]; a[11] -= xr;
    xr = a[21]; a[21] += a[10]; a[10] -= xr;
    xr = a[12]; a[12] += a[19]; a[19] -= xr;
    xr = a[13]; a[13] += a[18]; a[18] -= xr;
    xr = a[28]; a[28] += a[3]; a[3] -= xr;
    xr = a[29]; a[29] += a[2]; a[2] -= xr;
  }
}

static void mdct_short(F32* inout, int pos) {
  for (int l = 0; l < 3; l++) {
    double tc0, tc1, tc2, ts0, ts1, ts2;
    ts0 = inout[pos + 2 * 3] * WS(0) - inout[pos + 5 * 3];
    tc0 = inout[pos + 0 * 3] * WS(2) - inout[pos + 3 * 3];
    tc1 = ts0 + tc0;
    tc2 = ts0 - tc0;

    ts0 = inout[pos + 5 * 3] * WS(0) + inout[pos + 2 * 3];
    tc0 = inout[pos + 3 * 3] * WS(2) + inout[pos + 0 * 3];
    ts1 = ts0 + tc0;
    ts2 = -ts0 + tc0;

    tc0 = (inout[pos + 1 * 3] * WS(1) - inout[pos + 4 * 3]) * 2.069978111953089e-11;
    ts0 = (inout[pos + 4 * 3] * WS(1) + inout[pos + 1 * 3]) * 2.069978111953089e-11;

    inout[pos + 3 * 0] = tc1 * 1.907525191737280e-11 + tc0;
    inout[pos + 3 * 5] = -ts1 * 1.907525191737280e-11 + ts0;

    tc2 = tc2 * 0.86602540378443870761 * 1.907525191737281e-11;
    ts1 = ts1 * 0.5 * 1.907525191737281e-11 + ts0;
    inout[pos + 3 * 1] = tc2 - ts1;
    inout[pos + 3 * 2] = tc2 + ts1;

    tc1 = tc1 * 0.5 * 1.907525191737281e-11 - tc0;
    ts2 = ts2 * 0.86602540378443870761 * 1.907525191737281e-11;
    inout[pos + 3 * 3] = tc1 + ts2;
    inout[pos + 3 * 4] = tc1 - ts2;
    pos++;
  }
}

#define CX(i) WS(12 + (i))
static void mdct_long(F32* out, int op, const F32* in) {
  double ct, st;
  {
    double tc1, tc2, tc3, tc4, ts5, ts6, ts7, ts8;
    tc1 = in[17] - in[9];
    tc3 = in[15] - in[11];
    tc4 = in[14] - in[12];
    ts5 = in[0] + in[8];
    ts6 = in[1] + in[7];
    ts7 = in[2] + in[6];
    ts8 = in[3] + in[5];

    out[op + 17] = (ts5 + ts7 - ts8) - (ts6 - in[4]);
    st = (ts5 + ts7 - ts8) * CX(7) + (ts6 - in[4]);
    ct = (tc1 - tc3 - tc4) * CX(6);
    out[op + 5] = ct + st;
    out[op + 6] = ct - st;

    tc2 = (in[16] - in[10]) * CX(6);
    ts6 = ts6 * CX(7) + in[4];
    ct = tc1 * CX(0) + tc2 + tc3 * CX(1) + tc4 * CX(2);
    st = -ts5 * CX(4) + ts6 - ts7 * CX(5) + ts8 * CX(3);
    out[op + 1] = ct + st;
    out[op + 2] = ct - st;

    ct = tc1 * CX(1) - tc2 - tc3 * CX(2) + tc4 * CX(0);
    st = -ts5 * CX(5) + ts6 - ts7 * CX(3) + ts8 * CX(4);
    out[op + 9] = ct + st;
    out[op + 10] = ct - st;

    ct = tc1 * CX(2) - tc2 + tc3 * CX(0) - tc4 * CX(1);
    st = ts5 * CX(3) - ts6 + ts7 * CX(4) - ts8 * CX(5);
    out[op + 13] = ct + st;
    out[op + 14] = ct - st;
  }
  {
    double ts1, ts2, ts3, ts4, tc5, tc6, tc7, tc8;
    ts1 = in[8] - in[0];
    ts3 = in[6] - in[2];
    ts4 = in[5] - in[3];
    tc5 = in[17] + in[9];
    tc6 = in[16] + in[10];
    tc7 = in[15] + in[11];
    tc8 = in[14] + in[12];

    out[op + 0] = (tc5 + tc7 + tc8) + (tc6 + in[13]);
    ct = (tc5 + tc7 + tc8) * CX(7) - (tc6 + in[13]);
    st = (ts1 - ts3 + ts4) * CX(6);
    out[op + 11] = ct + st;
    out[op + 12] = ct - st;

    ts2 = (in[7] - in[1]) * CX(6);
    tc6 = in[13] - tc6 * CX(7);
    ct = tc5 * CX(3) - tc6 + tc7 * CX(4) + tc8 * CX(5);
    st = ts1 * CX(2) + ts2 + ts3 * CX(0) + ts4 * CX(1);
    out[op + 3] = ct + st;
    out[op + 4] = ct - st;

    ct = -tc5 * CX(5) + tc6 - tc7 * CX(3) - tc8 * CX(4);
    st = ts1 * CX(1) + ts2 - ts3 * CX(2) - ts4 * CX(0);
    out[op + 7] = ct + st;
    out[op + 8] = ct - st;

    ct = -tc5 * CX(4) + tc6 - tc7 * CX(5) - tc8 * CX(3);
    st = ts1 * CX(0) - ts2 + ts3 * CX(1) - ts4 * CX(2);
    out[op + 15] = ct + st;
    out[op + 16] = ct - st;
  }
}

void lj_mdct_sub48(LjEnc* e, const F32* w0, const F32* w1) {
  const F32* wk = w0;
  int wkPos = 286;
  for (int ch = 0; ch < e->channels_out; ch++) {
    for (int gr = 0; gr < e->mode_gr; gr++) {
      GrInfo* gi = &e->tt[gr][ch];
      F32* mdct_enc = gi->xr;
      int mp = 0;
      F32(*samp)[SBLIMIT] = e->sb_sample[ch][1 - gr];
      int sampPos = 0;
      for (int k = 0; k < 18 / 2; k++) {
        window_subband(wk, wkPos, samp[sampPos]);
        window_subband(wk, wkPos + 32, samp[sampPos + 1]);
        sampPos += 2;
        wkPos += 64;
        for (int band = 1; band < 32; band += 2) samp[sampPos - 1][band] *= -1;
      }
      for (int band = 0; band < 32; band++, mp += 18) {
        int type = gi->block_type;
        F32(*band0)[SBLIMIT] = e->sb_sample[ch][gr];
        F32(*band1)[SBLIMIT] = e->sb_sample[ch][1 - gr];
        const int ob = LJ_SB_ORDER[band];
        if (gi->mixed_block_flag != 0 && band < 2) type = 0;
        if (e->amp_filter[band] < 1e-12) {
          for (int k = 0; k < 18; k++) mdct_enc[mp + k] = 0;
        } else {
          if (e->amp_filter[band] < 1.0) {
            for (int k = 0; k < 18; k++) band1[k][ob] *= e->amp_filter[band];
          }
          if (type == SHORT_TYPE) {
            for (int k = -12 / 4; k < 0; k++) {
              double w = WS(k + 3);
              mdct_enc[mp + k * 3 + 9] = band0[9 + k][ob] * w - band0[8 - k][ob];
              mdct_enc[mp + k * 3 + 18] = band0[14 - k][ob] * w + band0[15 + k][ob];
              mdct_enc[mp + k * 3 + 10] = band0[15 + k][ob] * w - band0[14 - k][ob];
              mdct_enc[mp + k * 3 + 19] = band1[2 - k][ob] * w + band1[3 + k][ob];
              mdct_enc[mp + k * 3 + 11] = band1[3 + k][ob] * w - band1[2 - k][ob];
              mdct_enc[mp + k * 3 + 20] = band1[8 - k][ob] * w + band1[9 + k][ob];
            }
            mdct_short(mdct_enc, mp);
          } else {
            F32 work[18];
            for (int k = -36 / 4; k < 0; k++) {
              double a, b;
              a = WIN(type, k + 27) * band1[k + 9][ob] + WIN(type, k + 36) * band1[8 - k][ob];
              b = WIN(type, k + 9) * band0[k + 9][ob] - WIN(type, k + 18) * band0[8 - k][ob];
              work[k + 9] = a - b * WS(3 + k + 9);
              work[k + 18] = a * WS(3 + k + 9) + b;
            }
            mdct_long(mdct_enc, mp, work);
          }
        }
        if (type != SHORT_TYPE && band != 0) {
          for (int k = 7; k >= 0; --k) {
            double bu, bd;
            bu = mdct_enc[mp + k] * WS(20 + k) + mdct_enc[mp + -1 - k] * WS(28 + k);
            bd = mdct_enc[mp + k] * WS(28 + k) - mdct_enc[mp + -1 - k] * WS(20 + k);
            mdct_enc[mp + -1 - k] = bu;
            mdct_enc[mp + k] = bd;
          }
        }
      }
    }
    wk = w1;
    wkPos = 286;
    if (e->mode_gr == 1)   /* NewMDCT.js:1154-1159 */
      for (int i = 0; i < 18; i++) memcpy(e->sb_sample[ch][0][i], e->sb_sample[ch][1][i], sizeof(F32) * 32);
  }
}
