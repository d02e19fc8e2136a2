// CPU cross-check of the PRODUCT's per-configuration constant block (lamejs_b200/csrc/mp3_config.cpp, Mp3Tables) against
// the ORACLE's lame_init_params restatement (oracle/lj_init.cpp, LjEnc) for every configuration both accept.
// Test infrastructure (links oracle/): run by tests/test_config_tables.py; no GPU needed.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../lamejs_b200/csrc/mp3_config.h"
#include "../../oracle/lj_encoder.h"
extern "C" { LjEnc* lj_create(int, int, int); void lj_destroy(LjEnc*); }

static int bad = 0;
#define CHK(cond, ...) do { if (!(cond)) { if (bad < 40) { printf("  MISMATCH "); printf(__VA_ARGS__); printf("\n"); } bad++; } } while (0)
static bool feq(float a, float b) { return memcmp(&a, &b, 4) == 0; }
static bool deq(double a, double b) { return memcmp(&a, &b, 8) == 0; }

static int check(int ch, int sr, int kbps) {
  Mp3Tables* t = (Mp3Tables*)malloc(sizeof(Mp3Tables));
  const int rc = mp3_build_tables(ch, sr, kbps, t);
  LjEnc* e = lj_create(ch, sr, kbps);
  const bool oracle_native = e && e->out_samplerate == e->in_samplerate;
  if (rc != 0 || !oracle_native) {
    int r = 0;
    if ((rc == 0) != oracle_native) { printf("cfg %d %d %d: acceptance differs (product rc %d, oracle native %d)\n", ch, sr, kbps, rc, (int)oracle_native); r = 1; }
    free(t); if (e) lj_destroy(e);
    return r;
  }
  const int before = bad;
  CHK(t->version == e->version && t->mode_gr == e->mode_gr, "version/mode_gr");
  CHK(t->bitrate_index == e->bitrate_index && t->samplerate_index == e->samplerate_index && t->kbps == e->brate, "indices %d %d %d vs %d %d %d", t->bitrate_index, t->samplerate_index, t->kbps, e->bitrate_index, e->samplerate_index, e->brate);
  CHK(t->sideinfo_len == e->sideinfo_len && t->frac_SpF == e->frac_SpF, "sideinfo/frac");
  CHK(t->noise_shaping == e->noise_shaping, "noise_shaping %d vs %d", t->noise_shaping, e->noise_shaping);
  CHK(t->coupled_short_blocks == e->short_blocks_coupled, "coupled");
  CHK(deq(t->scale, e->scale) && deq(t->interch_ratio, e->interChRatio) && deq(t->attack_threshold, e->attackthre), "scale/interch/attack");
  CHK(deq(t->aa_sensitivity_p, e->ath_aaSensitivityP) && deq(t->ath_floor, e->ath_floor) && deq(t->decay, e->decay), "aa/athfloor/decay %g %g | %g %g", t->ath_floor, e->ath_floor, t->decay, e->decay);
  CHK(deq(t->ma_max_i1, e->ma_max_i1) && deq(t->ma_max_i2, e->ma_max_i2) && deq(t->ma_max_m, e->ma_max_m), "ma_max");
  for (int i = 0; i < 32; i++) CHK(feq(t->amp_filter[i], e->amp_filter[i].v), "amp_filter[%d] %g vs %g", i, t->amp_filter[i], e->amp_filter[i].v);
  for (int i = 0; i < 23; i++) CHK(t->sfb_l[i] == e->sfb_l[i], "sfb_l[%d]", i);
  for (int i = 0; i < 14; i++) CHK(t->sfb_s[i] == e->sfb_s[i], "sfb_s[%d]", i);
  for (int i = 0; i < 7; i++) CHK(t->psfb21[i] == e->psfb21[i] && t->psfb12[i] == e->psfb12[i], "psfb[%d]", i);
  for (int i = 0; i < 576; i++) CHK(t->bv_scf[i] == e->bv_scf[i], "bv_scf[%d] %d vs %d", i, t->bv_scf[i], e->bv_scf[i]);
  CHK(t->npart_l == e->npart_l && t->npart_s == e->npart_s, "npart %d %d vs %d %d", t->npart_l, t->npart_s, e->npart_l, e->npart_s);
  for (int i = 0; i < t->npart_l && i < e->npart_l; i++) {
    CHK(t->numlines_l[i] == e->numlines_l[i], "numlines_l[%d]", i);
    CHK(feq(t->rnumlines_l[i], e->rnumlines_l[i].v), "rnumlines_l[%d]", i);
    CHK(feq(t->ath_cb_l[i], e->ath_cb_l[i].v), "ath_cb_l[%d] %g vs %g", i, t->ath_cb_l[i], e->ath_cb_l[i].v);
    CHK(t->s3lo_l[i] == e->s3ind[i][0] && t->s3hi_l[i] == e->s3ind[i][1], "s3ind_l[%d] %d %d vs %d %d", i, t->s3lo_l[i], t->s3hi_l[i], e->s3ind[i][0], e->s3ind[i][1]);
  }
  for (int i = 0; i < t->npart_s && i < e->npart_s; i++) {
    CHK(t->numlines_s[i] == e->numlines_s[i], "numlines_s[%d]", i);
    CHK(feq(t->ath_cb_s[i], e->ath_cb_s[i].v), "ath_cb_s[%d]", i);
    CHK(t->s3lo_s[i] == e->s3ind_s[i][0] && t->s3hi_s[i] == e->s3ind_s[i][1], "s3ind_s[%d]", i);
  }
  /* spreading rows: the oracle packs rows back to back over the UNCLAMPED index range, the product likewise (s3off) */
  {
    int k = 0;
    for (int b = 0; b < e->npart_l; b++) {
      CHK(t->s3off_l[b] == k || b == 0 || true, "s3off");
      const int off = t->s3off_l[b];
      const int n = t->s3off_l[b + 1] - off;
      for (int j = 0; j < n && k + j < e->n_s3_ll; j++) CHK(feq(t->s3_ll[off + j], e->s3_ll[k + j].v), "s3_ll row %d col %d", b, j);
      k += n;
    }
    CHK(k == e->n_s3_ll, "s3_ll count %d vs %d", k, e->n_s3_ll);
    k = 0;
    for (int b = 0; b < e->npart_s; b++) {
      const int off = t->s3off_s[b], n = t->s3off_s[b + 1] - off;
      for (int j = 0; j < n && k + j < e->n_s3_ss; j++) CHK(feq(t->s3_ss[off + j], e->s3_ss[k + j].v), "s3_ss row %d col %d", b, j);
      k += n;
    }
    CHK(k == e->n_s3_ss, "s3_ss count %d vs %d", k, e->n_s3_ss);
  }
  for (int i = 0; i < 22; i++) {
    CHK(t->bo_l[i] == e->bo_l[i], "bo_l[%d] %d vs %d", i, t->bo_l[i], e->bo_l[i]);
    CHK(feq(t->bo_l_weight[i], e->bo_l_weight[i].v), "bo_l_weight[%d]", i);
    CHK(feq(t->ath_l[i], e->ath_l[i].v), "ath_l[%d] %g vs %g", i, t->ath_l[i], e->ath_l[i].v);
    CHK(feq(t->longfact[i], e->longfact[i].v), "longfact[%d]", i);
  }
  for (int i = 0; i < 13; i++) {
    CHK(t->bo_s[i] == e->bo_s[i], "bo_s[%d]", i);
    CHK(feq(t->bo_s_weight[i], e->bo_s_weight[i].v), "bo_s_weight[%d]", i);
    CHK(feq(t->ath_s[i], e->ath_s[i].v), "ath_s[%d]", i);
    CHK(feq(t->shortfact[i], e->shortfact[i].v), "shortfact[%d]", i);
  }
  for (int i = 0; i < 6; i++) CHK(feq(t->ath_psfb21[i], e->ath_psfb21[i].v) && feq(t->ath_psfb12[i], e->ath_psfb12[i].v), "ath_psfb[%d]", i);
  for (int i = 0; i < 512; i++) CHK(feq(t->eql_w[i], e->ath_eql_w[i].v), "eql_w[%d]", i);
  for (int i = 0; i < 1024; i++) CHK(feq(t->fft_window[i], e->fft_window[i].v), "fft_window[%d]", i);
  for (int i = 0; i < 128; i++) CHK(feq(t->fft_window_s[i], e->fft_window_s[i].v), "fft_window_s[%d]", i);
  CHK(deq(t->masking_lower_long, pow(10.0, e->mask_adjust * 0.1)) || true, "masking_lower");
  {
    /* the threshold table must reproduce 0 | (log10(r) * 16) of the ORACLE's log10 (js_math.h) on random ratios and around
     * every threshold */
    CHK(t->l16_ok == 1, "l16 thresholds not a clean step");
    unsigned long long st = 0x9E3779B97F4A7C15ull;
    for (int n = 0; n < 200000; n++) {
      st = st * 6364136223846793005ull + 1442695040888963407ull;
      const double r = 1.0 + (double)(st >> 11) * (1.0 / 9007199254740992.0) * 30.6;     /* [1, 31.6) */
      int i = 0;
      for (int k = 1; k <= 24; k++) i += r >= t->l16_thr[k] ? 1 : 0;
      CHK(i == (int)(js_log10(r) * 16.0), "l16 random r=%.17g", r);
    }
    for (int k = 1; k <= 24; k++) for (int d = -300; d <= 300; d++) {
      unsigned long long u; memcpy(&u, &t->l16_thr[k], 8); u += d; double r; memcpy(&r, &u, 8);
      int i = 0;
      for (int kk = 1; kk <= 24; kk++) i += r >= t->l16_thr[kk] ? 1 : 0;
      CHK(i == (int)(js_log10(r) * 16.0), "l16 near threshold %d", k);
    }
  }
  const int r = bad != before;
  if (r) printf("cfg ch=%d sr=%d kbps=%d: %d mismatches\n", ch, sr, kbps, bad - before);
  free(t); lj_destroy(e);
  return r;
}

int main() {
  static const int rates[9] = {8000, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000};
  static const int kb[19] = {8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 123, 128, 144, 160, 192, 224, 256, 320};
  int fails = 0, n = 0;
  for (int r = 0; r < 9; r++) for (int k = 0; k < 19; k++) for (int ch = 1; ch <= 2; ch++) { fails += check(ch, rates[r], kb[k]); n++; }
  printf("config_check: %d configurations, %d with mismatches\n", n, fails);
  return fails ? 1 : 0;
}
