/* oracle/rdoq_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's rate-distortion optimised quantiser for one transform unit,
 * TComTrQuant::xRateDistOptQuant (HM_dl/source/Lib/TLibCommon/TComTrQuant.cpp:2119-2670) with its helpers
 * xGetCodedLevel (:2812), xGetICRate (:2881), xGetRateLast (:2972), getSigCtxInc (:2708), calcPatternSigCtx (:2680),
 * getSigCoeffGroupCtxInc (:3023), the scan tables of TComRom.cpp:116-258 (ScanGenerator) and the context selection of
 * TComChromaFormat.cpp:96-160 / TComChromaFormat.h:233-262, at the reference's operating point (8-bit, no scaling lists,
 * no extended precision, no persistent Rice adaptation).  All cost arithmetic is IEEE double in the reference's order of
 * operations (the decisions compare doubles; nothing here may be re-associated or fused).
 * Pinned by tests/golden/tq_rdoq_192x128_qp32.npz: inputs and outputs of every sampled call, dumped by the reference encoder
 * itself (oracle/_ref/TAppEncoder_rdoqtrace, oracle/rdoq_dump.h; tools/gen_golden_tq.py).
 *
 * est: the reference's estBitsSbacStruct as 224 int32 (TComTrQuant.h:60-75):
 *   [0]   significantCoeffGroupBits[2][2]   [4]   significantBits[44][2]   [92]  lastXBits[2][10]   [112] lastYBits[2][10]
 *   [132] greaterOneBits[24][2]             [180] levelAbsBits[6][2]       [192] blockCbpBits[10][2]  [212] blockRootCbpBits[4][2]
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { E_SIGCG = 0, E_SIG = 4, E_LASTX = 92, E_LASTY = 112, E_G1 = 132, E_ABS = 180, E_CBP = 192, E_ROOT = 212 };
enum { SCAN_DIAG = 0, SCAN_HOR = 1, SCAN_VER = 2 };

static const int kQuantScales[6] = {26214, 23302, 20560, 18396, 16384, 14564};
static const int kInvQuantScales[6] = {40, 45, 51, 57, 64, 72};
static const uint8_t kCtxIndMap4x4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};
static const uint8_t kGroupIdx[32] = {0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9};
static const int kSigSetStart[2][4] = {{0, 9, 21, 27}, {0, 9, 12, 15}};   /* [channel][4x4, 8x8, NxN, single] */

/* scan of a w x h block (raster index with the given stride, offset ox, oy), one of the three HEVC scan types */
static void gen_scan(int type, int w, int h, int stride, int ox, int oy, uint16_t *out) {
  int line = 0, col = 0;
  for (int i = 0; i < w * h; i++) {
    out[i] = (uint16_t)((line + oy) * stride + col + ox);
    if (type == SCAN_DIAG) {
      if (col == w - 1 || line == 0) {
        line += col + 1; col = 0;
        if (line >= h) { col += line - (h - 1); line = h - 1; }
      } else { col++; line--; }
    } else if (type == SCAN_HOR) {
      if (col == w - 1) { line++; col = 0; } else col++;
    } else {
      if (line == h - 1) { col++; line = 0; } else line++;
    }
  }
}
/* coefficient scan grouped by 4x4 coefficient groups + the scan of the groups themselves */
void oracle_rdoq_scans(int type, int log2n, uint16_t *scan, uint16_t *scan_cg) {
  const int n = 1 << log2n, g = n >> 2;
  gen_scan(type, g, g, g, 0, 0, scan_cg);
  for (int k = 0; k < g * g; k++) gen_scan(type, 4, 4, n, (scan_cg[k] % g) * 4, (scan_cg[k] / g) * 4, scan + 16 * k);
}

typedef struct {
  const int32_t *est;
  double lambda;
  int ch;              /* 0 luma, 1 chroma */
} RateModel;

static double icost(const RateModel *m, double rate) { return m->lambda * rate; }

/* bits (<< 15) of coding absolute level `lvl` given the contexts and counters of the current coefficient group */
static int level_rate(const RateModel *m, unsigned lvl, int ctx_one, int ctx_abs, int rice, unsigned c1idx, unsigned c2idx) {
  int rate = 32768;                                                /* sign */
  const unsigned base = c1idx < 8 ? 2 + (c2idx < 1) : 1;
  if (lvl >= base) {
    unsigned sym = lvl - base, len;
    if (sym < (3u << rice)) {
      len = sym >> rice;
      rate += (int)(len + 1 + rice) << 15;
    } else {
      len = rice;
      sym -= 3u << rice;
      while (sym >= (1u << len)) sym -= 1u << (len++);
      rate += (int)(3 + len + 1 - rice + len) << 15;
    }
    if (c1idx < 8) {
      rate += m->est[E_G1 + 2 * ctx_one + 1];
      if (c2idx < 1) rate += m->est[E_ABS + 2 * ctx_abs + 1];
    }
  } else if (lvl == 1) rate += m->est[E_G1 + 2 * ctx_one + 0];
  else if (lvl == 2) rate += m->est[E_G1 + 2 * ctx_one + 1] + m->est[E_ABS + 2 * ctx_abs + 0];
  else rate = 0;
  return rate;
}

static double last_rate(const RateModel *m, unsigned px, unsigned py) {
  const unsigned cx = kGroupIdx[px], cy = kGroupIdx[py];
  double c = m->est[E_LASTX + 10 * m->ch + cx] + m->est[E_LASTY + 10 * m->ch + cy];
  if (cx > 3) c += 32768.0 * ((cx - 2) >> 1);
  if (cy > 3) c += 32768.0 * ((cy - 2) >> 1);
  return icost(m, c);
}

/* context increment of significant_coeff_flag at raster position pos */
static int sig_ctx_inc(int pattern, int first_ctx, int pos, int log2n, int ch) {
  if (first_ctx == kSigSetStart[ch][3]) return first_ctx;
  const int py = pos >> log2n, px = pos - (py << log2n);
  if (px + py == 0) return 0;
  int off;
  if (log2n == 2) off = kCtxIndMap4x4[4 * py + px];
  else {
    int cnt;
    const int xs = px & 3, ys = py & 3;
    if (pattern == 0) cnt = (xs + ys >= 3) ? 0 : ((xs + ys >= 1) ? 1 : 2);
    else if (pattern == 1) cnt = (ys >= 2) ? 0 : ((ys >= 1) ? 1 : 2);
    else if (pattern == 2) cnt = (xs >= 2) ? 0 : ((xs >= 1) ? 1 : 2);
    else cnt = 2;
    const int not_first = ((px >> 2) + (py >> 2)) > 0;
    off = (not_first ? (ch == 0 ? 3 : 0) : 0) + cnt;
  }
  return first_ctx + off;
}

static int ctx_set_index(int ch, int subset, int found_gt1) { return (ch == 0 ? 0 : 4) + ((ch == 0 && subset > 0) ? 2 : 0) + (found_gt1 ? 1 : 0); }

/* One TU.  coeff: transform output (row-major n x n).  Returns uiAbsSum; level_out receives the signed levels. */
uint32_t oracle_rdoq(const int32_t *coeff, int log2n, int ch, int scan_type, int qp, int tskip, double lambda, const int32_t *est,
                     int ctx_cbf, int is_intra, int tr_idx_zero, int sdh, int32_t *level_out) {
  const int n = 1 << log2n, n2 = n * n, ncg = n2 >> 4, wg = n >> 2;
  static __thread uint16_t scan[1024], scan_cg[64];
  static __thread double cost_coeff[1024], cost_sig[1024], cost_coeff0[1024];
  static __thread int rate_up[1024], rate_down[1024], sig_delta[1024], delta_u[1024];
  oracle_rdoq_scans(scan_type, log2n, scan, scan_cg);
  RateModel rm = {est, lambda, ch};
  int tshift = 15 - 8 - log2n;
  (void)tskip;                                   /* transform skip changes the shift only with extended precision (off) */
  const int per = qp / 6, rem = qp % 6, qbits = 14 + per + tshift;
  const int qscale = kQuantScales[rem];
  double err_scale = (double)(1 << 15);          /* SCALE_BITS, then the forward-transform scaling 2^(-2 shift) */
  for (int i = 0; i < 2 * tshift; i++) err_scale = err_scale / 2.0;      /* pow(2, -2*shift): exact */
  for (int i = 0; i < -2 * tshift; i++) err_scale = err_scale * 2.0;
  err_scale = err_scale / qscale / qscale / 1;
  int first_ctx;                                 /* getTUEntropyCodingParameters */
  if (n == 4) first_ctx = kSigSetStart[ch][0];
  else if (n == 8) first_ctx = kSigSetStart[ch][1] + (scan_type != SCAN_DIAG ? (ch == 0 ? 6 : 0) : 0);
  else first_ctx = kSigSetStart[ch][2];
  const int sig_off = ch == 0 ? 0 : 28;

  memset(cost_coeff, 0, sizeof(double) * n2); memset(cost_sig, 0, sizeof(double) * n2);
  memset(rate_up, 0, sizeof(int) * n2); memset(rate_down, 0, sizeof(int) * n2);
  memset(sig_delta, 0, sizeof(int) * n2); memset(delta_u, 0, sizeof(int) * n2);
  double cost_cg_sig[64];
  unsigned cg_flag[64];
  memset(cost_cg_sig, 0, sizeof cost_cg_sig); memset(cg_flag, 0, sizeof cg_flag);
  memset(level_out, 0, sizeof(int32_t) * n2);

  double uncoded = 0, base_cost = 0;
  int last_pos = -1, cg_last = -1;
  unsigned ctx_set = 0, c1idx = 0, c2idx = 0;
  int c1 = 1, c2 = 0, rice = 0;

  for (int cg = ncg - 1; cg >= 0; cg--) {
    const int cg_blk = scan_cg[cg], cgy = cg_blk / wg, cgx = cg_blk - cgy * wg;
    int nnz_before0 = 0;
    double st_coded = 0, st_uncoded = 0, st_sig = 0, st_sig0 = 0;
    int pattern = 0;
    if (wg > 1) {
      const int r = cgx < wg - 1 ? (cg_flag[cgy * wg + cgx + 1] != 0) : 0, b = cgy < wg - 1 ? (cg_flag[(cgy + 1) * wg + cgx] != 0) : 0;
      pattern = r + (b << 1);
    }
    for (int k = 15; k >= 0; k--) {
      const int sp = cg * 16 + k, bp = scan[sp];
      const int64_t t = (int64_t)abs(coeff[bp]) * qscale;
      const int64_t cap = (int64_t)0x7fffffff - ((int64_t)1 << (qbits - 1));
      const int32_t lvl_d = (int32_t)(t < cap ? t : cap);
      unsigned max_abs = (unsigned)((lvl_d + (1 << (qbits - 1))) >> qbits);
      if (max_abs > 32767u) max_abs = 32767u;
      const double e0 = (double)lvl_d;
      cost_coeff0[sp] = e0 * e0 * err_scale;
      uncoded += cost_coeff0[sp];
      level_out[bp] = (int32_t)max_abs;
      if (max_abs > 0 && last_pos < 0) { last_pos = sp; ctx_set = ctx_set_index(ch, sp >> 4, 0); cg_last = cg; }
      if (last_pos >= 0) {
        const int ctx_one = 4 * ctx_set + c1, ctx_abs = ctx_set + c2;
        const int is_last = sp == last_pos;
        int ctx_sig = 0;
        if (!is_last) ctx_sig = sig_off + sig_ctx_inc(pattern, first_ctx, bp, log2n, ch);
        /* best level among {max_abs, max_abs - 1 (, 0)} */
        unsigned best = 0;
        double cur_sig = 0;
        int decided = 0;
        if (!is_last && max_abs < 3) {
          cost_sig[sp] = icost(&rm, est[E_SIG + 2 * ctx_sig + 0]);
          cost_coeff[sp] = cost_coeff0[sp] + cost_sig[sp];
          if (max_abs == 0) decided = 1;
        } else cost_coeff[sp] = 1.7976931348623158e+308;
        if (!decided) {
          if (!is_last) cur_sig = icost(&rm, est[E_SIG + 2 * ctx_sig + 1]);
          const unsigned min_abs = max_abs > 1 ? max_abs - 1 : 1;
          for (int a = (int)max_abs; a >= (int)min_abs; a--) {
            const double e = (double)(lvl_d - (int32_t)((uint32_t)a << qbits));
            double c = e * e * err_scale + icost(&rm, level_rate(&rm, (unsigned)a, ctx_one, ctx_abs, rice, c1idx, c2idx));
            c += cur_sig;
            if (c < cost_coeff[sp]) { best = (unsigned)a; cost_coeff[sp] = c; cost_sig[sp] = cur_sig; }
          }
        }
        if (!is_last) sig_delta[bp] = est[E_SIG + 2 * ctx_sig + 1] - est[E_SIG + 2 * ctx_sig + 0];
        delta_u[bp] = (int)((lvl_d - (int32_t)((uint32_t)best << qbits)) >> (qbits - 8));
        if (best > 0) {
          const int now = level_rate(&rm, best, ctx_one, ctx_abs, rice, c1idx, c2idx);
          rate_up[bp] = level_rate(&rm, best + 1, ctx_one, ctx_abs, rice, c1idx, c2idx) - now;
          rate_down[bp] = level_rate(&rm, best - 1, ctx_one, ctx_abs, rice, c1idx, c2idx) - now;
        } else rate_up[bp] = est[E_G1 + 2 * ctx_one + 0];
        level_out[bp] = (int32_t)best;
        base_cost += cost_coeff[sp];
        const unsigned base_level = c1idx < 8 ? 2 + (c2idx < 1) : 1;
        if (best >= base_level && best > 3u * (1u << rice)) rice = rice + 1 < 4 ? rice + 1 : 4;
        if (best >= 1) c1idx++;
        if (best > 1) { c1 = 0; c2 += c2 < 2; c2idx++; }
        else if (c1 < 3 && c1 > 0 && best) c1++;
        if ((sp & 15) == 0 && sp > 0) {            /* entering the next coefficient group */
          ctx_set = ctx_set_index(ch, (sp - 1) >> 4, c1 == 0);
          c1 = 1; c2 = 0; c1idx = 0; c2idx = 0; rice = 0;
        }
      } else base_cost += cost_coeff0[sp];
      st_sig += cost_sig[sp];
      if (k == 0) st_sig0 = cost_sig[sp];
      if (level_out[bp]) {
        cg_flag[cg_blk] = 1;
        st_coded += cost_coeff[sp] - cost_sig[sp];
        st_uncoded += cost_coeff0[sp];
        if (k != 0) nnz_before0++;
      }
    }
    if (cg_last >= 0) {
      if (cg) {
        int sr = cgx < wg - 1 ? (cg_flag[cgy * wg + cgx + 1] != 0) : 0, sb = cgy < wg - 1 ? (cg_flag[(cgy + 1) * wg + cgx] != 0) : 0;
        const int cctx = (sr + sb) != 0;
        if (cg_flag[cg_blk] == 0) {
          base_cost += icost(&rm, est[E_SIGCG + 2 * cctx + 0]) - st_sig;
          cost_cg_sig[cg] = icost(&rm, est[E_SIGCG + 2 * cctx + 0]);
        } else if (cg < cg_last) {
          if (nnz_before0 == 0) { base_cost -= st_sig0; st_sig -= st_sig0; }
          double zero_cost = base_cost;
          base_cost += icost(&rm, est[E_SIGCG + 2 * cctx + 1]);
          zero_cost += icost(&rm, est[E_SIGCG + 2 * cctx + 0]);
          cost_cg_sig[cg] = icost(&rm, est[E_SIGCG + 2 * cctx + 1]);
          zero_cost += st_uncoded;
          zero_cost -= st_coded;
          zero_cost -= st_sig;
          if (zero_cost < base_cost) {
            cg_flag[cg_blk] = 0;
            base_cost = zero_cost;
            cost_cg_sig[cg] = icost(&rm, est[E_SIGCG + 2 * cctx + 0]);
            for (int k = 15; k >= 0; k--) {
              const int sp = cg * 16 + k, bp = scan[sp];
              if (level_out[bp]) { level_out[bp] = 0; cost_coeff[sp] = cost_coeff0[sp]; cost_sig[sp] = 0; }
            }
          }
        }
      } else cg_flag[cg_blk] = 1;
    }
  }
  if (last_pos < 0) return 0;

  /* position of the last coded coefficient */
  double best_cost;
  int best_last_p1 = 0;
  if (!is_intra && ch == 0 && tr_idx_zero) {
    best_cost = uncoded + icost(&rm, est[E_ROOT + 0]);
    base_cost += icost(&rm, est[E_ROOT + 1]);
  } else {
    best_cost = uncoded + icost(&rm, est[E_CBP + 2 * ctx_cbf + 0]);
    base_cost += icost(&rm, est[E_CBP + 2 * ctx_cbf + 1]);
  }
  int found = 0;
  for (int cg = cg_last; cg >= 0 && !found; cg--) {
    const int cg_blk = scan_cg[cg];
    base_cost -= cost_cg_sig[cg];
    if (!cg_flag[cg_blk]) continue;
    for (int k = 15; k >= 0; k--) {
      const int sp = cg * 16 + k;
      if (sp > last_pos) continue;
      const int bp = scan[sp];
      if (level_out[bp]) {
        const unsigned py = (unsigned)bp >> log2n, px = (unsigned)bp - (py << log2n);
        const double cl = scan_type == SCAN_VER ? last_rate(&rm, py, px) : last_rate(&rm, px, py);
        const double total = base_cost + cl - cost_sig[sp];
        if (total < best_cost) { best_last_p1 = sp + 1; best_cost = total; }
        if (level_out[bp] > 1) { found = 1; break; }
        base_cost -= cost_coeff[sp];
        base_cost += cost_coeff0[sp];
      } else base_cost -= cost_sig[sp];
    }
  }
  uint32_t abs_sum = 0;
  for (int sp = 0; sp < best_last_p1; sp++) {
    const int bp = scan[sp];
    const int32_t l = level_out[bp];
    abs_sum += (uint32_t)l;
    level_out[bp] = coeff[bp] < 0 ? -l : l;
  }
  for (int sp = best_last_p1; sp <= last_pos; sp++) level_out[scan[sp]] = 0;

  /* sign-bit hiding inside RDOQ (:2520-2668) */
  if (sdh && abs_sum >= 2) {
    const double iq = (double)kInvQuantScales[rem];
    const int64_t rd_factor = (int64_t)(iq * iq * (1 << (2 * per)) / lambda / 16 / 1 + 0.5);
    int last_cg = -1;
    for (int sub = (n2 - 1) >> 4; sub >= 0; sub--) {
      const int sp0 = sub << 4;
      int first_nz = 16, last_nz = -1, asum = 0, k;
      for (k = 15; k >= 0; --k) if (level_out[scan[k + sp0]]) { last_nz = k; break; }
      for (k = 0; k < 16; k++) if (level_out[scan[k + sp0]]) { first_nz = k; break; }
      for (k = first_nz; k <= last_nz; k++) asum += level_out[scan[k + sp0]];
      if (last_nz >= 0 && last_cg == -1) last_cg = 1;
      if (last_nz - first_nz >= 4) {
        const unsigned signbit = level_out[scan[sp0 + first_nz]] > 0 ? 0 : 1;
        if (signbit != (unsigned)(asum & 1)) {
          int64_t min_inc = INT64_MAX, cur = INT64_MAX;
          int min_pos = -1, final_change = 0, cur_change = 0;
          for (k = (last_cg == 1 ? last_nz : 15); k >= 0; --k) {
            const int bp = scan[k + sp0];
            if (level_out[bp] != 0) {
              const int64_t up = rd_factor * (-delta_u[bp]) + rate_up[bp];
              int64_t down = rd_factor * (delta_u[bp]) + rate_down[bp] - ((abs(level_out[bp]) == 1) ? sig_delta[bp] : 0);
              if (last_cg == 1 && last_nz == k && abs(level_out[bp]) == 1) down -= 4 << 15;
              if (up < down) { cur = up; cur_change = 1; }
              else {
                cur_change = -1;
                cur = (k == first_nz && abs(level_out[bp]) == 1) ? INT64_MAX : down;
              }
            } else {
              cur = rd_factor * (-(abs(delta_u[bp]))) + (1 << 15) + rate_up[bp] + sig_delta[bp];
              cur_change = 1;
              if (k < first_nz) {
                const unsigned s = coeff[bp] >= 0 ? 0 : 1;
                if (s != signbit) cur = INT64_MAX;
              }
            }
            if (cur < min_inc) { min_inc = cur; final_change = cur_change; min_pos = bp; }
          }
          if (level_out[min_pos] == 32767 || level_out[min_pos] == -32768) final_change = -1;
          if (coeff[min_pos] >= 0) level_out[min_pos] += final_change;
          else level_out[min_pos] -= final_change;
        }
      }
      if (last_cg == 1) last_cg = 0;
    }
  }
  return abs_sum;
}
