// tq_rdoq.cuh -- rate-distortion optimised quantisation of one transform unit on the device, bit-exact with the
// reference's TComTrQuant::xRateDistOptQuant (HM TLibCommon/TComTrQuant.cpp:2119-2670; helpers :2680-3050, scans
// TComRom.cpp:116-258, context selection TComChromaFormat.cpp:96-160) at the reference's operating point.
//
// The reference decides every level by comparing double-precision costs built in a fixed order of operations, so every
// double operation here is an explicit round-to-nearest intrinsic (__dmul_rn / __dadd_rn / __dsub_rn / __ddiv_rn): the
// compiler must not contract a*b+c into an FMA, which rounds once where the reference rounds twice.
// The coefficient-by-coefficient context state (c1, c2, Rice parameter, last position) makes the level decision serial
// inside a TU: one lane walks the scan, the warp shares the data-parallel preparation; TUs are independent, one per warp.
// Parity first -- this is not on the benchmarked step.
#pragma once
#include "common.cuh"

namespace hevcdl {

// offsets (int32) into the reference's estBitsSbacStruct (HM TLibCommon/TComTrQuant.h:60-75)
enum : int { E_SIGCG = 0, E_SIG = 4, E_LASTX = 92, E_LASTY = 112, E_G1 = 132, E_ABS = 180, E_CBP = 192, E_ROOT = 212, EST_INTS = 224 };

struct RdoqScratch {                      // per warp, global memory
  double cost_coeff[1024], cost_sig[1024], cost_coeff0[1024], cost_cg_sig[64];
  int rate_up[1024], rate_down[1024], sig_delta[1024], delta_u[1024], level[1024], lvl_d[1024];
  uint32_t cg_flag[64];
  uint16_t scan[1024], scan_cg[64];
};

__device__ __constant__ uint8_t c_ctx_ind_map_4x4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};
__device__ __constant__ uint8_t c_group_idx[32] = {0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9};
__device__ __constant__ uint8_t c_sig_set_start[2][4] = {{0, 9, 21, 27}, {0, 9, 12, 15}};

// scan of a w x h block: position i of the walk of ScanGenerator (TComRom.cpp:116-198)
__device__ inline void rdoq_gen_scan(int type, int w, int h, int stride, int ox, int oy, uint16_t *out) {
  int line = 0, col = 0;
  for (int i = 0; i < w * h; i++) {
    out[i] = (uint16_t)((line + oy) * stride + col + ox);
    if (type == 0) {
      if (col == w - 1 || line == 0) {
        line += col + 1; col = 0;
        if (line >= h) { col += line - (h - 1); line = h - 1; }
      } else { col++; line--; }
    } else if (type == 1) {
      if (col == w - 1) { line++; col = 0; } else col++;
    } else {
      if (line == h - 1) { col++; line = 0; } else line++;
    }
  }
}

struct RdoqRate {
  const int *est;
  double lambda;
  int ch;
  __device__ __forceinline__ double icost(double rate) const { return __dmul_rn(lambda, rate); }
  __device__ int level_rate(unsigned lvl, int ctx_one, int ctx_abs, int rice, unsigned c1idx, unsigned c2idx) const {
    int rate = 32768;
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
        rate += est[E_G1 + 2 * ctx_one + 1];
        if (c2idx < 1) rate += est[E_ABS + 2 * ctx_abs + 1];
      }
    } else if (lvl == 1) rate += est[E_G1 + 2 * ctx_one + 0];
    else if (lvl == 2) rate += est[E_G1 + 2 * ctx_one + 1] + est[E_ABS + 2 * ctx_abs + 0];
    else rate = 0;
    return rate;
  }
  __device__ double last_rate(unsigned px, unsigned py) const {
    const unsigned cx = c_group_idx[px], cy = c_group_idx[py];
    double c = (double)(est[E_LASTX + 10 * ch + cx] + est[E_LASTY + 10 * ch + cy]);
    if (cx > 3) c = __dadd_rn(c, __dmul_rn(32768.0, (double)((cx - 2) >> 1)));
    if (cy > 3) c = __dadd_rn(c, __dmul_rn(32768.0, (double)((cy - 2) >> 1)));
    return icost(c);
  }
};

__device__ __forceinline__ int rdoq_sig_ctx_inc(int pattern, int first_ctx, int pos, int log2n, int ch) {
  if (first_ctx == c_sig_set_start[ch][3]) return first_ctx;
  const int py = pos >> log2n, px = pos - (py << log2n);
  if (px + py == 0) return 0;
  int off;
  if (log2n == 2) off = c_ctx_ind_map_4x4[4 * py + px];
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
__device__ __forceinline__ int rdoq_ctx_set(int ch, int subset, int found_gt1) {
  return (ch == 0 ? 0 : 4) + ((ch == 0 && subset > 0) ? 2 : 0) + (found_gt1 ? 1 : 0);
}

// One TU by one warp.  coeff: the transform output (row-major, shared memory); the signed levels land in S.level; returns
// uiAbsSum in every lane.  rflags: bit0 sign-bit hiding, bit1 intra, bit2 transform index == 0.
__device__ inline uint32_t rdoq_tu(const int32_t *__restrict__ coeff, int log2n, int ch, int scan_type, int qp, double lambda,
                                   const int *__restrict__ est, int ctx_cbf, int rflags, RdoqScratch &S, int lane) {
  const int n = 1 << log2n, n2 = n * n, ncg = n2 >> 4, wg = n >> 2;
  const int tshift = 15 - 8 - log2n;
  const int per = qp / 6, rem = qp - 6 * per, qbits = 14 + per + tshift;
  const int qscale = c_tq_qscale[rem];
  // ---- data-parallel preparation: scans, zeroed tables, the unquantised level and the all-zero cost of every position ----
  if (lane == 0) rdoq_gen_scan(scan_type, wg, wg, wg, 0, 0, S.scan_cg);
  __syncwarp();
  for (int k = lane; k < ncg; k += 32) rdoq_gen_scan(scan_type, 4, 4, n, (S.scan_cg[k] % wg) * 4, (S.scan_cg[k] / wg) * 4, S.scan + 16 * k);
  double err_scale = 32768.0;                    // SCALE_BITS, then 2^(-2 shift) (exact), then / scale / scale
  for (int i = 0; i < 2 * tshift; i++) err_scale = __dmul_rn(err_scale, 0.5);
  for (int i = 0; i < -2 * tshift; i++) err_scale = __dmul_rn(err_scale, 2.0);
  err_scale = __ddiv_rn(__ddiv_rn(err_scale, (double)qscale), (double)qscale);
  __syncwarp();
  for (int sp = lane; sp < n2; sp += 32) {
    const int bp = S.scan[sp];
    const long long t = (long long)abs(coeff[bp]) * qscale;
    const long long cap = 0x7fffffffLL - (1LL << (qbits - 1));
    const int ld = (int)(t < cap ? t : cap);
    S.lvl_d[sp] = ld;
    const double e0 = (double)ld;
    S.cost_coeff0[sp] = __dmul_rn(__dmul_rn(e0, e0), err_scale);
    S.cost_coeff[sp] = 0.0; S.cost_sig[sp] = 0.0;
    S.rate_up[sp] = 0; S.rate_down[sp] = 0; S.sig_delta[sp] = 0; S.delta_u[sp] = 0; S.level[sp] = 0;
  }
  for (int k = lane; k < 64; k += 32) { S.cost_cg_sig[k] = 0.0; S.cg_flag[k] = 0; }
  __syncwarp();
  uint32_t abs_sum = 0;
  if (lane == 0) {
    RdoqRate rm{est, lambda, ch};
    int first_ctx;
    if (n == 4) first_ctx = c_sig_set_start[ch][0];
    else if (n == 8) first_ctx = c_sig_set_start[ch][1] + (scan_type != 0 ? (ch == 0 ? 6 : 0) : 0);
    else first_ctx = c_sig_set_start[ch][2];
    const int sig_off = ch == 0 ? 0 : 28;
    double uncoded = 0.0, base_cost = 0.0;
    int last_pos = -1, cg_last = -1;
    unsigned ctx_set = 0, c1idx = 0, c2idx = 0;
    int c1 = 1, c2 = 0, rice = 0;
    // rate_up / rate_down / sig_delta / delta_u / level are indexed by RASTER position below, as in the reference
    for (int cg = ncg - 1; cg >= 0; cg--) {
      const int cg_blk = S.scan_cg[cg], cgy = cg_blk / wg, cgx = cg_blk - cgy * wg;
      int nnz_before0 = 0;
      double st_coded = 0.0, st_uncoded = 0.0, st_sig = 0.0, st_sig0 = 0.0;
      int pattern = 0;
      if (wg > 1) {
        const int r = cgx < wg - 1 ? (S.cg_flag[cgy * wg + cgx + 1] != 0) : 0, b = cgy < wg - 1 ? (S.cg_flag[(cgy + 1) * wg + cgx] != 0) : 0;
        pattern = r + (b << 1);
      }
      for (int k = 15; k >= 0; k--) {
        const int sp = cg * 16 + k, bp = S.scan[sp];
        const int ld = S.lvl_d[sp];
        unsigned max_abs = (unsigned)((ld + (1 << (qbits - 1))) >> qbits);
        if (max_abs > 32767u) max_abs = 32767u;
        const double c0 = S.cost_coeff0[sp];
        uncoded = __dadd_rn(uncoded, c0);
        S.level[bp] = (int)max_abs;
        if (max_abs > 0 && last_pos < 0) { last_pos = sp; ctx_set = rdoq_ctx_set(ch, sp >> 4, 0); cg_last = cg; }
        double csig = 0.0;                          // cost_sig[sp]
        if (last_pos >= 0) {
          const int ctx_one = 4 * ctx_set + c1, ctx_abs = ctx_set + c2;
          const bool is_last = sp == last_pos;
          int ctx_sig = 0;
          if (!is_last) ctx_sig = sig_off + rdoq_sig_ctx_inc(pattern, first_ctx, bp, log2n, ch);
          unsigned best = 0;
          double cur_sig = 0.0, ccoef;
          bool decided = false;
          if (!is_last && max_abs < 3) {
            csig = rm.icost((double)est[E_SIG + 2 * ctx_sig + 0]);
            ccoef = __dadd_rn(c0, csig);
            if (max_abs == 0) decided = true;
          } else ccoef = 1.7e+308;
          if (!decided) {
            if (!is_last) cur_sig = rm.icost((double)est[E_SIG + 2 * ctx_sig + 1]);
            const unsigned min_abs = max_abs > 1 ? max_abs - 1 : 1;
            for (int a = (int)max_abs; a >= (int)min_abs; a--) {
              const double e = (double)(ld - (int)((unsigned)a << qbits));
              double c = __dadd_rn(__dmul_rn(__dmul_rn(e, e), err_scale),
                                   rm.icost((double)rm.level_rate((unsigned)a, ctx_one, ctx_abs, rice, c1idx, c2idx)));
              c = __dadd_rn(c, cur_sig);
              if (c < ccoef) { best = (unsigned)a; ccoef = c; csig = cur_sig; }
            }
          }
          S.cost_coeff[sp] = ccoef;
          if (!is_last) S.sig_delta[bp] = est[E_SIG + 2 * ctx_sig + 1] - est[E_SIG + 2 * ctx_sig + 0];
          S.delta_u[bp] = (ld - (int)(best << qbits)) >> (qbits - 8);
          if (best > 0) {
            const int now = rm.level_rate(best, ctx_one, ctx_abs, rice, c1idx, c2idx);
            S.rate_up[bp] = rm.level_rate(best + 1, ctx_one, ctx_abs, rice, c1idx, c2idx) - now;
            S.rate_down[bp] = rm.level_rate(best - 1, ctx_one, ctx_abs, rice, c1idx, c2idx) - now;
          } else S.rate_up[bp] = est[E_G1 + 2 * ctx_one + 0];
          S.level[bp] = (int)best;
          base_cost = __dadd_rn(base_cost, ccoef);
          const unsigned base_level = c1idx < 8 ? 2 + (c2idx < 1) : 1;
          if (best >= base_level && best > 3u * (1u << rice)) rice = rice + 1 < 4 ? rice + 1 : 4;
          if (best >= 1) c1idx++;
          if (best > 1) { c1 = 0; c2 += c2 < 2; c2idx++; }
          else if (c1 < 3 && c1 > 0 && best) c1++;
          if ((sp & 15) == 0 && sp > 0) {
            ctx_set = rdoq_ctx_set(ch, (sp - 1) >> 4, c1 == 0);
            c1 = 1; c2 = 0; c1idx = 0; c2idx = 0; rice = 0;
          }
        } else base_cost = __dadd_rn(base_cost, c0);
        S.cost_sig[sp] = csig;
        st_sig = __dadd_rn(st_sig, csig);
        if (k == 0) st_sig0 = csig;
        if (S.level[bp]) {
          S.cg_flag[cg_blk] = 1;
          st_coded = __dadd_rn(st_coded, __dsub_rn(S.cost_coeff[sp], csig));
          st_uncoded = __dadd_rn(st_uncoded, c0);
          if (k != 0) nnz_before0++;
        }
      }
      if (cg_last >= 0) {
        if (cg) {
          const int sr = cgx < wg - 1 ? (S.cg_flag[cgy * wg + cgx + 1] != 0) : 0, sb = cgy < wg - 1 ? (S.cg_flag[(cgy + 1) * wg + cgx] != 0) : 0;
          const int cctx = (sr + sb) != 0;
          const double r0 = rm.icost((double)est[E_SIGCG + 2 * cctx + 0]), r1 = rm.icost((double)est[E_SIGCG + 2 * cctx + 1]);
          if (S.cg_flag[cg_blk] == 0) {
            base_cost = __dadd_rn(base_cost, __dsub_rn(r0, st_sig));
            S.cost_cg_sig[cg] = r0;
          } else if (cg < cg_last) {
            if (nnz_before0 == 0) { base_cost = __dsub_rn(base_cost, st_sig0); st_sig = __dsub_rn(st_sig, st_sig0); }
            double zero_cost = base_cost;
            base_cost = __dadd_rn(base_cost, r1);
            zero_cost = __dadd_rn(zero_cost, r0);
            S.cost_cg_sig[cg] = r1;
            zero_cost = __dadd_rn(zero_cost, st_uncoded);
            zero_cost = __dsub_rn(zero_cost, st_coded);
            zero_cost = __dsub_rn(zero_cost, st_sig);
            if (zero_cost < base_cost) {
              S.cg_flag[cg_blk] = 0;
              base_cost = zero_cost;
              S.cost_cg_sig[cg] = r0;
              for (int k = 15; k >= 0; k--) {
                const int sp = cg * 16 + k, bp = S.scan[sp];
                if (S.level[bp]) { S.level[bp] = 0; S.cost_coeff[sp] = S.cost_coeff0[sp]; S.cost_sig[sp] = 0.0; }
              }
            }
          }
        } else S.cg_flag[cg_blk] = 1;
      }
    }
    if (last_pos >= 0) {
      double best_cost;
      int best_last_p1 = 0;
      if (!(rflags & 2) && ch == 0 && (rflags & 4)) {
        best_cost = __dadd_rn(uncoded, rm.icost((double)est[E_ROOT + 0]));
        base_cost = __dadd_rn(base_cost, rm.icost((double)est[E_ROOT + 1]));
      } else {
        best_cost = __dadd_rn(uncoded, rm.icost((double)est[E_CBP + 2 * ctx_cbf + 0]));
        base_cost = __dadd_rn(base_cost, rm.icost((double)est[E_CBP + 2 * ctx_cbf + 1]));
      }
      bool found = false;
      for (int cg = cg_last; cg >= 0 && !found; cg--) {
        const int cg_blk = S.scan_cg[cg];
        base_cost = __dsub_rn(base_cost, S.cost_cg_sig[cg]);
        if (!S.cg_flag[cg_blk]) continue;
        for (int k = 15; k >= 0; k--) {
          const int sp = cg * 16 + k;
          if (sp > last_pos) continue;
          const int bp = S.scan[sp];
          if (S.level[bp]) {
            const unsigned py = (unsigned)bp >> log2n, px = (unsigned)bp - (py << log2n);
            const double cl = scan_type == 2 ? rm.last_rate(py, px) : rm.last_rate(px, py);
            const double total = __dsub_rn(__dadd_rn(base_cost, cl), S.cost_sig[sp]);
            if (total < best_cost) { best_last_p1 = sp + 1; best_cost = total; }
            if (S.level[bp] > 1) { found = true; break; }
            base_cost = __dsub_rn(base_cost, S.cost_coeff[sp]);
            base_cost = __dadd_rn(base_cost, S.cost_coeff0[sp]);
          } else base_cost = __dsub_rn(base_cost, S.cost_sig[sp]);
        }
      }
      for (int sp = 0; sp < best_last_p1; sp++) {
        const int bp = S.scan[sp];
        const int l = S.level[bp];
        abs_sum += (uint32_t)l;
        S.level[bp] = coeff[bp] < 0 ? -l : l;
      }
      for (int sp = best_last_p1; sp <= last_pos; sp++) S.level[S.scan[sp]] = 0;
      if ((rflags & 1) && abs_sum >= 2) {           // sign-bit hiding inside RDOQ (:2520-2668)
        const double iq = (double)c_tq_iqscale[rem];
        const long long rd_factor =
            (long long)__dadd_rn(__ddiv_rn(__ddiv_rn(__dmul_rn(__dmul_rn(iq, iq), (double)(1 << (2 * per))), lambda), 16.0), 0.5);
        int last_cg = -1;
        for (int sub = (n2 - 1) >> 4; sub >= 0; sub--) {
          const int sp0 = sub << 4;
          int first_nz = 16, last_nz = -1, asum = 0, k;
          for (k = 15; k >= 0; --k) if (S.level[S.scan[k + sp0]]) { last_nz = k; break; }
          for (k = 0; k < 16; k++) if (S.level[S.scan[k + sp0]]) { first_nz = k; break; }
          for (k = first_nz; k <= last_nz; k++) asum += S.level[S.scan[k + sp0]];
          if (last_nz >= 0 && last_cg == -1) last_cg = 1;
          if (last_nz - first_nz >= 4) {
            const unsigned signbit = S.level[S.scan[sp0 + first_nz]] > 0 ? 0 : 1;
            if (signbit != (unsigned)(asum & 1)) {
              long long min_inc = 0x7fffffffffffffffLL, cur = 0x7fffffffffffffffLL;
              int min_pos = -1, final_change = 0, cur_change = 0;
              for (k = (last_cg == 1 ? last_nz : 15); k >= 0; --k) {
                const int bp = S.scan[k + sp0];
                const int lv = S.level[bp];
                if (lv != 0) {
                  const long long up = rd_factor * (-S.delta_u[bp]) + S.rate_up[bp];
                  long long down = rd_factor * (S.delta_u[bp]) + S.rate_down[bp] - ((abs(lv) == 1) ? S.sig_delta[bp] : 0);
                  if (last_cg == 1 && last_nz == k && abs(lv) == 1) down -= 4 << 15;
                  if (up < down) { cur = up; cur_change = 1; }
                  else {
                    cur_change = -1;
                    cur = (k == first_nz && abs(lv) == 1) ? 0x7fffffffffffffffLL : down;
                  }
                } else {
                  cur = rd_factor * (-(long long)abs(S.delta_u[bp])) + (1 << 15) + S.rate_up[bp] + S.sig_delta[bp];
                  cur_change = 1;
                  if (k < first_nz) {
                    const unsigned s = coeff[bp] >= 0 ? 0 : 1;
                    if (s != signbit) cur = 0x7fffffffffffffffLL;
                  }
                }
                if (cur < min_inc) { min_inc = cur; final_change = cur_change; min_pos = bp; }
              }
              if (S.level[min_pos] == 32767 || S.level[min_pos] == -32768) final_change = -1;
              if (coeff[min_pos] >= 0) S.level[min_pos] += final_change;
              else S.level[min_pos] -= final_change;
            }
          }
          if (last_cg == 1) last_cg = 0;
        }
      }
    }
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, abs_sum, 0);
}

}  // namespace hevcdl
