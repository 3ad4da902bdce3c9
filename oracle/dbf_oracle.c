/* oracle/dbf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's deblocking filter for ALL-INTRA pictures:
 * TComLoopFilter::loopFilterPic (HM_dl/source/Lib/TLibCommon/TComLoopFilter.cpp:130-158) -> xDeblockCU (:170-238: edges on
 * the 8x8 luma grid, chroma on the 8x8 chroma grid), xSetEdgefilterTU / PU (:274-360: an edge is a TU, PU or CU boundary
 * inside the picture), xGetBoundaryStrengthSingle (:416-555: Bs = 2 when either side is intra -- always, here),
 * xEdgeFilterLuma (:557-674), xEdgeFilterChroma (:676-828), xPelFilterLuma / Chroma, xUseStrongFiltering, xCalcDP / DQ
 * (:830-953), tables sm_tcTable / sm_betaTable (:59-67), chroma QP mapping g_aucChromaScale (TComRom.cpp:532-539).
 * Restrictions = the reference's operating point: 8-bit 4:2:0, one slice, no tiles, no PCM / lossless / transquant bypass.
 * All vertical edges of the picture are filtered first, then all horizontal ones (the two CTU loops of loopFilterPic).
 * Pinned by tests/golden/dbf_192x128.npz: reconstructed pictures dumped by the reference encoder itself right before and
 * right after its own loopFilterPic (oracle/_ref/TAppEncoder_dbftrace, oracle/dbf_dump.h), two QPs.
 */
#include <stdint.h>
#include <stdlib.h>

static const uint8_t kTc[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4,
                                5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};
static const uint8_t kBeta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24, 26,
                                  28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};
static const uint8_t kChromaQp420[58] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28,
                                         29, 29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48,
                                         49, 50, 51};
static int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static int clip8(int v) { return clip3(0, 255, v); }

/* is there a transform / coding block boundary between luma sample (x-1, y) and (x, y) [vertical] or (x, y-1) and (x, y)? */
static int is_edge(const uint8_t *tu, int w4, int x, int y, int vertical) {
  const int s = 1 << tu[(y >> 2) * w4 + (x >> 2)];
  return ((vertical ? x : y) & (s - 1)) == 0;
}

/* one 4-sample luma segment; p points at q0 of its first line, `across` steps across the edge, `along` along it */
static void luma_segment(int16_t *p, int across, int along, int tc, int beta) {
#define S(line, k) p[(line) * along + (k) * across]     /* k = -4..3: p3..p0, q0..q3 */
  const int dp0 = abs(S(0, -3) - 2 * S(0, -2) + S(0, -1)), dq0 = abs(S(0, 0) - 2 * S(0, 1) + S(0, 2));
  const int dp3 = abs(S(3, -3) - 2 * S(3, -2) + S(3, -1)), dq3 = abs(S(3, 0) - 2 * S(3, 1) + S(3, 2));
  const int d0 = dp0 + dq0, d3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3, d = d0 + d3;
  if (d >= beta) return;
  const int side = (beta + (beta >> 1)) >> 3;
  const int fp = dp < side, fq = dq < side;
  int strong = 1;
  for (int l = 0; l < 4; l += 3) {
    const int dd = 2 * (l ? d3 : d0);
    strong &= (abs(S(l, -4) - S(l, -1)) + abs(S(l, 3) - S(l, 0)) < (beta >> 3)) && (dd < (beta >> 2)) && (abs(S(l, -1) - S(l, 0)) < ((tc * 5 + 1) >> 1));
  }
  for (int l = 0; l < 4; l++) {
    const int m0 = S(l, -4), m1 = S(l, -3), m2 = S(l, -2), m3 = S(l, -1), m4 = S(l, 0), m5 = S(l, 1), m6 = S(l, 2), m7 = S(l, 3);
    if (strong) {
      S(l, -1) = (int16_t)clip3(m3 - 2 * tc, m3 + 2 * tc, (m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3);
      S(l, 0) = (int16_t)clip3(m4 - 2 * tc, m4 + 2 * tc, (m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3);
      S(l, -2) = (int16_t)clip3(m2 - 2 * tc, m2 + 2 * tc, (m1 + m2 + m3 + m4 + 2) >> 2);
      S(l, 1) = (int16_t)clip3(m5 - 2 * tc, m5 + 2 * tc, (m3 + m4 + m5 + m6 + 2) >> 2);
      S(l, -3) = (int16_t)clip3(m1 - 2 * tc, m1 + 2 * tc, (2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3);
      S(l, 2) = (int16_t)clip3(m6 - 2 * tc, m6 + 2 * tc, (m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3);
    } else {
      int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
      if (abs(delta) < tc * 10) {
        delta = clip3(-tc, tc, delta);
        S(l, -1) = (int16_t)clip8(m3 + delta);
        S(l, 0) = (int16_t)clip8(m4 - delta);
        const int tc2 = tc >> 1;
        if (fp) S(l, -2) = (int16_t)clip8(m2 + clip3(-tc2, tc2, (((m1 + m3 + 1) >> 1) - m2 + delta) >> 1));
        if (fq) S(l, 1) = (int16_t)clip8(m5 + clip3(-tc2, tc2, (((m6 + m4 + 1) >> 1) - m5 - delta) >> 1));
      }
    }
  }
#undef S
}

static void chroma_line(int16_t *p, int across, int tc) {
  const int m2 = p[-2 * across], m3 = p[-across], m4 = p[0], m5 = p[across];
  const int delta = clip3(-tc, tc, ((((m4 - m3) << 2) + m2 - m5 + 4) >> 3));
  p[-across] = (int16_t)clip8(m3 + delta);
  p[0] = (int16_t)clip8(m4 - delta);
}

/* In place.  Y: W x H, U / V: W/2 x H/2 (int16 samples, 8-bit content); tu_log2, qp: one entry per 4x4 luma unit. */
void oracle_deblock_frame(int16_t *Y, int sy, int16_t *U, int16_t *V, int sc, int W, int H, const uint8_t *tu_log2, const int8_t *qp,
                          int beta_off_div2, int tc_off_div2, int cb_qp_off, int cr_qp_off) {
  const int w4 = W >> 2;
  for (int dir = 0; dir < 2; dir++) {                       /* 0: vertical edges (filter across x), 1: horizontal edges */
    const int vertical = dir == 0;
    for (int y = vertical ? 0 : 8; y < H; y += vertical ? 4 : 8)
      for (int x = vertical ? 8 : 0; x < W; x += vertical ? 8 : 4) {
        if (!is_edge(tu_log2, w4, x, y, vertical)) continue;
        const int qq = qp[(y >> 2) * w4 + (x >> 2)], qpp = vertical ? qp[(y >> 2) * w4 + ((x - 1) >> 2)] : qp[((y - 1) >> 2) * w4 + (x >> 2)];
        const int q = (qq + qpp + 1) >> 1;
        {                                                   /* luma, Bs = 2 */
          const int tc = kTc[clip3(0, 53, q + 2 + (tc_off_div2 << 1))], beta = kBeta[clip3(0, 51, q + (beta_off_div2 << 1))];
          luma_segment(Y + y * sy + x, vertical ? 1 : sy, vertical ? sy : 1, tc, beta);
        }
        if (((vertical ? x : y) & 15) == 0) {               /* chroma: edges on the 8x8 chroma sample grid, two chroma lines per luma unit */
          for (int c = 0; c < 2; c++) {
            int qc = q + (c ? cr_qp_off : cb_qp_off);
            if (qc >= 58) qc -= 6; else if (qc >= 0) qc = kChromaQp420[qc];
            const int tc = kTc[clip3(0, 53, qc + 2 + (tc_off_div2 << 1))];
            int16_t *pl = (c ? V : U) + (y >> 1) * sc + (x >> 1);
            for (int l = 0; l < 2; l++) chroma_line(pl + l * (vertical ? sc : 1), vertical ? 1 : sc, tc);
          }
        }
      }
  }
}
