/* oracle/tq_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic core of the reference's RD pass for one transform unit
 * (SURVEY.md 8(f) row 1): TEncSearch::xIntraCodingTUBlock (HM_dl/source/Lib/TLibEncoder/TEncSearch.cpp:1129-1424) calls
 *   TComTrQuant::transformNxN     (Lib/TLibCommon/TComTrQuant.cpp:1450-1534): xT -> xTrMxN (:860-925, partial butterflies
 *                                 :388-858, DST :414-439) or xTransformSkip (:2010-2052), then xQuant (:1126-1249, the
 *                                 non-RDOQ branch; sign-bit hiding :991-1124 is not restated: parity is pinned with
 *                                 SignHideFlag=0)
 *   TComTrQuant::invTransformNxN  (:1537-1666): xDeQuant (:1308-1423, flat scaling), xIT -> xITrMxN (:927-988) or
 *                                 xITransformSkip (:2060-2104)
 * at the reference's operating point: 8-bit video, maxLog2TrDynamicRange 15, no scaling lists, extended precision off.
 * The partial butterflies are exact factorisations of the integer matrix products, so the products are written directly;
 * the matrices are those of the HEVC specification (8.6.4.2), generated from their 31 distinct magnitudes.
 * Pinned by tests/golden/tq_*.npz: (a) the reference's own xTrMxN / xITrMxN linked from oracle/_ref/libhmref.a and fed random
 * blocks, (b) per-TU dumps of the reference encoder built with DEBUG_TRANSFORM_AND_QUANTISE (tools/gen_golden_tq.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TQ_FLAG_DST 1       /* 4x4 intra luma: DST-VII instead of the DCT (TComTU::useDST) */
#define TQ_FLAG_TSKIP 2     /* transform skip (4x4 only in the reference's configuration) */
#define TQ_FLAG_INTER 4     /* rounding offset 85 instead of 171 (P/B slices; all-intra runs never set it) */

static const int kCos[32] = {0, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                             61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4};
static const int kDst[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};

/* entry (k, n) of the N-point HEVC core transform matrix (TComRom.cpp:368-520 as macros) */
int oracle_tq_matrix(int N, int k, int n) {
  if (k == 0) return 64;
  const int k32 = k * (32 / N);
  int r = ((2 * n + 1) * k32) % 128;
  if (r > 64) r = 128 - r;
  return r > 32 ? -kCos[64 - r] : kCos[r];
}

static int tmat(int N, int dst, int k, int n) { return dst ? kDst[k][n] : oracle_tq_matrix(N, k, n); }
static int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* forward 2-D transform of an N x N residual block (row-major) -> coefficients (row-major), xTrMxN */
void oracle_tq_forward(const int16_t *resi, int N, int dst, int32_t *coeff) {
  int lg = 0;
  while ((1 << lg) < N) lg++;
  const int s1 = lg + 8 + 6 - 15, s2 = lg + 6;
  const int a1 = s1 > 0 ? 1 << (s1 - 1) : 0, a2 = 1 << (s2 - 1);
  int32_t tmp[32 * 32];
  for (int j = 0; j < N; j++)          /* rows of the block: tmp[k][j] */
    for (int k = 0; k < N; k++) {
      int32_t s = 0;
      for (int n = 0; n < N; n++) s += tmat(N, dst, k, n) * resi[j * N + n];
      tmp[k * N + j] = (s + a1) >> s1;
    }
  for (int j = 0; j < N; j++)          /* rows of tmp: coeff[k][j] */
    for (int k = 0; k < N; k++) {
      int32_t s = 0;
      for (int n = 0; n < N; n++) s += tmat(N, dst, k, n) * tmp[j * N + n];
      coeff[k * N + j] = (s + a2) >> s2;
    }
}

/* inverse 2-D transform, xITrMxN: coefficients -> residual (int16) */
void oracle_tq_inverse(const int32_t *coeff, int N, int dst, int16_t *resi) {
  const int s1 = 7, s2 = 6 + 15 - 1 - 8;
  int32_t tmp[32 * 32];
  for (int j = 0; j < N; j++)          /* columns of coeff: tmp[j][n] */
    for (int n = 0; n < N; n++) {
      int32_t s = 0;
      for (int k = 0; k < N; k++) s += tmat(N, dst, k, n) * coeff[k * N + j];
      tmp[j * N + n] = clip3(-32768, 32767, (s + (1 << (s1 - 1))) >> s1);
    }
  for (int j = 0; j < N; j++)          /* columns of tmp: block[j][n] */
    for (int n = 0; n < N; n++) {
      int32_t s = 0;
      for (int k = 0; k < N; k++) s += tmat(N, dst, k, n) * tmp[k * N + j];
      resi[j * N + n] = (int16_t)clip3(-32768, 32767, (s + (1 << (s2 - 1))) >> s2);
    }
}

static const int kQuantScales[6] = {26214, 23302, 20560, 18396, 16384, 14564};   /* TComRom.cpp:354-362 */
static const int kInvQuantScales[6] = {40, 45, 51, 57, 64, 72};

/* One TU through transformNxN + invTransformNxN.  resi: N*N int16 in; outputs (any may be NULL): coeff (transform
 * output), level (quantised), deq (dequantised), rec (reconstructed residual).  Returns uiAbsSum. */
uint32_t oracle_tq_tu(const int16_t *resi, int log2n, int qp, int flags, int32_t *coeff, int32_t *level, int32_t *deq, int16_t *rec) {
  const int N = 1 << log2n, n2 = N * N;
  int32_t c[32 * 32], q[32 * 32], d[32 * 32];
  int16_t r[32 * 32];
  const int tshift = 15 - 8 - log2n;                               /* getTransformShift */
  if (flags & TQ_FLAG_TSKIP) for (int i = 0; i < n2; i++) c[i] = (int32_t)resi[i] << tshift;
  else oracle_tq_forward(resi, N, (flags & TQ_FLAG_DST) && N == 4, c);
  const int per = qp / 6, rem = qp % 6;
  const int qbits = 14 + per + tshift;
  const int64_t add = (int64_t)((flags & TQ_FLAG_INTER) ? 85 : 171) << (qbits - 9);
  uint32_t abs_sum = 0;
  for (int i = 0; i < n2; i++) {
    const int64_t t = (int64_t)abs(c[i]) * kQuantScales[rem];
    const int32_t mag = (int32_t)((t + add) >> qbits);
    abs_sum += (uint32_t)mag;
    q[i] = clip3(-32768, 32767, c[i] < 0 ? -mag : mag);
  }
  const int rs = 6 - (tshift + per);                                /* IQUANT_SHIFT - (transformShift + per) */
  int tib = 32 + rs - 7;                                            /* targetInputBitDepth */
  if (tib > 16) tib = 16;
  const int imin = -(1 << (tib - 1)), imax = (1 << (tib - 1)) - 1;
  for (int i = 0; i < n2; i++) {
    const int32_t cq = clip3(imin, imax, q[i]);
    int32_t v;
    if (rs > 0) v = (cq * kInvQuantScales[rem] + (1 << (rs - 1))) >> rs;
    else v = (int32_t)((uint32_t)(cq * kInvQuantScales[rem]) << (-rs));
    d[i] = clip3(-32768, 32767, v);
  }
  if (flags & TQ_FLAG_TSKIP) {
    const int off = tshift == 0 ? 0 : 1 << (tshift - 1);
    for (int i = 0; i < n2; i++) r[i] = (int16_t)((d[i] + off) >> tshift);
  } else oracle_tq_inverse(d, N, (flags & TQ_FLAG_DST) && N == 4, r);
  if (coeff) memcpy(coeff, c, n2 * sizeof(int32_t));
  if (level) memcpy(level, q, n2 * sizeof(int32_t));
  if (deq) memcpy(deq, d, n2 * sizeof(int32_t));
  if (rec) memcpy(rec, r, n2 * sizeof(int16_t));
  return abs_sum;
}
