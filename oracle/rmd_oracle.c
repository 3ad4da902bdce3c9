/* oracle/rmd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, integer) of the first ("RMD") pass of the reference's luma intra
 * search for one PU: reference-sample construction, smoothing, the 35 predictors, Hadamard
 * SATD, cost and the candidate list, plus the label-driven PU enumeration of the pruned
 * quadtree.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.  Citations are relative to /root/reference/HM_dl/source/Lib.
 *
 * Parity status: PINNED two ways (tests/test_oracle_rmd.py):
 *   (1) tests/golden/rmd_trace_*.npz -- per-mode SAD / mode-bits printed by the reference encoder
 *       itself (oracle/_ref/TAppEncoder_trace, i.e. DEBUG_INTRA_SEARCH_COSTS at
 *       TLibEncoder/TEncSearch.cpp:2315) together with its reconstruction, replayed here;
 *   (2) oracle/ref_harness.cpp links the reference's own TComPrediction / TComRdCost objects and
 *       compares predictors and SATD on random data (run by the same test when oracle/_ref exists).
 *
 * Reference-sample "line" layout used throughout (n = PU size, length 4n+1):
 *   line[0]        = below-left-most sample  (x=-1, y=2n-1)
 *   line[2n-1]     = left sample at y=0      (x=-1, y=0)
 *   line[2n]       = corner                  (x=-1, y=-1)
 *   line[2n+1+k]   = above sample            (x=k,  y=-1), k=0..2n-1
 * This is the substitution scan order of TLibCommon/TComPattern.cpp:386-541.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 64

/* z-order index of a 4x4 unit inside a 64x64 CTU (TLibCommon/TComRom.cpp:284-352 g_auiRasterToZscan) */
static int zidx(int ux, int uy)
{
  int z = 0;
  for (int b = 0; b < 4; b++) z |= ((ux >> b) & 1) << (2 * b) | ((uy >> b) & 1) << (2 * b + 1);
  return z;
}

/* A neighbouring 4x4 unit is available iff it lies inside the picture and precedes the current
 * unit in coding order (CTU raster, then z-order): the net effect of getPULeft/Above/AboveLeft/
 * AboveRight/BelowLeft (TLibCommon/TComDataCU.cpp:1000-1200) for one slice, no tiles,
 * constrained_intra_pred off (TLibCommon/TComPattern.cpp:572-749). */
static int unit_available(int xn, int yn, int xc, int yc, int W, int H)
{
  if (xn < 0 || yn < 0 || xn >= W || yn >= H) return 0;
  int cw = (W + 63) / 64;
  long on = ((long)(yn / 64) * cw + xn / 64) * 256 + zidx((xn % 64) / 4, (yn % 64) / 4);
  long oc = ((long)(yc / 64) * cw + xc / 64) * 256 + zidx((xc % 64) / 4, (yc % 64) / 4);
  return on < oc;
}

/* Build the unfiltered reference line of a PU at (x0,y0), size n, from picture `pic` (the
 * reconstruction in the reference: TComPattern.cpp:169; the original picture in the batched
 * original-reference mode).  Follows fillReferenceSamples, TComPattern.cpp:326-543. */
void oracle_build_ref_line(const uint8_t *pic, int stride, int W, int H, int x0, int y0, int n,
                           int16_t *line /* 4n+1 */)
{
  int units = n / 4, total = 4 * units + 1;      /* below-left, left, corner, above, above-right */
  int avail[2 * MAXN / 4 * 2 + 1];
  int navail = 0;
  for (int u = 0; u < 2 * units; u++) {          /* left column, bottom to top */
    int yn = y0 + (2 * units - 1 - u) * 4;
    avail[u] = unit_available(x0 - 1, yn, x0, y0, W, H);
    navail += avail[u];
  }
  avail[2 * units] = unit_available(x0 - 1, y0 - 1, x0, y0, W, H);
  navail += avail[2 * units];
  for (int u = 0; u < 2 * units; u++) {          /* above row, left to right */
    avail[2 * units + 1 + u] = unit_available(x0 + u * 4, y0 - 1, x0, y0, W, H);
    navail += avail[2 * units + 1 + u];
  }
  if (navail == 0) {                             /* :347-358 */
    for (int i = 0; i < 4 * n + 1; i++) line[i] = 128;
    return;
  }
  /* gather available samples */
  for (int i = 0; i < 2 * n; i++) {
    int y = y0 + (2 * n - 1 - i);
    line[i] = avail[i / 4] ? pic[y * stride + x0 - 1] : -1;
  }
  line[2 * n] = avail[2 * units] ? pic[(y0 - 1) * stride + x0 - 1] : -1;
  for (int k = 0; k < 2 * n; k++)
    line[2 * n + 1 + k] = avail[2 * units + 1 + k / 4] ? pic[(y0 - 1) * stride + x0 + k] : -1;
  /* substitution (:484-541): if the first unit is missing take the first available sample in
   * scan order; every other missing sample copies its predecessor. */
  if (line[0] < 0) {
    int j = 0;
    while (line[j] < 0) j++;
    line[0] = line[j];
  }
  for (int i = 1; i < 4 * n + 1; i++)
    if (line[i] < 0) line[i] = line[i - 1];
  (void)total;
}

/* [1 2 1] smoothing or strong (bilinear) smoothing for n==32 (TComPattern.cpp:203-294;
 * strong_intra_smoothing on, 8-bit => threshold 1<<3). */
void oracle_filter_ref_line(const int16_t *line, int n, int16_t *filt)
{
  int len = 4 * n + 1;
  int bl = line[0], tl = line[2 * n], tr = line[4 * n];
  int strong = 0;
  if (n >= 32) {
    int a = bl + tl - 2 * line[n], b = tl + tr - 2 * line[3 * n];
    if (a < 0) a = -a;
    if (b < 0) b = -b;
    strong = (a < 8) && (b < 8);
  }
  filt[0] = line[0];
  filt[len - 1] = line[len - 1];
  if (strong) {
    int shift = 0;
    while ((1 << shift) < 2 * n) shift++;
    for (int i = 1; i < 2 * n; i++) {
      filt[i] = (int16_t)(((2 * n - i) * bl + i * tl + n) >> shift);
      filt[2 * n + i] = (int16_t)(((2 * n - i) * tl + i * tr + n) >> shift);
    }
    filt[2 * n] = line[2 * n];
  } else {
    for (int i = 1; i < len - 1; i++) filt[i] = (int16_t)((line[i - 1] + 2 * line[i] + line[i + 1] + 2) >> 2);
  }
}

/* filtered references are used iff min(|m-10|,|m-26|) > thr[size]; never for DC
 * (TComPattern.cpp:545-570, table TComPrediction.cpp:50-58) */
int oracle_mode_uses_filter(int mode, int n)
{
  static const int thr[5] = {10, 7, 1, 0, 10};
  int s = n == 4 ? 0 : n == 8 ? 1 : n == 16 ? 2 : n == 32 ? 3 : 4;
  if (mode == 1) return 0;
  int a = mode - 10, b = mode - 26;
  if (a < 0) a = -a;
  if (b < 0) b = -b;
  return (a < b ? a : b) > thr[s];
}

static inline int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* One predictor.  left(y) = line[2n-1-y], top(x) = line[2n+1+x], corner = line[2n].
 * planar: TComPrediction.cpp:731-781; DC + edge filter: :183-201,794-817;
 * angular: :229-388 (angle tables :265-266). */
void oracle_predict(const int16_t *line, int n, int mode, int16_t *pred /* n*n */)
{
#define LEFT(y) ((int)line[2 * n - 1 - (y)])
#define TOP(x) ((int)line[2 * n + 1 + (x)])
  int lg = 0;
  while ((1 << lg) < n) lg++;
  if (mode == 0) {
    int bl = LEFT(n), tr = TOP(n);
    for (int y = 0; y < n; y++)
      for (int x = 0; x < n; x++) {
        int hor = (LEFT(y) << lg) + n + (x + 1) * (tr - LEFT(y));
        int ver = (TOP(x) << lg) + (y + 1) * (bl - TOP(x));
        pred[y * n + x] = (int16_t)((hor + ver) >> (lg + 1));
      }
    return;
  }
  if (mode == 1) {
    int sum = 0;
    for (int i = 0; i < n; i++) sum += TOP(i) + LEFT(i);
    int dc = (sum + n) / (2 * n);
    for (int i = 0; i < n * n; i++) pred[i] = (int16_t)dc;
    if (n <= 16) {
      pred[0] = (int16_t)((TOP(0) + LEFT(0) + 2 * dc + 2) >> 2);
      for (int x = 1; x < n; x++) pred[x] = (int16_t)((TOP(x) + 3 * dc + 2) >> 2);
      for (int y = 1; y < n; y++) pred[y * n] = (int16_t)((LEFT(y) + 3 * dc + 2) >> 2);
    }
    return;
  }
  static const int ang_tab[9] = {0, 2, 5, 9, 13, 17, 21, 26, 32};
  static const int inv_tab[9] = {0, 4096, 1638, 910, 630, 482, 390, 315, 256};
  int ver = mode >= 18;
  int am = ver ? mode - 26 : -(mode - 10);
  int aabs = am < 0 ? -am : am;
  int angle = (am < 0 ? -1 : 1) * ang_tab[aabs];
  int16_t buf[3 * MAXN + 2];
  int16_t *ref = buf + MAXN; /* ref[-n .. 2n], ref[0] = corner */
  /* main = above for vertical modes, left for horizontal; side = the other one */
  for (int i = 0; i <= 2 * n; i++) ref[i] = (int16_t)(i == 0 ? line[2 * n] : (ver ? TOP(i - 1) : LEFT(i - 1)));
  if (angle < 0) {
    int last = (n * angle) >> 5, acc = 128;
    for (int k = -1; k > last; k--) {
      acc += inv_tab[aabs];
      int s = acc >> 8; /* side[s], side[0] = corner */
      ref[k] = (int16_t)(s == 0 ? line[2 * n] : (ver ? LEFT(s - 1) : TOP(s - 1)));
    }
  }
  for (int j = 0; j < n; j++) {      /* j runs along the prediction direction (rows if vertical) */
    int pos = (j + 1) * angle, di = pos >> 5, df = pos & 31;
    for (int i = 0; i < n; i++) {
      int v = df ? (((32 - df) * ref[i + di + 1] + df * ref[i + di + 2] + 16) >> 5) : ref[i + di + 1];
      if (angle == 0 && n <= 16 && i == 0) {          /* pure V/H edge filter, :334-340 */
        int side = ver ? LEFT(j) : TOP(j);
        v = clip8(v + ((side - line[2 * n]) >> 1));
      }
      if (ver) pred[j * n + i] = (int16_t)v;
      else pred[i * n + j] = (int16_t)v;
    }
  }
#undef LEFT
#undef TOP
}

/* Sum of |2-D Hadamard| over b x b blocks (b = 8, or 4 for a 4x4 PU); 8x8: (s+2)>>2, 4x4:
 * (s+1)>>1 (TLibCommon/TComRdCost.cpp:1549-1824).  The sum of absolute values does not depend
 * on the butterfly ordering, so a textbook in-place Walsh-Hadamard is exact. */
uint32_t oracle_satd(const uint8_t *org, int ostride, const int16_t *pred, int n)
{
  int b = n >= 8 ? 8 : 4;
  uint32_t total = 0;
  for (int by = 0; by < n; by += b)
    for (int bx = 0; bx < n; bx += b) {
      int m[64];
      for (int y = 0; y < b; y++)
        for (int x = 0; x < b; x++) m[y * b + x] = (int)org[(by + y) * ostride + bx + x] - pred[(by + y) * n + bx + x];
      for (int y = 0; y < b; y++)                         /* rows */
        for (int h = 1; h < b; h <<= 1)
          for (int i = 0; i < b; i += 2 * h)
            for (int j = i; j < i + h; j++) {
              int a = m[y * b + j], c = m[y * b + j + h];
              m[y * b + j] = a + c; m[y * b + j + h] = a - c;
            }
      for (int x = 0; x < b; x++)                         /* columns */
        for (int h = 1; h < b; h <<= 1)
          for (int i = 0; i < b; i += 2 * h)
            for (int j = i; j < i + h; j++) {
              int a = m[j * b + x], c = m[(j + h) * b + x];
              m[j * b + x] = a + c; m[(j + h) * b + x] = a - c;
            }
      uint32_t s = 0;
      for (int i = 0; i < b * b; i++) s += (uint32_t)(m[i] < 0 ? -m[i] : m[i]);
      total += b == 8 ? (s + 2) >> 2 : (s + 1) >> 1;
    }
  return total;
}

/* 35 SATDs of one PU given its reference line (TLibEncoder/TEncSearch.cpp:2296-2320). */
void oracle_pu_satd35(const uint8_t *org, int ostride, const int16_t *line, int n, uint32_t *satd)
{
  int16_t *filt = (int16_t *)malloc(sizeof(int16_t) * (4 * n + 1));
  int16_t *pred = (int16_t *)malloc(sizeof(int16_t) * n * n);
  oracle_filter_ref_line(line, n, filt);
  for (int m = 0; m < 35; m++) {
    oracle_predict(oracle_mode_uses_filter(m, n) ? filt : line, n, m, pred);
    satd[m] = oracle_satd(org, ostride, pred, n);
  }
  free(filt); free(pred);
}

int oracle_num_rd_modes(int n) { return n >= 16 ? 3 : 8; } /* TLibCommon/TComRom.cpp:545-553 */

/* Candidate list: cost = satd + bits*sqrtLambda in double, insertion from the worst slot with
 * strict '<' (TEncSearch.cpp:2313,5562-5585); then the first `n_mpm_add` MPMs (1 if left==above
 * else 2, TEncSearch.cpp:2322-2345 with TLibCommon/TComDataCU.cpp:1362-1445) not yet present are
 * appended.  Returns the list length. */
int oracle_cand_list(const uint32_t *satd, const uint32_t *bits, double sqrt_lambda, int n,
                     const int *mpm, int n_mpm_add, uint8_t *modes /* <=10 */, double *costs /* <=8, may be 0 */)
{
  int keep = oracle_num_rd_modes(n);
  double cl[8];
  unsigned ml[10];
  for (int i = 0; i < keep; i++) { cl[i] = 1.7e308; ml[i] = 0; }
  for (int m = 0; m < 35; m++) {
    double c = (double)satd[m] + (double)bits[m] * sqrt_lambda;
    int shift = 0;
    while (shift < keep && c < cl[keep - 1 - shift]) shift++;
    if (shift) {
      for (int i = 1; i < shift; i++) { ml[keep - i] = ml[keep - 1 - i]; cl[keep - i] = cl[keep - 1 - i]; }
      ml[keep - shift] = (unsigned)m; cl[keep - shift] = c;
    }
  }
  int len = keep;
  for (int j = 0; j < n_mpm_add; j++) {
    int inc = 0;
    for (int i = 0; i < len; i++) inc |= (mpm[j] == (int)ml[i]);  /* HM scans the grown list too */
    if (!inc) ml[len++] = (unsigned)mpm[j];
  }
  for (int i = 0; i < len; i++) modes[i] = (uint8_t)ml[i];
  if (costs) for (int i = 0; i < keep; i++) costs[i] = cl[i];
  return len;
}

/* MPM derivation from the left / above luma modes (-1 = neighbour unavailable, or above lies in
 * another CTU => DC) (TComDataCU.cpp:1362-1445).  Returns how many MPMs the RMD list may gain. */
int oracle_mpm(int left, int above, int *mpm)
{
  if (left < 0) left = 1;
  if (above < 0) above = 1;
  if (left == above) {
    if (left > 1) { mpm[0] = left; mpm[1] = ((left + 29) % 32) + 2; mpm[2] = ((left - 1) % 32) + 2; }
    else { mpm[0] = 0; mpm[1] = 1; mpm[2] = 26; }
    return 1;
  }
  mpm[0] = left; mpm[1] = above;
  if (left && above) mpm[2] = 0;
  else mpm[2] = (left + above) < 2 ? 26 : 1;
  return 2;
}

/* ---- pruned quadtree -> PU list (TLibEncoder/TEncCu.cpp:496-520,815-834,945-965) ------------
 * Emits, in the encoder's visiting order, every PU whose RMD pass runs for one CTU:
 * one 2Nx2N PU per CU evaluated at its label depth and, for 8x8 CUs, the four 4x4 PUs of the
 * NxN trial (TEncCu.cpp:819-826).  pu[i] = {x, y, size, part(0=2Nx2N,1..4=NxN idx+1)}.
 * CUs crossing the picture edge are never evaluated at that depth (bBoundary, :574-576) and are
 * descended only if label>depth (the reference's fact-6 behaviour); children starting outside
 * the picture are skipped (:929-946). */
static int enum_cu(const uint8_t *label, int x, int y, int depth, int W, int H, int *pu, int cnt)
{
  int size = 64 >> depth;
  if (x >= W || y >= H) return cnt;
  int boundary = (x + size > W) || (y + size > H);
  int p = label[4 * ((y % 64) / 16) + (x % 64) / 16];
  if (p == depth && !boundary) {
    pu[4 * cnt] = x; pu[4 * cnt + 1] = y; pu[4 * cnt + 2] = size; pu[4 * cnt + 3] = 0; cnt++;
    if (depth == 3)
      for (int k = 0; k < 4; k++) {
        pu[4 * cnt] = x + (k & 1) * 4; pu[4 * cnt + 1] = y + (k >> 1) * 4; pu[4 * cnt + 2] = 4; pu[4 * cnt + 3] = k + 1; cnt++;
      }
  } else if (p > depth && depth < 3) {
    for (int k = 0; k < 4; k++)
      cnt = enum_cu(label, x + (k & 1) * (size / 2), y + (k >> 1) * (size / 2), depth + 1, W, H, pu, cnt);
  }
  return cnt;
}

int oracle_enum_ctu_pus(const uint8_t *label, int ctu_x, int ctu_y, int W, int H, int *pu /* [<=320][4] */)
{
  return enum_cu(label, ctu_x * 64, ctu_y * 64, 0, W, H, pu, 0);
}

/* Batched original-reference RMD of a whole frame (the throughput mode of the product): for
 * every PU of every CTU in [ctu_begin,ctu_end) compute the 35 SATDs with references taken from
 * `pic` itself.  out_pu [npu][4], out_satd [npu][35]; returns npu. */
int oracle_frame_rmd(const uint8_t *pic, int W, int H, const uint8_t *labels, int ctu_begin, int ctu_end,
                     int *out_pu, uint32_t *out_satd)
{
  int cw = (W + 63) / 64, npu = 0;
  int pus[4 * 340];
  for (int a = ctu_begin; a < ctu_end; a++) {            /* phase 1: enumerate (cheap, ordered) */
    int k = oracle_enum_ctu_pus(labels + (size_t)a * 16, a % cw, a / cw, W, H, pus);
    memcpy(out_pu + (size_t)npu * 4, pus, (size_t)k * 4 * sizeof(int));
    npu += k;
  }
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < npu; i++) {                        /* phase 2: PUs are independent */
    int16_t line[4 * MAXN + 1];
    int x = out_pu[4 * i], y = out_pu[4 * i + 1], n = out_pu[4 * i + 2];
    oracle_build_ref_line(pic, W, W, H, x, y, n, line);
    oracle_pu_satd35(pic + (size_t)y * W + x, W, line, n, out_satd + (size_t)i * 35);
  }
  return npu;
}
