// sao.cuh -- SAO statistics of a deblocked picture on the device: the data pass of the reference's SAO
// parameter estimation, TEncSampleAdaptiveOffset::getStatistics / getBlkStats (HM TLibEncoder/TEncSampleAdaptiveOffset.cpp:
// 295-341, 943-1345) for deblocked samples (isCalculatePreDeblockSamples = false), one slice, no tiles, 8-bit 4:2:0, 64x64
// CTUs.  Per CTU, colour component and SAO type (edge offset 0 / 90 / 135 / 45 degrees, band offset): per class the number of
// samples and the sum of (original - deblocked).  The parameter decision (an RD search with CABAC bit estimates over these
// 7.7 KB per CTU) stays the reference's.
//
// Byte work bound by HBM: one block per (CTU, component) reads the CTU's samples of both pictures once (plus a one-sample
// ring of the deblocked picture) into shared memory, every thread classifies its samples for all five types and adds into
// per-warp shared-memory histograms (one packed word per class: count << 20 + sum of differences), which are unpacked, summed
// and written out as the reference's int64 arrays.  What a CTU counts excludes the columns / rows its right / lower neighbours have not deblocked yet (5 / 4 luma,
// 3 / 2 chroma) and, for the edge types, samples whose neighbour lies outside the picture.
#pragma once
#include "common.cuh"

namespace hevcdl {

struct SaoParams {
  const int16_t *org[3], *src[3];
  int W, H, ctu_w;
  long long *out;            // [nctu][3][5][2][32]
};

__global__ void __launch_bounds__(256)
k_sao_stats(const SaoParams P) {
  __shared__ int16_t s_src[66][66];                          // the CTU's deblocked samples with a one-sample ring
  __shared__ int s_hist[8][5][32];                           // per warp, type and class: count << 20 + sum of differences
  const int a = blockIdx.x / 3, c = blockIdx.x - 3 * a;
  const int xp = (a % P.ctu_w) * 64, yp = (a / P.ctu_w) * 64;
  const int hl = yp + 64 > P.H ? P.H - yp : 64, wl = xp + 64 > P.W ? P.W - xp : 64;
  const bool left = xp > 0, above = yp > 0, right = xp + 64 < P.W, below = yp + 64 < P.H;
  const int sh = c ? 1 : 0, stride = P.W >> sh, pw = P.W >> sh, ph = P.H >> sh, width = wl >> sh, height = hl >> sh;
  const int x0p = xp >> sh, y0p = yp >> sh;
  const int16_t *__restrict__ src = P.src[c], *__restrict__ org = P.org[c];
  for (int i = threadIdx.x; i < 8 * 5 * 32; i += blockDim.x) (&s_hist[0][0][0])[i] = 0;
  int (*hist)[32] = s_hist[threadIdx.x >> 5];
  {
    // one warp per row, a lane per column: every load of the thread (<= 9 rows x 3 column steps) is issued before the first
    // store, so the block pays one global-memory latency for its tile instead of one per row
    int16_t t[9][3];
#pragma unroll
    for (int k = 0; k < 9; k++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int ly = (threadIdx.x >> 5) + 8 * k, lx = (threadIdx.x & 31) + 32 * j;
        const int gy = y0p + ly - 1, gx = x0p + lx - 1;
        t[k][j] = (ly < height + 2 && lx < width + 2 && gy >= 0 && gy < ph && gx >= 0 && gx < pw) ? src[(size_t)gy * stride + gx] : (int16_t)0;   // ring samples outside the picture are never used
      }
#pragma unroll
    for (int k = 0; k < 9; k++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int ly = (threadIdx.x >> 5) + 8 * k, lx = (threadIdx.x & 31) + 32 * j;
        if (ly < height + 2 && lx < width + 2) s_src[ly][lx] = t[k][j];
      }
  }
  __syncthreads();
  const int skip_r = c ? 3 : 5, skip_b = c ? 2 : 4;
  // per-type ranges (getBlkStats): x in [x0, x1), y in [y0, y1); the diagonal types restrict their first line further
  const int ex0 = left ? 0 : 1, ex1 = right ? width - skip_r : width - 1;          // types that look left / right
  const int fx1 = right ? width - skip_r : width;                                   // types that do not (EO 90, BO)
  const int ey1 = below ? height - skip_b : height - 1, fy1 = below ? height - skip_b : height;
  // a thread owns one column and every `bands`-th row; a warp's (count, sum) of a class is ONE word -- count << 20 + sum,
  // |sum| <= 512 samples x 255 < 2^19 -- so a sample costs one shared-memory atomic per type
  const int cols = width > 32 ? 64 : 32, bands = 256 / cols;
  const int x = threadIdx.x % cols;
  const bool inx = x >= ex0 && x < ex1;
  int16_t og[16];                               // the thread's original samples (<= 16 rows), loaded ahead of the loop
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int y = threadIdx.x / cols + bands * k;
    og[k] = (x < width && y < height) ? org[(size_t)(y0p + y) * stride + x0p + x] : (int16_t)0;
  }
  if (x < width) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int y = threadIdx.x / cols + bands * k;
      if (y >= height) break;
      const int v = s_src[y + 1][x + 1];
      const int w = (1 << 20) + (int)og[k] - v;
      auto sgn = [&](int ly, int lx) { const int n = s_src[ly][lx]; return (v > n) - (v < n); };
      if (inx && y < fy1) atomicAdd(&hist[0][2 + sgn(y + 1, x) + sgn(y + 1, x + 2)], w);
      if (y < ey1) {
        if (x < fx1 && y >= (above ? 0 : 1)) atomicAdd(&hist[1][2 + sgn(y, x + 1) + sgn(y + 2, x + 1)], w);
        const bool in135 = y == 0 ? (x >= ((left && above) ? 0 : 1) && x < (above ? ex1 : 1)) : inx;
        if (in135) atomicAdd(&hist[2][2 + sgn(y, x) + sgn(y + 2, x + 2)], w);
        const bool in45 = y == 0 ? (above && inx) : inx;
        if (in45) atomicAdd(&hist[3][2 + sgn(y, x + 2) + sgn(y + 2, x)], w);
      }
      if (x < fx1 && y < fy1) atomicAdd(&hist[4][v >> 3], w);
    }
  }
  __syncthreads();
  long long *o = P.out + ((size_t)a * 3 + c) * 5 * 64;
  for (int i = threadIdx.x; i < 5 * 2 * 32; i += blockDim.x) {
    const int t = i >> 6, which = (i >> 5) & 1, k = i & 31;
    long long r = 0;
#pragma unroll
    for (int wq = 0; wq < 8; wq++) {
      const int word = s_hist[wq][t][k], n = (word + (1 << 19)) >> 20;
      r += which ? n : word - (n << 20);
    }
    o[i] = r;
  }
}

// ---- SAO application -----------------------------------------------------------------------------------------------
// TComSampleAdaptiveOffset::offsetCTU / offsetBlock (HM TLibCommon/TComSampleAdaptiveOffset.cpp:554-611, 313-552) for every
// CTU of the picture: res = clip(src + offset[class]) for the samples the CTU's SAO type may touch (edge types leave out
// samples whose neighbour lies outside the picture -- one slice, no tiles -- with the reference's first / last line ranges
// of the diagonal types), every other sample keeps its deblocked value.  Classification reads only `src`, so all CTUs are
// independent: one block per (CTU, component), one warp per row (the row's valid range is decided once), neighbours straight
// from global memory (each sample is read at most three times, from L1 / L2).  Byte work: two pictures' worth of int16 traffic.
struct SaoApplyParams {
  const int16_t *src[3];
  int16_t *res[3];
  int W, H, ctu_w;
  const int8_t *type;        // [nctu][3]: -1 off, 0..3 edge offset 0 / 90 / 135 / 45 degrees, 4 band offset
  const int8_t *offset;      // [nctu][3][32]
};

__global__ void __launch_bounds__(256)
k_sao_apply(const SaoApplyParams P) {
  __shared__ int s_off[32];
  const int a = blockIdx.x / 3, c = blockIdx.x - 3 * a;
  const int xp = (a % P.ctu_w) * 64, yp = (a / P.ctu_w) * 64;
  const int hl = yp + 64 > P.H ? P.H - yp : 64, wl = xp + 64 > P.W ? P.W - xp : 64;
  const int sh = c ? 1 : 0, stride = P.W >> sh, width = wl >> sh, height = hl >> sh;
  const int16_t *__restrict__ s = P.src[c] + (size_t)(yp >> sh) * stride + (xp >> sh);
  int16_t *__restrict__ r = P.res[c] + (size_t)(yp >> sh) * stride + (xp >> sh);
  const int t = P.type[a * 3 + c];
  if (threadIdx.x < 32) s_off[threadIdx.x] = P.offset[(size_t)(a * 3 + c) * 32 + threadIdx.x];
  __syncthreads();
  const bool L = xp > 0, A = yp > 0, R = xp + 64 < P.W, B = yp + 64 < P.H, AL = L && A, AR = A && R, BL = B && L, BR = B && R;
  const int sx = L ? 0 : 1, ex = R ? width : width - 1;
  const int dx = t == 1 ? 0 : (t == 3 ? -1 : 1), dy = t == 0 ? 0 : 1;   // second neighbour at (+dx, +dy), first at (-dx, -dy)
  // one warp per row, a lane per column (two for 64-wide luma rows); the thread's centre samples (<= 8 rows x 2 column steps)
  // are all loaded before the first one is used, so the row loop does not pay one global-memory latency per row
  int16_t cv[8][2];
#pragma unroll
  for (int k = 0; k < 8; k++)
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int y = (threadIdx.x >> 5) + 8 * k, x = (threadIdx.x & 31) + 32 * j;
      cv[k][j] = (y < height && x < width) ? s[(size_t)y * stride + x] : (int16_t)0;
    }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int y = (threadIdx.x >> 5) + 8 * k;
    if (y >= height) break;
    int xa = 0, xb = 0;                           // t < 0: nothing is offset; the row's valid range is decided once
    if (t == 4) { xb = width; }
    else if (t == 0) { xa = sx; xb = ex; }
    else if (t == 1) { xb = (y < (A ? 0 : 1) || y >= (B ? height : height - 1)) ? 0 : width; }
    else if (t == 2) {
      if (y == 0) { xa = AL ? 0 : 1; xb = A ? ex : 1; }
      else if (y == height - 1) { xa = B ? sx : width - 1; xb = BR ? width : width - 1; }
      else { xa = sx; xb = ex; }
    } else if (t == 3) {
      if (y == 0) { xa = A ? sx : width - 1; xb = AR ? width : width - 1; }
      else if (y == height - 1) { xa = BL ? 0 : 1; xb = B ? ex : 1; }
      else { xa = sx; xb = ex; }
    }
    const int16_t *__restrict__ row = s + (size_t)y * stride;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int x = (threadIdx.x & 31) + 32 * j;
      if (x >= width) break;
      const int v = cv[k][j];
      int out = v;
      if (x >= xa && x < xb) {
        int cls;
        if (t == 4) cls = v >> 3;
        else {
          const int n0 = row[-dy * stride + x - dx], n1 = row[dy * stride + x + dx];
          cls = 2 + ((v > n0) - (v < n0)) + ((v > n1) - (v < n1));
        }
        out = min(255, max(0, v + s_off[cls]));
      }
      r[(size_t)y * stride + x] = (int16_t)out;
    }
  }
}

}  // namespace hevcdl
