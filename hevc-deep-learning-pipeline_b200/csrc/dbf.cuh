// dbf.cuh -- deblocking filter of an all-intra picture on the device, bit-exact with the reference's
// TComLoopFilter::loopFilterPic (HM TLibCommon/TComLoopFilter.cpp:130-158; edge selection :170-360, Bs :416-555 -- 2 everywhere in
// an all-intra picture --, luma filter :557-674 / :830-899, chroma filter :676-828 / :901-931, decisions :933-953, tables :59-67,
// chroma QP mapping TComRom.cpp:532-539) at the reference's operating point: 8-bit 4:2:0, one slice, no tiles, no PCM /
// lossless blocks.
//
// The filter is byte work bound by HBM: every sample is read and written once per pass.  All vertical edges of the picture
// are filtered first, then all horizontal ones (the two CTU loops of loopFilterPic); inside a pass the edges lie 8 samples
// apart and touch at most 3 samples either side, so every 4-sample edge segment is independent: one thread per segment,
// threads consecutive along x (vertical edges: each lane reads 16 contiguous bytes per line, a warp one contiguous row span;
// horizontal edges: 8 bytes per lane and row).  A thread also filters the two chroma lines of its segment on the 8x8 chroma
// grid.  An edge exists where a transform / coding block starts: position % (TU size at that 4x4 unit) == 0.
#pragma once
#include "common.cuh"

namespace hevcdl {

__device__ __constant__ uint8_t c_dbf_tc[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4,
                                                4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};
__device__ __constant__ uint8_t c_dbf_beta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24,
                                                  26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};
__device__ __constant__ uint8_t c_dbf_qpc[58] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28,
                                                 29, 29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47,
                                                 48, 49, 50, 51};

__device__ __forceinline__ int dbf_clip3(int lo, int hi, int v) { return min(hi, max(lo, v)); }

struct DbfParams {
  int16_t *Y, *U, *V;
  int sy, sc, W, H;
  const uint8_t *tu_log2;
  const int8_t *qp;
  int beta_off, tc_off, cb_off, cr_off;
};

// VERT = true: vertical edges (filtering across x); one thread per 4-sample segment of the 8x8 edge grid
template <bool VERT>
__global__ void __launch_bounds__(256)
k_dbf(const DbfParams P) {
  const int w4 = P.W >> 2;
  // segment grid: vertical edges x = 8, 16, ... (W/8 - 1 per row of 4 lines); horizontal edges y = 8, 16, ... for every 4 columns
  const int nx = VERT ? (P.W >> 3) - 1 : (P.W >> 2), ny = VERT ? (P.H >> 2) : (P.H >> 3) - 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nx * ny) return;
  const int ix = idx % nx, iy = idx / nx;
  const int x = VERT ? 8 * (ix + 1) : 4 * ix, y = VERT ? 4 * iy : 8 * (iy + 1);
  const int tsz = 1 << P.tu_log2[(y >> 2) * w4 + (x >> 2)];
  if (((VERT ? x : y) & (tsz - 1)) != 0) return;            // not a transform / coding block boundary
  const int qq = P.qp[(y >> 2) * w4 + (x >> 2)], qp_p = VERT ? P.qp[(y >> 2) * w4 + ((x - 1) >> 2)] : P.qp[((y - 1) >> 2) * w4 + (x >> 2)];
  const int q = (qq + qp_p + 1) >> 1;
  {
    const int tc = c_dbf_tc[dbf_clip3(0, 53, q + 2 + (P.tc_off << 1))], beta = c_dbf_beta[dbf_clip3(0, 51, q + (P.beta_off << 1))];
    int16_t *p = P.Y + (size_t)y * P.sy + x;
    const int across = VERT ? 1 : P.sy, along = VERT ? P.sy : 1;
    int s[4][8];                                            // [line][p3 p2 p1 p0 q0 q1 q2 q3]
#pragma unroll
    for (int l = 0; l < 4; l++)
#pragma unroll
      for (int k = 0; k < 8; k++) s[l][k] = p[l * along + (k - 4) * across];
    const int dp0 = abs(s[0][1] - 2 * s[0][2] + s[0][3]), dq0 = abs(s[0][4] - 2 * s[0][5] + s[0][6]);
    const int dp3 = abs(s[3][1] - 2 * s[3][2] + s[3][3]), dq3 = abs(s[3][4] - 2 * s[3][5] + s[3][6]);
    const int d0 = dp0 + dq0, d3 = dp3 + dq3, d = d0 + d3;
    if (d < beta) {
      const int side = (beta + (beta >> 1)) >> 3;
      const bool fp = dp0 + dp3 < side, fq = dq0 + dq3 < side;
      const bool strong = (abs(s[0][0] - s[0][3]) + abs(s[0][7] - s[0][4]) < (beta >> 3)) && (2 * d0 < (beta >> 2)) && (abs(s[0][3] - s[0][4]) < ((tc * 5 + 1) >> 1)) &&
                          (abs(s[3][0] - s[3][3]) + abs(s[3][7] - s[3][4]) < (beta >> 3)) && (2 * d3 < (beta >> 2)) && (abs(s[3][3] - s[3][4]) < ((tc * 5 + 1) >> 1));
#pragma unroll
      for (int l = 0; l < 4; l++) {
        const int m0 = s[l][0], m1 = s[l][1], m2 = s[l][2], m3 = s[l][3], m4 = s[l][4], m5 = s[l][5], m6 = s[l][6], m7 = s[l][7];
        int16_t *pl = p + l * along;
        if (strong) {
          pl[-across] = (int16_t)dbf_clip3(m3 - 2 * tc, m3 + 2 * tc, (m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3);
          pl[0] = (int16_t)dbf_clip3(m4 - 2 * tc, m4 + 2 * tc, (m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3);
          pl[-2 * across] = (int16_t)dbf_clip3(m2 - 2 * tc, m2 + 2 * tc, (m1 + m2 + m3 + m4 + 2) >> 2);
          pl[across] = (int16_t)dbf_clip3(m5 - 2 * tc, m5 + 2 * tc, (m3 + m4 + m5 + m6 + 2) >> 2);
          pl[-3 * across] = (int16_t)dbf_clip3(m1 - 2 * tc, m1 + 2 * tc, (2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3);
          pl[2 * across] = (int16_t)dbf_clip3(m6 - 2 * tc, m6 + 2 * tc, (m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3);
        } else {
          int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
          if (abs(delta) < tc * 10) {
            delta = dbf_clip3(-tc, tc, delta);
            pl[-across] = (int16_t)dbf_clip3(0, 255, m3 + delta);
            pl[0] = (int16_t)dbf_clip3(0, 255, m4 - delta);
            const int tc2 = tc >> 1;
            if (fp) pl[-2 * across] = (int16_t)dbf_clip3(0, 255, m2 + dbf_clip3(-tc2, tc2, (((m1 + m3 + 1) >> 1) - m2 + delta) >> 1));
            if (fq) pl[across] = (int16_t)dbf_clip3(0, 255, m5 + dbf_clip3(-tc2, tc2, (((m6 + m4 + 1) >> 1) - m5 - delta) >> 1));
          }
        }
      }
    }
  }
  if (((VERT ? x : y) & 15) == 0) {                         // chroma: 8x8 chroma sample grid, two chroma lines per 4 luma lines
#pragma unroll
    for (int c = 0; c < 2; c++) {
      int qc = q + (c ? P.cr_off : P.cb_off);
      if (qc >= 58) qc -= 6; else if (qc >= 0) qc = c_dbf_qpc[qc];
      const int tc = c_dbf_tc[dbf_clip3(0, 53, qc + 2 + (P.tc_off << 1))];
      int16_t *pc = (c ? P.V : P.U) + (size_t)(y >> 1) * P.sc + (x >> 1);
      const int across = VERT ? 1 : P.sc, along = VERT ? P.sc : 1;
#pragma unroll
      for (int l = 0; l < 2; l++) {
        int16_t *pl = pc + l * along;
        const int m2 = pl[-2 * across], m3 = pl[-across], m4 = pl[0], m5 = pl[across];
        const int delta = dbf_clip3(-tc, tc, ((((m4 - m3) << 2) + m2 - m5 + 4) >> 3));
        pl[-across] = (int16_t)dbf_clip3(0, 255, m3 + delta);
        pl[0] = (int16_t)dbf_clip3(0, 255, m4 - delta);
      }
    }
  }
}

}  // namespace hevcdl
