// rmd.cuh -- K6: label-driven PU / work-item plan and the 35-mode intra SATD ("RMD") pass (batched over the frames
// of a launch), plus the exact single-PU entry point.
//
// Replaces, for every PU of a frame at once, the first pass of TEncSearch::estIntraPredLumaQT
// (HM TLibEncoder/TEncSearch.cpp:2266-2346): reference-sample construction and smoothing
// (HM TLibCommon/TComPattern.cpp:119-543), planar/DC/33 angular predictors
// (HM TLibCommon/TComPrediction.cpp:183-473,731-817), Hadamard SATD
// (HM TLibCommon/TComRdCost.cpp:1549-1824) and the candidate list (TEncSearch.cpp:5562-5585).
// Integer arithmetic throughout; results are bit-exact with the reference given the same inputs.
//
// Reference-sample line layout (n = PU size, 4n+1 samples):
//   line[0..2n-1] left column bottom-up (below-left first), line[2n] corner, line[2n+1..4n] above row.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace hevcdl {

constexpr int MAX_PU_CTU = 320;                 // 64 8x8 CUs x (1 + 4 NxN PUs)

__device__ __forceinline__ int zidx4(int ux, int uy) {   // z-order of a 4x4 unit in a CTU (TComRom.cpp:284); ux, uy in [0,16)
  uint32_t v = (uint32_t)ux | ((uint32_t)uy << 8);       // both coordinates spread at once, one per byte
  v = (v | (v << 2)) & 0x3333u;
  v = (v | (v << 1)) & 0x5555u;
  return (int)((v | (v >> 7)) & 0xFFu);                  // x bits on even positions, y bits on odd ones
}

// Neighbour unit available iff inside the picture and earlier in coding order (CTU raster, then
// z-order) -- the net effect of TComDataCU::getPU{Left,Above,AboveLeft,AboveRight,BelowLeft}
// (HM TLibCommon/TComDataCU.cpp:1000-1200) for one slice, no tiles, constrained intra pred off.
__device__ __forceinline__ bool unit_available(int xn, int yn, int xc, int yc, int W, int H, int ctu_w) {
  if (xn < 0 || yn < 0 || xn >= W || yn >= H) return false;
  const int on = ((yn >> 6) * ctu_w + (xn >> 6)) * 256 + zidx4((xn & 63) >> 2, (yn & 63) >> 2);
  const int oc = ((yc >> 6) * ctu_w + (xc >> 6)) * 256 + zidx4((xc & 63) >> 2, (yc & 63) >> 2);
  return on < oc;
}

// ---- reference samples ----------------------------------------------------------------------
// [1 2 1] or strong (bilinear, n == 32 only since 64x64 never uses filtered samples) smoothing of
// a line (HM TComPattern.cpp:203-294); one warp.
__device__ __forceinline__ void filter_line_warp(const int16_t *line, int16_t *filt, int n, int lane) {
  const int len = 4 * n + 1;
  const int bl = line[0], tl = line[2 * n], tr = line[4 * n];
  bool strong = false;
  if (n == 32) strong = (abs(bl + tl - 2 * line[n]) < 8) && (abs(tl + tr - 2 * line[3 * n]) < 8);
  for (int i = lane; i < len; i += 32) {
    int v;
    if (i == 0 || i == len - 1) v = line[i];
    else if (!strong) v = (line[i - 1] + 2 * line[i] + line[i + 1] + 2) >> 2;
    else if (i < 2 * n) v = ((2 * n - i) * bl + i * tl + n) >> 6;
    else if (i == 2 * n) v = tl;
    else v = ((4 * n - i) * tl + (i - 2 * n) * tr + n) >> 6;
    filt[i] = (int16_t)v;
  }
}

// filtered references are used iff min(|m-10|,|m-26|) > thr[size]; never for DC
// (HM TComPattern.cpp:545-570, table TComPrediction.cpp:50-58)
// the same rule as a 35-bit mask per PU size (bit m set: mode m predicts from the filtered line)
__host__ __device__ constexpr unsigned long long mode_filter_mask(int n) {
  unsigned long long m = 0;
  const int thr = n == 8 ? 7 : (n == 16 ? 1 : 0);
  for (int mode = 0; mode < 35; mode++) {
    const int d10 = mode > 10 ? mode - 10 : 10 - mode, d26 = mode > 26 ? mode - 26 : 26 - mode;
    if (mode != 1 && n != 4 && n != 64 && (d10 < d26 ? d10 : d26) > thr) m |= 1ull << mode;
  }
  return m;
}
__device__ __forceinline__ bool mode_uses_filter(int mode, int n) {
  if (mode == 1 || n == 4 || n == 64) return false;
  const int thr = n == 8 ? 7 : (n == 16 ? 1 : 0);
  return min(abs(mode - 10), abs(mode - 26)) > thr;
}

// ---- one SATD unit: one B x B block (B = 8, or 4 for a 4x4 PU) of one PU for one mode ---------
__device__ __constant__ int8_t c_ang[9] = {0, 2, 5, 9, 13, 17, 21, 26, 32};
__device__ __constant__ int16_t c_inv[9] = {0, 4096, 1638, 910, 630, 482, 390, 315, 256};

template <int B>
__device__ __forceinline__ uint32_t satd_unit(const uint8_t *__restrict__ org, int ostride,   // block origin
                                              const int16_t *__restrict__ ref,                // chosen line
                                              const int16_t *__restrict__ line,               // unfiltered (DC)
                                              int n, int lg, int mode, int bx, int by, int dc) {
  int d[B][B];
  const int16_t *c = ref + 2 * n;                 // c[0] corner, c[1+k] above k, c[-1-k] left k
  if (mode == 0) {                                // planar (TComPrediction.cpp:731-781)
    const int blv = c[-1 - n], trv = c[1 + n];
#pragma unroll
    for (int y = 0; y < B; y++) {
      const int l = c[-1 - (by + y)];
#pragma unroll
      for (int x = 0; x < B; x++) {
        const int t = c[1 + bx + x];
        const int hor = (l << lg) + n + (bx + x + 1) * (trv - l);
        const int ver = (t << lg) + (by + y + 1) * (blv - t);
        d[y][x] = (int)org[y * ostride + x] - ((hor + ver) >> (lg + 1));
      }
    }
  } else if (mode == 1) {                         // DC + edge filter for n <= 16 (:183-201,794-817)
    const int16_t *u = line + 2 * n;
#pragma unroll
    for (int y = 0; y < B; y++)
#pragma unroll
      for (int x = 0; x < B; x++) {
        int p = dc;
        if (n <= 16) {
          const int gx = bx + x, gy = by + y;
          if (gx == 0 && gy == 0) p = (u[1] + u[-1] + 2 * dc + 2) >> 2;
          else if (gy == 0) p = (u[1 + gx] + 3 * dc + 2) >> 2;
          else if (gx == 0) p = (u[-1 - gy] + 3 * dc + 2) >> 2;
        }
        d[y][x] = (int)org[y * ostride + x] - p;
      }
  } else {                                        // angular (:229-388)
    const bool ver = mode >= 18;
    const int am = ver ? mode - 26 : 10 - mode;
    const int aabs = abs(am);
    const int angle = am < 0 ? -(int)c_ang[aabs] : (int)c_ang[aabs];
    const int inv = c_inv[aabs];
    const int sg = ver ? 1 : -1;                  // main array runs along +line for vertical modes
    // main(i) for i>=0: c[sg*i]; for i<0 projected from the side array: c[-sg*((128 + (-i)*inv) >> 8)]
    const int j0 = ver ? by : bx, i0 = ver ? bx : by;   // j along the prediction direction
    const bool edge = (angle == 0) && (n <= 16);
#pragma unroll
    for (int j = 0; j < B; j++) {
      const int pos = (j0 + j + 1) * angle, di = pos >> 5, df = pos & 31;
#pragma unroll
      for (int i = 0; i < B; i++) {
        const int k = i0 + i + di + 1;
        const int a = k >= 0 ? c[sg * k] : c[-sg * ((128 - k * inv) >> 8)];
        int p = a;
        if (df) {
          const int k1 = k + 1;
          const int b = k1 >= 0 ? c[sg * k1] : c[-sg * ((128 - k1 * inv) >> 8)];
          p = ((32 - df) * a + df * b + 16) >> 5;
        }
        if (edge && (i0 + i) == 0) p = clip255(p + ((c[-sg * (j0 + j + 1)] - c[0]) >> 1));
        const int yy = ver ? j : i, xx = ver ? i : j;
        d[yy][xx] = (int)org[yy * ostride + xx] - p;
      }
    }
  }
  // 2-D Hadamard, sum of magnitudes (ordering-independent): rows then columns
#pragma unroll
  for (int y = 0; y < B; y++) {
#pragma unroll
    for (int h = 1; h < B; h <<= 1)
#pragma unroll
      for (int i = 0; i < B; i += 2 * h)
#pragma unroll
        for (int j = i; j < i + h; j++) {
          const int a = d[y][j], b = d[y][j + h];
          d[y][j] = a + b; d[y][j + h] = a - b;
        }
  }
  uint32_t s = 0;
#pragma unroll
  for (int x = 0; x < B; x++) {
#pragma unroll
    for (int h = 1; h < B; h <<= 1)
#pragma unroll
      for (int i = 0; i < B; i += 2 * h)
#pragma unroll
        for (int j = i; j < i + h; j++) {
          const int a = d[j][x], b = d[j + h][x];
          d[j][x] = a + b; d[j + h][x] = a - b;
        }
#pragma unroll
    for (int y = 0; y < B; y++) s += (uint32_t)abs(d[y][x]);
  }
  return B == 8 ? (s + 2) >> 2 : (s + 1) >> 1;    // TComRdCost.cpp:1739-1749 / :1636-1640
}

__device__ __forceinline__ int num_rd_modes(int n) { return n >= 16 ? 3 : 8; }  // TComRom.cpp:545-553

// Candidate list by cost = satd + bits*sqrt_lambda (double), strict '<' insertion from the worst
// slot (TEncSearch.cpp:2313,5562-5585).  bits may be null (cost = satd).  Returns list length
// after appending missing MPMs (TEncSearch.cpp:2322-2345) when mpm != null.
__device__ inline int cand_list(const uint32_t *satd, const uint32_t *bits, double sqrt_lambda, int n,
                                const int8_t *mpm, int mpm_add, uint8_t *modes) {
  const int keep = num_rd_modes(n);
  double cl[8];
  uint8_t ml[10];
  for (int i = 0; i < keep; i++) { cl[i] = 1.7e308; ml[i] = 0; }
  for (int m = 0; m < 35; m++) {
    const double c = (double)satd[m] + (bits ? (double)bits[m] * sqrt_lambda : 0.0);
    int shift = 0;
    while (shift < keep && c < cl[keep - 1 - shift]) shift++;
    if (shift) {
      for (int i = 1; i < shift; i++) { ml[keep - i] = ml[keep - 1 - i]; cl[keep - i] = cl[keep - 1 - i]; }
      ml[keep - shift] = (uint8_t)m; cl[keep - shift] = c;
    }
  }
  int len = keep;
  if (mpm)
    for (int j = 0; j < mpm_add; j++) {
      bool inc = false;
      for (int i = 0; i < len; i++) inc |= (mpm[j] == (int8_t)ml[i]);
      if (!inc) ml[len++] = (uint8_t)mpm[j];
    }
  for (int i = 0; i < len; i++) modes[i] = ml[i];
  return len;
}

__device__ __forceinline__ int ilog2(int n) { return 31 - __clz(n); }

// ---- device helpers of the batched path ----------------------------------------------------------

__device__ __forceinline__ void mma_f16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                              uint32_t b1) {
  asm(  // not volatile: a pure function of its operands, so the compiler may interleave the chains of neighbouring slabs
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Two neighbouring predicted pixels of one PU for one mode, packed (first | second << 16).
// (X, Yc): position of the first pixel inside the PU; the second is (X+1, Yc) for vertical-type
// modes (planar, DC, 18..34) and (X, Yc+1) for horizontal modes (2..17).  c = centre (corner) of
// the reference line chosen for this mode, u = centre of the unfiltered line (DC edge filter).
__device__ __forceinline__ uint32_t predict_pair(const int16_t *__restrict__ c, const int16_t *__restrict__ u, int n, int lg,
                                                 int mode, int X, int Yc, int dc) {
  if (mode >= 2) {                                // angular (TComPrediction.cpp:229-388)
    const bool ver = mode >= 18;
    const int am = ver ? mode - 26 : 10 - mode;
    const int aabs = abs(am);
    const int angle = am < 0 ? -(int)c_ang[aabs] : (int)c_ang[aabs];
    const int inv = c_inv[aabs];
    const int sg = ver ? 1 : -1;
    const int jj = ver ? Yc : X, ii = ver ? X : Yc;     // jj: distance from the main reference, ii: along it
    const int pos = (jj + 1) * angle, di = pos >> 5, df = pos & 31;
    const int k = ii + di + 1;
    int k0 = sg * k, k1 = k0 + sg, k2 = k1 + sg;
    if (angle < 0) {                              // warp-uniform: negative angles project the side reference for k < 0
      if (k < 0) k0 = -sg * ((128 - k * inv) >> 8);
      if (k + 1 < 0) k1 = -sg * ((128 - (k + 1) * inv) >> 8);
      if (k + 2 < 0) k2 = -sg * ((128 - (k + 2) * inv) >> 8);
    }
    const uint32_t s0 = (uint16_t)c[k0], s1 = (uint16_t)c[k1], s2 = (uint16_t)c[k2];
    const uint32_t A = s0 | (s1 << 16), B = s1 | (s2 << 16);
    uint32_t P = (((32 - df) * A + df * B + 0x00100010u) >> 5) & 0x07FF07FFu;
    if (angle == 0 && n <= 16 && ii == 0) {        // pure V/H edge filter on the first column along the reference
      const int p0 = clip255((int)(P & 0xFFFF) + ((c[-sg * (jj + 1)] - c[0]) >> 1));
      P = (P & 0xFFFF0000u) | (uint32_t)p0;
    }
    return P;
  }
  int p0, p1;
  const int X1 = X + 1;                            // vertical-type pairing: second pixel to the right
  if (mode == 0) {                                 // planar (TComPrediction.cpp:731-781)
    const int blv = c[-1 - n], trv = c[1 + n], l = c[-1 - Yc];
    const int t0 = c[1 + X], t1 = c[1 + X1];
    p0 = ((l << lg) + n + (X + 1) * (trv - l) + (t0 << lg) + (Yc + 1) * (blv - t0)) >> (lg + 1);
    p1 = ((l << lg) + n + (X1 + 1) * (trv - l) + (t1 << lg) + (Yc + 1) * (blv - t1)) >> (lg + 1);
  } else {                                         // DC + edge filter for n <= 16 (:183-201,794-817)
    p0 = p1 = dc;
    if (n <= 16) {
      if (Yc == 0) {
        p0 = X == 0 ? (u[1] + u[-1] + 2 * dc + 2) >> 2 : (u[1 + X] + 3 * dc + 2) >> 2;
        p1 = (u[1 + X1] + 3 * dc + 2) >> 2;
      } else if (X == 0) {
        p0 = (u[-1 - Yc] + 3 * dc + 2) >> 2;
      }
    }
  }
  return (uint32_t)p0 | ((uint32_t)p1 << 16);
}

__device__ __constant__ int8_t c_mode_angle[35] = {0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26,
                                                   -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32};
__device__ __constant__ int16_t c_mode_inv[35] = {0, 0, 256, 315, 390, 482, 630, 910, 1638, 4096, 0, 4096, 1638, 910, 630, 482, 390, 315,
                                                  256, 315, 390, 482, 630, 910, 1638, 4096, 0, 4096, 1638, 910, 630, 482, 390, 315, 256};

// What block_large needs to know about (PU size, mode), one word: bit 0 filtered references (TComPattern.cpp:545-570),
// bit 1 horizontal class (modes 2..17), bit 2 not served by the table interpolation (planar, DC, pure H/V with the edge
// filter), bit 3 negative angle; bits 8..15 the angle, bits 16..31 the inverse angle.  Index 0/1/2 = PU size 16/32/64.
struct ModeInfoTab { uint32_t v[3][35]; };
__host__ __device__ constexpr ModeInfoTab make_mode_info() {
  ModeInfoTab t{};
  constexpr int ang[35] = {0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26,
                           -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32};
  constexpr int inv[35] = {0, 0, 256, 315, 390, 482, 630, 910, 1638, 4096, 0, 4096, 1638, 910, 630, 482, 390, 315,
                           256, 315, 390, 482, 630, 910, 1638, 4096, 0, 4096, 1638, 910, 630, 482, 390, 315, 256};
  for (int si = 0; si < 3; si++) {
    const int n = 16 << si;
    const unsigned long long fm = mode_filter_mask(n);
    for (int m = 0; m < 35; m++) {
      const uint32_t flt = (uint32_t)((fm >> m) & 1ull), hor = (m >= 2 && m < 18) ? 1u : 0u;
      const uint32_t slow = (m < 2 || (ang[m] == 0 && n <= 16)) ? 1u : 0u, neg = ang[m] < 0 ? 1u : 0u;
      t.v[si][m] = flt | (hor << 1) | (slow << 2) | (neg << 3) | (((uint32_t)ang[m] & 0xFFu) << 8) | ((uint32_t)inv[m] << 16);
    }
  }
  return t;
}
__device__ __constant__ ModeInfoTab c_mode_info = make_mode_info();

// Everything about (PU, mode) that does not depend on the pixel: computed once per work item.
struct ModeK {
  const int16_t *c;      // centre of the reference line used by this mode (filtered or not)
  const int16_t *u;      // centre of the unfiltered line
  int mode, n, lg, dc, angle, inv, sg;
  bool hor, edge;
};
__device__ __forceinline__ ModeK make_modek(const int16_t *line, int n, int mode, int dc) {
  ModeK k;
  const int16_t *ref = mode_uses_filter(mode, n) ? line + ((4 * n + 2) & ~1) : line;
  k.c = ref + 2 * n; k.u = line + 2 * n;
  k.mode = mode; k.n = n; k.lg = ilog2(n); k.dc = dc;
  k.angle = c_mode_angle[mode]; k.inv = c_mode_inv[mode];
  k.hor = mode >= 2 && mode < 18;
  k.sg = k.hor ? -1 : 1;
  k.edge = mode >= 2 && k.angle == 0 && n <= 16;
  return k;
}
// Two neighbouring predicted pixels, packed as exact fp16 integers (1024 + value each): first pixel (X, Yc),
// second (X+1, Yc) for vertical-type modes / (X, Yc+1) for horizontal modes.
__device__ __forceinline__ uint32_t predict_pair_k(const ModeK &k, int X, int Yc) {
  if (k.mode >= 2) {
    const int jj = k.hor ? X : Yc, ii = k.hor ? Yc : X;
    const int pos = (jj + 1) * k.angle, di = pos >> 5, df = pos & 31;
    const int kk = ii + di + 1;
    int k0 = k.sg * kk, k1 = k0 + k.sg, k2 = k1 + k.sg;
    if (k.angle < 0) {                            // warp-uniform: negative angles project the side reference for kk < 0
      if (kk < 0) k0 = -k.sg * ((128 - kk * k.inv) >> 8);
      if (kk + 1 < 0) k1 = -k.sg * ((128 - (kk + 1) * k.inv) >> 8);
      if (kk + 2 < 0) k2 = -k.sg * ((128 - (kk + 2) * k.inv) >> 8);
    }
    const uint32_t s0 = (uint16_t)k.c[k0], s1 = (uint16_t)k.c[k1], s2 = (uint16_t)k.c[k2];
    const uint32_t A = s0 | (s1 << 16), B = s1 | (s2 << 16);
    uint32_t P = (((32 - df) * A + df * B + 0x00100010u) >> 5) & 0x07FF07FFu;
    if (k.edge && ii == 0) {
      const int p0 = clip255((int)(P & 0xFFFF) + ((k.c[-k.sg * (jj + 1)] - k.c[0]) >> 1));
      P = (P & 0xFFFF0000u) | (uint32_t)p0;
    }
    return P | 0x64006400u;
  }
  return predict_pair(k.c, k.u, k.n, k.lg, k.mode, X, Yc, k.dc) | 0x64006400u;
}
// Residual pair (original - prediction) as an exact half2; org points at the first pixel in the staged block (row pitch opitch).
__device__ __forceinline__ uint32_t resid_pair(const uint8_t *org, int opitch, bool hor, uint32_t Pm) {
  uint32_t Om;
  if (!hor) Om = __byte_perm((uint32_t)*reinterpret_cast<const uint16_t *>(org), 0x64u, 0x4140);   // 0x64 b1 0x64 b0
  else Om = ((uint32_t)org[0] | ((uint32_t)org[opitch] << 16)) | 0x64006400u;
  const __half2 d = __hsub2(*reinterpret_cast<__half2 *>(&Om), *reinterpret_cast<__half2 *>(&Pm));
  return *reinterpret_cast<const uint32_t *>(&d);
}


// ---- K6 (batched, references taken from the original picture) --------------------------------
// Two launches per frame:
//
//   k_rmd_plan  : one warp per CTU walks the pruned quadtree (TEncCu.cpp:496-520) and writes the PU
//                 descriptors in encoder visiting order plus a queue of WORK ITEMS, big ones first:
//                   64x64 PU -> 4 items (one per 32x32 quadrant), 32x32 PU -> 1 item   [front of the queue]
//                   16x16 PU -> 1 item, 8x8 CU -> 1 item (2Nx2N 8x8 + its four NxN 4x4 PUs)
//                 Offsets come from the per-CTU counts the label kernel wrote (ctu_plan_counts()).
//   k_rmd_items : persistent 4-warp blocks pull items from a global counter (first round static).  The
//                 block stages the PU (or quadrant) and its transpose in shared memory, builds the
//                 reference lines straight from the picture in L2, then its warps split the 35 modes.
//                 A "unit" is one 8x8 block for one mode; a "slab" is two units.  Each lane predicts two
//                 neighbouring pixels per unit with one packed 16-bit interpolation, forms the residual
//                 as an exact fp16 pair, and the 2-D Hadamard transform of both units is two chained
//                 mma.sync (A = diag(H8,H8) or diag(H4 x4), entries +-1): stage 1 gives H*D (|v| <= 2040,
//                 exact in fp16), the accumulator fragment re-read as the next B operand is its
//                 transpose, stage 2 gives (H*D*H^T)^T (|v| <= 16320, exact in fp32).  Sum of magnitudes
//                 and the per-block rounding are those of TComRdCost.cpp:1549-1750.
//                 PUs >= 16 read the main reference array as a table of sample PAIRS (two 32-bit loads and
//                 one packed multiply-add per pixel pair): four block-wide tables (vertical/horizontal x
//                 unfiltered/filtered) serve every mode with a non-negative angle; a negative-angle mode
//                 gets a per-warp table with the projected side samples in front
//                 (TComPrediction.cpp:278-300).  Horizontal modes run the same code on the transposed
//                 block with the roles of the two reference arms swapped (SATD is transpose-invariant).
//                 When a PU's 35 SATDs are complete the block ranks them (ties -> lower mode,
//                 TEncSearch.cpp:5562-5585); for a 64x64 PU the last of its four quadrant items does.
struct RmdItem {
  uint32_t pu;          // index of the PU in the frame's list (kind 1: the 2Nx2N 8x8 PU; its 4x4 PUs follow)
  uint8_t quad;         // 64x64 PUs: 32x32 quadrant handled by this item
  uint8_t kind;         // 0: PU >= 16 (table path), 1: 8x8 CU
  uint16_t frame;       // frame of the launch batch
};
constexpr int MAX_ITEMS_CTU = 64;
#ifndef HEVCDL_RMD_BW
#define HEVCDL_RMD_BW 4
#endif
constexpr int RMD_BW = HEVCDL_RMD_BW;           // warps per block of k_rmd_items (tuning builds: -DHEVCDL_RMD_BW=1|2)
constexpr int ORG_P = 40;                       // row pitch of the staged block: 10 words -> conflict-free 16-bit reads
constexpr int RED_P = 36;                       // 16-byte aligned rows, conflict-free column writes and float4 row reads
constexpr int PTAB_P = 132;

// ctrl[0] = work counter of k_rmd_items (starts at its grid size: the first item of a block is blockIdx.x),
// ctrl[1] = number of items of the frame
__global__ void __launch_bounds__(256)
k_rmd_plan(const RmdBatch rb, FrameGeom geo, int items_grid, RmdItem *__restrict__ items, int *__restrict__ ctrl) {
  const int lane = threadIdx.x & 31, gctu = blockIdx.x * 8 + (threadIdx.x >> 5);
  pdl_launch_dependents();
  TL_BEGIN();
  pdl_wait();
  if (threadIdx.x == 0) { TL_WAITED(); }
  if (blockIdx.x == 0 && threadIdx.x == 0) ctrl[0] = items_grid;
  const int total = geo.nctu * rb.n;
  // PUs of the preceding CTUs of this frame; big / small items of everything before this CTU in (frame, CTU) order,
  // and the batch totals (big items of all frames go first in the queue).  The block's 256 threads share one pass over the
  // counts (round 1: every warp re-added all counts before its CTU, 3.6 M warp instructions per 4-frame launch): sums over
  // the CTUs before the block's first one (A*), the PUs before the first frame this block touches (B), the totals (T*);
  // each warp then adds the <= 7 CTUs of its own block that precede it.
  __shared__ uint32_t s_red[8][8];
  const int gb = blockIdx.x * 8, f0 = gb / geo.nctu, lo0 = f0 * geo.nctu;
  uint32_t Apu = 0, Abig = 0, Asm = 0, Bpu = 0, bigT = 0, smT = 0;
  for (int g = threadIdx.x; g < total; g += 256) {
    const int f = g / geo.nctu;
    const uint32_t v = __ldg(rb.ctu_cnt[f] + (g - f * geo.nctu));
    const uint32_t p = v & 0xFFFFu, bb = (v >> 16) & 15u, ss = v >> 20;
    bigT += bb; smT += ss;
    if (g < gb) { Apu += p; Abig += bb; Asm += ss; }
    if (g < lo0) Bpu += p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Apu += __shfl_xor_sync(0xffffffffu, Apu, o); Abig += __shfl_xor_sync(0xffffffffu, Abig, o); Asm += __shfl_xor_sync(0xffffffffu, Asm, o);
    Bpu += __shfl_xor_sync(0xffffffffu, Bpu, o); bigT += __shfl_xor_sync(0xffffffffu, bigT, o); smT += __shfl_xor_sync(0xffffffffu, smT, o);
  }
  const int wid = threadIdx.x >> 5;
  if (lane == 0) { s_red[wid][0] = Apu; s_red[wid][1] = Abig; s_red[wid][2] = Asm; s_red[wid][3] = Bpu; s_red[wid][4] = bigT; s_red[wid][5] = smT; }
  __syncthreads();
  Apu = Abig = Asm = Bpu = bigT = smT = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) { Apu += s_red[k][0]; Abig += s_red[k][1]; Asm += s_red[k][2]; Bpu += s_red[k][3]; bigT += s_red[k][4]; smT += s_red[k][5]; }
  if (gctu >= total) return;
  const int fr = gctu / geo.nctu, ctu = gctu - fr * geo.nctu;
  uint32_t pu0, big0 = Abig, sm0 = Asm;
  {
    // in-block predecessors: lane i < wid looks at global CTU gb + i
    uint32_t p = 0, bb = 0, ss = 0, pf = 0;        // pf: PUs of in-block predecessors that belong to this warp's frame
    if (lane < wid) {
      const int g = gb + lane, f = g / geo.nctu;
      const uint32_t v = __ldg(rb.ctu_cnt[f] + (g - f * geo.nctu));
      p = v & 0xFFFFu; bb = (v >> 16) & 15u; ss = v >> 20;
      pf = f == fr ? p : 0u;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      bb += __shfl_xor_sync(0xffffffffu, bb, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); pf += __shfl_xor_sync(0xffffffffu, pf, o);
      p += __shfl_xor_sync(0xffffffffu, p, o);
    }
    bb = __shfl_sync(0xffffffffu, bb, 0); ss = __shfl_sync(0xffffffffu, ss, 0); pf = __shfl_sync(0xffffffffu, pf, 0);
    big0 += bb; sm0 += ss;
    // PUs before this CTU inside its frame: the block's first frame starts at lo0 (prefix Bpu), a later frame starts inside the block
    pu0 = fr == f0 ? Apu - Bpu + pf : pf;
  }
  int *__restrict__ ctu_off = rb.ctu_off[fr];
  hevcdl_pu *__restrict__ pus = rb.pus[fr];
  uint32_t *__restrict__ satd = rb.satd[fr];
  uint8_t *__restrict__ cand = rb.cand[fr];
  if (lane == 0) {
    ctu_off[ctu] = (int)pu0;
    if (ctu == geo.nctu - 1) ctu_off[geo.nctu] = (int)(pu0 + (__ldg(rb.ctu_cnt[fr] + ctu) & 0xFFFFu));
    if (gctu == total - 1) ctrl[1] = (int)(bigT + smT);
  }
  uint32_t it_big = big0, it_small = bigT + sm0;
  const int W = geo.W, H = geo.H;
  const int x0 = (ctu % geo.ctu_w) * 64, y0 = (ctu / geo.ctu_w) * 64;
  const uint4 pk = *reinterpret_cast<const uint4 *>(rb.labels[fr] + (size_t)ctu * 16);
  const uint32_t lw[4] = {pk.x, pk.y, pk.z, pk.w};
  const uint16_t f16 = (uint16_t)fr;
  for (int round = 0; round < 2; round++) {
    const int i3 = round * 32 + lane;                                  // z-order index of an 8x8 position
    int esize = 0, ex = 0, ey = 0;
    for (int d = 0; d < 4; d++) {
      const int span = 64 >> (2 * d), o = i3 & ~(span - 1);
      int bx = 0, by = 0;
#pragma unroll
      for (int b = 0; b < 3; b++) { bx |= ((o >> (2 * b)) & 1) << b; by |= ((o >> (2 * b + 1)) & 1) << b; }
      const int x = x0 + bx * 8, y = y0 + by * 8, size = 64 >> d;
      if (x >= W || y >= H) break;                                     // CU outside the picture: skipped (TEncCu.cpp:929-946)
      const int li = 4 * ((y & 63) >> 4) + ((x & 63) >> 4);
      const int pl = (lw[li >> 2] >> (8 * (li & 3))) & 255;
      const bool boundary = (x + size > W) || (y + size > H);          // never evaluated (TEncCu.cpp:574-576)
      if (pl == d && !boundary) {
        if (o == i3) { esize = size; ex = x; ey = y; }
        break;
      }
      if (!(pl > d && d < 3)) break;                                   // pruned
    }
    const int np = esize == 0 ? 0 : (esize == 8 ? 5 : 1);
    const int nb = esize == 64 ? 4 : (esize == 32 ? 1 : 0);
    const int ns = (esize == 16 || esize == 8) ? 1 : 0;
    int sp = np, sb = nb, ss = ns;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, sp, o), b = __shfl_up_sync(0xffffffffu, sb, o), c = __shfl_up_sync(0xffffffffu, ss, o);
      if (lane >= o) { sp += a; sb += b; ss += c; }
    }
    const uint32_t ppos = pu0 + sp - np, bpos = it_big + sb - nb, spos = it_small + ss - ns;
    if (esize) {
      pus[ppos] = hevcdl_pu{(uint16_t)ex, (uint16_t)ey, (uint8_t)esize, 0, (uint16_t)ctu};
      if (esize == 8) {
        for (int k = 0; k < 4; k++)
          pus[ppos + 1 + k] = hevcdl_pu{(uint16_t)(ex + (k & 1) * 4), (uint16_t)(ey + (k >> 1) * 4), 4, (uint8_t)(k + 1), (uint16_t)ctu};
        items[spos] = RmdItem{ppos, 0, 1, f16};
      } else if (esize == 16) {
        items[spos] = RmdItem{ppos, 0, 0, f16};
      } else if (esize == 32) {
        items[bpos] = RmdItem{ppos, 0, 0, f16};
      } else {
        for (int k = 0; k < 4; k++) items[bpos + k] = RmdItem{ppos, (uint8_t)k, 0, f16};
        for (int m = 0; m < 35; m++) satd[(size_t)ppos * 35 + m] = 0;   // quadrant items accumulate with atomicAdd
        *reinterpret_cast<uint32_t *>(cand + (size_t)ppos * 8) = 0;     // ... and count themselves here until the last one ranks
      }
    }
    pu0 += __shfl_sync(0xffffffffu, sp, 31);
    it_big += __shfl_sync(0xffffffffu, sb, 31);
    it_small += __shfl_sync(0xffffffffu, ss, 31);
  }
  TL_END(5);
}

// Test hook support: per-CTU counts for labels that did not come from the CNN kernels (which write them themselves).
__global__ void __launch_bounds__(256)
k_rmd_counts(const uint8_t *__restrict__ labels, FrameGeom geo, uint32_t *__restrict__ ctu_cnt) {
  const int ctu = blockIdx.x * blockDim.x + threadIdx.x;
  if (ctu >= geo.nctu) return;
  uint8_t lab[16];
  *reinterpret_cast<uint4 *>(lab) = *reinterpret_cast<const uint4 *>(labels + (size_t)ctu * 16);
  ctu_cnt[ctu] = ctu_plan_counts(lab, ctu % geo.ctu_w, ctu / geo.ctu_w, geo.W, geo.H);
}

struct RmdWarpS {
  uint32_t tab[200];              // pair table of a negative-angle mode: tab[k + n] = ext[k] | ext[k+1] << 16
  float red[16][RED_P];           // per-lane block partial sums of one reduction group (16 blocks)
};
struct RmdBlockS {
  uint8_t org[32 * ORG_P];        // the PU (or 32x32 quadrant of a 64x64 PU), row pitch ORG_P
  uint8_t orgT[32 * ORG_P];       // its transpose (horizontal modes)
  int16_t line[2][260];           // [0] unfiltered, [1] filtered reference line (8x8 CU items: see block_small)
  uint32_t ptab[4][PTAB_P];       // [hor*2 + filtered][k] = main[k] | main[k+1] << 16, k in [0, 2n)
  uint32_t satd[5][36];           // finished SATDs of the item's PU(s), for the ranking
  int16_t dcs[8];
  int next_item;
  int grp_ctr;                    // next reduction group of the current item (block_large)
  int pad_[2];
  RmdWarpS w[RMD_BW];
};

// Entries [e0, e1) of the reference line of one PU from the ORIGINAL picture, with HM's availability rule and
// substitution scan (TComPattern.cpp:326-543): an unavailable 4-sample unit takes the last sample of the nearest
// available unit before it in scan order, else the first sample of the first available unit after it.  One warp;
// every calling warp evaluates the availability of all <= 65 units itself (ballots), then fills its own entries.
__device__ __noinline__ void build_line_part(const uint8_t *__restrict__ Y, int pitch, int W, int H, int ctu_w, int px, int py, int n,
                                                int16_t *line, int e0, int e1, int lane) {
  const int nu = n >> 2, ntot = 4 * nu + 1;     // units: [0,2nu) left bottom-up, 2nu corner, (2nu, 4nu] above
  unsigned long long mlo = 0;
  bool top = false;
  for (int base = 0; base < ntot; base += 32) {
    const int u = base + lane;
    bool a = false;
    if (u < ntot) {
      int xn, yn;
      if (u < 2 * nu) { xn = px - 1; yn = py + (2 * nu - 1 - u) * 4; }
      else if (u == 2 * nu) { xn = px - 1; yn = py - 1; }
      else { xn = px + (u - 2 * nu - 1) * 4; yn = py - 1; }
      a = unit_available(xn, yn, px, py, W, H, ctu_w);
    }
    const uint32_t b = __ballot_sync(0xffffffffu, a);
    if (base < 64) mlo |= (unsigned long long)b << base;
    else top = b & 1u;
  }
  auto sample = [&](int i) -> int {             // picture sample at line index i (only called for available units)
    int gx, gy;
    if (i < 2 * n) { gx = px - 1; gy = py + 2 * n - 1 - i; }
    else if (i == 2 * n) { gx = px - 1; gy = py - 1; }
    else { gx = px + i - 2 * n - 1; gy = py - 1; }
    return Y[(size_t)gy * pitch + gx];
  };
  for (int i = e0 + lane; i < e1; i += 32) {
    const int u = i < 2 * n ? (i >> 2) : (i == 2 * n ? 2 * nu : 2 * nu + 1 + ((i - 2 * n - 1) >> 2));
    int s;
    {
      const unsigned long long upto = u >= 63 ? ~0ull : ((2ull << u) - 1);
      const unsigned long long below = mlo & upto, above = mlo & ~upto;
      if (u == 64 && top) s = 64;
      else if (below) s = 63 - __clzll((long long)below);
      else if (above) s = __ffsll((long long)above) - 1;
      else s = top ? 64 : -1;
    }
    int v;
    if (s < 0) v = 128;
    else if (s == u) v = sample(i);
    else {
      const int firsti = s < 2 * nu ? 4 * s : (s == 2 * nu ? 2 * n : 2 * n + 1 + 4 * (s - 2 * nu - 1));
      const int lasti = s == 2 * nu ? 2 * n : firsti + 3;
      v = sample(s < u ? lasti : firsti);
    }
    line[i] = (int16_t)v;
  }
}

// DC value of a line (TComPrediction.cpp:183-201); all lanes return it.
__device__ __forceinline__ int line_dc_warp(const int16_t *line, int n, int lane) {
  int sum = 0;
  for (int i = lane; i < n; i += 32) sum += line[2 * n + 1 + i] + line[2 * n - 1 - i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  return (sum + n) / (2 * n);
}

// SATD-ranked candidates of one PU: cand[8], the first 3 (size >= 16) or 8 valid, the rest 255.  Rank by
// (satd, mode): the strict '<' insertion from the worst slot of xUpdateCandList (TEncSearch.cpp:5562-5585)
// keeps the earlier (lower) mode ahead on equal cost.  s0: SATD of mode `lane`; s1: of mode 32 + lane (lane < 3).
__device__ __forceinline__ void rank35_warp(uint32_t s0, uint32_t s1, int keep, uint8_t *__restrict__ cand8, int lane) {
  // key = satd << 6 | mode (satd < 2^24: at most 64 blocks of (64*64*255 + 2) >> 2); repeated warp-wide minimum
  uint32_t k0 = (s0 << 6) | (uint32_t)lane, k1 = lane < 3 ? ((s1 << 6) | (uint32_t)(32 + lane)) : 0xFFFFFFFFu;
  uint32_t lo = 0xFFFFFFFFu, hi = 0xFFFFFFFFu;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    if (k < keep) {                             // warp-uniform
      const uint32_t m = __reduce_min_sync(0xffffffffu, min(k0, k1));
      if (k0 == m) k0 = 0xFFFFFFFFu;
      if (k1 == m) k1 = 0xFFFFFFFFu;
      const uint32_t mode = m & 63u;
      if (k < 4) lo = (lo & ~(0xFFu << (8 * k))) | (mode << (8 * k));
      else hi = (hi & ~(0xFFu << (8 * (k - 4)))) | (mode << (8 * (k - 4)));
    }
  }
  if (lane == 0) *reinterpret_cast<uint2 *>(cand8) = make_uint2(lo, hi);
}

// PUs >= 16: the staged region is the 16x16 PU, the 32x32 PU, or one 32x32 quadrant of a 64x64 PU.  Rolled loops
// on purpose: the hot loop is a few hundred instructions, so the warps of an SM share the instruction cache.
template <int RS>
__device__ __forceinline__ void block_large(RmdBlockS &S, const uint8_t *__restrict__ Y, int pitch, const FrameGeom &geo, const hevcdl_pu pu,
                                            const RmdItem item, uint32_t a8, uint32_t *__restrict__ satd_out, uint8_t *__restrict__ cand_out) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  RmdWarpS &Wp = S.w[wid];
  const int n = pu.size, px = pu.x, py = pu.y, lg = ilog2(n);
  constexpr int rs = RS;                        // region side: 16 (16x16 PU) or 32
  constexpr int lgnb = rs == 32 ? 2 : 1;        // log2(8x8 blocks per region row)
  constexpr int spm = rs == 32 ? 8 : 2;         // slabs per mode; a reduction group is up to 8 slabs = 16 blocks
  constexpr int ngroups = rs == 32 ? 35 : 11;   // 32: one mode per group; 16: eight groups of 4 modes, then 32, 33, 34 alone
  const uint32_t *__restrict__ minfo = c_mode_info.v[n == 16 ? 0 : (n == 32 ? 1 : 2)];
  const int rx0 = n == 64 ? (item.quad & 1) * 32 : 0, ry0 = n == 64 ? (item.quad >> 1) * 32 : 0;
  const int g = lane >> 2, t = lane & 3;
  const bool has_flt = n == 16 || n == 32;
  // ---- phase 1: stage the region and its transpose; every warp fills a quarter of the reference line ----------
  if (tid == 0) S.grp_ctr = RMD_BW;
  {
    constexpr int wpr = rs >> 2, lgw = rs == 32 ? 3 : 2;
    const uint8_t *src = Y + (size_t)(py + ry0) * pitch + px + rx0;
#pragma unroll
    for (int i = tid; i < rs * wpr; i += RMD_BW * 32) {
      const int r = i >> lgw, cw = i & (wpr - 1);
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)r * pitch + 4 * cw));
      *reinterpret_cast<uint32_t *>(&S.org[r * ORG_P + 4 * cw]) = v;
#pragma unroll
      for (int k = 0; k < 4; k++) S.orgT[(4 * cw + k) * ORG_P + r] = (uint8_t)(v >> (8 * k));
    }
    const int len = 4 * n + 1, q = (len + RMD_BW - 1) / RMD_BW;
    build_line_part(Y, pitch, geo.W, geo.H, geo.ctu_w, px, py, n, S.line[0], wid * q, min(len, wid * q + q), lane);
  }
  __syncthreads();
  if (has_flt) {                                // [1 2 1] / strong smoothing (TComPattern.cpp:203-294), whole block
    const int16_t *line = S.line[0];
    const int len = 4 * n + 1;
    const int bl = line[0], tl = line[2 * n], tr = line[4 * n];
    const bool strong = n == 32 && (abs(bl + tl - 2 * line[n]) < 8) && (abs(tl + tr - 2 * line[3 * n]) < 8);
    for (int i = tid; i < len; i += RMD_BW * 32) {
      int v;
      if (i == 0 || i == len - 1) v = line[i];
      else if (!strong) v = (line[i - 1] + 2 * line[i] + line[i + 1] + 2) >> 2;
      else if (i < 2 * n) v = ((2 * n - i) * bl + i * tl + n) >> 6;
      else if (i == 2 * n) v = tl;
      else v = ((4 * n - i) * tl + (i - 2 * n) * tr + n) >> 6;
      S.line[1][i] = (int16_t)v;
    }
  }
  const int dc = line_dc_warp(S.line[0], n, lane);
  __syncthreads();
  // pair tables of the non-negative angles: main[k] | main[k+1] << 16 for the four (orientation, filter) variants
  for (int idx = tid; idx < 8 * n; idx += RMD_BW * 32) {
    const int v = idx / (2 * n), k = idx - v * 2 * n;
    if ((v & 1) && !has_flt) continue;
    const int16_t *c = S.line[v & 1] + 2 * n;
    const int sg = (v & 2) ? -1 : 1;
    S.ptab[v][k] = (uint32_t)c[sg * k] | ((uint32_t)c[sg * (k + 1)] << 16);
  }
  __syncthreads();

  // ---- phase 2: the warps pull reduction groups (1 or 4 modes) from a block-wide counter ------------------------
  int q = wid;                                  // first group static (the counter starts at RMD_BW), then dynamic
  while (q < ngroups) {
    int nq = 0;
    if (lane == 0) nq = atomicAdd(&S.grp_ctr, 1);   // claim the next group now: the round trip hides behind this one
    const int mfirst = rs == 32 ? q : (q < 8 ? 4 * q : 24 + q);
    const int nm = rs == 32 ? 1 : (q < 8 ? 4 : 1);
    for (int mi = 0; mi < nm; mi++) {
      const int mode = mfirst + mi;
      const uint32_t info = minfo[mode];
      const bool flt = info & 1u, hor = info & 2u;
      const int16_t *c = S.line[flt ? 1 : 0] + 2 * n;
      const int angle = (int)(int8_t)(info >> 8), sg = hor ? -1 : 1;
      const uint8_t *o = (hor ? S.orgT : S.org) + g * ORG_P + 2 * t;
      const int i0 = (hor ? ry0 : rx0) + 2 * t, j0 = (hor ? rx0 : ry0) + g;
      const bool slow = info & 4u;
      const uint32_t *tbase = S.ptab[((info >> 1) & 1u) * 2 + (info & 1u)];
      if (info & 8u) {                          // negative angle: projected side samples in front of the main array
        const int inv = (int)(info >> 16);
        const int kmin = ((n * angle) >> 5) + 1;
        __syncwarp();                           // readers of the previous table are done
        for (int k = kmin + lane; k <= n; k += 32) {
          uint32_t e;
          if (k >= 0) e = tbase[k];
          else {
            const int k1 = k + 1;
            const int e0 = c[-sg * ((128 - k * inv) >> 8)];
            const int e1 = k1 < 0 ? c[-sg * ((128 - k1 * inv) >> 8)] : c[0];
            e = (uint32_t)e0 | ((uint32_t)e1 << 16);
          }
          Wp.tab[k + n] = e;
        }
        __syncwarp();
        tbase = Wp.tab + n;
      }
      // slabs: two horizontally adjacent 8x8 blocks each, row of blocks by row of blocks.  The interpolation position
      // depends on the row only; the common angular case gets its own loop so that nothing is decided per slab.
      auto finish = [&](int slab, uint32_t oa, uint32_t ob, uint32_t Pa, uint32_t Pb) {
        // residuals as exact half2 (fp16 1024+v on both sides)
        const uint32_t Oa = __byte_perm(oa, 0x64u, 0x4140), Ob = __byte_perm(ob, 0x64u, 0x4140);
        Pa |= 0x64006400u; Pb |= 0x64006400u;
        const __half2 da = __hsub2(*reinterpret_cast<const __half2 *>(&Oa), *reinterpret_cast<const __half2 *>(&Pa));
        const __half2 db = __hsub2(*reinterpret_cast<const __half2 *>(&Ob), *reinterpret_cast<const __half2 *>(&Pb));
        float c1[4], c2[4];
        mma_f16_16816(c1, a8, 0u, 0u, a8, *reinterpret_cast<const uint32_t *>(&da), *reinterpret_cast<const uint32_t *>(&db));
        mma_f16_16816(c2, a8, 0u, 0u, a8, pack_h2(c1[0], c1[1]), pack_h2(c1[2], c1[3]));
        float *rr = &Wp.red[slab * 2][lane];
        rr[0] = fabsf(c2[0]) + fabsf(c2[1]);
        rr[RED_P] = fabsf(c2[2]) + fabsf(c2[3]);
      };
      if (!slow) {
        // two slabs at a time (32: one row of blocks, 16: both rows): all shared-memory loads of the pair are issued
        // before the first slab's stores, so the two dependent chains (loads, interpolation, two MMAs) overlap
        const uint32_t *trow = tbase + (i0 + 1);
#pragma unroll 1
        for (int pr = 0; pr < spm / 2; pr++) {
          uint32_t oa[2], ob[2], ta[2][2], tb[2][2], dfv[2];
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int by = rs == 32 ? pr : h, bx = rs == 32 ? h : 0;
            const int pos = (j0 + 1 + 8 * by) * angle, di = pos >> 5;
            dfv[h] = pos & 31;
            const uint32_t *tp = trow + di + 16 * bx;
            const uint8_t *op = o + by * 8 * ORG_P + 16 * bx;
            oa[h] = *reinterpret_cast<const uint16_t *>(op); ob[h] = *reinterpret_cast<const uint16_t *>(op + 8);
            ta[h][0] = tp[0]; ta[h][1] = tp[1]; tb[h][0] = tp[8]; tb[h][1] = tp[9];
          }
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const uint32_t df = dfv[h], w0 = 32 - df;
            const uint32_t Pa = ((w0 * ta[h][0] + df * ta[h][1] + 0x00100010u) >> 5) & 0x07FF07FFu;
            const uint32_t Pb = ((w0 * tb[h][0] + df * tb[h][1] + 0x00100010u) >> 5) & 0x07FF07FFu;
            finish(mi * spm + 2 * pr + h, oa[h], ob[h], Pa, Pb);
          }
        }
      } else {
#pragma unroll 1
        for (int sl = 0; sl < spm; sl++) {
          const int by8 = (sl >> (lgnb - 1)) * 8, bx8 = ((2 * sl) & ((1 << lgnb) - 1)) * 8;
          const uint8_t *op = o + by8 * ORG_P + bx8;
          const uint32_t oa = *reinterpret_cast<const uint16_t *>(op), ob = *reinterpret_cast<const uint16_t *>(op + 8);
          uint32_t Pa, Pb;
          if (mode < 2) {
            Pa = predict_pair(c, S.line[0] + 2 * n, n, lg, mode, (rx0 + 2 * t) + bx8, (ry0 + g) + by8, dc);
            Pb = predict_pair(c, S.line[0] + 2 * n, n, lg, mode, (rx0 + 2 * t) + bx8 + 8, (ry0 + g) + by8, dc);
          } else {                              // pure H/V (angle 0, no interpolation) with the edge filter (n <= 16)
            const int i = i0 + bx8, j = j0 + by8;
            Pa = tbase[i + 1]; Pb = tbase[i + 9];
            if (i == 0) {
              const int p0 = clip255((int)(Pa & 0xFFFF) + ((c[-sg * (j + 1)] - c[0]) >> 1));
              Pa = (Pa & 0xFFFF0000u) | (uint32_t)p0;
            }
          }
          finish(mi * spm + sl, oa, ob, Pa, Pb);
        }
      }
    }
    __syncwarp();
    // block totals: lane l sums half of row (l & 15), the two halves meet through one shuffle
    {
      const float4 *row = reinterpret_cast<const float4 *>(&Wp.red[lane & 15][(lane >> 4) * 16]);
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; k++) { const float4 r4 = row[k]; acc += (r4.x + r4.y) + (r4.z + r4.w); }
      acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      uint32_t v = ((uint32_t)acc + 2) >> 2;    // per-block rounding (TComRdCost.cpp:1739-1749)
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (rs == 32) {
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        if (lane == 0) {
          if (n == 64) atomicAdd(&satd_out[(size_t)item.pu * 35 + mfirst], v);
          else { satd_out[(size_t)item.pu * 35 + mfirst] = v; S.satd[0][mfirst] = v; }
        }
      } else {
        const int mode = mfirst + (lane >> 2);  // lanes 4m..4m+3 hold the four blocks of the group's m-th mode
        if ((lane >> 2) < nm && (lane & 3) == 0) { satd_out[(size_t)item.pu * 35 + mode] = v; S.satd[0][mode] = v; }
      }
    }
    __syncwarp();
    q = __shfl_sync(0xffffffffu, nq, 0);
  }
  // ---- phase 3: rank ------------------------------------------------------------------------------------------
  if (n == 64) __threadfence();                 // this thread's atomicAdds are visible before the block counts itself
  __syncthreads();
  if (wid == 0) {
    if (n != 64) {
      rank35_warp(S.satd[0][lane], lane < 3 ? S.satd[0][32 + lane] : 0xFFFFFFFFu, 3, cand_out + (size_t)item.pu * 8, lane);
    } else {
      uint32_t done = 0;
      if (lane == 0) { __threadfence(); done = atomicAdd(reinterpret_cast<uint32_t *>(cand_out + (size_t)item.pu * 8), 1u); }
      done = __shfl_sync(0xffffffffu, done, 0);
      if (done == 3) {                          // the other three quadrants are complete: rank the totals
        __threadfence();
        const uint32_t *sp = satd_out + (size_t)item.pu * 35;
        rank35_warp(__ldcg(sp + lane), lane < 3 ? __ldcg(sp + 32 + lane) : 0xFFFFFFFFu, 3, cand_out + (size_t)item.pu * 8, lane);
      }
    }
  }
}

// One 8x8 CU: the 2Nx2N 8x8 PU and the four 4x4 PUs of its NxN trial (TEncCu.cpp:819-826).  A slab is one 8x8
// area for two modes; in the NxN slabs every 4x4 quadrant is predicted from its own PU's references and
// transformed with diag(H4 x4).  The 36 slabs are dealt round-robin to the block's warps.
__device__ __forceinline__ void block_small(RmdBlockS &S, const uint8_t *__restrict__ Y, int pitch, const FrameGeom &geo, const hevcdl_pu pu,
                                            const RmdItem item, uint32_t a8, uint32_t a4, uint32_t *__restrict__ satd_out,
                                            uint8_t *__restrict__ cand_out) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int px = pu.x, py = pu.y;
  const int g = lane >> 2, t = lane & 3;
  // lines: 8x8 at L[0..33), its filtered copy at [34..67) (make_modek's convention), 4x4 PU k at [68 + 20k ..)
  int16_t *L = S.line[0];
  if (wid == 0 && lane < 16) {
    const int r = lane >> 1, cw = lane & 1;
    *reinterpret_cast<uint32_t *>(&S.org[r * ORG_P + 4 * cw]) =
        __ldg(reinterpret_cast<const uint32_t *>(Y + (size_t)(py + r) * pitch + px + 4 * cw));
  }
  if constexpr (RMD_BW == 4) {
    int16_t *l4 = L + 68 + 20 * wid;            // warp k: the line of 4x4 PU k, then a quarter of the 8x8 line
    build_line_part(Y, pitch, geo.W, geo.H, geo.ctu_w, px + (wid & 1) * 4, py + (wid >> 1) * 4, 4, l4, 0, 17, lane);
    build_line_part(Y, pitch, geo.W, geo.H, geo.ctu_w, px, py, 8, L, wid * 9, min(33, wid * 9 + 9), lane);
    __syncwarp();
    const int dc = line_dc_warp(l4, 4, lane);
    if (lane == 0) S.dcs[1 + wid] = (int16_t)dc;
  } else {                                      // tuning builds with fewer warps per block
    for (int k = wid; k < 4; k += RMD_BW) {
      int16_t *l4 = L + 68 + 20 * k;
      build_line_part(Y, pitch, geo.W, geo.H, geo.ctu_w, px + (k & 1) * 4, py + (k >> 1) * 4, 4, l4, 0, 17, lane);
      __syncwarp();
      const int dc = line_dc_warp(l4, 4, lane);
      if (lane == 0) S.dcs[1 + k] = (int16_t)dc;
    }
    constexpr int q8 = (33 + RMD_BW - 1) / RMD_BW;
    build_line_part(Y, pitch, geo.W, geo.H, geo.ctu_w, px, py, 8, L, wid * q8, min(33, wid * q8 + q8), lane);
  }
  __syncthreads();
  if (wid == 0) {
    filter_line_warp(L, L + 34, 8, lane);
    const int dc = line_dc_warp(L, 8, lane);
    if (lane == 0) S.dcs[0] = (int16_t)dc;
  }
  __syncthreads();
  const size_t p = item.pu;
  for (int q = wid; q < 36; q += RMD_BW) {
    const int r = q >> 1;
    const bool small = q & 1;
    uint32_t bf[2] = {0u, 0u};
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int mode = 2 * r + e;
      if (mode >= 35) continue;                 // warp-uniform
      const bool hor = mode >= 2 && mode < 18;
      const int ux = hor ? g : 2 * t, uy = hor ? 2 * t : g;
      const int sub = (uy >> 2) * 2 + (ux >> 2);
      const ModeK k = small ? make_modek(L + 68 + 20 * sub, 4, mode, S.dcs[1 + sub]) : make_modek(L, 8, mode, S.dcs[0]);
      const int X = small ? ux & 3 : ux, Yc = small ? uy & 3 : uy;
      bf[e] = resid_pair(&S.org[uy * ORG_P + ux], ORG_P, hor, predict_pair_k(k, X, Yc));
    }
    const uint32_t a = small ? a4 : a8;
    float c1[4], c2[4];
    mma_f16_16816(c1, a, 0u, 0u, a, bf[0], bf[1]);
    mma_f16_16816(c2, a, 0u, 0u, a, pack_h2(c1[0], c1[1]), pack_h2(c1[2], c1[3]));
    float sA = fabsf(c2[0]) + fabsf(c2[1]), sB = fabsf(c2[2]) + fabsf(c2[3]);
    if (!small) {
      const bool upper = lane & 16;
      float v = (upper ? sB : sA) + __shfl_xor_sync(0xffffffffu, upper ? sA : sB, 16);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const int mode = 2 * r + (upper ? 1 : 0);
      if ((lane & 15) == 0 && mode < 35) {
        const uint32_t sv = ((uint32_t)v + 2) >> 2;
        satd_out[p * 35 + mode] = sv; S.satd[0][mode] = sv;
      }
    } else {
      // 4x4 blocks: lanes sharing (g>>2, t>>1) hold one block of each unit: reduce over lane bits 0, 2, 3
#pragma unroll
      for (int o = 1; o <= 8; o <<= 1) {
        if (o == 2) continue;
        sA += __shfl_xor_sync(0xffffffffu, sA, o);
        sB += __shfl_xor_sync(0xffffffffu, sB, o);
      }
      if ((lane & 13) == 0) {                   // lanes 0, 2, 16, 18
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int mode = 2 * r + e;
          if (mode >= 35) continue;
          // the result is transposed: its row block (g>>2) follows the slab's n index, its column block (t>>1) the k index
          const bool hor = mode >= 2 && mode < 18;
          const int sub = hor ? (t >> 1) * 2 + (g >> 2) : (g >> 2) * 2 + (t >> 1);
          const uint32_t sv = ((uint32_t)(e ? sB : sA) + 1) >> 1;   // TComRdCost.cpp:1636-1640
          satd_out[(p + 1 + sub) * 35 + mode] = sv; S.satd[1 + sub][mode] = sv;
        }
      }
    }
  }
  __syncthreads();
  for (int k = wid; k < 5; k += RMD_BW)
    rank35_warp(S.satd[k][lane], lane < 3 ? S.satd[k][32 + lane] : 0xFFFFFFFFu, 8, cand_out + (p + k) * 8, lane);
}

#ifndef HEVCDL_RMD_MINB
#define HEVCDL_RMD_MINB (HEVCDL_RMD_BW == 4 ? 9 : 32 / HEVCDL_RMD_BW)   // 9 blocks of 4 warps (54 registers, no spill): 64.9 vs 65.6 us per 1080p frame with 8; 10 / 11 blocks spill and are slower (69.2 / 72.2 us), profiles/r02j_k6_blocks_per_sm.log
#endif
__global__ void __launch_bounds__(RMD_BW * 32, HEVCDL_RMD_MINB)
k_rmd_items(const RmdBatch rb, FrameGeom geo, int pitch, const RmdItem *__restrict__ items, int *__restrict__ ctrl) {
  __shared__ __align__(16) RmdBlockS S;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // A fragments (m16n8k16 row-major A): only the diagonal 8x8 blocks are non-zero
  uint32_t a8, a4;
  {
    const uint32_t one = 0x3C00u, neg = 0xBC00u;
    const int c0 = 2 * t, c1 = 2 * t + 1;
    a8 = ((__popc(g & c0) & 1) ? neg : one) | (((__popc(g & c1) & 1) ? neg : one) << 16);
    const bool on = (g >> 2) == (t >> 1);
    a4 = on ? (((__popc((g & 3) & (c0 & 3)) & 1) ? neg : one) | (((__popc((g & 3) & (c1 & 3)) & 1) ? neg : one) << 16)) : 0u;
  }
  pdl_launch_dependents();
  TL_BEGIN();
  pdl_wait();
  if (threadIdx.x == 0) { TL_WAITED(); }
  const int nitems = ctrl[1];
  int it = blockIdx.x;                          // first round is static: k_rmd_plan started the counter at gridDim.x
  while (it < nitems) {
    int nxt = 0;
    if (threadIdx.x == 0) nxt = atomicAdd(&ctrl[0], 1);   // claim the next item now: its latency hides behind this one
    const RmdItem item = items[it];
    const int f = item.frame;
    const uint8_t *__restrict__ Y = rb.Y[f];
    uint32_t *__restrict__ satd_out = rb.satd[f];
    uint8_t *__restrict__ cand_out = rb.cand[f];
    const hevcdl_pu pu = rb.pus[f][item.pu];
    if (item.kind == 1) block_small(S, Y, pitch, geo, pu, item, a8, a4, satd_out, cand_out);
    else if (pu.size == 16) block_large<16>(S, Y, pitch, geo, pu, item, a8, satd_out, cand_out);
    else block_large<32>(S, Y, pitch, geo, pu, item, a8, satd_out, cand_out);
    if (threadIdx.x == 0) S.next_item = nxt;
    __syncthreads();                            // also: every warp is done with the item's shared memory
    it = S.next_item;
  }
  TL_END(6);
}

// ---- exact mode: explicit original blocks, reference lines and mode bits ----------------------
// One CTA per PU.  org_off/line_off: exclusive prefix offsets computed on the host.
__global__ void __launch_bounds__(128, 4)
k_rmd_exact(int npu, const uint8_t *__restrict__ sizes, const uint8_t *__restrict__ org, const int *__restrict__ org_off,
            const int16_t *__restrict__ lines, const int *__restrict__ line_off, const uint32_t *__restrict__ bits,
            const int8_t *__restrict__ mpm, const uint8_t *__restrict__ mpm_add, double sqrt_lambda,
            uint32_t *__restrict__ satd_out, uint8_t *__restrict__ cand_out, uint8_t *__restrict__ ncand_out) {
  __shared__ __align__(16) uint8_t s_org[64 * 64];
  __shared__ int16_t s_line[260], s_filt[260];
  __shared__ uint32_t s_satd[35];
  __shared__ int s_dc;
  const int tid = threadIdx.x, lane = tid & 31;
  for (int p = blockIdx.x; p < npu; p += gridDim.x) {
    const int n = sizes[p];
    for (int i = tid; i < n * n; i += blockDim.x) s_org[i] = org[org_off[p] + i];
    for (int i = tid; i < 4 * n + 1; i += blockDim.x) s_line[i] = lines[line_off[p] + i];
    if (tid < 35) s_satd[tid] = 0;
    __syncthreads();
    if (tid < 32) {
      if (n == 8 || n == 16 || n == 32) filter_line_warp(s_line, s_filt, n, lane);
      int sum = 0;
      for (int i = lane; i < n; i += 32) sum += s_line[2 * n + 1 + i] + s_line[2 * n - 1 - i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) s_dc = (sum + n) / (2 * n);
    }
    __syncthreads();
    const int nb = n >= 8 ? n >> 3 : 1, nblk = nb * nb;
    for (int r = tid; r < 35 * nblk; r += blockDim.x) {
      const int mode = r / nblk, blk = r % nblk;
      const int16_t *ref = mode_uses_filter(mode, n) ? s_filt : s_line;
      uint32_t v;
      if (n >= 8) {
        const int bx = (blk % nb) * 8, by = (blk / nb) * 8;
        v = satd_unit<8>(&s_org[by * n + bx], n, ref, s_line, n, ilog2(n), mode, bx, by, s_dc);
      } else {
        v = satd_unit<4>(s_org, 4, ref, s_line, 4, 2, mode, 0, 0, s_dc);
      }
      atomicAdd(&s_satd[mode], v);
    }
    __syncthreads();
    if (tid < 35 && satd_out) satd_out[(size_t)p * 35 + tid] = s_satd[tid];
    if (tid == 0 && cand_out) {
      uint8_t modes[10];
      for (int i = 0; i < 10; i++) modes[i] = 255;
      const int len = cand_list(s_satd, bits ? bits + (size_t)p * 35 : nullptr, sqrt_lambda, n,
                                mpm ? mpm + (size_t)p * 3 : nullptr, mpm_add ? mpm_add[p] : 0, modes);
      for (int i = 0; i < 10; i++) cand_out[(size_t)p * 10 + i] = modes[i];
      if (ncand_out) ncand_out[p] = (uint8_t)len;
    }
    __syncthreads();
  }
}

}  // namespace hevcdl
