// rmd.cuh -- K6: label-driven PU enumeration and the 35-mode intra SATD ("RMD") pass.
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

constexpr int RMD_THREADS = 256;
constexpr int MAX_PU_CTU = 320;                 // 64 8x8 CUs x (1 + 4 NxN PUs)
constexpr int TILE_P = 144;                     // staged luma pitch: x0-16 .. x0+127
constexpr int TILE_H = 65;                      // y0-1 .. y0+63
constexpr int LINE_POOL = 2 * (64 * 33 + 256 * 17);  // worst case: unfiltered + filtered lines of a CTU
constexpr int MAX_SLABS = 64 * 18 * 2;          // 64 8x8 CUs x (2Nx2N + NxN) x 36 units / 2

__device__ __forceinline__ int zidx4(int ux, int uy) {   // z-order of a 4x4 unit in a CTU (TComRom.cpp:284)
  int z = 0;
#pragma unroll
  for (int b = 0; b < 4; b++) z |= (((ux >> b) & 1) << (2 * b)) | (((uy >> b) & 1) << (2 * b + 1));
  return z;
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

// ---- PU enumeration ------------------------------------------------------------------------
// Visits the pruned quadtree of one CTU exactly as TEncCu::xCompressCU does (HM
// TLibEncoder/TEncCu.cpp:496-520: evaluate a CU only where label == depth, descend only where
// label > depth; :574-576 CUs crossing the picture edge are never evaluated; :929-946 children
// starting outside the picture are skipped; :819-826 8x8 CUs also try NxN = four 4x4 PUs).
// emit == nullptr: count only.
__device__ inline int enum_ctu_pus(const uint8_t *label, int ctu, int ctu_x, int ctu_y, int W, int H, hevcdl_pu *emit) {
  int cnt = 0;
  // explicit z-order walk: depth-3 index i3 in 0..63 enumerates 8x8 blocks in z-order
  for (int i3 = 0; i3 < 64;) {
    // position of this 8x8 block
    int bx = 0, by = 0;
#pragma unroll
    for (int b = 0; b < 3; b++) { bx |= ((i3 >> (2 * b)) & 1) << b; by |= ((i3 >> (2 * b + 1)) & 1) << b; }
    const int x = ctu_x * 64 + bx * 8, y = ctu_y * 64 + by * 8;
    // largest aligned CU starting here: depth d is possible iff i3 % (64 >> 2d) == 0
    int step = 1;
    bool done = false;
    for (int d = 0; d < 4 && !done; d++) {
      const int span = 64 >> (2 * d);             // number of 8x8 blocks covered by a depth-d CU
      if (i3 % span) continue;
      const int size = 64 >> d;
      if (x >= W || y >= H) { step = span; done = true; break; }   // whole CU outside: skipped
      const int p = label[4 * ((y & 63) >> 4) + ((x & 63) >> 4)];
      const bool boundary = (x + size > W) || (y + size > H);
      if (p == d && !boundary) {
        if (emit) emit[cnt] = hevcdl_pu{(uint16_t)x, (uint16_t)y, (uint8_t)size, 0, (uint16_t)ctu};
        cnt++;
        if (d == 3) {
          for (int k = 0; k < 4; k++) {
            if (emit) emit[cnt] = hevcdl_pu{(uint16_t)(x + (k & 1) * 4), (uint16_t)(y + (k >> 1) * 4), 4, (uint8_t)(k + 1), (uint16_t)ctu};
            cnt++;
          }
        }
        step = span; done = true;
      } else if (p > d && d < 3) {
        continue;                                  // descend: try the next depth at the same origin
      } else {
        step = span; done = true;                  // pruned: nothing evaluated inside this CU
      }
    }
    i3 += step;
  }
  return cnt;
}

// per-CTU PU counts -> exclusive offsets (one block; nctu <= 8160 at 8K).  The descriptors themselves are
// written by k_rmd_batched, which enumerates its CTU again with one thread per 8x8 position.
__global__ void __launch_bounds__(1024, 1)
k_enum_pus(const uint8_t *__restrict__ labels, FrameGeom geo, int *__restrict__ ctu_off /* nctu+1 */) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < geo.nctu; base += blockDim.x) {
    const int ctu = base + threadIdx.x;
    int cnt = 0;
    if (ctu < geo.nctu) {
      uint8_t lab[16];
      *reinterpret_cast<uint4 *>(lab) = *reinterpret_cast<const uint4 *>(labels + (size_t)ctu * 16);
      cnt = enum_ctu_pus(lab, ctu, ctu % geo.ctu_w, ctu / geo.ctu_w, geo.W, geo.H, nullptr);
    }
    int v = cnt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int carry = carry_s;
    if (ctu < geo.nctu) ctu_off[ctu] = carry + (warp ? warp_tot[warp - 1] : 0) + v - cnt;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) ctu_off[geo.nctu] = carry_s;
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

// ---- K6 (batched, references taken from the staged picture itself) ---------------------------
// Work decomposition: a "unit" is one 8x8 block of one PU for one mode (for the four 4x4 PUs of an
// 8x8 CU: the CU's 8x8 area for one mode, each quadrant predicted from its own PU's references);
// a warp processes "slabs" of two units.  Each lane predicts two neighbouring pixels per unit
// with one packed 16-bit interpolation, forms the residual as an exact fp16 pair, and the 2-D
// Hadamard transform of both units is two chained mma.sync (A = diag(H8,H8) or diag(H4 x4), entries
// +-1): stage 1 gives H*D (|v| <= 2040, exact in fp16), the accumulator fragment re-read as the
// next B operand is its transpose, stage 2 gives (H*D*H^T)^T (|v| <= 16320, exact in fp32).
// Sum of magnitudes and the per-block rounding are those of TComRdCost.cpp:1549-1750.
struct RmdSmem {
  uint8_t tile[TILE_H * TILE_P];                // luma rows y0-1..y0+63, cols x0-16..x0+127
  int16_t lines[LINE_POOL + 8];
  uint32_t satd[MAX_PU_CTU * 35];
  hevcdl_pu pu[MAX_PU_CTU];
  int line_off[MAX_PU_CTU + 1];
  int unit_off[MAX_PU_CTU + 1];
  int16_t dc[MAX_PU_CTU];
  uint8_t avail[RMD_THREADS / 32][68];
  int8_t src[RMD_THREADS / 32][68];
  uint16_t slab_pu[MAX_SLABS];                  // PU owning each slab (slabs never straddle PUs: unit counts are padded to even)
  int npu_w0;
};

__device__ __forceinline__ void mma_f16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                              uint32_t b1) {
  asm volatile(
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
// Residual pair (original - prediction) as an exact half2; org points at the first pixel in the staged tile.
__device__ __forceinline__ uint32_t resid_pair(const uint8_t *org, bool hor, uint32_t Pm) {
  uint32_t Om;
  if (!hor) Om = __byte_perm((uint32_t)*reinterpret_cast<const uint16_t *>(org), 0x64u, 0x4140);   // 0x64 b1 0x64 b0
  else Om = ((uint32_t)org[0] | ((uint32_t)org[TILE_P] << 16)) | 0x64006400u;
  const __half2 d = __hsub2(*reinterpret_cast<__half2 *>(&Om), *reinterpret_cast<__half2 *>(&Pm));
  return *reinterpret_cast<const uint32_t *>(&d);
}

__global__ void __launch_bounds__(RMD_THREADS, 2)
k_rmd_batched(const uint8_t *__restrict__ Y, FrameGeom geo, int pitch, const uint8_t *__restrict__ labels,
              const int *__restrict__ ctu_off, hevcdl_pu *__restrict__ pus_out, uint32_t *__restrict__ satd_out,
              uint8_t *__restrict__ cand_out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  RmdSmem &S = *reinterpret_cast<RmdSmem *>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = geo.W, H = geo.H;
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

  for (int ctu = blockIdx.x; ctu < geo.nctu; ctu += gridDim.x) {
    const int first = ctu_off[ctu], npu = ctu_off[ctu + 1] - first;
    if (npu == 0) continue;                     // uniform per block
    const int ctu_x = ctu % geo.ctu_w, ctu_y = ctu / geo.ctu_w;
    const int x0 = ctu_x * 64, y0 = ctu_y * 64;
    // stage luma: 16-byte vectors, zero outside the picture (never read: availability masks it)
    for (int i = tid; i < TILE_H * (TILE_P / 16); i += RMD_THREADS) {
      const int r = i / (TILE_P / 16), cv = i % (TILE_P / 16);
      const int gy = y0 - 1 + r, gx = x0 - 16 + cv * 16;
      uint4 v = make_uint4(0, 0, 0, 0);       // rows are pitch-aligned (128 B): whole vectors stay inside the row
      if (gy >= 0 && gy < H && gx >= 0 && gx < pitch) v = *reinterpret_cast<const uint4 *>(Y + (size_t)gy * pitch + gx);
      *reinterpret_cast<uint4 *>(&S.tile[r * TILE_P + cv * 16]) = v;
    }
    // ---- PU enumeration: one thread per 8x8 position (z-order index), TEncCu.cpp:496-520 ----------
    if (tid < 64) {
      const int i3 = tid;
      const uint4 pk = *reinterpret_cast<const uint4 *>(labels + (size_t)ctu * 16);
      const uint32_t lw[4] = {pk.x, pk.y, pk.z, pk.w};
      int emit = 0, esize = 0, ex = 0, ey = 0;
      for (int d = 0; d < 4; d++) {
        const int span = 64 >> (2 * d), o = i3 & ~(span - 1);
        int bx = 0, by = 0;
#pragma unroll
        for (int b = 0; b < 3; b++) { bx |= ((o >> (2 * b)) & 1) << b; by |= ((o >> (2 * b + 1)) & 1) << b; }
        const int x = x0 + bx * 8, y = y0 + by * 8, size = 64 >> d;
        if (x >= W || y >= H) break;                                   // CU outside the picture: skipped
        const int li = 4 * ((y & 63) >> 4) + ((x & 63) >> 4);
        const int pl = (lw[li >> 2] >> (8 * (li & 3))) & 255;
        const bool boundary = (x + size > W) || (y + size > H);
        if (pl == d && !boundary) {
          if (o == i3) { emit = d == 3 ? 5 : 1; esize = size; ex = x; ey = y; }
          break;
        }
        if (!(pl > d && d < 3)) break;                                 // pruned
      }
      const uint32_t m1 = __ballot_sync(0xffffffffu, emit >= 1), m5 = __ballot_sync(0xffffffffu, emit == 5);
      const uint32_t lt = (1u << lane) - 1;
      int pos = __popc(m1 & lt) + 4 * __popc(m5 & lt);
      if (tid == 31) S.npu_w0 = pos + emit;
      asm volatile("bar.sync 2, 64;" ::: "memory");
      if (warp == 1) pos += S.npu_w0;
      if (emit) {
        S.pu[pos] = hevcdl_pu{(uint16_t)ex, (uint16_t)ey, (uint8_t)esize, 0, (uint16_t)ctu};
        if (emit == 5)
          for (int k = 0; k < 4; k++)
            S.pu[pos + 1 + k] = hevcdl_pu{(uint16_t)(ex + (k & 1) * 4), (uint16_t)(ey + (k >> 1) * 4), 4, (uint8_t)(k + 1), (uint16_t)ctu};
      }
    }
    __syncthreads();
    if (warp == 0) {                            // offsets of each PU's lines and SATD units (warp scan)
      int lo_run = 0, uo_run = 0;
      for (int base = 0; base < npu; base += 32) {
        const int i = base + lane;
        int lo = 0, uo = 0;
        if (i < npu) {
          const int n = S.pu[i].size;
          lo = (4 * n + 1 + 1) & ~1;
          if (n == 8 || n == 16 || n == 32) lo *= 2;
          uo = n >= 16 ? 35 : ((n == 8 || S.pu[i].part == 1) ? 18 : 0);   // work items: (PU, mode) or (PU, mode pair)
        }
        int li = lo, ui = uo;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int a = __shfl_up_sync(0xffffffffu, li, o), b = __shfl_up_sync(0xffffffffu, ui, o);
          if (lane >= o) { li += a; ui += b; }
        }
        if (i < npu) { S.line_off[i] = lo_run + li - lo; S.unit_off[i] = uo_run + ui - uo; }
        lo_run += __shfl_sync(0xffffffffu, li, 31);
        uo_run += __shfl_sync(0xffffffffu, ui, 31);
      }
      if (lane == 0) { S.line_off[npu] = lo_run; S.unit_off[npu] = uo_run; }
    }
    __syncthreads();
    for (int p = warp; p < npu; p += RMD_THREADS / 32)
      for (int sl = S.unit_off[p] + lane; sl < S.unit_off[p + 1]; sl += 32) S.slab_pu[sl] = (uint16_t)p;
    auto pix = [&](int gx, int gy) -> int { return S.tile[(gy - (y0 - 1)) * TILE_P + gx - (x0 - 16)]; };

    // reference lines: one warp per PU (HM TComPattern.cpp:326-543)
    for (int p = warp; p < npu; p += RMD_THREADS / 32) {
      const int n = S.pu[p].size, px = S.pu[p].x, py = S.pu[p].y;
      const int nu = n >> 2;                    // units: [0,2nu) left bottom-up, 2nu corner, (2nu, 4nu] above
      int16_t *line = S.lines + S.line_off[p];
      for (int u = lane; u <= 4 * nu; u += 32) {
        int xn, yn;
        if (u < 2 * nu) { xn = px - 1; yn = py + (2 * nu - 1 - u) * 4; }
        else if (u == 2 * nu) { xn = px - 1; yn = py - 1; }
        else { xn = px + (u - 2 * nu - 1) * 4; yn = py - 1; }
        S.avail[warp][u] = unit_available(xn, yn, px, py, W, H, geo.ctu_w);
      }
      __syncwarp();
      for (int u = lane; u <= 4 * nu; u += 32) {
        int s = -1;
        for (int v = u; v >= 0; v--) if (S.avail[warp][v]) { s = v; break; }
        if (s < 0) for (int v = u + 1; v <= 4 * nu; v++) if (S.avail[warp][v]) { s = v; break; }
        S.src[warp][u] = (int8_t)s;
      }
      __syncwarp();
      auto sample = [&](int i) -> int {         // picture sample at line index i
        if (i < 2 * n) return pix(px - 1, py + 2 * n - 1 - i);
        if (i == 2 * n) return pix(px - 1, py - 1);
        return pix(px + i - 2 * n - 1, py - 1);
      };
      for (int i = lane; i < 4 * n + 1; i += 32) {
        const int u = i < 2 * n ? (i >> 2) : (i == 2 * n ? 2 * nu : 2 * nu + 1 + ((i - 2 * n - 1) >> 2));
        const int s = S.src[warp][u];
        int v;
        if (s < 0) v = 128;
        else if (s == u) v = sample(i);
        else {
          // last sample (scan order) of an earlier unit, first sample of a later one
          const int firsti = s < 2 * nu ? 4 * s : (s == 2 * nu ? 2 * n : 2 * n + 1 + 4 * (s - 2 * nu - 1));
          const int lasti = s == 2 * nu ? 2 * n : firsti + 3;
          v = sample(s < u ? lasti : firsti);
        }
        line[i] = (int16_t)v;
      }
      __syncwarp();
      if (n == 8 || n == 16 || n == 32) filter_line_warp(line, line + ((4 * n + 2) & ~1), n, lane);
      // DC value (TComPrediction.cpp:183-201)
      int sum = 0;
      for (int i = lane; i < n; i += 32) sum += line[2 * n + 1 + i] + line[2 * n - 1 - i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) S.dc[p] = (int16_t)((sum + n) / (2 * n));
      __syncwarp();
    }
    __syncthreads();

    // ---- SATD: work items (PU, mode) for PUs >= 16, (PU, mode pair) for 8x8 PUs and 4x4 groups ---------
    const int nitems = S.unit_off[npu];
    for (int item = warp; item < nitems; item += RMD_THREADS / 32) {
      const int p = S.slab_pu[item], n = S.pu[p].size;
      const int r = item - S.unit_off[p];
      const uint8_t *porg = &S.tile[(S.pu[p].y - (y0 - 1)) * TILE_P + S.pu[p].x - (x0 - 16)];
      if (n >= 16) {
        // one mode, all 8x8 blocks of the PU, two per slab
        const ModeK k = make_modek(S.lines + S.line_off[p], n, r, S.dc[p]);
        const int ux = k.hor ? g : 2 * t, uy = k.hor ? 2 * t : g;
        const int lgnb = k.lg - 3, nb = 1 << lgnb, nslabs = 1 << (2 * lgnb - 1);
        auto slab = [&](int sl, float &sA, float &sB) {
          uint32_t bf[2];
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int blk = 2 * sl + e;
            const int X = ux + (blk & (nb - 1)) * 8, Yc = uy + (blk >> lgnb) * 8;
            bf[e] = resid_pair(porg + Yc * TILE_P + X, k.hor, predict_pair_k(k, X, Yc));
          }
          float c1[4], c2[4];
          mma_f16_16816(c1, a8, 0u, 0u, a8, bf[0], bf[1]);
          mma_f16_16816(c2, a8, 0u, 0u, a8, pack_h2(c1[0], c1[1]), pack_h2(c1[2], c1[3]));
          sA = fabsf(c2[0]) + fabsf(c2[1]); sB = fabsf(c2[2]) + fabsf(c2[3]);
        };
        uint32_t acc = 0;
        if (n == 16) {
#pragma unroll
          for (int sl = 0; sl < 2; sl++) {
            float sA, sB;
            slab(sl, sA, sB);
            const bool upper = lane & 16;
            float v = (upper ? sB : sA) + __shfl_xor_sync(0xffffffffu, upper ? sA : sB, 16);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc += ((uint32_t)v + 2) >> 2;        // lanes 0-15 hold block A's total, 16-31 block B's
          }
          acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        } else {
          for (int ch = 0; ch < nslabs; ch += 16) {
            float part[32];
#pragma unroll
            for (int sl = 0; sl < 16; sl++) {
              part[2 * sl] = 0.f; part[2 * sl + 1] = 0.f;
              if (ch + sl < nslabs) slab(ch + sl, part[2 * sl], part[2 * sl + 1]);   // warp-uniform
            }
            const float tot = warp_transpose_sum32(part, lane);                     // lane l: total of block l of this chunk
            acc += ((uint32_t)tot + 2) >> 2;                                        // per-block rounding (0 for unused slots)
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        }
        if (lane == 0) S.satd[p * 35 + r] = acc;
      } else {
        // two modes of one 8x8 PU, or of the four 4x4 PUs of one CU (p is part 1, parts 2-4 follow)
        const bool small = n < 8;
        uint32_t bf[2] = {0u, 0u};
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int mode = 2 * r + e;
          if (mode >= 35) continue;               // warp-uniform
          const bool hor = mode >= 2 && mode < 18;
          const int ux = hor ? g : 2 * t, uy = hor ? 2 * t : g;
          const int q = small ? p + (uy >> 2) * 2 + (ux >> 2) : p;
          const ModeK k = make_modek(S.lines + S.line_off[q], small ? 4 : 8, mode, S.dc[q]);
          const int X = small ? ux & 3 : ux, Yc = small ? uy & 3 : uy;
          bf[e] = resid_pair(porg + uy * TILE_P + ux, hor, predict_pair_k(k, X, Yc));
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
          if ((lane & 15) == 0 && mode < 35) S.satd[p * 35 + mode] = ((uint32_t)v + 2) >> 2;
        } else {
          // 4x4 blocks: lanes sharing (g>>2, t>>1) hold one block of each unit: reduce over lane bits 0, 2, 3
#pragma unroll
          for (int o = 1; o <= 8; o <<= 1) {
            if (o == 2) continue;
            sA += __shfl_xor_sync(0xffffffffu, sA, o);
            sB += __shfl_xor_sync(0xffffffffu, sB, o);
          }
          if ((lane & 13) == 0) {                 // lanes 0, 2, 16, 18
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int mode = 2 * r + e;
              if (mode >= 35) continue;
              // the result is transposed: its row block (g>>2) follows the slab's n index, its column block (t>>1) the k index
              const bool hor = mode >= 2 && mode < 18;
              const int sub = hor ? (t >> 1) * 2 + (g >> 2) : (g >> 2) * 2 + (t >> 1);
              S.satd[(p + sub) * 35 + mode] = ((uint32_t)(e ? sB : sA) + 1) >> 1;
            }
          }
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < npu; i += RMD_THREADS) pus_out[first + i] = S.pu[i];
    for (int i = tid; i < npu * 35; i += RMD_THREADS) satd_out[(size_t)first * 35 + i] = S.satd[i];
    for (int p = tid; p < npu; p += RMD_THREADS) {
      uint8_t modes[10];
      for (int i = 0; i < 8; i++) modes[i] = 255;
      cand_list(&S.satd[p * 35], nullptr, 0.0, S.pu[p].size, nullptr, 0, modes);
      uint2 pk;
      pk.x = modes[0] | (modes[1] << 8) | (modes[2] << 16) | ((uint32_t)modes[3] << 24);
      pk.y = modes[4] | (modes[5] << 8) | (modes[6] << 16) | ((uint32_t)modes[7] << 24);
      *reinterpret_cast<uint2 *>(cand_out + (size_t)(first + p) * 8) = pk;
    }
    __syncthreads();
  }
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
