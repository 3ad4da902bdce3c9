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
#include "common.cuh"

namespace hevcdl {

constexpr int RMD_THREADS = 256;
constexpr int MAX_PU_CTU = 320;                 // 64 8x8 CUs x (1 + 4 NxN PUs)
constexpr int TILE_P = 144;                     // staged luma pitch: x0-16 .. x0+127
constexpr int TILE_H = 65;                      // y0-1 .. y0+63
constexpr int LINE_POOL = 2 * (64 * 33 + 256 * 17);  // worst case: unfiltered + filtered lines of a CTU

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

// counts -> exclusive offsets -> descriptors, one launch, one block (nctu <= 8160 at 8K)
__global__ void __launch_bounds__(1024, 1)
k_enum_pus(const uint8_t *__restrict__ labels, FrameGeom geo, int *__restrict__ ctu_off /* nctu+1 */,
           hevcdl_pu *__restrict__ pus) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < geo.nctu; base += blockDim.x) {
    const int ctu = base + threadIdx.x;
    int cnt = 0;
    uint8_t lab[16];
    if (ctu < geo.nctu) {
      const uint4 pk = *reinterpret_cast<const uint4 *>(labels + (size_t)ctu * 16);
      *reinterpret_cast<uint4 *>(lab) = pk;
      cnt = enum_ctu_pus(lab, ctu, ctu % geo.ctu_w, ctu / geo.ctu_w, geo.W, geo.H, nullptr);
    }
    // block exclusive scan
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
    const int excl = carry + (warp ? warp_tot[warp - 1] : 0) + v - cnt;
    if (ctu < geo.nctu) {
      ctu_off[ctu] = excl;
      enum_ctu_pus(lab, ctu, ctu % geo.ctu_w, ctu / geo.ctu_w, geo.W, geo.H, pus + excl);
    }
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
struct RmdSmem {
  uint8_t tile[TILE_H * TILE_P];                // luma rows y0-1..y0+63, cols x0-16..x0+127
  int16_t lines[LINE_POOL];
  uint32_t satd[MAX_PU_CTU * 35];
  hevcdl_pu pu[MAX_PU_CTU];
  int line_off[MAX_PU_CTU + 1];
  int unit_off[MAX_PU_CTU + 1];
  int16_t dc[MAX_PU_CTU];
  uint8_t avail[RMD_THREADS / 32][68];
  int8_t src[RMD_THREADS / 32][68];
};

__global__ void __launch_bounds__(RMD_THREADS, 2)
k_rmd_batched(const uint8_t *__restrict__ Y, FrameGeom geo, int pitch, const int *__restrict__ ctu_off,
              const hevcdl_pu *__restrict__ pus, uint32_t *__restrict__ satd_out, uint8_t *__restrict__ cand_out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  RmdSmem &S = *reinterpret_cast<RmdSmem *>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = geo.W, H = geo.H;

  for (int ctu = blockIdx.x; ctu < geo.nctu; ctu += gridDim.x) {
    const int first = ctu_off[ctu], npu = ctu_off[ctu + 1] - first;
    if (npu == 0) continue;                     // uniform per block
    const int x0 = (ctu % geo.ctu_w) * 64, y0 = (ctu / geo.ctu_w) * 64;
    // stage luma: 16-byte vectors, zero outside the picture (never read: availability masks it)
    for (int i = tid; i < TILE_H * (TILE_P / 16); i += RMD_THREADS) {
      const int r = i / (TILE_P / 16), cv = i % (TILE_P / 16);
      const int gy = y0 - 1 + r, gx = x0 - 16 + cv * 16;
      uint4 v = make_uint4(0, 0, 0, 0);       // rows are pitch-aligned (128 B): whole vectors stay inside the row
      if (gy >= 0 && gy < H && gx >= 0 && gx < pitch) v = *reinterpret_cast<const uint4 *>(Y + (size_t)gy * pitch + gx);
      *reinterpret_cast<uint4 *>(&S.tile[r * TILE_P + cv * 16]) = v;
    }
    for (int i = tid; i < npu; i += RMD_THREADS) S.pu[i] = pus[first + i];
    for (int i = tid; i < npu * 35; i += RMD_THREADS) S.satd[i] = 0;
    __syncthreads();
    if (tid == 0) {                             // offsets of each PU's lines and SATD units
      int lo = 0, uo = 0;
      for (int i = 0; i < npu; i++) {
        const int n = S.pu[i].size;
        S.line_off[i] = lo; S.unit_off[i] = uo;
        lo += (4 * n + 1 + 1) & ~1;             // unfiltered
        if (n == 8 || n == 16 || n == 32) lo += (4 * n + 1 + 1) & ~1;
        uo += 35 * (n >= 8 ? (n >> 3) * (n >> 3) : 1);
      }
      S.line_off[npu] = lo; S.unit_off[npu] = uo;
    }
    __syncthreads();
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

    // SATD units, flattened over (PU, mode, block)
    const int nunits = S.unit_off[npu];
    for (int uidx = tid; uidx < nunits; uidx += RMD_THREADS) {
      int lo = 0, hi = npu - 1;                 // last p with unit_off[p] <= uidx
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (S.unit_off[mid] <= uidx) lo = mid; else hi = mid - 1;
      }
      const int p = lo, n = S.pu[p].size, r = uidx - S.unit_off[p];
      const int nb = n >= 8 ? n >> 3 : 1, nblk = nb * nb;
      const int mode = r / nblk, blk = r % nblk;
      const int16_t *line = S.lines + S.line_off[p];
      const int16_t *ref = mode_uses_filter(mode, n) ? line + ((4 * n + 2) & ~1) : line;
      const int lx = S.pu[p].x - (x0 - 16), ly = S.pu[p].y - (y0 - 1);
      uint32_t v;
      if (n >= 8) {
        const int bx = (blk % nb) * 8, by = (blk / nb) * 8;
        v = satd_unit<8>(&S.tile[(ly + by) * TILE_P + lx + bx], TILE_P, ref, line, n, ilog2(n), mode, bx, by, S.dc[p]);
      } else {
        v = satd_unit<4>(&S.tile[ly * TILE_P + lx], TILE_P, ref, line, 4, 2, mode, 0, 0, S.dc[p]);
      }
      atomicAdd(&S.satd[p * 35 + mode], v);
    }
    __syncthreads();
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
