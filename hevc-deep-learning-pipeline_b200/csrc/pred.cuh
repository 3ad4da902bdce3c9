// pred.cuh -- intra prediction of single blocks from explicit reference samples (the predictor the
// RD pass calls for every luma and chroma transform block), bit-exact with the reference's TComPrediction::predIntraAng
// (HM TLibCommon/TComPrediction.cpp:390-472): planar :731-781, DC :183-201 + boundary smoothing :794-817, angular :229-388
// incl. the projection of the side reference for negative angles and the first-column filter of the pure vertical /
// horizontal modes.  The caller passes the reference line HM prepared (filtered or not, TComPattern.cpp:119-570) in the
// layout of the RMD kernels: 4N+1 samples = left column bottom-up, corner, top row.  Edge filters apply to luma blocks of
// N <= 16 only (flag bit 0); chroma blocks never have them.
// One warp per request, one lane per pixel pair: latency-bound by design -- inside HM this is one synchronous call per
// block (a parity demonstration of the component, like hevcdl_tu_code); the throughput path is k_rmd_items.
#pragma once
#include "common.cuh"
#include "rmd.cuh"

namespace hevcdl {

struct IntraPredReq {      // mirrors hevcdl_pred_req (include/hevcdl.h)
  uint8_t log2_size, mode, flags, reserved;
  uint32_t line_offset;    // into lines[], in samples
  uint32_t pred_offset;    // into pred[], in samples (N*N written, dense)
};

constexpr int PRED_WARPS = 4;

// One predicted sample.  c = centre (corner) of the reference line: c[1+k] above sample k, c[-1-k] left sample k.
__device__ __forceinline__ int pred_pixel(const int16_t *__restrict__ c, int n, int lg, int mode, int x, int y, int dc, bool edge) {
  if (mode == 0) {                                // planar
    const int l = c[-1 - y], t = c[1 + x], blv = c[-1 - n], trv = c[1 + n];
    return ((l << lg) + n + (x + 1) * (trv - l) + (t << lg) + (y + 1) * (blv - t)) >> (lg + 1);
  }
  if (mode == 1) {                                // DC (+ boundary smoothing)
    if (!edge) return dc;
    if (x == 0 && y == 0) return (c[1] + c[-1] + 2 * dc + 2) >> 2;
    if (y == 0) return (c[1 + x] + 3 * dc + 2) >> 2;
    if (x == 0) return (c[-1 - y] + 3 * dc + 2) >> 2;
    return dc;
  }
  const bool ver = mode >= 18;
  const int am = ver ? mode - 26 : 10 - mode, aabs = abs(am);
  const int angle = am < 0 ? -(int)c_ang[aabs] : (int)c_ang[aabs], inv = c_inv[aabs];
  const int sg = ver ? 1 : -1;                    // the main reference runs along +line for vertical modes
  const int jj = ver ? y : x, ii = ver ? x : y;   // jj: distance from the main reference, ii: position along it
  const int pos = (jj + 1) * angle, di = pos >> 5, df = pos & 31;
  const int k = ii + di + 1;
  auto main_ref = [&](int i) -> int { return i >= 0 ? c[sg * i] : c[-sg * ((128 - i * inv) >> 8)]; };   // i < 0: projected side sample
  int v = main_ref(k);
  if (df) v = ((32 - df) * v + df * main_ref(k + 1) + 16) >> 5;
  if (edge && angle == 0 && ii == 0) v = clip255(v + ((c[-sg * (jj + 1)] - c[0]) >> 1));
  return v;
}

__global__ void __launch_bounds__(PRED_WARPS * 32)
k_intra_pred(int nreq, const IntraPredReq *__restrict__ reqs, const int16_t *__restrict__ lines, int16_t *__restrict__ pred) {
  __shared__ int16_t s_line[PRED_WARPS][260];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * PRED_WARPS + warp; r < nreq; r += gridDim.x * PRED_WARPS) {
    const IntraPredReq q = reqs[r];
    const int lg = q.log2_size, n = 1 << lg, mode = q.mode;
    const bool edge = (q.flags & 1) && n <= 16;
    int16_t *line = s_line[warp];
    for (int i = lane; i < 4 * n + 1; i += 32) line[i] = lines[q.line_offset + i];
    __syncwarp();
    const int16_t *c = line + 2 * n;
    int dc = 0;
    if (mode == 1) {
      int sum = 0;
      for (int i = lane; i < n; i += 32) sum += c[1 + i] + c[-1 - i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      dc = (sum + n) / (2 * n);
    }
    int16_t *out = pred + q.pred_offset;
    for (int p = lane; p < n * n; p += 32) out[p] = (int16_t)pred_pixel(c, n, lg, mode, p & (n - 1), p >> lg, dc, edge);
    __syncwarp();
  }
}

}  // namespace hevcdl
