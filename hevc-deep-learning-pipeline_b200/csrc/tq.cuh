// tq.cuh -- transform-unit coding core on the device (first slice of the RD pass):
// forward core transform (DCT 4..32 / DST 4 / transform skip) -> flat quantiser -> dequantiser -> inverse transform for a
// batch of TUs, bit-exact with the reference's TComTrQuant::transformNxN / invTransformNxN
// (HM TLibCommon/TComTrQuant.cpp:1450-1666: xT -> xTrMxN :860, xQuant :1126 non-RDOQ branch, xDeQuant :1308, xIT -> xITrMxN :927)
// at the reference's operating point (8-bit video, dynamic range 15, no scaling lists).  The rate-distortion optimised
// quantiser (xRateDistOptQuant :2119, with its sign-bit hiding) is tq_rdoq.cuh; the flat quantiser's own sign-bit hiding
// (signBitHidingHDQ :991) is not on the device.
//
// Every 1-D pass is a small dense integer contraction  C = A x B,  A = the core matrix (|a| <= 90) or its transpose,
// B = 16..19-bit data.  It runs on the tensor cores as mma.sync m16n8k16 with fp16 operands and fp32 accumulation, EXACTLY:
// the data are split into a signed high part (v >> 8, |hi| <= 2047) and an unsigned low byte, both exact in fp16; every partial
// sum is an integer below 2^24, exact in fp32; the two accumulators are recombined as hi * 256 + lo in int32 before the
// reference's rounding shift and clip.  One warp codes one TU; the intermediate matrices live in shared memory.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace hevcdl {

// hevcdl_tu.flags
constexpr int TQ_DST = 1, TQ_TSKIP = 2, TQ_INTER = 4, TQ_RDOQ = 8, TQ_COEFF_IN = 16;

__device__ __constant__ int8_t c_tq_cos[32] = {0, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                                               61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4};
__device__ __constant__ int8_t c_tq_dst[16] = {29, 55, 74, 84, 74, 74, 0, -74, 84, -29, -74, 55, 55, -84, 74, -29};
__device__ __constant__ int c_tq_qscale[6] = {26214, 23302, 20560, 18396, 16384, 14564};   // g_quantScales, TComRom.cpp:354
__device__ __constant__ int c_tq_iqscale[6] = {40, 45, 51, 57, 64, 72};                    // g_invQuantScales

}  // namespace hevcdl
#include "tq_rdoq.cuh"
namespace hevcdl {

// entry (k, n) of the N-point core transform matrix (HEVC 8.6.4.2; TComRom.cpp:368-520 spells it out as macros)
__device__ __forceinline__ int tq_matrix(int lgN, int k, int n) {
  if (k == 0) return 64;
  int r = ((2 * n + 1) * (k << (5 - lgN))) & 127;
  if (r > 64) r = 128 - r;
  return r > 32 ? -(int)c_tq_cos[64 - r] : (int)c_tq_cos[r];
}

// shared-memory tables: for lgN = 2..5 the matrix T [k][n] and its transpose, as halves; then DST and DST^T
__host__ __device__ constexpr int tq_tab_off(int lg) { return lg == 2 ? 0 : (lg == 3 ? 32 : (lg == 4 ? 160 : 672)); }   // halves; size lg holds 2 * N * N entries
constexpr int TQ_TAB_DST = 672 + 2048, TQ_TAB_HALVES = TQ_TAB_DST + 32;
constexpr int TQ_WARPS = 4;

struct TqWarpS { int32_t a[1024], b[1024]; };
struct TqBlockS { __half tab[TQ_TAB_HALVES]; TqWarpS w[TQ_WARPS]; };

__device__ __forceinline__ uint32_t tq_pack(int lo, int hi) {
  const __half2 h = __halves2half2(__int2half_rn(lo), __int2half_rn(hi));
  return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ void tq_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One 1-D pass of one TU by one warp:  C[r][j] = sum_n A[r][n] * S[j][n]  (A: N x N halves row-major, S: N x N int32
// row-major in shared memory), then v = clip((C + add) >> shift) and
//   TRANSPOSE_OUT = false: out[r * N + j] = v      TRANSPOSE_OUT = true: out[j * N + r] = v
// OutT = int32_t (shared memory) or int16_t (global memory).
template <bool TRANSPOSE_OUT, typename OutT>
__device__ __forceinline__ void tq_pass(const __half *__restrict__ A, const int32_t *__restrict__ S, int N, int add, int shift, int lo_clip,
                                        int hi_clip, OutT *__restrict__ out, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int mt_n = (N + 15) >> 4, nt_n = (N + 7) >> 3, ks_n = (N + 15) >> 4;
  for (int mt = 0; mt < mt_n; mt++) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    for (int nt = 0; nt < nt_n; nt++) {
      const int j = nt * 8 + g;
      float chi[4] = {0.f, 0.f, 0.f, 0.f}, clo[4] = {0.f, 0.f, 0.f, 0.f};
      for (int ks = 0; ks < ks_n; ks++) {
        const int c0 = ks * 16 + 2 * t, c1 = c0 + 8;
        uint32_t a[4];
        a[0] = (r0 < N && c0 < N) ? *reinterpret_cast<const uint32_t *>(A + r0 * N + c0) : 0u;
        a[1] = (r1 < N && c0 < N) ? *reinterpret_cast<const uint32_t *>(A + r1 * N + c0) : 0u;
        a[2] = (r0 < N && c1 < N) ? *reinterpret_cast<const uint32_t *>(A + r0 * N + c1) : 0u;
        a[3] = (r1 < N && c1 < N) ? *reinterpret_cast<const uint32_t *>(A + r1 * N + c1) : 0u;
        int2 v0 = make_int2(0, 0), v1 = make_int2(0, 0);
        if (j < N && c0 < N) v0 = *reinterpret_cast<const int2 *>(S + j * N + c0);
        if (j < N && c1 < N) v1 = *reinterpret_cast<const int2 *>(S + j * N + c1);
        tq_mma(chi, a, tq_pack(v0.x >> 8, v0.y >> 8), tq_pack(v1.x >> 8, v1.y >> 8));
        tq_mma(clo, a, tq_pack(v0.x & 255, v0.y & 255), tq_pack(v1.x & 255, v1.y & 255));
      }
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int r = (e & 2) ? r1 : r0, jj = nt * 8 + 2 * t + (e & 1);
        if (r < N && jj < N) {
          int v = (__float2int_rn(chi[e]) * 256 + __float2int_rn(clo[e]) + add) >> shift;
          v = v < lo_clip ? lo_clip : (v > hi_clip ? hi_clip : v);
          out[TRANSPOSE_OUT ? jj * N + r : r * N + jj] = (OutT)v;
        }
      }
    }
  }
}

// tus: {log2 size, qp, flags, 0, element offset}; resi: int16 blocks (row-major) at the offsets; outputs at the same offsets
// (coeff / deq may be null); per TU abs_sum (uiAbsSum of transformNxN: 0 = no coded coefficient) and ssd = sum (resi - rec)^2.
__global__ void __launch_bounds__(TQ_WARPS * 32)
k_tu_code(int n, const hevcdl_tu *__restrict__ tus, const int16_t *__restrict__ resi, int32_t *__restrict__ coeff_out,
          int16_t *__restrict__ level_out, int32_t *__restrict__ deq_out, int16_t *__restrict__ rec_out, uint32_t *__restrict__ abs_sum_out,
          uint64_t *__restrict__ ssd_out, const hevcdl_tu_rdoq *__restrict__ rq, const int *__restrict__ est_tabs, RdoqScratch *__restrict__ scratch) {
  extern __shared__ __align__(16) uint8_t tq_smem[];
  TqBlockS &S = *reinterpret_cast<TqBlockS *>(tq_smem);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < TQ_TAB_HALVES; i += blockDim.x) {
    int v;
    if (i >= TQ_TAB_DST) { const int e = i - TQ_TAB_DST, tr = e >> 4, r = (e >> 2) & 3, c = e & 3; v = tr ? c_tq_dst[c * 4 + r] : c_tq_dst[r * 4 + c]; }
    else {
      const int lg = i < 32 ? 2 : (i < 160 ? 3 : (i < 672 ? 4 : 5)), N = 1 << lg;
      const int e = i - tq_tab_off(lg), tr = e >= N * N, q = e - tr * N * N, r = q >> lg, c = q & (N - 1);
      v = tr ? tq_matrix(lg, c, r) : tq_matrix(lg, r, c);
    }
    S.tab[i] = __int2half_rn(v);
  }
  __syncthreads();
  TqWarpS &W = S.w[wid];
  for (int tu = blockIdx.x * TQ_WARPS + wid; tu < n; tu += gridDim.x * TQ_WARPS) {
    const hevcdl_tu d = tus[tu];
    const int lg = d.log2_size, N = 1 << lg, n2 = N * N, flags = d.flags;
    const size_t off = d.offset;
    const bool dst = (flags & TQ_DST) && lg == 2, tskip = flags & TQ_TSKIP;
    const __half *T = dst ? S.tab + TQ_TAB_DST : S.tab + tq_tab_off(lg);
    const __half *Tt = T + (dst ? 16 : n2);
    const int tshift = 15 - 8 - lg;                                    // getTransformShift
    __syncwarp();
    for (int i = lane; i < n2; i += 32) W.a[i] = resi[off + i];
    __syncwarp();
    if (flags & TQ_COEFF_IN) {
      // the caller hands over transform coefficients: quantiser, dequantiser and inverse transform only
    } else if (tskip) {                                                // xTransformSkip, TComTrQuant.cpp:2010-2052
      for (int i = lane; i < n2; i += 32) W.a[i] = W.a[i] << tshift;
    } else {                                                           // xTrMxN, :860-925
      const int s1 = lg + 8 + 6 - 15, s2 = lg + 6;
      tq_pass<false>(T, W.a, N, s1 > 0 ? 1 << (s1 - 1) : 0, s1, -0x7fffffff - 1, 0x7fffffff, W.b, lane);
      __syncwarp();
      tq_pass<false>(T, W.b, N, 1 << (s2 - 1), s2, -0x7fffffff - 1, 0x7fffffff, W.a, lane);
    }
    __syncwarp();
    // xQuant (flat, :1184-1238) and xDeQuant (:1385-1420); the dequantised block is kept transposed for the next pass
    const int per = d.qp / 6, rem = d.qp - 6 * per;
    const int qbits = 14 + per + tshift;
    const long long qadd = (long long)((flags & TQ_INTER) ? 85 : 171) << (qbits - 9);
    const int qs = c_tq_qscale[rem], iqs = c_tq_iqscale[rem];
    const int rs = 6 - (tshift + per);
    const int tib = min(16, 32 + rs - 7);
    const int imin = -(1 << (tib - 1)), imax = (1 << (tib - 1)) - 1;
    uint32_t asum = 0;
    // the quantiser: xQuant's flat branch, or (HEVCDL_TU_RDOQ) the rate-distortion optimised one of tq_rdoq.cuh
    const bool use_rdoq = rq != nullptr && (flags & TQ_RDOQ);
    RdoqScratch *RS = use_rdoq ? scratch + (blockIdx.x * TQ_WARPS + wid) : nullptr;
    if (use_rdoq) {
      const hevcdl_tu_rdoq r = rq[tu];
      asum = rdoq_tu(W.a, lg, r.channel, r.scan_type, d.qp, r.lambda, est_tabs + (size_t)r.est_index * EST_INTS, r.ctx_cbf, r.flags, *RS, lane);
    }
    for (int i = lane; i < n2; i += 32) {
      const int c = W.a[i];
      int q;
      if (use_rdoq) q = RS->level[i];
      else {
        const int mag = (int)(((long long)abs(c) * qs + qadd) >> qbits);
        asum += (uint32_t)mag;
        q = c < 0 ? -mag : mag;
        q = max(-32768, min(32767, q));
      }
      const int cq = max(imin, min(imax, q));
      int v = rs > 0 ? (cq * iqs + (1 << (rs - 1))) >> rs : (int)((uint32_t)(cq * iqs) << (-rs));
      v = max(-32768, min(32767, v));
      if (coeff_out) coeff_out[off + i] = c;
      level_out[off + i] = (int16_t)q;
      if (deq_out) deq_out[off + i] = v;
      W.b[tskip ? i : (i & (N - 1)) * N + (i >> lg)] = v;
    }
    if (!use_rdoq) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
    }
    __syncwarp();
    if (tskip) {                                                       // xITransformSkip, :2060-2104
      const int offs = tshift == 0 ? 0 : 1 << (tshift - 1);
      for (int i = lane; i < n2; i += 32) W.b[i] = (int)(int16_t)((W.b[i] + offs) >> tshift);
    } else {                                                           // xITrMxN, :927-988
      tq_pass<false>(Tt, W.b, N, 64, 7, -32768, 32767, W.a, lane);
      __syncwarp();
      tq_pass<true>(Tt, W.a, N, 1 << 11, 12, -32768, 32767, W.b, lane);
    }
    __syncwarp();
    unsigned long long ssd = 0;
    for (int i = lane; i < n2; i += 32) {
      const int r = W.b[i], e = (int)resi[off + i] - r;
      rec_out[off + i] = (int16_t)r;
      ssd += (unsigned long long)(e * e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssd += __shfl_xor_sync(0xffffffffu, ssd, o);
    if (lane == 0) { abs_sum_out[tu] = asum; if (ssd_out) ssd_out[tu] = ssd; }
  }
}

}  // namespace hevcdl
