// cnn_tc.cuh -- tcgen05 (5th-gen tensor core, TMEM accumulators) implementation of the
// depth-prediction CNN (use_model.py:16-58, BatchNorm in TRAINING mode per sample) for a whole
// frame: four persistent kernels chained through L2-resident bf16 intermediates.
//
//   K1 k_tc_l1   : CTU staging (Y + co-sited Cb/Cr -> RGB, zero outside the picture) + conv64 and the
//                  four per-quadrant conv1 as implicit GEMMs over 4x2-pixel "super-pixels"
//                  (M = 128 super-pixels, N = 2 channel halves x 8 positions x 8 channels, K = 8-pixel window rows of
//                  the (R,G) / (B,0) planes), batch-stat BN + ReLU + max-pool in the epilogue -> cat
//   K2 k_tc_conv2: conv2 (M = 8x16 pixels, N = 64 channels, K = 9 taps x 32 channels)      -> a2
//   K3 k_tc_conv3: conv3 (M = 128 channels, N = 256 = 4 samples x 8x8 pixels, K = 9 x 64)  -> features
//   K4 k_tc_fc   : fc1/fc2 on tensor cores for 32 samples per CTA, fc3 + argmax + label rules + per-CTU PU/item counts
// One launch covers the CTUs of up to MAX_BATCH frames (FrameBatch, common.cuh); every kernel starts with its prologue
// under programmatic dependent launch and waits for its predecessor only before touching the frame's data.
//
// All activations operands are read by the tensor core straight out of padded planes through
// sliding-window shared-memory descriptors (K-major, no swizzle: 16-byte units = one pixel x 8
// channels, LBO selects the second 8 channels / next pixels, SBO = plane row pitch) -- no im2col
// buffer is ever materialised.  Layouts, descriptor parameters and the packed weight blob are
// replayed on the CPU by tools/tc_emulate.py (tests/test_tc_layout_cpu.py).
#pragma once
#include <string>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace hevcdl {

constexpr bool TC_BUILT = true;

// ---- packed weight blob (tools/tc_pack.py) -------------------------------------------------
constexpr int SZ_L1W = 24 * 4096, SZ_W2 = 18 * 2048, SZ_W3 = 36 * 4096, SZ_FC1 = 32 * 32768, SZ_FC2 = 128 * 256 * 2;
constexpr int OFF_L1W = 0, OFF_W2 = OFF_L1W + SZ_L1W, OFF_W3 = OFF_W2 + SZ_W2, OFF_FC1 = OFF_W3 + SZ_W3,
              OFF_FC2 = OFF_FC1 + SZ_FC1, OFF_F32 = OFF_FC2 + SZ_FC2;
constexpr int N_F32 = 32 + 32 + 128 + 256 + 256 + 64 + 1024 + 16;
constexpr int SZ_HDLT = OFF_F32 + 4 * N_F32;
// float offsets inside the fp32 section
constexpr int F_G64 = 0, F_B64 = 16, F_G1 = 32, F_B1 = 48, F_G2 = 64, F_B2 = 128, F_G3 = 192, F_B3 = 320, F_F1B = 448,
              F_F2B = 704, F_F3W = 768, F_F3B = 1792;

// ---- global intermediate layouts ---------------------------------------------------------------
constexpr int CAT_PLANE = 18 * 18 * 16, CAT_BYTES = 10 * CAT_PLANE;   // 10 planes [18][18][8 ch] bf16
constexpr int A2_PLANE = 40 * 10 * 16, A2_BYTES = 8 * A2_PLANE;       // 8 planes [4*(y+1)+s][10][8 ch] bf16

struct TcParams {
  const uint8_t *blob = nullptr;   // device copy of the HDLT payload
  uint8_t *cat = nullptr, *a2 = nullptr, *feats = nullptr;
  float *logits_scratch = nullptr;
  int npad = 0;                    // samples padded to a multiple of 128
};

constexpr int TC_THREADS = 288;    // K3, K4: warps 0-7 epilogue, warp 8 loads + MMA issue (K1: + 3 staging warps, K2: 16 + 1 warps)
#define EPI_BAR_SYNC() asm volatile("bar.sync 1, 256;" ::: "memory")
#define EPI2_BAR_SYNC() asm volatile("bar.sync 1, 512;" ::: "memory")   // kernels with 16 epilogue warps

// Sum of v[e] over the 32 lanes, for all e in [0,32): lane l returns the total of element l.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool up = lane & off;
#pragma unroll
    for (int j = 0; j < n / 2; j++) {
      const float send = up ? v[j] : v[j + n / 2];
      const float keep = up ? v[j + n / 2] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// Sum of v[e] over the 32 lanes for e in [0,16): every lane l returns the total of element l & 15.
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int off = 8, n = 16; off >= 1; off >>= 1, n >>= 1) {
    const bool up = lane & off;
#pragma unroll
    for (int j = 0; j < n / 2; j++) {
      const float send = up ? v[j] : v[j + n / 2];
      const float keep = up ? v[j + n / 2] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// Train-mode BatchNorm of one channel of one sample (use_model.py:16-46): totals of x and
// x^2 over the sample -> y = x*sc + sh with the BIASED variance.  fp32: the sample sizes are powers of two, so
// the mean is exact up to the rounding of the total; the cancellation error of E[x^2]-mean^2 is ~6e-8*(1+mean^2/var)
// relative, orders below the bf16 operand rounding of this path (the fp32 parity path is cnn_fp32.cuh).
__device__ __forceinline__ void bn_scale_shift(float sum, float sumsq, float rcnt, float eps, float gamma, float beta,
                                               float &sc, float &sh) {
  const float mean = sum * rcnt;
  const float var = fmaxf(fmaf(-mean, mean, sumsq * rcnt), 0.f);
  sc = gamma * rsqrtf(var + eps);
  sh = fmaf(-mean, sc, beta);
}

__device__ __forceinline__ uint4 pack8_bf16(const float *y) {
  return make_uint4(tc::pack_bf16(y[0], y[1]), tc::pack_bf16(y[2], y[3]), tc::pack_bf16(y[4], y[5]), tc::pack_bf16(y[6], y[7]));
}

// ================================================================================================
// K1: staging + conv64 + conv1
// ================================================================================================
constexpr int P64_PITCH = 68, P64_BYTES = 68 * 68 * 4;   // (R,G) or (B,0) bf16 pairs, 2-pixel zero halo
constexpr int P1_PITCH = 36, P1_BYTES = 36 * 36 * 4;
// raw CTU tiles as the TMA engine delivers them, double-buffered: Y 64x64, Cb 32x32, Cr 32x32 (u8)
constexpr int RAW_Y = 0, RAW_U = 4096, RAW_V = 5120, RAW_BYTES = 6144;
constexpr int K1_W = 0, K1_P64 = K1_W + SZ_L1W, K1_P1 = K1_P64 + 2 * P64_BYTES, K1_RED = K1_P1 + 8 * P1_BYTES,
              K1_BAR = K1_RED + 2 * 8 * 16 * 4, K1_RAW = (K1_BAR + 128 + 127) / 128 * 128, K1_SMEM = K1_RAW + 2 * RAW_BYTES;
#define STAGE_BAR_SYNC() asm volatile("bar.sync 2, 96;" ::: "memory")   // the three staging warps of K1

// 8 epilogue warps (warp = TMEM lane quarter + 4 * channel half) + 1 TMA/MMA warp + 3 staging warps.  A 16-epilogue-warp
// variant (4 channels per thread) was measured slower (40.9 vs 37 us per 1080p frame): twice the TMEM load instructions and
// duplicated reduction work outweigh the extra latency hiding.
// The three roles run decoupled, one CTU apart: the staging warps fill the input planes of CTU j+1 as soon as the MMAs of
// CTU j have finished reading them (conv64 planes after tile pair 1, quadrant planes after pair 3), the MMA thread starts
// CTU j+1 as soon as planes and a TMEM slot are available, so the epilogue warps -- the bottleneck -- never wait for
// staging or for the first pair's MMA latency.
constexpr int K1_THREADS = 12 * 32, K1_MMA_WARP = 8, K1_STAGE_THREADS = 96, K1_STAGE_ITERS = (1024 + K1_STAGE_THREADS - 1) / K1_STAGE_THREADS;

__global__ void __launch_bounds__(K1_THREADS, 1)
k_tc_l1(const FrameBatch fb, const __grid_constant__ TmapBatch tm, FrameGeom geo, int pitch, int cpitch, const uint8_t *__restrict__ blob,
        uint8_t *__restrict__ cat) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tmem_slot;
  uint64_t *bar_full = reinterpret_cast<uint64_t *>(sm + K1_BAR);   // [2] accumulators of a tile pair complete
  uint64_t *bar_empty = bar_full + 2;                               // [2] ... read back by the 8 epilogue warps
  uint64_t *bar_w = bar_full + 4;
  uint64_t *bar_pfull = bar_full + 5;                               // [2] conv64 planes / quadrant planes staged (3 staging warps)
  uint64_t *bar_pfree = bar_full + 7;                               // [2] ... no longer read by the tensor core
  uint64_t *bar_raw = bar_full + 9;                                 // [2] raw Y / Cb / Cr tiles of a CTU landed (TMA)
  float *red = reinterpret_cast<float *>(sm + K1_RED);              // [2][8 warps][16]
  const float *fp = reinterpret_cast<const float *>(blob + OFF_F32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  TL_BEGIN();

  if (tid == 0) {
    mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1);
    mbar_init(&bar_empty[0], 8); mbar_init(&bar_empty[1], 8);
    mbar_init(bar_w, 1);
    mbar_init(&bar_pfull[0], K1_STAGE_THREADS / 32); mbar_init(&bar_pfull[1], K1_STAGE_THREADS / 32);
    mbar_init(&bar_pfree[0], 1); mbar_init(&bar_pfree[1], 1);
    mbar_init(&bar_raw[0], 1); mbar_init(&bar_raw[1], 1);
    mbar_init_fence();
  }
  if (warp == K1_MMA_WARP) tmem_alloc(&tmem_slot, 512);
  for (int i = tid; i < (2 * P64_BYTES + 8 * P1_BYTES) / 16; i += K1_THREADS) reinterpret_cast<uint4 *>(sm + K1_P64)[i] = make_uint4(0, 0, 0, 0);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_slot;
  if (warp == K1_MMA_WARP && elect_one()) {
    mbar_expect_tx(bar_w, SZ_L1W);
    for (int i = 0; i < 24; i++) bulk_g2s(sm + K1_W + i * 4096, blob + OFF_L1W + i * 4096, 4096, bar_w);
  }
  const uint32_t idesc = idesc_bf16(128, 128);
  const int total = geo.nctu * fb.n;            // CTUs of all frames of this launch
  pdl_wait();                                   // prologue done; from here on global memory of the frame is touched
  if (tid == 0) { TL_WAITED(); }
  const long long trace_t0 = clock64(); (void)trace_t0;

  if (warp > K1_MMA_WARP) {
    // ---- staging: (R,G) and (B,0) planes of the CTU and of its four zero-padded quadrants ----------
    // The raw tiles come in by TMA (cp.async.bulk.tensor, one 64x64 Y box and two 32x32 chroma boxes per CTU, zeros
    // outside the picture), one CTU ahead of the conversion: replaces the frame dump + PIL crops of the reference
    // (gen_frames.py:21, use_model.py:78-95).  -DHEVCDL_K1_LDG: the former __ldg staging, kept for the A/B measurement.
    const int stid = tid - (K1_MMA_WARP + 1) * 32;
    uint32_t j = 0;                               // CTUs done by this CTA
#ifndef HEVCDL_K1_LDG
    auto issue_raw = [&](int ctu, uint32_t buf) {   // one elected thread: the three boxes of a CTU
      const int f = ctu / geo.nctu, ctu_l = ctu - f * geo.nctu;
      const int cx = ctu_l % geo.ctu_w, cy = ctu_l / geo.ctu_w;
      uint8_t *raw = sm + K1_RAW + buf * RAW_BYTES;
      mbar_expect_tx(&bar_raw[buf], RAW_BYTES);
      tma_load_2d(raw + RAW_Y, &tm.y[f], cx * 64, cy * 64, &bar_raw[buf]);
      tma_load_2d(raw + RAW_U, &tm.u[f], cx * 32, cy * 32, &bar_raw[buf]);
      tma_load_2d(raw + RAW_V, &tm.v[f], cx * 32, cy * 32, &bar_raw[buf]);
    };
    if (stid == 0 && (int)blockIdx.x < total) issue_raw(blockIdx.x, 0);
#endif
#pragma unroll 1
    for (int ctu = blockIdx.x; ctu < total; ctu += gridDim.x, j++) {
      const int f = ctu / geo.nctu, ctu_l = ctu - f * geo.nctu;
      const int ctu_x = ctu_l % geo.ctu_w, ctu_y = ctu_l / geo.ctu_w;
      uint32_t yv[K1_STAGE_ITERS], uv[K1_STAGE_ITERS], vv[K1_STAGE_ITERS];
#ifndef HEVCDL_K1_LDG
      STAGE_BAR_SYNC();                           // every staging thread has read the other raw buffer (CTU j-1)
      if (stid == 0 && ctu + (int)gridDim.x < total) issue_raw(ctu + gridDim.x, (j + 1) & 1);
      MBAR_WAIT(&bar_raw[j & 1], (j >> 1) & 1, 23);
      const uint8_t *raw = sm + K1_RAW + (j & 1) * RAW_BYTES;
#pragma unroll
      for (int k = 0; k < K1_STAGE_ITERS; k++) {
        const int it = stid + K1_STAGE_THREADS * k, y = it >> 4, x4 = (it & 15) * 4;
        yv[k] = 0; uv[k] = 0; vv[k] = 0;
        if (it < 1024) {
          yv[k] = *reinterpret_cast<const uint32_t *>(raw + RAW_Y + y * 64 + x4);
          uv[k] = *reinterpret_cast<const uint16_t *>(raw + RAW_U + (y >> 1) * 32 + (x4 >> 1));
          vv[k] = *reinterpret_cast<const uint16_t *>(raw + RAW_V + (y >> 1) * 32 + (x4 >> 1));
        }
      }
#else
      const uint8_t *__restrict__ Y = fb.Y[f], *__restrict__ U = fb.U[f], *__restrict__ V = fb.V[f];
#pragma unroll
      for (int k = 0; k < K1_STAGE_ITERS; k++) {  // raw samples first: the loads fly while the planes are still being read
        const int it = stid + K1_STAGE_THREADS * k, y = it >> 4, x4 = (it & 15) * 4;
        const int gy = ctu_y * 64 + y, gx = ctu_x * 64 + x4;
        yv[k] = 0; uv[k] = 0; vv[k] = 0;
        if (it < 1024 && gy < geo.H && gx < geo.W) {   // W is a multiple of 8: 4 pixels are in or out together
          yv[k] = __ldg(reinterpret_cast<const uint32_t *>(Y + (size_t)gy * pitch + gx));
          uv[k] = __ldg(reinterpret_cast<const uint16_t *>(U + (size_t)(gy >> 1) * cpitch + (gx >> 1)));
          vv[k] = __ldg(reinterpret_cast<const uint16_t *>(V + (size_t)(gy >> 1) * cpitch + (gx >> 1)));
        }
      }
#endif
      uint32_t rg[K1_STAGE_ITERS][4], b0[K1_STAGE_ITERS][4];
#pragma unroll
      for (int k = 0; k < K1_STAGE_ITERS; k++) {
        const int it = stid + K1_STAGE_THREADS * k, y = it >> 4, x4 = (it & 15) * 4;
        const bool in = it < 1024 && ctu_y * 64 + y < geo.H && ctu_x * 64 + x4 < geo.W;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          int r = 0, g = 0, b = 0;
          if (in) yuv2rgb((yv[k] >> (8 * i)) & 255, (uv[k] >> (8 * (i >> 1))) & 255, (vv[k] >> (8 * (i >> 1))) & 255, r, g, b);
          rg[k][i] = pack_bf16((float)r, (float)g);
          b0[k][i] = pack_bf16((float)b, 0.f);
        }
      }
      if (j) MBAR_WAIT(&bar_pfree[0], (j - 1) & 1, 20);
#pragma unroll
      for (int k = 0; k < K1_STAGE_ITERS; k++) {
        const int it = stid + K1_STAGE_THREADS * k, y = it >> 4, x4 = (it & 15) * 4;
        if (it < 1024) {
          uint8_t *d64 = sm + K1_P64 + ((y + 2) * P64_PITCH + x4 + 2) * 4;
          *reinterpret_cast<uint2 *>(d64) = make_uint2(rg[k][0], rg[k][1]);
          *reinterpret_cast<uint2 *>(d64 + 8) = make_uint2(rg[k][2], rg[k][3]);
          *reinterpret_cast<uint2 *>(d64 + P64_BYTES) = make_uint2(b0[k][0], b0[k][1]);
          *reinterpret_cast<uint2 *>(d64 + P64_BYTES + 8) = make_uint2(b0[k][2], b0[k][3]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_pfull[0]);
      if (j) MBAR_WAIT(&bar_pfree[1], (j - 1) & 1, 21);
#pragma unroll
      for (int k = 0; k < K1_STAGE_ITERS; k++) {
        const int it = stid + K1_STAGE_THREADS * k, y = it >> 4, x4 = (it & 15) * 4;
        if (it < 1024) {
          const int q = (y >> 5) * 2 + (x4 >> 5);
          uint8_t *d1 = sm + K1_P1 + q * 2 * P1_BYTES + (((y & 31) + 2) * P1_PITCH + (x4 & 31) + 2) * 4;
          *reinterpret_cast<uint2 *>(d1) = make_uint2(rg[k][0], rg[k][1]);
          *reinterpret_cast<uint2 *>(d1 + 8) = make_uint2(rg[k][2], rg[k][3]);
          *reinterpret_cast<uint2 *>(d1 + P1_BYTES) = make_uint2(b0[k][0], b0[k][1]);
          *reinterpret_cast<uint2 *>(d1 + P1_BYTES + 8) = make_uint2(b0[k][2], b0[k][3]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_pfull[1]);
    }
  } else if (warp == K1_MMA_WARP) {
    // ---- MMA issue: 4 tile pairs per CTU (conv64 quarters 0-1, 2-3; conv1 quadrants 0-1, 2-3) --------
    if (elect_one()) {
      MBAR_WAIT(bar_w, 0, 0);
      const uint32_t sb = smem_u32(sm);
      uint32_t j = 0, npair = 0;                  // CTUs done, running count of tile pairs (same sequence in the epilogue)
#pragma unroll 1
      for (int ctu = blockIdx.x; ctu < total; ctu += gridDim.x, j++, npair += 4) {
#pragma unroll 1
        for (int pi = 0; pi < 4; pi++) {
          const uint32_t n = npair + pi, p = n & 1, use = n >> 1;
          if ((pi & 1) == 0) MBAR_WAIT(&bar_pfull[pi >> 1], j & 1, 22);
          MBAR_WAIT(&bar_empty[p], (use & 1) ^ 1, 1);
          fence_after_sync();
          const bool c1 = pi >= 2;
          uint32_t a0, a1, plane_stride, row_stride;
          if (!c1) {   // quarters t = 2*pi, 2*pi+1 -> qy = pi, qx = 0 / 1
            a0 = sb + K1_P64 + ((pi * 32) * P64_PITCH) * 4;
            a1 = a0 + 32 * 4;
            plane_stride = P64_BYTES; row_stride = P64_PITCH * 4;
          } else {
            a0 = sb + K1_P1 + ((pi - 2) * 2) * 2 * P1_BYTES;
            a1 = a0 + 2 * P1_BYTES;
            plane_stride = P1_BYTES; row_stride = P1_PITCH * 4;
          }
          const uint64_t da0 = smem_desc(a0, 16, 2 * row_stride), da1 = smem_desc(a1, 16, 2 * row_stride);
          const uint64_t db = smem_desc(sb + K1_W + (c1 ? 12 * 4096 : 0), 128, 256);
          const uint32_t d0 = tbase + (2 * p) * 128, d1 = d0 + 128;
#pragma unroll
          for (int kb = 0; kb < 12; kb++) {
            const uint64_t aofs = (uint64_t)(((kb & 1) * plane_stride + (kb >> 1) * row_stride) >> 4);
            const uint64_t bofs = (uint64_t)((kb * 4096) >> 4);
            mma_bf16_ss(d0, da0 + aofs, db + bofs, idesc, kb ? 1u : 0u);
            mma_bf16_ss(d1, da1 + aofs, db + bofs, idesc, kb ? 1u : 0u);
          }
          mma_commit(&bar_full[p]);
          if (pi & 1) mma_commit(&bar_pfree[pi >> 1]);   // the planes this half of the CTU read may be overwritten
        }
      }
    }
    __syncwarp();
  } else {
    // ---- epilogue: warp = (lane quarter, channel half) -------------------------------------------
    const int lq = warp & 3, h = warp >> 2;
    const int m = lq * 32 + lane, g = m >> 3, ii = m & 7;
    const float g64r = __ldg(fp + F_G64 + 8 * h + (lane & 7)), b64r = __ldg(fp + F_B64 + 8 * h + (lane & 7));
    const float g1r = __ldg(fp + F_G1 + 8 * h + (lane & 7)), b1r = __ldg(fp + F_B1 + 8 * h + (lane & 7));
    uint32_t rb = 0, npair = 0;
#pragma unroll 1
    for (int ctu = blockIdx.x; ctu < total; ctu += gridDim.x, npair += 4) {
      float s[8], q[8], pool64[4][8];
#pragma unroll
      for (int c = 0; c < 8; c++) { s[c] = 0.f; q[c] = 0.f; }
      // One tile of accumulators: load, hand the pair's slots back after its second tile, statistics, then either the
      // conv64 4x4 max-pool into pool64[T] (T compile-time: registers, no select chains) or -- conv1 -- the sample's
      // batch statistics, normalisation, 2x2 max-pool and store.  conv64's four tiles are unrolled; conv1's four run as a
      // loop over one body (the instruction footprint of the three roles of this kernel has to stay cache-resident).
      uint8_t *cbase = cat + (size_t)ctu * CAT_BYTES;
      auto tile = [&](auto Tc, const int t) {
        constexpr int T = decltype(Tc)::value;      // 0..3: conv64 tile T; 4: a conv1 tile (t = 4..7 at run time)
        const uint32_t pi = t >> 1, e = t & 1, p = pi & 1, use = (npair + pi) >> 1;   // npair is a multiple of 4
        if (e == 0) { MBAR_WAIT(&bar_full[p], use & 1, 2); fence_after_sync(); }
        float v[8][8];                              // [position in the 4x2 super-pixel][channel 8h + c]: 64 contiguous columns
        tmem_ld32(tmem_addr(tbase, lq * 32, (2 * p + e) * 128 + 64 * h), &v[0][0]);
        tmem_ld32(tmem_addr(tbase, lq * 32, (2 * p + e) * 128 + 64 * h + 32), &v[4][0]);
        tmem_ld_wait();
        if (e == 1) {   // both accumulators of the pair are in registers: hand the TMEM slots back
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_empty[p]);
        }
#ifdef HEVCDL_ABLATE_EPI
        if (v[0][0] != 12345.f) return;
#endif
        // sums and sums of squares, two channels per instruction (packed fp32 add / fma of sm_100)
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          float2 s2 = make_float2(s[c], s[c + 1]), q2 = make_float2(q[c], q[c + 1]);
#pragma unroll
          for (int pos = 0; pos < 8; pos++) {
            const float2 x = make_float2(v[pos][c], v[pos][c + 1]);
            s2 = __fadd2_rn(s2, x);
            q2 = __ffma2_rn(x, x, q2);
          }
          s[c] = s2.x; s[c + 1] = s2.y; q[c] = q2.x; q[c + 1] = q2.y;
        }
        if (T < 4) {
#pragma unroll
          for (int c = 0; c < 8; c++) {
            float mx = v[0][c];
#pragma unroll
            for (int pos = 1; pos < 8; pos++) mx = fmaxf(mx, v[pos][c]);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));     // other row pair of the 4x4 window
            pool64[T & 3][c] = mx;
          }
        }
        if (T >= 3) {
          // per-channel totals over the sample: transposing reduction inside the warp (lane l & 15 ends with value
          // l: sums of channels 0..7, then sums of squares), the 4 warps of this channel half through shared memory
          float pv[16];
#pragma unroll
          for (int c = 0; c < 8; c++) { pv[c] = s[c]; pv[8 + c] = q[c]; }
          const float wt = warp_transpose_sum16(pv, lane);
          if (lane < 16) red[(rb * 8 + warp) * 16 + lane] = wt;
          EPI_BAR_SYNC();
          float tot = 0.f;
#pragma unroll
          for (int k = 0; k < 4; k++) tot += red[(rb * 8 + h * 4 + k) * 16 + (lane & 15)];
          const float totq = __shfl_down_sync(0xffffffffu, tot, 8);       // lanes 0..7: sum of squares of channel `lane`
          float sc1, sh1;                                                  // scale/shift of channel lane & 7 (valid in lanes 0..7)
          bn_scale_shift(tot, totq, T == 3 ? 1.f / 4096.f : 1.f / 1024.f, 1e-5f * 255.f * 255.f, T == 3 ? g64r : g1r, T == 3 ? b64r : b1r, sc1, sh1);
          float sc[8], sh[8];
#pragma unroll
          for (int c = 0; c < 8; c++) { sc[c] = __shfl_sync(0xffffffffu, sc1, c); sh[c] = __shfl_sync(0xffffffffu, sh1, c); }
          rb ^= 1;
          if (T == 3) {
            if ((g & 1) == 0) {
#pragma unroll
              for (int tt = 0; tt < 4; tt++) {
                float yv[8];
#pragma unroll
                for (int c = 0; c < 8; c++) yv[c] = fmaxf(fmaf(pool64[tt][c], sc[c], sh[c]), 0.f);
                const int py = (tt >> 1) * 8 + (g >> 1), px = (tt & 1) * 8 + ii;
                *reinterpret_cast<uint4 *>(cbase + (8 + h) * CAT_PLANE + ((py + 1) * 18 + px + 1) * 16) = pack8_bf16(yv);
              }
            }
          } else {
            const int smp = t - 4;
            float y0[8], y1[8];
#pragma unroll
            for (int c = 0; c < 8; c++) {
              const float w0 = fmaxf(fmaxf(v[0][c], v[1][c]), fmaxf(v[4][c], v[5][c]));
              const float w1 = fmaxf(fmaxf(v[2][c], v[3][c]), fmaxf(v[6][c], v[7][c]));
              y0[c] = fmaxf(fmaf(w0, sc[c], sh[c]), 0.f);
              y1[c] = fmaxf(fmaf(w1, sc[c], sh[c]), 0.f);
            }
            uint8_t *d = cbase + (smp * 2 + h) * CAT_PLANE + ((g + 1) * 18 + 2 * ii + 1) * 16;
            *reinterpret_cast<uint4 *>(d) = pack8_bf16(y0);
            *reinterpret_cast<uint4 *>(d + 16) = pack8_bf16(y1);
          }
#pragma unroll
          for (int c = 0; c < 8; c++) { s[c] = 0.f; q[c] = 0.f; }
        }
      };
      tile(std::integral_constant<int, 0>{}, 0);
      tile(std::integral_constant<int, 1>{}, 1);
      tile(std::integral_constant<int, 2>{}, 2);
      tile(std::integral_constant<int, 3>{}, 3);
#pragma unroll 1
      for (int t = 4; t < 8; t++) tile(std::integral_constant<int, 4>{}, t);
    }
  }
  fence_before_sync();
  __syncthreads();
  TRACE_TOTAL(16, trace_t0);
  TL_END(1);
  if (warp == 8) tmem_dealloc(tbase, 512);
}

// ================================================================================================
// K2: conv2
// ================================================================================================
constexpr int K2_W = 0, K2_CAT = K2_W + SZ_W2, K2_RED = K2_CAT + 2 * CAT_BYTES, K2_BAR = K2_RED + 2 * 8 * 64 * 4,
              K2_SMEM = K2_BAR + 128;

// 16 epilogue warps (warp = TMEM lane quarter + 4 * 16-channel group) + 1 TMA/MMA warp: the epilogue is
// CUDA-core work over 64 K accumulator values per CTU and was issue-bound with 8 warps (2 per scheduler).
constexpr int K2_THREADS = 17 * 32, K2_MMA_WARP = 16;

__global__ void __launch_bounds__(K2_THREADS, 1)
k_tc_conv2(FrameGeom geo, const uint8_t *__restrict__ blob, const uint8_t *__restrict__ cat, uint8_t *__restrict__ a2) {
  using namespace tc;
  const long long trace_t0 = clock64(); (void)trace_t0;
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tmem_slot;
  uint64_t *bar_full = reinterpret_cast<uint64_t *>(sm + K2_BAR);   // [4] per sample
  uint64_t *bar_empty = bar_full + 4;                               // [4]
  uint64_t *bar_cfull = bar_full + 8;                               // [2] cat buffer landed
  uint64_t *bar_cfree = bar_full + 10;                              // [2] cat buffer no longer read
  uint64_t *bar_w = bar_full + 12;
  float *red = reinterpret_cast<float *>(sm + K2_RED);              // [2][16 warps][16 sums + 16 sums of squares]
  const float *fp = reinterpret_cast<const float *>(blob + OFF_F32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  TL_BEGIN();

  if (tid == 0) {
    for (int i = 0; i < 4; i++) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 16); }
    for (int i = 0; i < 2; i++) { mbar_init(&bar_cfull[i], 1); mbar_init(&bar_cfree[i], 1); }
    mbar_init(bar_w, 1);
    mbar_init_fence();
  }
  if (warp == K2_MMA_WARP) tmem_alloc(&tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_slot;
  const uint32_t idesc = idesc_bf16(128, 64);

  if (warp == K2_MMA_WARP) {
    if (elect_one()) {
      mbar_expect_tx(bar_w, SZ_W2);
      for (int i = 0; i < 9; i++) bulk_g2s(sm + K2_W + i * 4096, blob + OFF_W2 + i * 4096, 4096, bar_w);
      pdl_wait();                               // the weights are on their way; cat is the predecessor's output
      TL_WAITED();
      if ((int)blockIdx.x < geo.nctu) {
        mbar_expect_tx(&bar_cfull[0], CAT_BYTES);
        bulk_g2s(sm + K2_CAT, cat + (size_t)blockIdx.x * CAT_BYTES, CAT_BYTES, &bar_cfull[0]);
      }
      mbar_wait(bar_w, 0);
      const uint32_t sb = smem_u32(sm);
      uint32_t it = 0;
      for (int ctu = blockIdx.x; ctu < geo.nctu; ctu += gridDim.x, it++) {
        const uint32_t b = it & 1;
        const int next = ctu + gridDim.x;
        if (next < geo.nctu) {   // prefetch the next CTU's planes into the other buffer
          if (it >= 1) MBAR_WAIT(&bar_cfree[b ^ 1], ((it - 1) >> 1) & 1, 4);
          mbar_expect_tx(&bar_cfull[b ^ 1], CAT_BYTES);
          bulk_g2s(sm + K2_CAT + (b ^ 1) * CAT_BYTES, cat + (size_t)next * CAT_BYTES, CAT_BYTES, &bar_cfull[b ^ 1]);
        }
        MBAR_WAIT(&bar_cfull[b], (it >> 1) & 1, 5);
        const uint32_t cbuf = sb + K2_CAT + b * CAT_BYTES;
        const uint64_t db = smem_desc(sb + K2_W, 128, 256);
        const uint64_t d64 = smem_desc(cbuf + 8 * CAT_PLANE, CAT_PLANE, 288);   // shared conv64 planes
        // Two samples at a time = four independent accumulators in flight: an MMA that accumulates into the same TMEM
        // tile as its predecessor costs ~138 cycles whatever its size (tools/tc_probe.cu), so a lone chain of 64-column
        // MMAs runs the tensor pipe at a third of its rate.  The other two samples' tiles belong to the epilogue meanwhile.
#pragma unroll 1
        for (int sp = 0; sp < 4; sp += 2) {
          MBAR_WAIT(&bar_empty[sp], (it & 1) ^ 1, 6);
          MBAR_WAIT(&bar_empty[sp + 1], (it & 1) ^ 1, 6);
          fence_after_sync();
          const uint64_t d1a = smem_desc(cbuf + sp * 2 * CAT_PLANE, CAT_PLANE, 288);
          const uint64_t d1b = smem_desc(cbuf + (sp + 1) * 2 * CAT_PLANE, CAT_PLANE, 288);
          const uint32_t t0 = tbase + (2 * sp) * 64;     // tiles: sample sp left/right, sample sp+1 left/right
#pragma unroll
          for (int tap = 0; tap < 9; tap++) {
            const uint64_t aofs = (uint64_t)((tap / 3) * 18 + (tap % 3));   // 16-byte units
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const uint64_t daa = (j ? d64 : d1a) + aofs, dab = (j ? d64 : d1b) + aofs;
              const uint64_t bofs = (uint64_t)(((tap * 2 + j) * 2048) >> 4);
              const uint32_t acc = (tap | j) ? 1u : 0u;
              mma_bf16_ss(t0, daa, db + bofs, idesc, acc);
              mma_bf16_ss(t0 + 64, daa + 8, db + bofs, idesc, acc);     // right half: +8 pixels
              mma_bf16_ss(t0 + 128, dab, db + bofs, idesc, acc);
              mma_bf16_ss(t0 + 192, dab + 8, db + bofs, idesc, acc);
            }
          }
          mma_commit(&bar_full[sp]);
          mma_commit(&bar_full[sp + 1]);
        }
        mma_commit(&bar_cfree[b]);
      }
    }
    __syncwarp();
  } else {
    pdl_wait();                                   // every warp that stores to global memory orders itself behind the predecessor grid (a2 is still read by the previous batch's K3)
    const int lq = warp & 3, cq = warp >> 2;      // TMEM lane quarter, group of 16 output channels
    const int m = lq * 32 + lane, yy = m >> 3, xi = m & 7;
    const float g2r = __ldg(fp + F_G2 + 16 * cq + (lane & 15)), b2r = __ldg(fp + F_B2 + 16 * cq + (lane & 15));
    uint32_t it = 0, rb = 0;
    for (int ctu = blockIdx.x; ctu < geo.nctu; ctu += gridDim.x, it++) {
#pragma unroll 1
      for (int smp = 0; smp < 4; smp++) {
        MBAR_WAIT(&bar_full[smp], it & 1, 7);
        fence_after_sync();
        float v0[16], v1[16];
        tmem_ld16(tmem_addr(tbase, lq * 32, (2 * smp) * 64 + 16 * cq), v0);
        tmem_ld16(tmem_addr(tbase, lq * 32, (2 * smp + 1) * 64 + 16 * cq), v1);
        tmem_ld_wait();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[smp]);
#ifdef HEVCDL_ABLATE_EPI
        if (__all_sync(0xffffffffu, v0[0] != 12345.f)) continue;
#endif
        float s[16], q[16];
#pragma unroll
        for (int c = 0; c < 16; c += 2) {           // packed fp32: two channels per instruction
          const float2 x0 = make_float2(v0[c], v0[c + 1]), x1 = make_float2(v1[c], v1[c + 1]);
          const float2 s2 = __fadd2_rn(x0, x1), q2 = __ffma2_rn(x0, x0, __fmul2_rn(x1, x1));
          s[c] = s2.x; s[c + 1] = s2.y; q[c] = q2.x; q[c + 1] = q2.y;
        }
        const float ts = warp_transpose_sum16(s, lane), tq = warp_transpose_sum16(q, lane);   // lane l & 15: channel 16cq + (l & 15)
        if (lane < 16) { red[(rb * 16 + warp) * 32 + lane] = ts; red[(rb * 16 + warp) * 32 + 16 + lane] = tq; }
        // 2x2 max-pool: exchange halves with the x neighbour (lane^1) then the y neighbour (lane^8);
        // this thread ends with 4 channels of one window for each of the two tiles
        const bool b0 = lane & 1, b3 = lane & 8;
        float p0[8], p1[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const float keep0 = b0 ? v0[8 + c] : v0[c], send0 = b0 ? v0[c] : v0[8 + c];
          const float keep1 = b0 ? v1[8 + c] : v1[c], send1 = b0 ? v1[c] : v1[8 + c];
          p0[c] = fmaxf(keep0, __shfl_xor_sync(0xffffffffu, send0, 1));
          p1[c] = fmaxf(keep1, __shfl_xor_sync(0xffffffffu, send1, 1));
        }
        float w0[4], w1[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const float keep0 = b3 ? p0[4 + c] : p0[c], send0 = b3 ? p0[c] : p0[4 + c];
          const float keep1 = b3 ? p1[4 + c] : p1[c], send1 = b3 ? p1[c] : p1[4 + c];
          w0[c] = fmaxf(keep0, __shfl_xor_sync(0xffffffffu, send0, 8));
          w1[c] = fmaxf(keep1, __shfl_xor_sync(0xffffffffu, send1, 8));
        }
        EPI2_BAR_SYNC();
        // lane l finishes channel 16cq + (l & 15) (totals of the 4 warps of this channel group), then hands scale/shift
        // to the lanes that need it
        float a = 0.f, bq = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) { a += red[(rb * 16 + cq * 4 + k) * 32 + (lane & 15)]; bq += red[(rb * 16 + cq * 4 + k) * 32 + 16 + (lane & 15)]; }
        float sc1, sh1;
        bn_scale_shift(a, bq, 1.f / 256.f, 1e-5f, g2r, b2r, sc1, sh1);
        const int sub = 8 * (b0 ? 1 : 0) + 4 * (b3 ? 1 : 0);   // this thread's 4 channels inside the group of 16
        float y0[4], y1[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const float sc = __shfl_sync(0xffffffffu, sc1, sub + c), sh = __shfl_sync(0xffffffffu, sh1, sub + c);
          y0[c] = fmaxf(fmaf(w0[c], sc, sh), 0.f);
          y1[c] = fmaxf(fmaf(w1[c], sc, sh), 0.f);
        }
        rb ^= 1;
        // window (py, px): py = y/2, px = xi/2 (+4 for the right-half tile); plane row 4*(py+1)+smp, col px+1; the plane
        // holds channels [8*chunk, 8*chunk+8) per 16-byte pixel unit, this thread owns half of it
        const int chunk = 2 * cq + (b0 ? 1 : 0);
        uint8_t *d = a2 + (size_t)ctu * A2_BYTES + chunk * A2_PLANE + ((4 * ((yy >> 1) + 1) + smp) * 10 + (xi >> 1) + 1) * 16 + (b3 ? 8 : 0);
        *reinterpret_cast<uint2 *>(d) = make_uint2(tc::pack_bf16(y0[0], y0[1]), tc::pack_bf16(y0[2], y0[3]));
        *reinterpret_cast<uint2 *>(d + 4 * 16) = make_uint2(tc::pack_bf16(y1[0], y1[1]), tc::pack_bf16(y1[2], y1[3]));
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  TRACE_TOTAL(17, trace_t0);
  TL_END(2);
  if (warp == K2_MMA_WARP) tmem_dealloc(tbase, 512);
}

// ================================================================================================
// K3: conv3
// ================================================================================================
constexpr int K3_W = 0, K3_A2 = K3_W + SZ_W3, K3_RED = K3_A2 + A2_BYTES, K3_BAR = K3_RED + 2 * 2 * 8 * 128 * 4,
              K3_SMEM = K3_BAR + 128;

__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_conv3(FrameGeom geo, const uint8_t *__restrict__ blob, const uint8_t *__restrict__ a2, uint8_t *__restrict__ feats, int npad) {
  using namespace tc;
  const long long trace_t0 = clock64(); (void)trace_t0;
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tmem_slot;
  uint64_t *bar_full = reinterpret_cast<uint64_t *>(sm + K3_BAR);   // [2] accumulator ready
  uint64_t *bar_empty = bar_full + 2;                               // [2]
  uint64_t *bar_afull = bar_full + 4;                               // [4] plane pair landed
  uint64_t *bar_afree = bar_full + 8;                               // [4] plane pair no longer read
  uint64_t *bar_w = bar_full + 12;
  float *red = reinterpret_cast<float *>(sm + K3_RED);              // [2][2 halves][8][128]
  const float *fp = reinterpret_cast<const float *>(blob + OFF_F32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  TL_BEGIN();

  if (tid == 0) {
    for (int i = 0; i < 2; i++) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 8); }
    for (int i = 0; i < 4; i++) { mbar_init(&bar_afull[i], 1); mbar_init(&bar_afree[i], 1); }
    mbar_init(bar_w, 1);
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_slot;
  const uint32_t idesc = idesc_bf16(128, 256);

  if (warp == 8) {
    if (elect_one()) {
      mbar_expect_tx(bar_w, SZ_W3);
      for (int i = 0; i < 36; i++) bulk_g2s(sm + K3_W + i * 4096, blob + OFF_W3 + i * 4096, 4096, bar_w);
      pdl_wait();                               // a2 is the predecessor's output
      TL_WAITED();
      if ((int)blockIdx.x < geo.nctu)
        for (int j = 0; j < 4; j++) {
          mbar_expect_tx(&bar_afull[j], 2 * A2_PLANE);
          bulk_g2s(sm + K3_A2 + j * 2 * A2_PLANE, a2 + (size_t)blockIdx.x * A2_BYTES + j * 2 * A2_PLANE, 2 * A2_PLANE, &bar_afull[j]);
        }
      MBAR_WAIT(bar_w, 0, 8);
      const uint32_t sb = smem_u32(sm);
      const uint64_t dw = smem_desc(sb + K3_W, 128, 256);
      uint32_t it = 0;
      for (int ctu = blockIdx.x; ctu < geo.nctu; ctu += gridDim.x, it++) {
        const uint32_t t = it & 1;
        MBAR_WAIT(&bar_empty[t], ((it >> 1) & 1) ^ 1, 9);
        fence_after_sync();
        const uint32_t dacc = tbase + t * 256;
#pragma unroll 1
        for (int j = 0; j < 4; j++) {
          MBAR_WAIT(&bar_afull[j], it & 1, 10);
          const uint64_t dact = smem_desc(sb + K3_A2 + j * 2 * A2_PLANE, A2_PLANE, 160);
#pragma unroll
          for (int tap = 0; tap < 9; tap++) {
            const uint64_t aofs = (uint64_t)(4 * (tap / 3) * 10 + (tap % 3));   // 16-byte units
            mma_bf16_ss(dacc, dw + (uint64_t)(((j * 9 + tap) * 4096) >> 4), dact + aofs, idesc, (j | tap) ? 1u : 0u);
          }
          mma_commit(&bar_afree[j]);
        }
        mma_commit(&bar_full[t]);
        const int next = ctu + gridDim.x;
        if (next < geo.nctu)
          for (int j = 0; j < 4; j++) {   // refill each plane pair as soon as its MMAs have drained
            MBAR_WAIT(&bar_afree[j], it & 1, 11);
            mbar_expect_tx(&bar_afull[j], 2 * A2_PLANE);
            bulk_g2s(sm + K3_A2 + j * 2 * A2_PLANE, a2 + (size_t)next * A2_BYTES + j * 2 * A2_PLANE, 2 * A2_PLANE, &bar_afull[j]);
          }
      }
    }
    __syncwarp();
  } else {
    pdl_wait();                                   // feats is still read by the previous batch's K4: no store before the predecessor grid is complete
    const int lq = warp & 3, hy = warp >> 2;
    const int c = lq * 32 + lane;   // output channel
    uint32_t it = 0, rb = 0;
    const float gam = __ldg(fp + F_G3 + c), bet = __ldg(fp + F_B3 + c);
    for (int ctu = blockIdx.x; ctu < geo.nctu; ctu += gridDim.x, it++) {
      const uint32_t t = it & 1;
      MBAR_WAIT(&bar_full[t], (it >> 1) & 1, 12);
      fence_after_sync();
      float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
      float pooled[2][4][4];   // [py local][sample][px]
#pragma unroll
      for (int pr = 0; pr < 2; pr++) {   // pairs of rows y = 4*hy + 2*pr, +1
        float r0[32], r1[32];   // columns 8*s + x of one row
        tmem_ld32(tmem_addr(tbase, lq * 32, t * 256 + 128 * hy + 64 * pr), r0);
        tmem_ld32(tmem_addr(tbase, lq * 32, t * 256 + 128 * hy + 64 * pr + 32), r1);
        tmem_ld_wait();
#pragma unroll
        for (int smp = 0; smp < 4; smp++) {         // packed fp32: two pixels per instruction
          float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int x = 0; x < 8; x += 2) {
            const float2 a = make_float2(r0[8 * smp + x], r0[8 * smp + x + 1]), b = make_float2(r1[8 * smp + x], r1[8 * smp + x + 1]);
            s2 = __fadd2_rn(s2, __fadd2_rn(a, b));
            q2 = __ffma2_rn(a, a, __ffma2_rn(b, b, q2));
          }
          s[smp] += s2.x + s2.y;
          q[smp] += q2.x + q2.y;
        }
#pragma unroll
        for (int smp = 0; smp < 4; smp++)
#pragma unroll
          for (int px = 0; px < 4; px++)
            pooled[pr][smp][px] = fmaxf(fmaxf(r0[8 * smp + 2 * px], r0[8 * smp + 2 * px + 1]), fmaxf(r1[8 * smp + 2 * px], r1[8 * smp + 2 * px + 1]));
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_empty[t]);
#ifdef HEVCDL_ABLATE_EPI
      if (__all_sync(0xffffffffu, s[0] != 12345.f)) continue;
#endif
      float *rr = red + ((rb * 2 + hy) * 8) * 128;
#pragma unroll
      for (int smp = 0; smp < 4; smp++) { rr[smp * 128 + c] = s[smp]; rr[(4 + smp) * 128 + c] = q[smp]; }
      EPI_BAR_SYNC();
      const float *ro = red + ((rb * 2 + (hy ^ 1)) * 8) * 128;
      rb ^= 1;
#pragma unroll
      for (int smp = 0; smp < 4; smp++) {
        float sc, sh;
        bn_scale_shift(s[smp] + ro[smp * 128 + c], q[smp] + ro[(4 + smp) * 128 + c], 1.f / 64.f, 1e-5f, gam, bet, sc, sh);
        float yv[8];
#pragma unroll
        for (int pr = 0; pr < 2; pr++)
#pragma unroll
          for (int px = 0; px < 4; px++) yv[pr * 4 + px] = fmaxf(fmaf(pooled[pr][smp][px], sc, sh), 0.f);
        // feature k = c*16 + (2*hy+pr)*4 + px -> 8 consecutive k = one 16-byte core-matrix row
        const int n = 4 * ctu + smp, kcore = 2 * c + hy;
        const size_t ofs = ((size_t)(kcore >> 3) * (npad >> 3) + (n >> 3)) * 1024 + (kcore & 7) * 128 + (n & 7) * 16;
        *reinterpret_cast<uint4 *>(feats + ofs) = pack8_bf16(yv);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  TRACE_TOTAL(18, trace_t0);
  TL_END(3);
  if (warp == 8) tmem_dealloc(tbase, 512);
}

// ================================================================================================
// K4: fc1 + fc2 (tensor cores), fc3 + argmax + label rules; one CTA = FC_NT samples = FC_NT/4 CTUs.
// fc1 streams its 1 MB of bf16 weights through every CTA, so the sample tile is kept small (more CTAs in flight,
// 4x less MMA and epilogue time per CTA) rather than large.
// ================================================================================================
// Two instantiations: 32-sample tiles (a single frame: 64 CTAs, the tile count is what keeps the SMs busy) and 64-sample
// tiles (launch batches whose 32-sample tiles would not fit in one wave: half the weight traffic through L2).
template <int NT_, int NSTAGE_>
struct FcCfg {
  static constexpr int NT = NT_, NSTAGE = NSTAGE_;
  static constexpr int KPH = NT <= 32 ? 4 : 2;                    // K-phases = independent accumulators per output tile
  // fc2's accumulators sit behind fc1's, or -- when fc1's 2 * KPH tiles fill the 512 columns (NT = 128) -- over them: fc1's
  // are read out (and the CTA synchronised) before the first fc2 MMA is issued
  static constexpr int FC2_COL = 3 * KPH * NT <= 512 ? 2 * KPH * NT : 0;
  static constexpr int STAGE = 32768 + NT * 128;                  // fc1 weights [256 x 64] + features [NT x 64], bf16
  static constexpr int BAR = NSTAGE * STAGE;
  static constexpr int FC2W = (BAR + 128 + 1023) / 1024 * 1024;   // fc2 weights, loaded in the prologue
  static constexpr int F3W = FC2W + SZ_FC2;                       // fc3 weights transposed, fp32 [64][16]
  static constexpr int SMEM = F3W + 4096;
  // after the fc1 K loop the stage memory is reused:
  static constexpr int H1 = 0 /* bf16 fc2 operand [NT/8][32][8][8] */, H2 = H1 + NT * 512 /* fp32 [NT][65] */,
                       LG = H2 + (NT * 65 * 4 + 15) / 16 * 16 /* fp32 [NT][16] */;
  static_assert(SMEM <= 227 * 1024, "K4 shared memory");
  static_assert(2 * KPH * NT <= 512, "K4 tensor memory: 2*KPH fc1 accumulators (+ KPH fc2 accumulators, behind or over them)");
  static_assert(LG + NT * 16 * 4 <= BAR, "K4 epilogue scratch must fit in the stage memory");
  static_assert(NSTAGE <= 6, "barrier slots");
};
using FcSmall = FcCfg<32, 4>;
using FcLarge = FcCfg<64, 3>;
using FcXL = FcCfg<128, 3>;

template <class CFG>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_fc(FrameGeom geo, const FrameBatch fb, const uint8_t *__restrict__ blob, const uint8_t *__restrict__ feats, int npad, int boundary_fix) {
  using namespace tc;
  constexpr int FC_NT = CFG::NT, FC_KPH = CFG::KPH, K4_NSTAGE = CFG::NSTAGE, K4_STAGE = CFG::STAGE, K4_BAR = CFG::BAR, K4_FC2W = CFG::FC2W,
                K4_F3W = CFG::F3W, K4_H1 = CFG::H1, K4_H2 = CFG::H2, K4_LG = CFG::LG;
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tmem_slot;
  uint64_t *bar_full = reinterpret_cast<uint64_t *>(sm + K4_BAR);   // [K4_NSTAGE]
  uint64_t *bar_free = bar_full + 6;                                // [K4_NSTAGE]
  uint64_t *bar_done = bar_full + 12;                                // fc1 accumulators complete
  uint64_t *bar_w2 = bar_full + 13;                                 // fc2 weights landed
  uint64_t *bar_done2 = bar_full + 14;
  const float *fp = reinterpret_cast<const float *>(blob + OFF_F32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = blockIdx.x;   // sample tile
  pdl_launch_dependents();
  TL_BEGIN();
#ifdef HEVCDL_FC_TRACE
  long long tr[8];
#define FC_TR(i) tr[i] = clock64()
#else
#define FC_TR(i)
#endif
  FC_TR(0);

  if (tid == 0) {
    for (int i = 0; i < K4_NSTAGE; i++) { mbar_init(&bar_full[i], 1); mbar_init(&bar_free[i], 1); }
    mbar_init(bar_done, 1); mbar_init(bar_w2, 1); mbar_init(bar_done2, 1);
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, 512);   // fc1: 2 * KPH tiles of FC_NT columns, fc2: KPH tiles (FcCfg::FC2_COL)
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_slot;
  const uint32_t idesc = idesc_bf16(128, FC_NT);
  const uint32_t sb = smem_u32(sm);

  if (warp == 8) {
    if (elect_one()) {
      auto load_w = [&](int kc) {                // constant operand: fc1 weights of K chunk kc
        const int st = kc % K4_NSTAGE;
        mbar_expect_tx(&bar_full[st], K4_STAGE);
        bulk_g2s(sm + st * K4_STAGE, blob + OFF_FC1 + (size_t)kc * 32768, 32768, &bar_full[st]);
      };
      auto load_f = [&](int kc) {                // the predecessor's output: features of this sample tile
        const int st = kc % K4_NSTAGE;
        bulk_g2s(sm + st * K4_STAGE + 32768, feats + ((size_t)kc * (npad >> 3) + nt * (FC_NT / 8)) * 1024, FC_NT * 128, &bar_full[st]);
      };
      auto load = [&](int kc) { load_w(kc); load_f(kc); };
      mbar_expect_tx(bar_w2, SZ_FC2);
      for (int i = 0; i < 16; i++) bulk_g2s(sm + K4_FC2W + i * 4096, blob + OFF_FC2 + i * 4096, 4096, bar_w2);
      for (int kc = 0; kc < K4_NSTAGE; kc++) load_w(kc);
      pdl_wait();
      TL_WAITED();
      for (int kc = 0; kc < K4_NSTAGE; kc++) load_f(kc);
#pragma unroll 1
      for (int kc = 0; kc < 32; kc++) {
        const int st = kc % K4_NSTAGE;
        mbar_wait(&bar_full[st], (kc / K4_NSTAGE) & 1);
        fence_after_sync();
        const uint64_t da = smem_desc(sb + st * K4_STAGE, 128, 1024), db = smem_desc(sb + st * K4_STAGE + 32768, 128, 1024);
#pragma unroll
        for (int t = 0; t < 4; t++) {
          // eight independent accumulators (4 K-phases x 2 output tiles): a dependent MMA costs ~138 cycles whatever
          // its size (tools/tc_probe.cu), these are 16-cycle MMAs; the epilogue adds the four K-phases
          const uint32_t accp = (kc || t >= FC_KPH) ? 1u : 0u;
          mma_bf16_ss(tbase + (2 * (t % FC_KPH)) * FC_NT, da + (uint64_t)(t * 16), db + (uint64_t)(t * 16), idesc, accp);
          mma_bf16_ss(tbase + (2 * (t % FC_KPH) + 1) * FC_NT, da + (uint64_t)(1024 + t * 16), db + (uint64_t)(t * 16), idesc, accp);
        }
        mma_commit(&bar_free[st]);
        if (kc >= 1 && kc - 1 + K4_NSTAGE < 32) {   // refill the stage consumed one iteration ago
          mbar_wait(&bar_free[(kc - 1) % K4_NSTAGE], ((kc - 1) / K4_NSTAGE) & 1);
          load(kc - 1 + K4_NSTAGE);
        }
      }
      mma_commit(bar_done);
    }
    __syncwarp();
  } else {
    // fc1 epilogue: thread = output o, FC_NT sample columns -> relu -> bf16 -> fc2 B operand [n][k=o]
    const int lq = warp & 3, mh = warp >> 2;
    const int o = mh * 128 + lq * 32 + lane;
    const float bias = __ldg(fp + F_F1B + o);
    for (int f = tid; f < 16 * 64; f += 256)    // fc3 weights [16][64] -> [64][16]: conflict-free reads in the fc3 step
      reinterpret_cast<float *>(sm + K4_F3W)[(f & 63) * 16 + (f >> 6)] = __ldg(fp + F_F3W + f);
    FC_TR(1);
    pdl_wait();                                   // labels / logits / ctu_cnt are read by the previous batch's K6 and copies: no store before the predecessor grid is complete
    mbar_wait(bar_done, 0);
    fence_after_sync();
    FC_TR(2);
#pragma unroll 1
    for (int cb = 0; cb < FC_NT; cb += 32) {
      float v[32], w[32];
      tmem_ld32(tmem_addr(tbase, lq * 32, mh * FC_NT + cb), v);
#pragma unroll
      for (int t = 1; t < FC_KPH; t++) {        // the other K-phases
        tmem_ld32(tmem_addr(tbase, lq * 32, (2 * t + mh) * FC_NT + cb), w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] += w[j];
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < (FC_NT < 32 ? FC_NT : 32); j++) {
        const int n = cb + j;
        const __nv_bfloat16 hv = __float2bfloat16(fmaxf(v[j] + bias, 0.f));
        *reinterpret_cast<__nv_bfloat16 *>(sm + K4_H1 + (n >> 3) * 4096 + (o >> 3) * 128 + (n & 7) * 16 + (o & 7) * 2) = hv;
      }
    }
    fence_async_smem();
    fence_before_sync();
    FC_TR(3);
  }
  __syncthreads();
  fence_after_sync();
  if (warp == 8) {
    if (elect_one()) {
      mbar_wait(bar_w2, 0);
      const uint64_t da = smem_desc(sb + K4_FC2W, 128, 4096), db = smem_desc(sb + K4_H1, 128, 4096);
#pragma unroll
      for (int t = 0; t < 16; t++)
        mma_bf16_ss(tbase + CFG::FC2_COL + (t % FC_KPH) * FC_NT, da + (uint64_t)(t * 16), db + (uint64_t)(t * 16), idesc, t >= FC_KPH ? 1u : 0u);
      mma_commit(bar_done2);
    }
    __syncwarp();
  } else {
    const int lq = warp & 3;   // lanes = fc2 outputs (0..63 real): warps 0 and 1 read them back
    mbar_wait(bar_done2, 0);
    fence_after_sync();
    FC_TR(4);
    if (warp < 2) {
      const int o = lq * 32 + lane;
      const float bias = __ldg(fp + F_F2B + o);
      float *h2 = reinterpret_cast<float *>(sm + K4_H2);
#pragma unroll 1
      for (int cb = 0; cb < FC_NT; cb += 32) {
        float v[32], w[32];
        tmem_ld32(tmem_addr(tbase, lq * 32, CFG::FC2_COL + cb), v);
#pragma unroll
        for (int t = 1; t < FC_KPH; t++) {
          tmem_ld32(tmem_addr(tbase, lq * 32, CFG::FC2_COL + t * FC_NT + cb), w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] += w[j];
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < (FC_NT < 32 ? FC_NT : 32); j++) h2[(cb + j) * 65 + o] = fmaxf(v[j] + bias, 0.f);
      }
    }
    fence_before_sync();
  }
  __syncthreads();
  FC_TR(5);
  // fc3 (use_model.py:40,57): one thread per sample
  float *lg = reinterpret_cast<float *>(sm + K4_LG);
  if (tid < 256) {                              // 8 threads per sample, two logits each
#pragma unroll 1
    for (int idx = tid; idx < FC_NT * 8; idx += 256) {
      const int smp = idx >> 3, o2 = (idx & 7) * 2;
      const float *h2 = reinterpret_cast<const float *>(sm + K4_H2) + smp * 65;
      const float *w3 = reinterpret_cast<const float *>(sm + K4_F3W) + o2;
      float a0 = __ldg(fp + F_F3B + o2), a1 = __ldg(fp + F_F3B + o2 + 1);
#pragma unroll 8
      for (int i = 0; i < 64; i++) {
        const float x = h2[i];
        const float2 w = *reinterpret_cast<const float2 *>(w3 + i * 16);
        a0 = fmaf(w.x, x, a0); a1 = fmaf(w.y, x, a1);
      }
      const int n = nt * FC_NT + smp;            // global sample = 4 * global CTU + quadrant
      *reinterpret_cast<float2 *>(lg + smp * 16 + o2) = make_float2(a0, a1);
      if (n < 4 * geo.nctu * fb.n) {
        const int f = (n >> 2) / geo.nctu, nl = n - f * 4 * geo.nctu;
        if (fb.logits[f]) *reinterpret_cast<float2 *>(fb.logits[f] + (size_t)nl * 16 + o2) = make_float2(a0, a1);
      }
    }
  }
  __syncthreads();
  FC_TR(6);
  if (tid < FC_NT / 4) {
    const int ctu_g = nt * (FC_NT / 4) + tid;
    if (ctu_g < geo.nctu * fb.n) {
      const int f = ctu_g / geo.nctu, ctu = ctu_g - f * geo.nctu;
      uint8_t *__restrict__ labels = fb.labels[f];
      uint32_t *__restrict__ ctu_cnt = fb.ctu_cnt[f];
      uint8_t lab[16];
      logits_to_labels(lg + tid * 64, lab, ctu % geo.ctu_w, ctu / geo.ctu_w, geo.W, geo.H, boundary_fix);
      uint4 pk;
      uint32_t *pw = reinterpret_cast<uint32_t *>(&pk);
      for (int i = 0; i < 4; i++) pw[i] = lab[4 * i] | (lab[4 * i + 1] << 8) | (lab[4 * i + 2] << 16) | (lab[4 * i + 3] << 24);
      *reinterpret_cast<uint4 *>(labels + (size_t)ctu * 16) = pk;
      if (ctu_cnt) ctu_cnt[ctu] = ctu_plan_counts(lab, ctu % geo.ctu_w, ctu / geo.ctu_w, geo.W, geo.H);
    }
  }
  fence_before_sync();
  __syncthreads();
  FC_TR(7);
#ifdef HEVCDL_FC_TRACE
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
    printf("fc trace blk %d: setup %lld loop %lld fc1epi %lld fc2 %lld fc2epi %lld fc3 %lld labels %lld total %lld clk\n", (int)blockIdx.x,
           tr[1] - tr[0], tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4], tr[6] - tr[5], tr[7] - tr[6], tr[7] - tr[0]);
#endif
  TL_END(4);
  if (warp == 8) tmem_dealloc(tbase, 512);
}

// ---- host side -------------------------------------------------------------------------------------
inline int tc_prepare(const char *hdlt_path, int nctu /* CTUs of one launch: frames per batch x CTUs per frame */, TcParams *p, std::string &err) {
  FILE *f = fopen(hdlt_path, "rb");
  if (!f) { err = std::string("cannot open tensor-core weight blob: ") + hdlt_path; return HEVCDL_E_WEIGHTS; }
  std::vector<uint8_t> buf(SZ_HDLT);
  char magic[8], extra;
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "HDLT0001", 8) == 0 && fread(buf.data(), 1, SZ_HDLT, f) == (size_t)SZ_HDLT &&
            fread(&extra, 1, 1, f) == 0;
  fclose(f);
  if (!ok) { err = "malformed HDLT weight blob"; return HEVCDL_E_WEIGHTS; }
  uint8_t *d = nullptr;
  p->npad = ((4 * nctu + 127) / 128) * 128;
  const size_t feats_bytes = (size_t)p->npad * 4096;
  if (cudaMalloc(&d, SZ_HDLT) != cudaSuccess || cudaMemcpy(d, buf.data(), SZ_HDLT, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMalloc(&p->cat, (size_t)nctu * CAT_BYTES) != cudaSuccess || cudaMemset(p->cat, 0, (size_t)nctu * CAT_BYTES) != cudaSuccess ||
      cudaMalloc(&p->a2, (size_t)nctu * A2_BYTES) != cudaSuccess || cudaMemset(p->a2, 0, (size_t)nctu * A2_BYTES) != cudaSuccess ||
      cudaMalloc(&p->feats, feats_bytes) != cudaSuccess || cudaMemset(p->feats, 0, feats_bytes) != cudaSuccess) {
    err = std::string("tensor-core path allocation: ") + cudaGetErrorString(cudaGetLastError());
    return HEVCDL_E_CUDA;
  }
  p->blob = d;
  return HEVCDL_OK;
}

inline void tc_release(TcParams *p) {
  cudaFree(const_cast<uint8_t *>(p->blob)); cudaFree(p->cat); cudaFree(p->a2); cudaFree(p->feats);
  *p = TcParams{};
}

inline int tc_configure(std::string &err) {
  if (cudaFuncSetAttribute(k_tc_l1, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_tc_conv2, cudaFuncAttributeMaxDynamicSharedMemorySize, K2_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_tc_conv3, cudaFuncAttributeMaxDynamicSharedMemorySize, K3_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_tc_fc<FcSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, FcSmall::SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_tc_fc<FcLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize, FcLarge::SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_tc_fc<FcXL>, cudaFuncAttributeMaxDynamicSharedMemorySize, FcXL::SMEM) != cudaSuccess) {
    err = std::string("tensor-core kernels: shared-memory attribute: ") + cudaGetErrorString(cudaGetLastError());
    return HEVCDL_E_CUDA;
  }
  return HEVCDL_OK;
}

template <typename... KArgs, typename... Args>
inline cudaError_t tc_launch_pdl(void (*kernel)(KArgs...), int grid, int threads, int smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("HEVCDL_NO_PDL") != nullptr;   // timing experiments: plain stream order
  cfg.attrs = at; cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Queue the four CNN kernels of one frame (programmatic dependent launch: each kernel's prologue overlaps its
// predecessor's tail).  *launches += kernels launched; returns the first launch error.
inline cudaError_t tc_launch(const TcParams &p, const FrameBatch &fb, const TmapBatch &tm, FrameGeom g, int pitch, int cpitch, int boundary_fix, int num_sms,
                             cudaStream_t st, int *launches) {
  FrameGeom gt = g;                              // K2/K3 only loop over CTUs: give them the launch's total
  gt.nctu = g.nctu * fb.n;
  const int grid = gt.nctu < num_sms ? gt.nctu : num_sms;
  const int npad = ((4 * gt.nctu + 127) / 128) * 128;   // <= p.npad (sized for the largest batch); the feats layout follows the launch
  cudaError_t e;
  if ((e = tc_launch_pdl(k_tc_l1, grid, K1_THREADS, K1_SMEM, st, fb, tm, g, pitch, cpitch, p.blob, p.cat)) != cudaSuccess) return e;
  if ((e = tc_launch_pdl(k_tc_conv2, grid, K2_THREADS, K2_SMEM, st, gt, p.blob, (const uint8_t *)p.cat, p.a2)) != cudaSuccess) return e;
  if ((e = tc_launch_pdl(k_tc_conv3, grid, TC_THREADS, K3_SMEM, st, gt, p.blob, (const uint8_t *)p.a2, p.feats, npad)) != cudaSuccess) return e;
  // every CTA streams fc1's 1 MB of weights: as few sample tiles as still fill one wave of SMs
  if (npad / FcLarge::NT > num_sms)             // more 64-sample tiles than SMs (8-frame 1080p launches): 128-sample tiles
    e = tc_launch_pdl(k_tc_fc<FcXL>, npad / FcXL::NT, TC_THREADS, FcXL::SMEM, st, g, fb, p.blob, (const uint8_t *)p.feats, npad, boundary_fix);
  else if (npad / FcSmall::NT > num_sms)        // more 32-sample tiles than SMs: 64-sample tiles halve the weight traffic
    e = tc_launch_pdl(k_tc_fc<FcLarge>, npad / FcLarge::NT, TC_THREADS, FcLarge::SMEM, st, g, fb, p.blob, (const uint8_t *)p.feats, npad, boundary_fix);
  else
    e = tc_launch_pdl(k_tc_fc<FcSmall>, npad / FcSmall::NT, TC_THREADS, FcSmall::SMEM, st, g, fb, p.blob, (const uint8_t *)p.feats, npad, boundary_fix);
  if (e != cudaSuccess) return e;
  *launches += 4;
  return cudaSuccess;
}

}  // namespace hevcdl
