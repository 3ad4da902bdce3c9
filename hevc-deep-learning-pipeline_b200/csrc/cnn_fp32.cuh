// cnn_fp32.cuh -- fp32 CUDA-core implementation of the depth-prediction CNN, one CTA per CTU.
//
// Replaces the per-quadrant batch-1 torch forwards of the reference sidecar
// (use_model.py:86-101, ConvNet2 at :16-58) for a whole frame in one launch: K0 (tile staging +
// YUV->RGB), conv1/conv64/conv2/conv3 each with TRAINING-mode BatchNorm on the sample's own
// statistics, ReLU, max-pool, the three linear layers, the 4-way argmaxes and
// the label fix-ups (use_model.py:101-119).  conv64 is evaluated once per CTU and shared by its
// four quadrant samples (the reference recomputes it four times with identical inputs).
//
// BN + ReLU + max-pool are applied as  relu(bn(max_window(conv)))  -- identical to
// max_window(relu(bn(conv))) because bn is monotone (min_window is used when gamma < 0) -- so only
// the pooled raw conv outputs and per-channel (sum, sum of squares) need to be kept.
#pragma once
#include "common.cuh"

namespace hevcdl {

constexpr int FP32_THREADS = 512;
constexpr int IMG_P = 68;                          // padded 64x64 RGB tile pitch (pad 2)
constexpr int CAT_P = 18;                          // padded 16x16 (pad 1)
constexpr int A2_P = 10;                           // padded 8x8 (pad 1)
// shared memory plan (floats)
constexpr int SM_AC = 4 * 64 * A2_P * A2_P;        // conv2 out [4][64][10][10]; aliases img [3][68][68]
constexpr int SM_B = (4 * 16 + 16) * CAT_P * CAT_P;// conv1 out [4][16][18][18] + conv64 out [16][18][18]; later conv3 out [4][2048]
constexpr int SM_MISC = 4 * 256 + 4 * 64 + 4 * 16; // fc1 out, fc2 out, logits
constexpr int FP32_SMEM_BYTES = (SM_AC + SM_B + SM_MISC) * 4;
static_assert(SM_AC >= 3 * IMG_P * IMG_P, "img must fit in region AC");
static_assert(SM_B >= 4 * 2048, "conv3 out must fit in region B");

// One conv + BN(train) + ReLU + maxpool layer for NSAMP samples held in shared memory.
//   LoadIn(s, ci, y, x)   -> input value at padded coordinates (y,x in [0, S+K-1))
//   StoreOut(s, c, wy, wx, v)
// A task = (sample, CPT output channels); a warp owns a task: lanes = pooling windows, so the
// per-channel statistics finish with warp shuffles only (deterministic order).
template <int CIN, int COUT, int K, int S, int POOL, int CG, int NSAMP, class LoadIn, class StoreOut>
__device__ __forceinline__ void conv_bn_relu_pool(const float *__restrict__ wpk, const float *__restrict__ bias,
                                                  const float *__restrict__ gamma, const float *__restrict__ beta,
                                                  LoadIn load_in, StoreOut store_out) {
  constexpr int WPR = S / POOL, NW = WPR * WPR;
  constexpr int NWL = NW < 32 ? NW : 32;           // lanes that enumerate windows
  constexpr int NSUB = 32 / NWL;                   // channel sub-groups per warp
  constexpr int CPT = CG * NSUB;                   // channels per task
  constexpr int WPL = NW / NWL;                    // windows per lane
  constexpr int PW = POOL + K - 1;                 // input patch width
  constexpr int NTASK = NSAMP * (COUT / CPT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int wl = lane % NWL, sub = lane / NWL;

  for (int task = warp; task < NTASK; task += nwarps) {
    const int s = task / (COUT / CPT);
    const int c0 = (task % (COUT / CPT)) * CPT + sub * CG;
    const float *wbase = wpk + (size_t)(c0 / CG) * CIN * K * K * CG;
    double sum[CG], sq[CG];
    float pooled[WPL][CG];
    float sgn[CG], bs[CG];
#pragma unroll
    for (int c = 0; c < CG; c++) {
      sum[c] = 0.0; sq[c] = 0.0;
      sgn[c] = gamma[c0 + c] < 0.f ? -1.f : 1.f;
      bs[c] = bias[c0 + c];
    }
#pragma unroll
    for (int wi = 0; wi < WPL; wi++) {
      const int win = wl + wi * NWL, wy = win / WPR, wx = win % WPR;
      float acc[CG][POOL * POOL];
#pragma unroll
      for (int c = 0; c < CG; c++)
#pragma unroll
        for (int p = 0; p < POOL * POOL; p++) acc[c][p] = 0.f;
      for (int ci = 0; ci < CIN; ci++) {
        float patch[PW * PW];
#pragma unroll
        for (int py = 0; py < PW; py++)
#pragma unroll
          for (int px = 0; px < PW; px++) patch[py * PW + px] = load_in(s, ci, wy * POOL + py, wx * POOL + px);
        const float *wp = wbase + (size_t)ci * K * K * CG;
#pragma unroll
        for (int ky = 0; ky < K; ky++)
#pragma unroll
          for (int kx = 0; kx < K; kx++) {
            float wv[CG];
#pragma unroll
            for (int c = 0; c < CG; c++) wv[c] = __ldg(wp + (ky * K + kx) * CG + c);
#pragma unroll
            for (int c = 0; c < CG; c++)
#pragma unroll
              for (int py = 0; py < POOL; py++)
#pragma unroll
                for (int px = 0; px < POOL; px++)
                  acc[c][py * POOL + px] = fmaf(wv[c], patch[(py + ky) * PW + px + kx], acc[c][py * POOL + px]);
          }
      }
#pragma unroll
      for (int c = 0; c < CG; c++) {
        float m = -3.4e38f;
#pragma unroll
        for (int p = 0; p < POOL * POOL; p++) {
          float v = acc[c][p] + bs[c];
          sum[c] += (double)v;
          sq[c] += (double)v * (double)v;
          m = fmaxf(m, v * sgn[c]);
        }
        pooled[wi][c] = m * sgn[c];
      }
    }
    // statistics over the NWL window-lanes of this channel sub-group
#pragma unroll
    for (int c = 0; c < CG; c++) {
#pragma unroll
      for (int o = NWL / 2; o > 0; o >>= 1) {
        sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], o);
        sq[c] += __shfl_xor_sync(0xffffffffu, sq[c], o);
      }
      const double mean = sum[c] / (double)(S * S);
      double var = sq[c] / (double)(S * S) - mean * mean;   // biased variance (train-mode BN)
      var = var < 0.0 ? 0.0 : var;
      const double inv = rsqrt(var + 1e-5) * (double)gamma[c0 + c];
      const double bt = (double)beta[c0 + c];
#pragma unroll
      for (int wi = 0; wi < WPL; wi++) {
        const int win = wl + wi * NWL;
        float y = (float)(((double)pooled[wi][c] - mean) * inv + bt);
        store_out(s, c0 + c, win / WPR, win % WPR, y > 0.f ? y : 0.f);
      }
    }
  }
}

// Stage the CTU tile: Y 64x64 + co-sited Cb/Cr 32x32 -> RGB/255 in img[3][68][68] with a zero
// border of 2; samples outside the picture are RGB 0 (PIL crop black padding, use_model.py:92-93).
__device__ __forceinline__ void stage_ctu_rgb(const uint8_t *__restrict__ Y, const uint8_t *__restrict__ U,
                                              const uint8_t *__restrict__ V, int W, int H, int pitch, int cpitch,
                                              int ctu_x, int ctu_y, float *img) {
  for (int i = threadIdx.x; i < 3 * IMG_P * IMG_P; i += blockDim.x) img[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) {       // 4 pixels per item
    const int y = i >> 4, x4 = (i & 15) * 4;
    const int py = ctu_y * 64 + y, px = ctu_x * 64 + x4;
    if (py >= H || px >= W) continue;                             // W multiple of 8: all 4 in or out
    const uint32_t yv = *reinterpret_cast<const uint32_t *>(Y + (size_t)py * pitch + px);
    const uint16_t uv = *reinterpret_cast<const uint16_t *>(U + (size_t)(py >> 1) * cpitch + (px >> 1));
    const uint16_t vv = *reinterpret_cast<const uint16_t *>(V + (size_t)(py >> 1) * cpitch + (px >> 1));
#pragma unroll
    for (int k = 0; k < 4; k++) {
      int r, g, b;
      yuv2rgb((yv >> (8 * k)) & 255, (uv >> (8 * (k >> 1))) & 255, (vv >> (8 * (k >> 1))) & 255, r, g, b);
      const int o = (y + 2) * IMG_P + x4 + k + 2;
      img[o] = (float)r / 255.0f;                                 // torchvision ToTensor
      img[IMG_P * IMG_P + o] = (float)g / 255.0f;
      img[2 * IMG_P * IMG_P + o] = (float)b / 255.0f;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FP32_THREADS, 1)
k_cnn_fp32(const uint8_t *__restrict__ Y, const uint8_t *__restrict__ U, const uint8_t *__restrict__ V,
           FrameGeom geo, int pitch, int cpitch, Fp32Params P, int boundary_fix, uint8_t *__restrict__ labels,
           float *__restrict__ logits_out, uint32_t *__restrict__ ctu_cnt) {
  extern __shared__ float smem[];
  pdl_launch_dependents();
  pdl_wait();
  float *regAC = smem;                 // img, then conv2 out
  float *regB = smem + SM_AC;          // conv1/conv64 out, then conv3 out
  float *fc1o = regB + SM_B, *fc2o = fc1o + 4 * 256, *lgt = fc2o + 4 * 64;

  for (int ctu = blockIdx.x; ctu < geo.nctu; ctu += gridDim.x) {
    const int ctu_x = ctu % geo.ctu_w, ctu_y = ctu / geo.ctu_w;
    stage_ctu_rgb(Y, U, V, geo.W, geo.H, pitch, cpitch, ctu_x, ctu_y, regAC);
    for (int i = threadIdx.x; i < SM_B; i += blockDim.x) regB[i] = 0.f;   // zero borders of the cat planes
    __syncthreads();
    const float *img = regAC;
    float *cat1 = regB;                                  // [4][16][18][18]
    float *cat64 = regB + 4 * 16 * CAT_P * CAT_P;        // [16][18][18]

    // conv64: Conv2d(3,16,5,pad 2)+BN+ReLU+MaxPool(4) on the whole CTU (use_model.py:41-46)
    conv_bn_relu_pool<3, 16, 5, 64, 4, 1, 1>(
        P.c64w, P.c64b, P.g64, P.b64,
        [&](int, int ci, int y, int x) { return img[(ci * IMG_P + y) * IMG_P + x]; },
        [&](int, int c, int wy, int wx, float v) { cat64[(c * CAT_P + wy + 1) * CAT_P + wx + 1] = v; });
    // conv1: Conv2d(3,16,5,pad 2)+BN+ReLU+MaxPool(2) on each 32x32 quadrant crop (use_model.py:20-25);
    // taps leaving the crop read zero (the crop is padded on its own, not with its neighbours).
    conv_bn_relu_pool<3, 16, 5, 32, 2, 4, 4>(
        P.c1w, P.c1b, P.g1, P.b1,
        [&](int s, int ci, int y, int x) {
          const bool in = (y >= 2) & (y < 34) & (x >= 2) & (x < 34);
          return in ? img[(ci * IMG_P + (s >> 1) * 32 + y) * IMG_P + (s & 1) * 32 + x] : 0.f;
        },
        [&](int s, int c, int wy, int wx, float v) { cat1[((s * 16 + c) * CAT_P + wy + 1) * CAT_P + wx + 1] = v; });
    __syncthreads();
    for (int i = threadIdx.x; i < SM_AC; i += blockDim.x) regAC[i] = 0.f;   // img dead; zero conv2-out borders
    __syncthreads();
    // conv2 on cat([conv1, conv64]) (use_model.py:26-31,50-51)
    float *a2 = regAC;                                   // [4][64][10][10]
    conv_bn_relu_pool<32, 64, 3, 16, 2, 4, 4>(
        P.c2w, P.c2b, P.g2, P.b2,
        [&](int s, int ci, int y, int x) {
          return ci < 16 ? cat1[((s * 16 + ci) * CAT_P + y) * CAT_P + x] : cat64[((ci - 16) * CAT_P + y) * CAT_P + x];
        },
        [&](int s, int c, int wy, int wx, float v) { a2[((s * 64 + c) * A2_P + wy + 1) * A2_P + wx + 1] = v; });
    __syncthreads();
    // conv3 (use_model.py:32-37); output flattened in C,H,W order (view at :53)
    float *a3 = regB;                                    // [4][2048]
    conv_bn_relu_pool<64, 128, 3, 8, 2, 4, 4>(
        P.c3w, P.c3b, P.g3, P.b3,
        [&](int s, int ci, int y, int x) { return a2[((s * 64 + ci) * A2_P + y) * A2_P + x]; },
        [&](int s, int c, int wy, int wx, float v) { a3[s * 2048 + c * 16 + wy * 4 + wx] = v; });
    __syncthreads();
    // fc1 2048->256 + ReLU (use_model.py:38,54): thread = (half of K, output o), 4 samples each
    {
      const int o = threadIdx.x & 255, half = threadIdx.x >> 8;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float *wT = P.f1wT + (size_t)half * 1024 * 256 + o;
      const float *in = a3 + half * 1024;
#pragma unroll 4
      for (int i = 0; i < 1024; i++) {
        const float w = __ldg(wT + (size_t)i * 256);
#pragma unroll
        for (int s = 0; s < 4; s++) acc[s] = fmaf(w, in[s * 2048 + i], acc[s]);
      }
      float *part = regAC;                               // a2 dead
      __syncthreads();
#pragma unroll
      for (int s = 0; s < 4; s++) part[(half * 4 + s) * 256 + o] = acc[s];
      __syncthreads();
      if (half == 0) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
          float v = part[s * 256 + o] + part[(4 + s) * 256 + o] + __ldg(P.f1b + o);
          fc1o[s * 256 + o] = v > 0.f ? v : 0.f;
        }
      }
      __syncthreads();
    }
    // fc2 256->64 + ReLU (use_model.py:39,56)
    if (threadIdx.x < 256) {
      const int s = threadIdx.x >> 6, o = threadIdx.x & 63;
      float acc = __ldg(P.f2b + o);
      for (int i = 0; i < 256; i++) acc = fmaf(__ldg(P.f2wT + i * 64 + o), fc1o[s * 256 + i], acc);
      fc2o[s * 64 + o] = acc > 0.f ? acc : 0.f;
    }
    __syncthreads();
    // fc3 64->16 (use_model.py:40,57)
    if (threadIdx.x < 64) {
      const int s = threadIdx.x >> 4, o = threadIdx.x & 15;
      float acc = __ldg(P.f3b + o);
      for (int i = 0; i < 64; i++) acc = fmaf(__ldg(P.f3wT + i * 16 + o), fc2o[s * 64 + i], acc);
      lgt[s * 16 + o] = acc;
      if (logits_out) logits_out[(size_t)ctu * 64 + s * 16 + o] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint8_t lab[16];
      logits_to_labels(lgt, lab, ctu_x, ctu_y, geo.W, geo.H, boundary_fix);
      uint4 pk;
      uint32_t *pw = reinterpret_cast<uint32_t *>(&pk);
      for (int i = 0; i < 4; i++) pw[i] = lab[4 * i] | (lab[4 * i + 1] << 8) | (lab[4 * i + 2] << 16) | (lab[4 * i + 3] << 24);
      *reinterpret_cast<uint4 *>(labels + (size_t)ctu * 16) = pk;
      if (ctu_cnt) ctu_cnt[ctu] = ctu_plan_counts(lab, ctu_x, ctu_y, geo.W, geo.H);
    }
    __syncthreads();
  }
}

}  // namespace hevcdl
