// common.cuh -- shared definitions for libhevcdl.so (sm_100a only).
#pragma once
#include <cuda.h>           // CUtensorMap (type only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "../../include/hevcdl.h"

namespace hevcdl {

// Offsets (floats) into the HDLW blob; order fixed by tools/convert_weights.py
// (tensor names of rec/hevc_encoder_model.pt as loaded at use_model.py:62).
enum : int {
  O_C1W = 0, O_C1B = O_C1W + 16 * 3 * 25, O_BN1G = O_C1B + 16, O_BN1B = O_BN1G + 16,
  O_C64W = O_BN1B + 16, O_C64B = O_C64W + 16 * 3 * 25, O_BN64G = O_C64B + 16, O_BN64B = O_BN64G + 16,
  O_C2W = O_BN64B + 16, O_C2B = O_C2W + 64 * 32 * 9, O_BN2G = O_C2B + 64, O_BN2B = O_BN2G + 64,
  O_C3W = O_BN2B + 64, O_C3B = O_C3W + 128 * 64 * 9, O_BN3G = O_C3B + 128, O_BN3B = O_BN3G + 128,
  O_F1W = O_BN3B + 128, O_F1B = O_F1W + 256 * 2048,
  O_F2W = O_F1B + 256, O_F2B = O_F2W + 64 * 256,
  O_F3W = O_F2B + 64, O_F3B = O_F3W + 16 * 64,
  HDLW_NFLOATS = O_F3B + 16
};

// Device-side packed parameters of the fp32 path (built once in hevcdl_create).
struct Fp32Params {
  const float *c1w, *c64w, *c2w, *c3w;   // [COUT/CG][CIN][K*K][CG]
  const float *c1b, *c64b, *c2b, *c3b;   // conv biases
  const float *g1, *b1, *g64, *b64, *g2, *b2, *g3, *b3;  // BN gamma/beta
  const float *f1wT, *f1b;               // [2048][256]
  const float *f2wT, *f2b;               // [256][64]
  const float *f3wT, *f3b;               // [64][16]
};

// Programmatic dependent launch (PDL): every kernel of the frame pipeline lets its successor start early
// (pdl_launch_dependents at entry) and runs its own prologue -- barrier init, TMEM allocation, weight loads --
// before pdl_wait(), which returns once the predecessor grid has completed and its writes are visible.  Without
// the launch attribute both are no-ops.
// Tuning builds only (-DHEVCDL_TIMELINE): every CTA logs (kernel, block, entry, predecessor-wait passed, exit) in
// %globaltimer nanoseconds; hevcdl_destroy writes the log to $HEVCDL_TIMELINE_OUT (tools/timeline.py reads it).
#ifdef HEVCDL_TIMELINE
constexpr unsigned TL_CAP = 1u << 18;
__device__ unsigned long long g_tl[TL_CAP][4];
__device__ unsigned g_tl_n;
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TL_BEGIN() __shared__ unsigned long long tl_s[2]; if (threadIdx.x == 0) tl_s[0] = tl_now()
#define TL_WAITED() tl_s[1] = tl_now()
#define TL_END(kid)                                                                   \
  if (threadIdx.x == 0) {                                                              \
    const unsigned i_ = atomicAdd(&g_tl_n, 1u);                                        \
    if (i_ < TL_CAP) { g_tl[i_][0] = ((unsigned long long)(kid) << 32) | blockIdx.x; g_tl[i_][1] = tl_s[0]; g_tl[i_][2] = tl_s[1]; g_tl[i_][3] = tl_now(); } \
  }
#else
#define TL_BEGIN()
#define TL_WAITED()
#define TL_END(kid)
#endif

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct FrameGeom {
  int W, H;        // luma size
  int ctu_w, ctu_h, nctu;
};

// BT.601 limited-range YCbCr -> full-range RGB, 16.16 fixed point (the K0 input definition,
// DESIGN.md; replaces gen_frames.py:21 + PIL decode at use_model.py:78).
__host__ __device__ __forceinline__ int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
__host__ __device__ __forceinline__ void yuv2rgb(int y, int cb, int cr, int &r, int &g, int &b) {
  int c = 76309 * (y - 16), d = cb - 128, e = cr - 128;
  r = clip255((c + 104597 * e + 32768) >> 16);
  g = clip255((c - 25675 * d - 53279 * e + 32768) >> 16);
  b = clip255((c + 132201 * d + 32768) >> 16);
}

// Frames of one CNN launch.  All-intra frames are independent, so several frames of the same geometry can share a
// launch: 510 CTUs of one 1080p frame leave 148 persistent CTAs with 3 or 4 CTUs each (86 % balance) and the fc
// kernel with 64 CTAs; two frames give 6.9 -> 7 (98 %) and 128 CTAs.  Global CTU index g = frame * nctu + ctu.
constexpr int MAX_BATCH = 8;
struct FrameBatch {
  const uint8_t *Y[MAX_BATCH], *U[MAX_BATCH], *V[MAX_BATCH];
  uint8_t *labels[MAX_BATCH];
  float *logits[MAX_BATCH];
  uint32_t *ctu_cnt[MAX_BATCH];   // may be null (rmd = 0)
  int n;
};

// TMA descriptors of the frames of a launch (K1 stages every CTU's 64x64 Y tile and 32x32 Cb / Cr tiles with
// cp.async.bulk.tensor; samples outside the picture arrive as zeros).  Passed as a __grid_constant__ kernel parameter.
struct TmapBatch {
  CUtensorMap y[MAX_BATCH], u[MAX_BATCH], v[MAX_BATCH];
};

// The same for the RMD kernels: one plan + one items launch serve all frames of a batch (one work-item queue).
struct RmdBatch {
  const uint8_t *Y[MAX_BATCH];
  const uint8_t *labels[MAX_BATCH];
  const uint32_t *ctu_cnt[MAX_BATCH];
  int *ctu_off[MAX_BATCH];
  hevcdl_pu *pus[MAX_BATCH];
  uint32_t *satd[MAX_BATCH];
  uint8_t *cand[MAX_BATCH];
  int n;
};

// Logits of the 4 quadrant forwards -> 16 labels (use_model.py:101-119), optional boundary fix.
// lg: [4][16].  Runs in one thread.
__device__ __forceinline__ void logits_to_labels(const float *lg, uint8_t *label, int ctu_x, int ctu_y,
                                                 int W, int H, int boundary_fix) {
  const int scatter[4][4] = {{0, 1, 4, 5}, {2, 3, 6, 7}, {8, 9, 12, 13}, {10, 11, 14, 15}};
  for (int q = 0; q < 4; q++) {
    int pred[4];
    for (int g = 0; g < 4; g++) {
      const float *l = lg + q * 16 + g * 4;
      int best = 0;  // torch.argmax: first maximum
      for (int i = 1; i < 4; i++)
        if (l[i] > l[best]) best = i;
      pred[g] = best;
    }
    bool has0 = false, all0 = true;
    for (int g = 0; g < 4; g++) { has0 |= pred[g] == 0; all0 &= pred[g] == 0; }
    if (has0 && !all0)
      for (int g = 0; g < 4; g++) if (pred[g] == 0) pred[g] = 1;
    bool has1 = false, all1 = true;
    for (int g = 0; g < 4; g++) { has1 |= pred[g] == 1; all1 &= pred[g] == 1; }
    if (has1 && !all1)
      for (int g = 0; g < 4; g++) if (pred[g] == 1) pred[g] = 2;
    all0 = pred[0] == 0 && pred[1] == 0 && pred[2] == 0 && pred[3] == 0;
    int prev = q == 1 ? label[0] : q == 2 ? label[2] : q == 3 ? label[8] : 0;
    if (q > 0 && all0 && prev != 0) pred[0] = pred[1] = pred[2] = pred[3] = 1;
    for (int g = 0; g < 4; g++) label[scatter[q][g]] = (uint8_t)pred[g];
  }
  if (boundary_fix) {
    // make partial CTUs tile -- a CU crossing the picture edge must be split.
    int x0 = ctu_x * 64, y0 = ctu_y * 64;
    if (x0 + 64 > W || y0 + 64 > H) {
      for (int i = 0; i < 16; i++) {
        int bx = x0 + (i & 3) * 16, by = y0 + (i >> 2) * 16;
        int qx = x0 + ((i & 3) >> 1) * 32, qy = y0 + ((i >> 2) >> 1) * 32;
        int l = label[i] < 1 ? 1 : label[i];
        if (qx + 32 > W || qy + 32 > H) l = l < 2 ? 2 : l;
        if (bx + 16 > W || by + 16 > H) l = 3;
        label[i] = (uint8_t)l;
      }
    }
  }
}

// PU and work-item counts of one CTU, closed form of the walk in k_rmd_plan (rmd.cuh) (one thread; called by the
// label kernels).  Returns npu | nbig << 16 | nsmall << 20 (big items: 32x32 PUs and 64x64 quadrants;
// small items: 16x16 PUs and 8x8 CUs).
__device__ inline uint32_t ctu_plan_counts(const uint8_t *lab, int ctu_x, int ctu_y, int W, int H) {
  const int x0 = ctu_x * 64, y0 = ctu_y * 64;
  if (lab[0] == 0) return (x0 + 64 <= W && y0 + 64 <= H) ? (1u | (4u << 16)) : 0u;
  uint32_t npu = 0, nbig = 0, nsm = 0;
  for (int q = 0; q < 4; q++) {
    const int qx = x0 + (q & 1) * 32, qy = y0 + (q >> 1) * 32;
    if (qx >= W || qy >= H) continue;
    const int lq = lab[8 * (q >> 1) + 2 * (q & 1)];
    if (lq == 1) { if (qx + 32 <= W && qy + 32 <= H) { npu += 1; nbig += 1; } continue; }
    if (lq < 1) continue;
    for (int s = 0; s < 4; s++) {
      const int bx = qx + (s & 1) * 16, by = qy + (s >> 1) * 16;
      if (bx >= W || by >= H) continue;
      const int l = lab[4 * ((by & 63) >> 4) + ((bx & 63) >> 4)];
      if (l == 2) { if (bx + 16 <= W && by + 16 <= H) { npu += 1; nsm += 1; } continue; }
      if (l < 2) continue;
      for (int e = 0; e < 4; e++) {
        const int ex = bx + (e & 1) * 8, ey = by + (e >> 1) * 8;
        if (ex >= W || ey >= H) continue;
        if (l == 3) { npu += 5; nsm += 1; }       // W, H are multiples of 8: an 8x8 CU never straddles the edge
      }
    }
  }
  return npu | (nbig << 16) | (nsm << 20);
}

}  // namespace hevcdl
