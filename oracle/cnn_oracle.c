/* oracle/cnn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, double accumulation) of the depth-prediction path of
 * the reference sidecar.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  Each function cites the
 * reference lines it follows (paths relative to /root/reference).
 *
 * Parity status: PINNED against the reference's own ConvNet2 class (use_model.py:16-58,
 * train-mode BatchNorm, batch 1) run under torch on seeded inputs -- fixtures
 * tests/golden/cnn_golden.npz made by tools/gen_golden.py.  The JPEG stage of
 * gen_frames.py:21 is NOT reproduced (ffmpeg absent, settings unpinned): the CNN input is
 * defined by oracle_stage_ctu_rgb() below (see DESIGN.md "Input definition").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HDLW_NFLOATS 637264

/* offsets (in floats) into the flat HDLW weight blob, order fixed by tools/convert_weights.py */
enum {
  O_C1W = 0,                 O_C1B = O_C1W + 16*3*25,  O_BN1G = O_C1B + 16,  O_BN1B = O_BN1G + 16,
  O_C64W = O_BN1B + 16,      O_C64B = O_C64W + 16*3*25, O_BN64G = O_C64B + 16, O_BN64B = O_BN64G + 16,
  O_C2W = O_BN64B + 16,      O_C2B = O_C2W + 64*32*9,  O_BN2G = O_C2B + 64,  O_BN2B = O_BN2G + 64,
  O_C3W = O_BN2B + 64,       O_C3B = O_C3W + 128*64*9, O_BN3G = O_C3B + 128, O_BN3B = O_BN3G + 128,
  O_F1W = O_BN3B + 128,      O_F1B = O_F1W + 256*2048,
  O_F2W = O_F1B + 256,       O_F2B = O_F2W + 64*256,
  O_F3W = O_F2B + 64,        O_F3B = O_F3W + 16*64,
  O_END = O_F3B + 16
};

int oracle_hdlw_nfloats(void) { return O_END; }

/* ---- K0 definition: CTU tile staging + YUV420 -> RGB ---------------------------------------
 * Replaces gen_frames.py:21 (ffmpeg yuv420p -> jpg) + use_model.py:78,92-95 (PIL crop, zero pad
 * outside the picture, ToTensor).  BT.601 limited-range -> full-range RGB, nearest (co-sited
 * 2x2) chroma, 16.16 fixed point, arithmetic >> (floor), clip to [0,255].
 * out: rgb[3][64][64] u8 (R,G,B planes) for CTU (ctu_x, ctu_y); samples outside WxH are 0. */
static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

void oracle_yuv2rgb_px(int y, int cb, int cr, uint8_t *r, uint8_t *g, uint8_t *b)
{
  int c = 76309 * (y - 16), d = cb - 128, e = cr - 128;
  *r = (uint8_t)clip255((c + 104597 * e + 32768) >> 16);
  *g = (uint8_t)clip255((c - 25675 * d - 53279 * e + 32768) >> 16);
  *b = (uint8_t)clip255((c + 132201 * d + 32768) >> 16);
}

void oracle_stage_ctu_rgb(const uint8_t *Y, const uint8_t *U, const uint8_t *V, int W, int H,
                          int ctu_x, int ctu_y, uint8_t *rgb /* [3][64][64] */)
{
  int cw = W / 2;
  for (int y = 0; y < 64; y++)
    for (int x = 0; x < 64; x++) {
      int px = ctu_x * 64 + x, py = ctu_y * 64 + y;
      uint8_t r = 0, g = 0, b = 0;
      if (px < W && py < H)
        oracle_yuv2rgb_px(Y[py * W + px], U[(py / 2) * cw + px / 2], V[(py / 2) * cw + px / 2], &r, &g, &b);
      rgb[0 * 4096 + y * 64 + x] = r;
      rgb[1 * 4096 + y * 64 + x] = g;
      rgb[2 * 4096 + y * 64 + x] = b;
    }
}

/* ---- layers (use_model.py:16-58) ------------------------------------------------------------ */
/* Conv2d(cin,cout,k,padding=k/2) + BatchNorm2d in TRAINING mode on one sample (use_model.py never
 * calls .eval(): statistics are this sample's per-channel mean / biased variance, eps 1e-5)
 * + ReLU + MaxPool2d(pool).  in: [cin][s][s] float, out: [cout][s/pool][s/pool] float. */
static void conv_bn_relu_pool(const float *in, int cin, int s, const float *w, const float *bias,
                              const float *gamma, const float *beta, int cout, int k, int pool,
                              float *out)
{
  int pad = k / 2, so = s / pool;
  float *tmp = (float *)malloc(sizeof(float) * s * s);
  for (int co = 0; co < cout; co++) {
    double sum = 0.0;
    for (int y = 0; y < s; y++)
      for (int x = 0; x < s; x++) {
        double acc = 0.0;
        for (int ci = 0; ci < cin; ci++)
          for (int ky = 0; ky < k; ky++) {
            int iy = y + ky - pad;
            if (iy < 0 || iy >= s) continue;
            for (int kx = 0; kx < k; kx++) {
              int ix = x + kx - pad;
              if (ix < 0 || ix >= s) continue;
              acc += (double)in[(ci * s + iy) * s + ix] * (double)w[((co * cin + ci) * k + ky) * k + kx];
            }
          }
        float v = (float)(acc + (double)bias[co]);
        tmp[y * s + x] = v;
        sum += v;
      }
    double mean = sum / (s * s), var = 0.0;
    for (int i = 0; i < s * s; i++) { double d = tmp[i] - mean; var += d * d; }
    var /= (s * s);                                   /* biased, as F.batch_norm(training=True) */
    double inv = 1.0 / sqrt(var + 1e-5);
    for (int i = 0; i < s * s; i++) {
      float v = (float)((tmp[i] - mean) * inv * (double)gamma[co] + (double)beta[co]);
      tmp[i] = v > 0.f ? v : 0.f;
    }
    for (int y = 0; y < so; y++)
      for (int x = 0; x < so; x++) {
        float m = tmp[(y * pool) * s + x * pool];
        for (int py = 0; py < pool; py++)
          for (int px = 0; px < pool; px++) {
            float v = tmp[(y * pool + py) * s + x * pool + px];
            if (v > m) m = v;
          }
        out[(co * so + y) * so + x] = m;
      }
  }
  free(tmp);
}

static void linear(const float *in, int nin, const float *w, const float *b, int nout, int relu, float *out)
{
  for (int o = 0; o < nout; o++) {
    double acc = 0.0;
    for (int i = 0; i < nin; i++) acc += (double)in[i] * (double)w[o * nin + i];
    float v = (float)(acc + (double)b[o]);
    out[o] = (relu && v < 0.f) ? 0.f : v;
  }
}

/* ConvNet2.forward(x32, x64) for one sample (use_model.py:48-58).  Inputs are u8 RGB planes,
 * converted as torchvision ToTensor does (float32 value / 255).  logits[16]. */
void oracle_convnet2_forward(const float *wts, const uint8_t *x32 /*[3][32][32]*/,
                             const uint8_t *x64 /*[3][64][64]*/, float *logits)
{
  float *f32 = (float *)malloc(sizeof(float) * 3 * 32 * 32);
  float *f64 = (float *)malloc(sizeof(float) * 3 * 64 * 64);
  float *cat = (float *)malloc(sizeof(float) * 32 * 16 * 16);
  float *a2 = (float *)malloc(sizeof(float) * 64 * 8 * 8);
  float *a3 = (float *)malloc(sizeof(float) * 128 * 4 * 4);
  float h1[256], h2[64];
  for (int i = 0; i < 3 * 32 * 32; i++) f32[i] = (float)x32[i] / 255.0f;
  for (int i = 0; i < 3 * 64 * 64; i++) f64[i] = (float)x64[i] / 255.0f;
  /* torch.cat([conv1(x32), conv64(x64)], dim=1): conv1 channels first (use_model.py:50) */
  conv_bn_relu_pool(f32, 3, 32, wts + O_C1W, wts + O_C1B, wts + O_BN1G, wts + O_BN1B, 16, 5, 2, cat);
  conv_bn_relu_pool(f64, 3, 64, wts + O_C64W, wts + O_C64B, wts + O_BN64G, wts + O_BN64B, 16, 5, 4, cat + 16 * 256);
  conv_bn_relu_pool(cat, 32, 16, wts + O_C2W, wts + O_C2B, wts + O_BN2G, wts + O_BN2B, 64, 3, 2, a2);
  conv_bn_relu_pool(a2, 64, 8, wts + O_C3W, wts + O_C3B, wts + O_BN3G, wts + O_BN3B, 128, 3, 2, a3);
  linear(a3, 2048, wts + O_F1W, wts + O_F1B, 256, 1, h1);   /* view(in_size,-1): C,H,W order */
  linear(h1, 256, wts + O_F2W, wts + O_F2B, 64, 1, h2);
  linear(h2, 64, wts + O_F3W, wts + O_F3B, 16, 0, logits);
  free(f32); free(f64); free(cat); free(a2); free(a3);
}

/* ---- logits -> 16 labels (use_model.py:101-119) ---------------------------------------------
 * logits4: the four quadrant forwards of one CTU, [4][16].  torch.argmax = first maximum.
 * margins (optional, [16]): top-1 minus top-2 logit of each 4-way argmax, in label order. */
void oracle_ctu_labels(const float *logits4, uint8_t *label /*[16]*/, float *margins)
{
  static const int scatter[4][4] = {{0, 1, 4, 5}, {2, 3, 6, 7}, {8, 9, 12, 13}, {10, 11, 14, 15}};
  for (int q = 0; q < 4; q++) {
    int pred[4];
    for (int g = 0; g < 4; g++) {
      const float *l = logits4 + q * 16 + g * 4;
      int best = 0;
      for (int i = 1; i < 4; i++) if (l[i] > l[best]) best = i;
      pred[g] = best;
      if (margins) {
        float second = -INFINITY;
        for (int i = 0; i < 4; i++) if (i != best && l[i] > second) second = l[i];
        margins[scatter[q][g]] = l[best] - second;
      }
    }
    int has0 = 0, all0 = 1, has1 = 0, all1 = 1;
    for (int g = 0; g < 4; g++) { has0 |= pred[g] == 0; all0 &= pred[g] == 0; }
    if (has0 && !all0) for (int g = 0; g < 4; g++) if (pred[g] == 0) pred[g] = 1;      /* :102-103 */
    for (int g = 0; g < 4; g++) { has1 |= pred[g] == 1; all1 &= pred[g] == 1; }
    if (has1 && !all1) for (int g = 0; g < 4; g++) if (pred[g] == 1) pred[g] = 2;      /* :104-105 */
    all0 = pred[0] == 0 && pred[1] == 0 && pred[2] == 0 && pred[3] == 0;
    if (q == 1 && all0 && label[0] != 0) pred[0] = pred[1] = pred[2] = pred[3] = 1;    /* :109-110 */
    if (q == 2 && all0 && label[2] != 0) pred[0] = pred[1] = pred[2] = pred[3] = 1;    /* :113-114 */
    if (q == 3 && all0 && label[8] != 0) pred[0] = pred[1] = pred[2] = pred[3] = 1;    /* :117-118 */
    for (int g = 0; g < 4; g++) label[scatter[q][g]] = (uint8_t)pred[g];
  }
}

/* Whole frame: for every CTU in raster order (use_model.py:80-86) stage RGB, run the four
 * quadrant forwards, reduce to labels.  labels [nctu][16]; logits_out (optional) [nctu][4][16];
 * margins_out (optional) [nctu][16].  ctu_begin/ctu_end bound the work (bench samples). */
void oracle_frame_labels(const float *wts, const uint8_t *Y, const uint8_t *U, const uint8_t *V,
                         int W, int H, int ctu_begin, int ctu_end, uint8_t *labels,
                         float *logits_out, float *margins_out)
{
  int cw = (W + 63) / 64;
#pragma omp parallel for schedule(dynamic, 1)
  for (int a = ctu_begin; a < ctu_end; a++) {
    uint8_t *rgb = (uint8_t *)malloc(3 * 4096), *x32 = (uint8_t *)malloc(3 * 1024);
    float lg[4][16];
    oracle_stage_ctu_rgb(Y, U, V, W, H, a % cw, a / cw, rgb);
    for (int q = 0; q < 4; q++) {
      int ox = (q % 2) * 32, oy = (q / 2) * 32;          /* use_model.py:90-91 */
      for (int c = 0; c < 3; c++)
        for (int y = 0; y < 32; y++)
          memcpy(x32 + (c * 32 + y) * 32, rgb + c * 4096 + (oy + y) * 64 + ox, 32);
      oracle_convnet2_forward(wts, x32, rgb, lg[q]);
    }
    oracle_ctu_labels(&lg[0][0], labels + (size_t)a * 16, margins_out ? margins_out + (size_t)a * 16 : 0);
    if (logits_out) memcpy(logits_out + (size_t)a * 64, lg, sizeof(lg));
    free(rgb); free(x32);
  }
}
