/* oracle/sao_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's SAO statistics pass: TEncSampleAdaptiveOffset::getStatistics
 * (HM_dl/source/Lib/TLibEncoder/TEncSampleAdaptiveOffset.cpp:295-341) and getBlkStats (:943-1345) for deblocked samples
 * (isCalculatePreDeblockSamples = false; SAOLcuBoundary 0, encoder_intra_main.cfg:55), one slice, no tiles, 8-bit 4:2:0,
 * 64x64 CTUs.  For every CTU, colour component and SAO type it accumulates, per class, the number of samples and the sum of
 * (original - deblocked): the four edge-offset types classify a sample by the signs of its differences to the two
 * neighbours along the direction (class index 2 + sgn + sgn), band offset by the sample's upper five bits.  The reference
 * walks lines with running sign buffers; the classes do not depend on that, so they are written per sample here.  What a CTU
 * counts excludes the columns / rows its right / lower neighbours have not deblocked yet (skipLinesR = 5 luma / 3 chroma,
 * skipLinesB = 4 / 2, :126-133) and, for the edge types, samples whose neighbour lies outside the picture.
 * Pinned by tests/golden/sao_stats.npz: inputs and output of the reference's own getStatistics (TAppEncoder_saotrace).
 *
 * out: [nctu][3][5] x { int64 diff[32], int64 count[32] } exactly as the reference's SAOStatData arrays.
 */
#include <stdint.h>
#include <string.h>

static int sgn(int v) { return (v > 0) - (v < 0); }

void oracle_sao_stats(const int16_t *org[3], const int16_t *src[3], int W, int H, int64_t *out) {
  const int cw = (W + 63) / 64, ch = (H + 63) / 64;
  memset(out, 0, sizeof(int64_t) * (size_t)cw * ch * 3 * 5 * 64);
  for (int a = 0; a < cw * ch; a++) {
    const int xp = (a % cw) * 64, yp = (a / cw) * 64;
    const int hl = yp + 64 > H ? H - yp : 64, wl = xp + 64 > W ? W - xp : 64;
    const int left = xp > 0, above = yp > 0, right = xp + 64 < W, below = yp + 64 < H, above_left = left && above;
    for (int c = 0; c < 3; c++) {
      const int sh = c ? 1 : 0, stride = W >> sh, width = wl >> sh, height = hl >> sh;
      const int16_t *s = src[c] + (size_t)(yp >> sh) * stride + (xp >> sh), *o = org[c] + (size_t)(yp >> sh) * stride + (xp >> sh);
      const int skip_r = c ? 3 : 5, skip_b = c ? 2 : 4;
      for (int t = 0; t < 5; t++) {
        int64_t *diff = out + (((size_t)a * 3 + c) * 5 + t) * 64, *count = diff + 32;
        const int x0 = (t == 1 || t == 4) ? 0 : (left ? 0 : 1);
        const int x1 = right ? width - skip_r : ((t == 1 || t == 4) ? width : width - 1);
        const int y0 = (t == 1) ? (above ? 0 : 1) : 0;
        const int y1 = below ? height - skip_b : ((t == 0 || t == 4) ? height : height - 1);
        for (int y = y0; y < y1; y++) {
          int xa = x0, xb = x1;
          if (y == 0 && t == 2) { xa = above_left ? 0 : 1; xb = above ? x1 : 1; }         /* first line of the 135-degree type */
          if (y == 0 && t == 3) { xa = above ? x0 : x1; }                                 /* ... of the 45-degree type */
          for (int x = xa; x < xb; x++) {
            const int v = s[y * stride + x];
            int cls;
            if (t == 0) cls = 2 + sgn(v - s[y * stride + x - 1]) + sgn(v - s[y * stride + x + 1]);
            else if (t == 1) cls = 2 + sgn(v - s[(y - 1) * stride + x]) + sgn(v - s[(y + 1) * stride + x]);
            else if (t == 2) cls = 2 + sgn(v - s[(y - 1) * stride + x - 1]) + sgn(v - s[(y + 1) * stride + x + 1]);
            else if (t == 3) cls = 2 + sgn(v - s[(y - 1) * stride + x + 1]) + sgn(v - s[(y + 1) * stride + x - 1]);
            else cls = v >> 3;
            diff[cls] += o[y * stride + x] - v;
            count[cls]++;
          }
        }
      }
    }
  }
}

/* SAO application: TComSampleAdaptiveOffset::offsetCTU / offsetBlock (HM_dl/source/Lib/TLibCommon/TComSampleAdaptiveOffset.cpp:
 * 554-611, 313-552) for every CTU of a picture, as the encoder runs it from decideBlkParams (TEncSampleAdaptiveOffset.cpp:894)
 * with the merge candidates already resolved (reconstructBlkSAOParam): res = src + offset[class], clipped to 8 bits, for the
 * samples the type may touch -- edge types leave out samples whose neighbour lies outside the picture (one slice, no tiles:
 * a neighbouring CTU is available iff it exists), with the reference's special first / last line ranges of the diagonal
 * types; every other sample keeps its deblocked value.  All classification reads `src`, never `res`, so CTUs are independent.
 * type[nctu*3]: -1 = off, 0..3 edge offset 0 / 90 / 135 / 45 degrees, 4 band offset; offset[nctu*3][32]: entries 0..4 for the
 * edge classes (index 2 + sgn + sgn), 0..31 for the bands.  Pinned by tests/golden/sao_apply.npz (TAppEncoder_saoapplytrace). */
static int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

void oracle_sao_apply(const int16_t *src[3], int16_t *res[3], int W, int H, const int8_t *type, const int8_t *offset) {
  const int cw = (W + 63) / 64, ch = (H + 63) / 64;
  for (int c = 0; c < 3; c++) memcpy(res[c], src[c], sizeof(int16_t) * (size_t)(W >> (c ? 1 : 0)) * (H >> (c ? 1 : 0)));
  for (int a = 0; a < cw * ch; a++) {
    const int xp = (a % cw) * 64, yp = (a / cw) * 64;
    const int hl = yp + 64 > H ? H - yp : 64, wl = xp + 64 > W ? W - xp : 64;
    const int L = xp > 0, A = yp > 0, R = xp + 64 < W, B = yp + 64 < H, AL = L && A, AR = A && R, BL = B && L, BR = B && R;
    for (int c = 0; c < 3; c++) {
      const int t = type[a * 3 + c];
      if (t < 0) continue;
      const int8_t *of = offset + (size_t)(a * 3 + c) * 32;
      const int sh = c ? 1 : 0, stride = W >> sh, width = wl >> sh, height = hl >> sh;
      const int16_t *s = src[c] + (size_t)(yp >> sh) * stride + (xp >> sh);
      int16_t *r = res[c] + (size_t)(yp >> sh) * stride + (xp >> sh);
      const int sx = L ? 0 : 1, ex = R ? width : width - 1;
      for (int y = 0; y < height; y++) {
        int xa, xb;
        if (t == 4) { xa = 0; xb = width; }
        else if (t == 0) { xa = sx; xb = ex; }
        else if (t == 1) { xa = 0; xb = (y < (A ? 0 : 1) || y >= (B ? height : height - 1)) ? 0 : width; }
        else if (t == 2) {
          if (y == 0) { xa = AL ? 0 : 1; xb = A ? ex : 1; }
          else if (y == height - 1) { xa = B ? sx : width - 1; xb = BR ? width : width - 1; }
          else { xa = sx; xb = ex; }
        } else {
          if (y == 0) { xa = A ? sx : width - 1; xb = AR ? width : width - 1; }
          else if (y == height - 1) { xa = BL ? 0 : 1; xb = B ? ex : 1; }
          else { xa = sx; xb = ex; }
        }
        for (int x = xa; x < xb; x++) {
          const int v = s[y * stride + x];
          int cls;
          if (t == 0) cls = 2 + sgn(v - s[y * stride + x - 1]) + sgn(v - s[y * stride + x + 1]);
          else if (t == 1) cls = 2 + sgn(v - s[(y - 1) * stride + x]) + sgn(v - s[(y + 1) * stride + x]);
          else if (t == 2) cls = 2 + sgn(v - s[(y - 1) * stride + x - 1]) + sgn(v - s[(y + 1) * stride + x + 1]);
          else if (t == 3) cls = 2 + sgn(v - s[(y - 1) * stride + x + 1]) + sgn(v - s[(y + 1) * stride + x - 1]);
          else cls = v >> 3;
          r[y * stride + x] = (int16_t)clip8(v + of[cls]);
        }
      }
    }
  }
}
