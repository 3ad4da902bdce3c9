/* examples/hotpath_min.c -- the smallest host program over the C-ABI (include/hevcdl.h): what a compiled encoder does in place
 * of the reference's sidecar handshake (gen_frames.py:21, use_model.py:74-127, TEncCu.cpp:243-253).  Plain C, no CUDA headers:
 *     gcc -O2 -Iinclude examples/hotpath_min.c -Lhevc-deep-learning-pipeline_b200/csrc -lhevcdl \
 *         -Wl,-rpath,$PWD/hevc-deep-learning-pipeline_b200/csrc -o /tmp/hotpath_min
 *     /tmp/hotpath_min weights/hevc_encoder_model.hdlw 416 240 3
 * Submits N synthetic 4:2:0 frames, reads each CTU's 16 depth labels and the frame's PU list with the ranked first-pass
 * modes, prints a histogram.  Exit code 0 = every call succeeded. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hevcdl.h"

#define CK(call)                                                                                              \
  do {                                                                                                        \
    int rc_ = (call);                                                                                         \
    if (rc_) { fprintf(stderr, "%s -> %s (%s)\n", #call, hevcdl_status_str(rc_), hevcdl_last_error(ctx)); return 1; } \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s weights.hdlw width height [frames]\n", argv[0]); return 2; }
  const int W = atoi(argv[2]), H = atoi(argv[3]), N = argc > 4 ? atoi(argv[4]) : 2;
  hevcdl_ctx *ctx = NULL;
  hevcdl_cfg cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = HEVCDL_ABI_VERSION;
  cfg.width = W; cfg.height = H;
  cfg.slots = N;                          /* frames that may be in flight */
  cfg.precision = HEVCDL_PREC_BF16_TC;    /* tensor cores; HEVCDL_PREC_FP32 = the tight-parity mode */
  cfg.rmd = 1;                            /* also run the 35-mode SATD pass for the surviving PUs */
  cfg.batch = N < 8 ? N : 8;              /* all-intra frames are independent: they share launches */
  cfg.weights_path = argv[1];
  CK(hevcdl_create(&cfg, &ctx));

  uint8_t *y = malloc((size_t)W * H), *u = malloc((size_t)W * H / 4), *v = malloc((size_t)W * H / 4);
  for (int f = 0; f < N; f++) {           /* any 8-bit planar 4:2:0 picture; the planes may be reused once submit returns */
    for (int i = 0; i < W * H; i++) y[i] = (uint8_t)(((i % W) * 3 + (i / W) * 5 + 37 * f) ^ ((i / W / 16) * 29));
    for (int i = 0; i < W * H / 4; i++) { u[i] = (uint8_t)(128 + (i % 23) - f); v[i] = (uint8_t)(120 + (i % 17) + f); }
    CK(hevcdl_submit_frame_u8(ctx, f, y, W, u, v, W / 2));
  }
  const int nctu = ((W + 63) / 64) * ((H + 63) / 64);
  for (int f = 0; f < N; f++) {
    long hist[4] = {0, 0, 0, 0};
    for (int a = 0; a < nctu; a++) {      /* what TEncCu::compressCtu asks per CTU */
      uint8_t lab[16];
      CK(hevcdl_ctu_labels(ctx, f, a, lab));
      for (int i = 0; i < 16; i++) hist[lab[i] & 3]++;
    }
    int npu = 0;
    CK(hevcdl_frame_pu_count(ctx, f, &npu));
    hevcdl_pu *pus = malloc(sizeof(hevcdl_pu) * (size_t)(npu ? npu : 1));
    uint8_t *cand = malloc((size_t)8 * (npu ? npu : 1));
    CK(hevcdl_frame_pus(ctx, f, pus, NULL, cand));
    printf("frame %d: depth labels %ld / %ld / %ld / %ld, %d PUs", f, hist[0], hist[1], hist[2], hist[3], npu);
    if (npu) printf(", first PU %dx%d at (%d,%d): best first-pass modes %d %d %d", pus[0].size, pus[0].size, pus[0].x, pus[0].y, cand[0], cand[1], cand[2]);
    printf("\n");
    free(pus); free(cand);
    CK(hevcdl_release_frame(ctx, f));
  }
  free(y); free(u); free(v);
  hevcdl_destroy(ctx);
  return 0;
}
