/* include/hevcdl.h -- C ABI of libhevcdl.so, the B200-native (sm_100a) CNN-gated intra
 * CU-partition hot path.  Plain pointers and sizes only; no C++/torch types.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * reference checkout; HM = HM_dl/source/Lib).  The reference-side binding (the replacement
 * TEncCu::compressCtu) is in hm_plugin/ and described in INTEGRATION.md.
 *
 * Threading: a context is used from one encoder thread (as the reference's single CTU loop,
 * HM TLibEncoder/TEncSlice.cpp:792).  All functions return 0 or a negative hevcdl_status.
 * There is NO CPU fallback: without a CUDA device hevcdl_create fails with HEVCDL_E_NODEVICE.
 */
#ifndef HEVCDL_H
#define HEVCDL_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HEVCDL_ABI_VERSION 2

typedef enum {
  HEVCDL_OK = 0,
  HEVCDL_E_INVAL = -1,     /* bad argument */
  HEVCDL_E_NODEVICE = -2,  /* no CUDA device / wrong architecture (needs sm_100) */
  HEVCDL_E_CUDA = -3,      /* CUDA runtime error (hevcdl_last_error has the text) */
  HEVCDL_E_WEIGHTS = -4,   /* weight blob missing or malformed */
  HEVCDL_E_NOFRAME = -5,   /* frame id unknown / already released */
  HEVCDL_E_BUSY = -6,      /* no free frame slot: release or wait for an older frame */
  HEVCDL_E_NOMEM = -7
} hevcdl_status;

typedef enum {
  HEVCDL_PREC_FP32 = 0,    /* CUDA-core fp32 CNN (tightest parity with the torch fp32 oracle) */
  HEVCDL_PREC_BF16_TC = 1  /* tcgen05 tensor-core CNN, bf16 operands, fp32 accumulate + fp32 BN */
} hevcdl_precision;

typedef struct hevcdl_ctx hevcdl_ctx;

/* Replaces the sidecar's implicit configuration: bitstream.cfg parsed by line number
 * (use_model.py:65-71, gen_frames.py:4-16), DEVICE selection (use_model.py:60) and
 * torch.load of rec/hevc_encoder_model.pt (use_model.py:62). */
typedef struct {
  int32_t abi_version;     /* HEVCDL_ABI_VERSION */
  int32_t device;          /* CUDA ordinal */
  int32_t width, height;   /* luma samples, multiples of 8 (HM TAppEncCfg.cpp:2176) */
  int32_t slots;           /* frames in flight (>=1) */
  int32_t precision;       /* hevcdl_precision */
  int32_t rmd;             /* 1: run the batched 35-mode SATD pass (K6) after the labels */
  int32_t boundary_fix;    /* 1: raise labels of picture-edge CTUs so partial CTUs tile (the
                              reference leaves them inconsistent); 0 = reference */
  int32_t batch;           /* frames per CNN launch, 1..8 (0 = 1).  All-intra frames are independent; with batch > 1 the
                              kernels of a frame start once `batch` frames have been submitted or when a pending frame is
                              asked for, whichever comes first.  Results do not depend on it.  Tensor-core path only. */
  const char *weights_path;/* HDLW blob made by tools/convert_weights.py */
} hevcdl_cfg;

/* One prediction unit of the pruned quadtree (HM TLibEncoder/TEncCu.cpp:496-520,815-834). */
typedef struct {
  uint16_t x, y;           /* luma position in the picture */
  uint8_t size;            /* 64,32,16,8 (2Nx2N) or 4 (one PU of the NxN trial of an 8x8 CU) */
  uint8_t part;            /* 0 = 2Nx2N, 1..4 = NxN PU index + 1 */
  uint16_t ctu;            /* raster CTU address (HM getCtuRsAddr) */
} hevcdl_pu;

typedef struct {
  uint64_t frames, ctus, pus;
  double ms_cnn, ms_rmd;   /* device time (CUDA events) accumulated over frames; recorded with batch == 1, or with any
                              batch when HEVCDL_STAGE_TIMES is set in the environment (the events cost throughput) */
  uint64_t kernel_launches;
} hevcdl_stats_t;

int hevcdl_create(const hevcdl_cfg *cfg, hevcdl_ctx **out);
void hevcdl_destroy(hevcdl_ctx *ctx);
const char *hevcdl_last_error(const hevcdl_ctx *ctx);   /* text of the last failure */
const char *hevcdl_status_str(int status);

/* Replaces gen_frames.py:21 (ffmpeg frame dump) + the sidecar's per-frame loop
 * (use_model.py:74-127): hand one picture to the device; returns immediately after queueing
 * H2D + kernels + D2H on the context's stream.  `frame` is the id HM threads through
 * compressCtu (m_iFrame, HM TLibEncoder/TEncCu.cpp:234; TAppEncTop.cpp:634).
 * u8: planar 8-bit 4:2:0.  pel16: HM's Pel (int16) planes as held by TComPicYuv
 * (HM TLibCommon/TComPicYuv.h), 8-bit content. */
int hevcdl_submit_frame_u8(hevcdl_ctx *ctx, int frame, const uint8_t *y, int stride_y,
                           const uint8_t *u, const uint8_t *v, int stride_c);
int hevcdl_submit_frame_pel16(hevcdl_ctx *ctx, int frame, const int16_t *y, int stride_y,
                              const int16_t *u, const int16_t *v, int stride_c);

/* Replaces the busy-poll on ./pred/<frame>/ctu<addr>.txt (HM TEncCu.cpp:244-245). */
int hevcdl_wait_frame(hevcdl_ctx *ctx, int frame);

/* Replaces reading the 16 labels of one CTU from its text file (HM TEncCu.cpp:246-253;
 * written at use_model.py:121-125).  Blocks until the frame is done. */
int hevcdl_ctu_labels(hevcdl_ctx *ctx, int frame, int ctu_rs_addr, uint8_t out[16]);
/* Whole frame: labels [nctu*16]; logits (optional) [nctu*4*16] float, the CNN outputs before
 * argmax (use_model.py:100), for margin reporting. */
int hevcdl_frame_labels(hevcdl_ctx *ctx, int frame, uint8_t *labels, float *logits);

/* Batched original-reference RMD results (K6), replacing the first pass of
 * TEncSearch::estIntraPredLumaQT (HM TLibEncoder/TEncSearch.cpp:2266-2320) for every PU of the
 * frame at once.  PUs are ordered as the encoder visits them (CTU raster, z-order inside).
 * satd [npu*35] (mode-major per PU); cand [npu*8]: modes ranked by SATD, ties to the lower
 * mode (first 3 valid for size>=16, 8 otherwise).  Any output pointer may be NULL. */
int hevcdl_frame_pu_count(hevcdl_ctx *ctx, int frame, int *npu);
int hevcdl_frame_pus(hevcdl_ctx *ctx, int frame, hevcdl_pu *pus, uint32_t *satd, uint8_t *cand);
int hevcdl_ctu_pu_range(hevcdl_ctx *ctx, int frame, int ctu_rs_addr, int *first, int *count);

/* Zero-copy access to everything the device produced for one frame: pointers into the context's pinned host
 * buffers (the same bytes the copying getters above return), valid until hevcdl_release_frame(frame).  Blocks
 * until the frame is done; with want_pus != 0 also until the PU lists have arrived (rmd=1 contexts only,
 * otherwise pus/satd/cand are NULL and npu = 0).  This is what the sidecar's consumer would read instead of
 * parsing ./pred/<frame>/ctu<addr>.txt one file at a time (HM TEncCu.cpp:244-253). */
typedef struct {
  const uint8_t *labels;   /* [nctu*16] */
  const float *logits;     /* [nctu*4*16] */
  const int32_t *ctu_off;  /* [nctu+1] first PU of each CTU (NULL when rmd=0) */
  const hevcdl_pu *pus;    /* [npu] */
  const uint32_t *satd;    /* [npu*35] */
  const uint8_t *cand;     /* [npu*8] */
  int32_t nctu, npu;
} hevcdl_frame_view;
int hevcdl_frame_view_get(hevcdl_ctx *ctx, int frame, int want_pus, hevcdl_frame_view *out);

int hevcdl_release_frame(hevcdl_ctx *ctx, int frame);

/* Exact RMD for explicit inputs, the same device code as K6 fed what the reference feeds its
 * first pass: per PU the original block, the 4n+1 reference line built from RECONSTRUCTED
 * neighbours (HM TLibCommon/TComPattern.cpp:119-543; order: below-left..left (bottom-up),
 * corner, above..above-right) and the 35 mode-bit counts of xModeBitsIntra (HM TEncSearch.cpp:5530).
 * Outputs: satd [n*35]; cand [n*10] + ncand [n]: the reference's uiRdModeList (TEncSearch.cpp:
 * 2313-2345) -- cost = satd + bits*sqrt_lambda in double, strict '<', then the first mpm_add[i]
 * of mpm[i*3..] not yet listed.  org: concatenated size*size u8 blocks; lines: concatenated
 * (4*size+1) int16.  Synchronous. */
int hevcdl_rmd_exact(hevcdl_ctx *ctx, int n, const uint8_t *sizes, const uint8_t *org,
                     const int16_t *lines, const uint32_t *bits, const int8_t *mpm,
                     const uint8_t *mpm_add, double sqrt_lambda, uint32_t *satd, uint8_t *cand,
                     uint8_t *ncand);

/* Measurement: run the device pipeline `iters` times over planes already resident in the slots of
 * frames[0..nframes) (round-robin; no H2D/D2H), timed with CUDA events on the context's stream.
 * ms[0] = total of one pass timed by a single event pair; ms[1], ms[2] = CNN (K0-K5) and RMD
 * (enumeration + K6) stage totals from a second pass with per-stage events.  launches: kernels
 * launched per pass. */
int hevcdl_bench_resident(hevcdl_ctx *ctx, const int *frames, int nframes, int iters, float ms[3],
                          int *launches);

/* Measurement, end to end: `iters` frames through the public calls above -- hevcdl_submit_frame_u8 from the HOST
 * planes y/u/v[i % nbuf] (pinned or pageable, caller-owned), hevcdl_frame_view_get(want_pus) on the oldest frame
 * once `depth` frames are in flight, hevcdl_release_frame -- timed with the host's steady clock from the first
 * submit to the last view.  Frame ids first_id .. first_id+iters-1.  seconds: wall time; d2h_bytes: bytes of the
 * views read; checksum: a value folded from every view so the reads cannot be elided. */
int hevcdl_bench_e2e(hevcdl_ctx *ctx, int first_id, int iters, int depth, int nbuf, const uint8_t *const *y,
                     const uint8_t *const *u, const uint8_t *const *v, int stride_y, int stride_c, double *seconds,
                     uint64_t *d2h_bytes, uint64_t *checksum);

/* Test hook (tensor-core path only): copy one L2-resident intermediate of the most recent frame to
 * the host -- which = 0: conv1/conv64 output planes ("cat"), 1: conv2 output planes, 2: conv3
 * features in the fc1 operand layout.  *size receives the byte size; dst may be NULL to query it. */
int hevcdl_debug_copy(hevcdl_ctx *ctx, int which, void *dst, size_t nbytes, size_t *size);

/* Test hook: run the RMD pass (K6) of a finished frame again with caller-supplied labels [nctu*16] instead of the CNN's --
 * the reference's own interface hands labels over as files (use_model.py:121-125), and the CNN never predicts some
 * cases (64x64 CUs on ordinary content) that K6 must still handle.  Synchronous; afterwards the frame's label and PU
 * getters return the new labels and their PU lists. */
int hevcdl_debug_rerun_rmd(hevcdl_ctx *ctx, int frame, const uint8_t *labels);

int hevcdl_get_stats(hevcdl_ctx *ctx, hevcdl_stats_t *out);
/* CUDA stream handle (cudaStream_t) of the context, for callers that time with their own events */
void *hevcdl_stream(hevcdl_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* HEVCDL_H */
