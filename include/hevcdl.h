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

#define HEVCDL_ABI_VERSION 3

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

/* hevcdl_cfg.outputs: what besides the 16 labels per CTU (and, with rmd=1, the PU list and the ranked candidate
 * modes) is copied back to the host for every frame.  The 35 x u32 SATD table is 82 % of a frame's device-to-host
 * bytes and only a consumer that re-ranks with its own mode bits (the HM first-pass hook, hm_plugin/rmd_hook.h)
 * needs it; the logits are a diagnostic (argmax margins). */
typedef enum {
  HEVCDL_OUT_LOGITS = 1,   /* CNN outputs before argmax (use_model.py:100) */
  HEVCDL_OUT_SATD = 2      /* per-PU 35-entry SATD table of the RMD pass */
} hevcdl_output_flags;

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
  int32_t outputs;         /* hevcdl_output_flags: optional per-frame results copied to the host (0: labels, PU list and
                              candidate modes only).  Getters of an output that was not asked for return NULL / E_INVAL. */
  int32_t pinned_input;    /* 1: the caller promises that the planes handed to hevcdl_submit_frame_u8 live in page-locked
                              host memory and stay UNTOUCHED until hevcdl_wait_frame / hevcdl_ctu_labels / any getter has
                              returned for that frame: they are read by the copy engine after submit returns (no staging
                              copy).  0 (default): submit copies the planes into the context's own pinned staging buffer
                              before it returns; the caller may reuse its buffers at once. */
  int32_t numa_bind;       /* 1: before allocating, pin the CALLING thread to the CPUs of the device's NUMA node
                              (/sys/bus/pci/devices/<bdf>/numa_node), so that the context's pinned buffers and the staging
                              copies of that thread are local to the GPU's PCIe root (8 ranks copying through one socket's
                              memory is what bounds multi-GPU end-to-end throughput).  The affinity stays in force. */
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

/* Transform-unit coding core (first slice of the RD pass): for a batch of TUs the arithmetic of
 * TComTrQuant::transformNxN + invTransformNxN as TEncSearch::xIntraCodingTUBlock calls them
 * (HM TLibEncoder/TEncSearch.cpp:1129-1424; TLibCommon/TComTrQuant.cpp:1450-1666): forward core transform (DCT 4..32, DST for
 * 4x4 intra luma, or transform skip) -> xQuant's flat quantiser (:1126-1249 without RDOQ and sign-bit hiding) -> xDeQuant
 * (:1308-1423, no scaling lists) -> inverse transform, bit-exact at the reference's operating point (8-bit video, |residual|
 * <= 2047 accepted).  tus[i].offset: element offset of TU i's (1 << log2_size)^2 block, row-major, in resi and in every
 * output array (nelem elements each).  coeff (transform output) and deq (dequantised) may be NULL.  level: quantised
 * coefficients (TCoeff clipped to 16 bits by xQuant); rec: reconstructed residual; abs_sum[i]: uiAbsSum (0 = cbf 0);
 * ssd[i] (may be NULL): sum (resi - rec)^2, the SSE of xIntraCodingTUBlock when prediction + residual does not clip.
 * Synchronous. */
typedef struct {
  uint8_t log2_size;       /* 2..5 */
  uint8_t qp;              /* 0..51: luma QP, or the mapped chroma QP */
  uint8_t flags;           /* HEVCDL_TU_* */
  uint8_t reserved;
  uint32_t offset;
} hevcdl_tu;
#define HEVCDL_TU_DST 1    /* 4x4 intra luma: DST-VII (TComTU::useDST) */
#define HEVCDL_TU_TSKIP 2  /* transform skip */
#define HEVCDL_TU_INTER 4  /* rounding offset 85/512 instead of 171/512 (non-I slices) */
#define HEVCDL_TU_RDOQ 8   /* hevcdl_tu_code_rdoq only: quantise this TU with the rate-distortion optimised quantiser */
#define HEVCDL_TU_COEFF_IN 16 /* `resi` holds this TU's transform coefficients (16-bit by construction) instead of its residual: the
                                 forward transform is skipped; ssd[i] is meaningless for such a TU */
int hevcdl_tu_code(hevcdl_ctx *ctx, int n, const hevcdl_tu *tus, const int16_t *resi, size_t nelem, int32_t *coeff,
                   int16_t *level, int32_t *deq, int16_t *rec, uint32_t *abs_sum, uint64_t *ssd);

/* The same with the reference's rate-distortion optimised quantiser, TComTrQuant::xRateDistOptQuant (HM TLibCommon/
 * TComTrQuant.cpp:2119-2670, sign-bit hiding included), for the TUs flagged HEVCDL_TU_RDOQ -- the reference's operating point
 * (RDOQ 1, RDOQTS 1, SignHideFlag 1).  rdoq[i] carries what the reference reads from encoder state for TU i: the lambda of the
 * component (TComTrQuant::m_dLambda after selectLambda = TComSlice::getLambdas()[compID]), the scan type
 * (TComDataCU::getCoefScanIdx), the cbf context (getCtxQtCbf + getCBFContextOffset) and the index of the CABAC bit-estimate
 * table in force (estBitsSbacStruct as filled by TEncSbac::estBit before the call, TEncSearch.cpp:1282-1286): est holds n_est
 * tables of HEVCDL_EST_INTS int32 each, in the reference's own struct layout.  Bit-exact: every level decision compares
 * double-precision costs built in the reference's order of operations.  Synchronous. */
#define HEVCDL_EST_INTS 224
typedef struct {
  double lambda;
  uint32_t est_index;
  uint8_t channel;         /* 0 luma, 1 chroma */
  uint8_t scan_type;       /* 0 diagonal, 1 horizontal, 2 vertical */
  uint8_t ctx_cbf;
  uint8_t flags;           /* bit 0: sign-bit hiding enabled, bit 1: intra CU, bit 2: transform index of the CU is 0 */
} hevcdl_tu_rdoq;
int hevcdl_tu_code_rdoq(hevcdl_ctx *ctx, int n, const hevcdl_tu *tus, const hevcdl_tu_rdoq *rdoq, const int32_t *est, int n_est,
                        const int16_t *resi, size_t nelem, int32_t *coeff, int16_t *level, int32_t *deq, int16_t *rec,
                        uint32_t *abs_sum, uint64_t *ssd);

/* Deblocking filter of one ALL-INTRA reconstructed picture, in place: replaces TComLoopFilter::loopFilterPic
 * (HM TLibCommon/TComLoopFilter.cpp:130-158 with everything below it: edge selection, boundary strength -- 2 on every edge of
 * an all-intra picture --, the luma and chroma filters) as TEncGOP::compressGOP calls it after the last CTU of a picture
 * (HM TLibEncoder/TEncGOP.cpp:1742).  y / u / v: HM's Pel (int16) planes of the 8-bit 4:2:0 reconstruction, width x height luma
 * samples (multiples of 8), strides in samples.  tu_log2 / qp: one entry per 4x4 luma unit, raster order ((height/4) x (width/4)):
 * log2 of the size of the transform unit covering it (2..5; = 6 - CU depth - transform index) and its QP.  The offsets are the
 * slice's deblocking offsets (div 2) and the PPS chroma QP offsets.  Restrictions = the reference's operating point: every CU
 * intra, one slice, no tiles, no PCM / lossless blocks, deblocking enabled.  Bit-exact.  Synchronous. */
int hevcdl_deblock_frame(hevcdl_ctx *ctx, int16_t *y, int stride_y, int16_t *u, int16_t *v, int stride_c, int width, int height,
                         const uint8_t *tu_log2, const int8_t *qp, int beta_offset_div2, int tc_offset_div2, int cb_qp_offset,
                         int cr_qp_offset);

/* Deblocking and SAO statistics of one picture in ONE round trip, the picture staying on the device: what hevcdl_deblock_frame does
 * to y / u / v, then what hevcdl_sao_stats does with the deblocked picture against org_* -- one upload (reconstruction, maps,
 * original), one download (deblocked picture, statistics).  Afterwards the deblocked picture is RESIDENT in the context:
 * a following hevcdl_sao_apply with src_y = src_u = src_v = NULL takes it as its source, so that applying the offsets HM decides
 * from these statistics costs only the parameters up and the result down.  Any other in-loop entry point of the context
 * (hevcdl_deblock_frame, hevcdl_sao_stats, hevcdl_sao_apply with a source) drops the resident picture.  Same restrictions and
 * bit-exactness as the two calls it combines.  Synchronous. */
int hevcdl_inloop_frame(hevcdl_ctx *ctx, int16_t *y, int stride_y, int16_t *u, int16_t *v, int stride_c, int width, int height,
                        const uint8_t *tu_log2, const int8_t *qp, int beta_offset_div2, int tc_offset_div2, int cb_qp_offset,
                        int cr_qp_offset, const int16_t *org_y, const int16_t *org_u, const int16_t *org_v, int org_stride_y,
                        int org_stride_c, int64_t *stats);

/* Intra prediction of `n` blocks from explicit reference samples: the arithmetic of TComPrediction::predIntraAng (HM TLibCommon/
 * TComPrediction.cpp:390-472: planar, DC with boundary smoothing, the 33 angular predictors with the projected side reference and
 * the first-column filter of the pure vertical / horizontal modes), which the RD pass calls for every luma and chroma transform
 * block (TEncSearch.cpp:1208) and the first pass for every mode (:2303).  Per block: log2_size 2..6, mode 0..34, flags bit 0 =
 * luma edge filters enabled (what the reference applies for luma blocks up to 16x16; clear for chroma), line_offset = start, in
 * samples, of its reference line in `lines` -- 4*size+1 samples in the order of hevcdl_rmd_exact: left column bottom-up, corner,
 * top row (what getPredictorPtr holds after TComPattern's availability / substitution / smoothing steps, filtered or not as the
 * reference chose) --, pred_offset = where its size*size predicted samples go in `pred` (dense, row-major).  8-bit.  Bit-exact.
 * Synchronous. */
typedef struct hevcdl_pred_req {
  uint8_t log2_size, mode, flags, reserved;
  uint32_t line_offset;
  uint32_t pred_offset;
} hevcdl_pred_req;
#define HEVCDL_PRED_EDGE 1u
int hevcdl_intra_pred(hevcdl_ctx *ctx, int n, const hevcdl_pred_req *reqs, const int16_t *lines, size_t nline, int16_t *pred,
                      size_t npred);

/* SAO statistics of one deblocked picture: replaces the data pass of the reference's SAO parameter estimation,
 * TEncSampleAdaptiveOffset::getStatistics (HM TLibEncoder/TEncSampleAdaptiveOffset.cpp:295-341 -> getBlkStats :943-1345) as
 * SAOProcess calls it (:258) for deblocked samples.  org_* / rec_*: HM's Pel (int16) planes of the original and the deblocked
 * 8-bit 4:2:0 pictures, width x height luma samples, strides in samples.  stats receives, for every 64x64 CTU in raster order,
 * component (Y, Cb, Cr) and SAO type (edge offset 0 / 90 / 135 / 45 degrees, band offset): int64 diff[32] then int64 count[32]
 * -- the reference's SAOStatData arrays (TEncSampleAdaptiveOffset.h:66-70) laid end to end, nctu * 3 * 5 * 64 values.
 * Restrictions = the reference's operating point: one slice, no tiles, SAOLcuBoundary 0.  Bit-exact.  Synchronous. */
int hevcdl_sao_stats(hevcdl_ctx *ctx, const int16_t *org_y, const int16_t *org_u, const int16_t *org_v, int org_stride_y,
                     int org_stride_c, const int16_t *rec_y, const int16_t *rec_u, const int16_t *rec_v, int rec_stride_y,
                     int rec_stride_c, int width, int height, int64_t *stats);

/* SAO application over one picture: TComSampleAdaptiveOffset::offsetCTU for every CTU (HM TLibCommon/TComSampleAdaptiveOffset.cpp:
 * 554-611 -> offsetBlock :313-552) as the encoder runs it from decideBlkParams (TLibEncoder/TEncSampleAdaptiveOffset.cpp:894) once
 * a CTU's parameters are decided and its merge candidates resolved (reconstructBlkSAOParam).  src_*: the deblocked picture (HM's
 * m_tempPicYuv copy), res_*: receives every sample of the picture -- src + offset[class] clipped to 8 bits where the CTU's SAO
 * type applies, the deblocked value elsewhere.  params: for every 64x64 CTU in raster order and component (Y, Cb, Cr) the type
 * (-1 off; 0..3 edge offset 0 / 90 / 135 / 45 degrees; 4 band offset) and SAOOffset::offset[32] (entries 0..4 = the five edge
 * classes, 0..31 = the bands).  Restrictions = the reference's operating point: one slice, no tiles, 8-bit 4:2:0.  Bit-exact.
 * Synchronous. */
typedef struct hevcdl_sao_param {
  int8_t type;
  int8_t reserved[3];
  int8_t offset[32];
} hevcdl_sao_param;
int hevcdl_sao_apply(hevcdl_ctx *ctx, const int16_t *src_y, const int16_t *src_u, const int16_t *src_v, int src_stride_y,
                     int src_stride_c, int16_t *res_y, int16_t *res_u, int16_t *res_v, int res_stride_y, int res_stride_c, int width,
                     int height, const hevcdl_sao_param *params);

/* Page-locked host memory for frame planes handed over with hevcdl_cfg.pinned_input = 1 (any page-locked memory will do;
 * this is the allocator for callers without a CUDA runtime of their own).  write_combined != 0: cudaHostAllocWriteCombined --
 * not snooped during the transfer, which some hosts move faster over PCIe; the CPU should only WRITE such memory (reads
 * are uncached).  Returns NULL on failure. */
void *hevcdl_host_alloc(size_t bytes, int write_combined);
void hevcdl_host_free(void *p);

/* Page-lock memory the caller already owns (cudaHostRegister), e.g. the encoder's picture buffers: the in-loop entry points
 * (hevcdl_deblock_frame, hevcdl_inloop_frame, hevcdl_sao_stats, hevcdl_sao_apply) copy planes that lie in page-locked memory
 * straight between the caller's strided rows and the device (no packing pass through the context's staging buffer), planes in
 * ordinary memory through that buffer -- the results are the same.  Registering a range twice is not an error.  Unregister
 * before freeing the memory.  HEVCDL_E_CUDA if the range cannot be locked (the calls above still work, through staging). */
int hevcdl_host_register(void *p, size_t bytes);
int hevcdl_host_unregister(void *p);

/* Pin the calling thread to the CPUs of `device`'s NUMA node (what hevcdl_cfg.numa_bind does inside hevcdl_create),
 * for callers that allocate their own pinned frame buffers before creating a context.  Returns the node (>= 0), or a
 * negative hevcdl_status when the topology cannot be read (single-node hosts report node 0 or -1 in sysfs: no-op, 0). */
int hevcdl_numa_bind_thread(int device);

int hevcdl_get_stats(hevcdl_ctx *ctx, hevcdl_stats_t *out);

#ifdef __cplusplus
}
#endif
#endif /* HEVCDL_H */
