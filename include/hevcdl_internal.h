/* include/hevcdl_internal.h -- measurement and test hooks of libhevcdl.so.  NOT part of the drop-in boundary: nothing
 * the reference's encoder binds lives here (include/hevcdl.h is that surface).  Used by bench.py (through
 * hevc-deep-learning-pipeline_b200/host.py) and by tests/. */
#ifndef HEVCDL_INTERNAL_H
#define HEVCDL_INTERNAL_H
#include "hevcdl.h"

#ifdef __cplusplus
extern "C" {
#endif

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

/* Measurement: device time (CUDA events on the context's stream, around the kernels only -- the copies are outside) of the
 * most recent hevcdl_tu_code / hevcdl_tu_code_rdoq / hevcdl_deblock_frame / hevcdl_sao_stats call of this context.
 * HEVCDL_E_NOFRAME before the first such call. */
int hevcdl_last_aux_ms(hevcdl_ctx *ctx, float *ms);

/* CUDA stream handle (cudaStream_t) of the context, for callers that time with their own events */
void *hevcdl_stream(hevcdl_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* HEVCDL_INTERNAL_H */
