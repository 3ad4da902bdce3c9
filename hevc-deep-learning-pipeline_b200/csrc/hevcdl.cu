// hevcdl.cu -- C-ABI implementation (include/hevcdl.h) over the sm_100a kernels.
// Host side: frame slots, pinned staging, three streams (copies in, kernels, copies out) chained by CUDA events,
// so the H2D copy of frame i+1 and the D2H copies of frame i-1 overlap the kernels of frame i.
#include <cuda_runtime.h>
#include <sched.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cnn_fp32.cuh"
#include "cnn_tc.cuh"
#include "common.cuh"
#include "rmd.cuh"
#include "tq.cuh"
#include "dbf.cuh"
#include "sao.cuh"
#include "pred.cuh"
#include "../../include/hevcdl_internal.h"

using namespace hevcdl;

namespace {

// Kernel launch with the programmatic-dependent-launch attribute (common.cuh: pdl_wait / pdl_launch_dependents).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

enum SlotState { SLOT_FREE = 0, SLOT_PENDING = 1 /* copied in, kernels not launched yet */, SLOT_QUEUED = 2, SLOT_DONE = 3 };

struct Slot {
  int frame = -1;
  SlotState state = SLOT_FREE;
  // device
  uint8_t *dY = nullptr, *dU = nullptr, *dV = nullptr;
  uint8_t *dLabels = nullptr;
  float *dLogits = nullptr;
  int *dCtuOff = nullptr;
  uint32_t *dCtuCnt = nullptr;         // per-CTU npu | nitems << 16, written by the label kernel
  int *dCtrl = nullptr;                // [0] work counter, [1] item count (k_rmd_plan / k_rmd_items)
  RmdItem *dItems = nullptr;
  hevcdl_pu *dPus = nullptr;
  uint32_t *dSatd = nullptr;
  uint8_t *dCand = nullptr;
  // pinned host
  uint8_t *hPlanes = nullptr;          // staging Y|U|V (pitched like the device planes)
  uint8_t *hLabels = nullptr;
  float *hLogits = nullptr;
  int *hCtuOff = nullptr;
  hevcdl_pu *hPus = nullptr;
  uint32_t *hSatd = nullptr;
  uint8_t *hCand = nullptr;
  size_t hPuCap = 0;
  bool pusFetched = false;
  CUtensorMap tmY, tmU, tmV;           // TMA descriptors of the three device planes (tensor-core path: K1 stages by cp.async.bulk.tensor)
  bool timed = false;                  // head of a launch batch: its evT0..evT2 bracket the batch's stages
  cudaEvent_t evCnn = nullptr, evIn = nullptr, evLabels = nullptr, evRmd = nullptr, evT0 = nullptr, evT1 = nullptr, evT2 = nullptr;
};

}  // namespace

struct hevcdl_ctx {
  hevcdl_cfg cfg{};
  FrameGeom geo{};
  int pitch = 0, cpitch = 0;           // device plane pitches (bytes)
  size_t puCap = 0;
  cudaStream_t stream = nullptr, h2d = nullptr, d2h = nullptr;   // kernels / copies in / labels out
  cudaStream_t rmd = nullptr;          // K6 of batch k runs here, concurrently with the CNN of batch k+1 on `stream` (HEVCDL_RMD_STREAM=0: same stream)
  cudaStream_t d2hPu = nullptr;        // PU lists out, on demand (its own stream: must not queue behind later frames' label copies)
  std::vector<Slot> slots;
  float *dWeights = nullptr;           // raw HDLW blob
  float *dPacked = nullptr;            // fp32-path packed weights
  Fp32Params fp{};
  TcParams tc{};
  int numSMs = 0;
  int batch = 1;                       // frames per CNN launch (cfg.batch)
  bool stageTimes = true;              // bracket the CNN / RMD stages of every launch with events (hevcdl_get_stats: ms_cnn, ms_rmd)
  std::vector<Slot *> pending;         // submitted frames waiting for their batch to fill
  int rmdBlocks = 0;                   // persistent grid of k_rmd_items: resident blocks per SM x SMs
  std::string err;
  hevcdl_stats_t stats{};
  // device time of the most recent hevcdl_tu_code* / hevcdl_deblock_frame / hevcdl_sao_stats kernels (hevcdl_last_aux_ms)
  cudaEvent_t evAux0 = nullptr, evAux1 = nullptr;
  // scratch of the in-loop entry points (deblocking, SAO statistics / application)
  void *dDbf = nullptr;
  void *hDbf = nullptr;
  size_t dbfCap = 0;
  // the deblocked picture hevcdl_inloop_frame left at the start of dDbf (int16 Y, Cb, Cr, dense): source of a later
  // hevcdl_sao_apply(src = NULL); any other call that uses the scratch invalidates it
  bool loopValid = false;
  int loopW = 0, loopH = 0;
  // scratch for hevcdl_tu_code
  void *dRdoq = nullptr;               // RdoqScratch per resident warp of k_tu_code (RDOQ launches only)
  int rdoqWarps = 0;
  void *dTq = nullptr;
  void *hTq = nullptr;                 // pinned mirror of dTq
  size_t tqCap = 0;
  // scratch for hevcdl_rmd_exact
  void *dExact = nullptr;
  void *hExact = nullptr;              // pinned mirror of dExact
  size_t exactCap = 0;
};

namespace {

thread_local std::string g_create_err;

#define CK(call)                                                                            \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      char b_[512];                                                                         \
      snprintf(b_, sizeof b_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      ctx->err = b_;                                                                        \
      return HEVCDL_E_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

cudaError_t aux_begin(hevcdl_ctx *ctx) {          // start of the timed kernel region of an auxiliary entry point
  if (!ctx->evAux0) {
    cudaError_t e = cudaEventCreate(&ctx->evAux0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->evAux1);
    if (e != cudaSuccess) return e;
  }
  return cudaEventRecord(ctx->evAux0, ctx->stream);
}

Slot *find_slot(hevcdl_ctx *ctx, int frame) {
  for (auto &s : ctx->slots)
    if (s.state != SLOT_FREE && s.frame == frame) return &s;
  return nullptr;
}

// [COUT][CIN][K][K] -> [COUT/CG][CIN][K*K][CG]
void pack_conv(const float *w, int cout, int cin, int kk, int cg, float *out) {
  for (int co = 0; co < cout; co++)
    for (int ci = 0; ci < cin; ci++)
      for (int k = 0; k < kk; k++)
        out[(((size_t)(co / cg) * cin + ci) * kk + k) * cg + co % cg] = w[((size_t)co * cin + ci) * kk + k];
}
void transpose(const float *w, int rows, int cols, float *out) {  // [rows][cols] -> [cols][rows]
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) out[(size_t)c * rows + r] = w[(size_t)r * cols + c];
}

int load_weights(hevcdl_ctx *ctx) {
  FILE *f = fopen(ctx->cfg.weights_path, "rb");
  if (!f) { ctx->err = std::string("cannot open weights: ") + ctx->cfg.weights_path; return HEVCDL_E_WEIGHTS; }
  char magic[8];
  std::vector<float> w(HDLW_NFLOATS);
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "HDLW0001", 8) == 0 &&
            fread(w.data(), sizeof(float), HDLW_NFLOATS, f) == (size_t)HDLW_NFLOATS;
  char extra;
  ok = ok && fread(&extra, 1, 1, f) == 0;
  fclose(f);
  if (!ok) { ctx->err = "malformed HDLW weight blob"; return HEVCDL_E_WEIGHTS; }

  // fp32 path: packed convs + transposed fcs + the small vectors, one device allocation
  std::vector<float> pk;
  auto push = [&](size_t n) { size_t o = pk.size(); pk.resize(o + n); return o; };
  size_t o_c1 = push(16 * 3 * 25), o_c64 = push(16 * 3 * 25), o_c2 = push(64 * 32 * 9), o_c3 = push(128 * 64 * 9);
  size_t o_f1 = push(2048 * 256), o_f2 = push(256 * 64), o_f3 = push(64 * 16);
  pack_conv(&w[O_C1W], 16, 3, 25, 4, &pk[o_c1]);
  pack_conv(&w[O_C64W], 16, 3, 25, 1, &pk[o_c64]);
  pack_conv(&w[O_C2W], 64, 32, 9, 4, &pk[o_c2]);
  pack_conv(&w[O_C3W], 128, 64, 9, 4, &pk[o_c3]);
  transpose(&w[O_F1W], 256, 2048, &pk[o_f1]);
  transpose(&w[O_F2W], 64, 256, &pk[o_f2]);
  transpose(&w[O_F3W], 16, 64, &pk[o_f3]);
  CK(cudaMalloc(&ctx->dWeights, HDLW_NFLOATS * sizeof(float)));
  CK(cudaMemcpy(ctx->dWeights, w.data(), HDLW_NFLOATS * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&ctx->dPacked, pk.size() * sizeof(float)));
  CK(cudaMemcpy(ctx->dPacked, pk.data(), pk.size() * sizeof(float), cudaMemcpyHostToDevice));
  const float *W = ctx->dWeights, *P = ctx->dPacked;
  Fp32Params &p = ctx->fp;
  p.c1w = P + o_c1; p.c64w = P + o_c64; p.c2w = P + o_c2; p.c3w = P + o_c3;
  p.c1b = W + O_C1B; p.c64b = W + O_C64B; p.c2b = W + O_C2B; p.c3b = W + O_C3B;
  p.g1 = W + O_BN1G; p.b1 = W + O_BN1B; p.g64 = W + O_BN64G; p.b64 = W + O_BN64B;
  p.g2 = W + O_BN2G; p.b2 = W + O_BN2B; p.g3 = W + O_BN3G; p.b3 = W + O_BN3B;
  p.f1wT = P + o_f1; p.f1b = W + O_F1B; p.f2wT = P + o_f2; p.f2b = W + O_F2B; p.f3wT = P + o_f3; p.f3b = W + O_F3B;
  if (ctx->cfg.precision == HEVCDL_PREC_BF16_TC) {
    // pre-packed tensor-core operands live next to the fp32 blob: <name>.hdlw -> <name>.hdlt (tools/tc_pack.py)
    std::string tp = ctx->cfg.weights_path;
    const size_t dot = tp.rfind('.');
    tp = (dot == std::string::npos ? tp : tp.substr(0, dot)) + ".hdlt";
    return tc_prepare(tp.c_str(), ctx->geo.nctu * ctx->batch, &ctx->tc, ctx->err);
  }
  return HEVCDL_OK;
}

// 2-D u8 tensor map of one plane: dims (w, h), row pitch `pitch` bytes, box bw x bh, no swizzle, zero fill outside.
// cuTensorMapEncodeTiled is a driver entry point; it is looked up through the runtime so that libcuda is not a link
// dependency of this library.
int make_plane_tmap(hevcdl_ctx *ctx, CUtensorMap *tm, const uint8_t *base, int w, int h, int pitch, int bw, int bh) {
  typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiled enc = nullptr;
  if (!enc) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { ctx->err = "cuTensorMapEncodeTiled not available in this driver"; return HEVCDL_E_CUDA; }
    enc = (EncodeTiled)fn;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch};
  const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[128];
    snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (CUresult %d) for a %dx%d plane, pitch %d", (int)r, w, h, pitch);
    ctx->err = b;
    return HEVCDL_E_CUDA;
  }
  return HEVCDL_OK;
}

int alloc_slot(hevcdl_ctx *ctx, Slot &s) {
  const FrameGeom &g = ctx->geo;
  const size_t ybytes = (size_t)ctx->pitch * g.H, cbytes = (size_t)ctx->cpitch * (g.H / 2);
  CK(cudaMalloc(&s.dY, ybytes + 2 * cbytes));
  CK(cudaMemset(s.dY, 0, ybytes + 2 * cbytes));
  s.dU = s.dY + ybytes; s.dV = s.dU + cbytes;
  if (ctx->cfg.precision == HEVCDL_PREC_BF16_TC) {
    int rc;
    if ((rc = make_plane_tmap(ctx, &s.tmY, s.dY, g.W, g.H, ctx->pitch, 64, 64)) ||
        (rc = make_plane_tmap(ctx, &s.tmU, s.dU, g.W / 2, g.H / 2, ctx->cpitch, 32, 32)) ||
        (rc = make_plane_tmap(ctx, &s.tmV, s.dV, g.W / 2, g.H / 2, ctx->cpitch, 32, 32)))
      return rc;
  }
  CK(cudaMalloc(&s.dLabels, (size_t)g.nctu * 16));
  CK(cudaMalloc(&s.dLogits, (size_t)g.nctu * 64 * sizeof(float)));
  CK(cudaMalloc(&s.dCtuOff, ((size_t)g.nctu + 1) * sizeof(int)));
  CK(cudaMemset(s.dCtuOff, 0, ((size_t)g.nctu + 1) * sizeof(int)));
  CK(cudaMalloc(&s.dCtuCnt, (size_t)g.nctu * sizeof(uint32_t)));
  CK(cudaMalloc(&s.dCtrl, 2 * sizeof(int)));
  CK(cudaMemset(s.dCtrl, 0, 2 * sizeof(int)));
  CK(cudaMalloc(&s.dItems, (size_t)g.nctu * MAX_ITEMS_CTU * sizeof(RmdItem) * ctx->batch));   // the head slot of a batch holds its queue
  CK(cudaMalloc(&s.dPus, ctx->puCap * sizeof(hevcdl_pu)));
  CK(cudaMalloc(&s.dSatd, ctx->puCap * 35 * sizeof(uint32_t)));
  CK(cudaMalloc(&s.dCand, ctx->puCap * 8));
  CK(cudaMallocHost(&s.hPlanes, ybytes + 2 * cbytes));
  CK(cudaMallocHost(&s.hLabels, (size_t)g.nctu * 16));
  if (ctx->cfg.outputs & HEVCDL_OUT_LOGITS) CK(cudaMallocHost(&s.hLogits, (size_t)g.nctu * 64 * sizeof(float)));
  CK(cudaMallocHost(&s.hCtuOff, ((size_t)g.nctu + 1) * sizeof(int)));
  CK(cudaEventCreateWithFlags(&s.evIn, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&s.evCnn, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&s.evLabels, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&s.evRmd, cudaEventDisableTiming));
  CK(cudaEventCreate(&s.evT0)); CK(cudaEventCreate(&s.evT1)); CK(cudaEventCreate(&s.evT2));
  return HEVCDL_OK;
}

int ensure_host_pu_cap(hevcdl_ctx *ctx, Slot &s, size_t n) {
  if (n <= s.hPuCap) return HEVCDL_OK;
  size_t cap = s.hPuCap ? s.hPuCap : 4096;
  while (cap < n) cap *= 2;
  if (s.hPus) cudaFreeHost(s.hPus);
  if (s.hSatd) cudaFreeHost(s.hSatd);
  if (s.hCand) cudaFreeHost(s.hCand);
  s.hPus = nullptr; s.hSatd = nullptr; s.hCand = nullptr; s.hPuCap = 0;
  CK(cudaMallocHost(&s.hPus, cap * sizeof(hevcdl_pu)));
  if (ctx->cfg.outputs & HEVCDL_OUT_SATD) CK(cudaMallocHost(&s.hSatd, cap * 35 * sizeof(uint32_t)));
  CK(cudaMallocHost(&s.hCand, cap * 8));
  s.hPuCap = cap;
  return HEVCDL_OK;
}

// Order `st` behind `ev` -- unless the event has already completed: a wait node between two kernels costs front-end time
// and keeps the next kernel's CTAs from starting in the shadow of its predecessor (programmatic dependent launch).
static void wait_unless_done(cudaStream_t st, cudaEvent_t ev) {
  if (cudaEventQuery(ev) == cudaSuccess) return;
  cudaGetLastError();                             // cudaErrorNotReady is not an error
  cudaStreamWaitEvent(st, ev, 0);
}

// Queue the device pipeline of n <= cfg.batch slots on ctx->stream: one CNN launch for all of them (tensor-core path),
// then the RMD kernels per frame.  *launches += kernels launched; every launch / record return code is checked.
int launch_pipeline(hevcdl_ctx *ctx, Slot *const *sl, int n, bool timed, int *launches_out) {
  const FrameGeom g = ctx->geo;
  int launches = 0;
  Slot &head = *sl[0];
  if (ctx->rmd)
    for (int i = 0; i < n; i++) wait_unless_done(ctx->stream, sl[i]->evRmd);   // a slot's earlier K6 (other stream) is done
  if (timed) CK(cudaEventRecord(head.evT0, ctx->stream));
  if (ctx->cfg.precision == HEVCDL_PREC_BF16_TC) {
    FrameBatch fb{};
    TmapBatch tm;
    fb.n = n;
    for (int i = 0; i < MAX_BATCH; i++) {         // unused entries repeat the head's descriptors (never dereferenced)
      const Slot &s = *sl[i < n ? i : 0];
      tm.y[i] = s.tmY; tm.u[i] = s.tmU; tm.v[i] = s.tmV;
    }
    for (int i = 0; i < n; i++) {
      Slot &s = *sl[i];
      fb.Y[i] = s.dY; fb.U[i] = s.dU; fb.V[i] = s.dV;
      fb.labels[i] = s.dLabels; fb.logits[i] = s.dLogits; fb.ctu_cnt[i] = ctx->cfg.rmd ? s.dCtuCnt : nullptr;
    }
    CK(tc_launch(ctx->tc, fb, tm, g, ctx->pitch, ctx->cpitch, ctx->cfg.boundary_fix, ctx->numSMs, ctx->stream, &launches));
  } else {
    const int grid = g.nctu < 4 * ctx->numSMs ? g.nctu : 4 * ctx->numSMs;
    for (int i = 0; i < n; i++) {
      Slot &s = *sl[i];
      CK(launch_pdl(k_cnn_fp32, grid, FP32_THREADS, FP32_SMEM_BYTES, ctx->stream, (const uint8_t *)s.dY, (const uint8_t *)s.dU, (const uint8_t *)s.dV, g,
                    ctx->pitch, ctx->cpitch, ctx->fp, ctx->cfg.boundary_fix, s.dLabels, s.dLogits, ctx->cfg.rmd ? s.dCtuCnt : nullptr));
      launches++;
    }
  }
  if (timed) CK(cudaEventRecord(head.evT1, ctx->stream));
  // K6 is CUDA-core work with small blocks, the CNN kernels are one big tensor-core CTA per SM: on its own stream the
  // RMD pass of this batch shares the SMs with K1/K2 of the next batch instead of waiting in line behind them.
  cudaStream_t rs = ctx->rmd ? ctx->rmd : ctx->stream;
  if (ctx->cfg.rmd) {                             // one plan + one items launch for the whole batch (queue in the head slot)
    if (rs != ctx->stream) {
      CK(cudaEventRecord(head.evCnn, ctx->stream));
      CK(cudaStreamWaitEvent(rs, head.evCnn, 0));
    }
    RmdBatch rb{};
    rb.n = n;
    for (int i = 0; i < n; i++) {
      Slot &s = *sl[i];
      rb.Y[i] = s.dY; rb.labels[i] = s.dLabels; rb.ctu_cnt[i] = s.dCtuCnt; rb.ctu_off[i] = s.dCtuOff;
      rb.pus[i] = s.dPus; rb.satd[i] = s.dSatd; rb.cand[i] = s.dCand;
    }
    CK(launch_pdl(k_rmd_plan, (g.nctu * n + 7) / 8, 256, 0, rs, rb, g, ctx->rmdBlocks, head.dItems, head.dCtrl));
    CK(launch_pdl(k_rmd_items, ctx->rmdBlocks, RMD_BW * 32, 0, rs, rb, g, ctx->pitch, (const RmdItem *)head.dItems, head.dCtrl));
    launches += 2;
  } else {
    rs = ctx->stream;
  }
  for (int i = 0; i < n; i++) CK(cudaEventRecord(sl[i]->evRmd, rs));   // every kernel of these frames is done
  if (timed) CK(cudaEventRecord(head.evT2, rs));
  if (launches_out) *launches_out += launches;
  return HEVCDL_OK;
}

// Launch the kernels of the frames submitted so far (a full batch, or fewer when somebody asks for a pending frame)
// and queue their label copies behind them.
int flush_pending(hevcdl_ctx *ctx) {
  const int n = (int)ctx->pending.size();
  if (n == 0) return HEVCDL_OK;
  const FrameGeom &g = ctx->geo;
  for (Slot *s : ctx->pending) wait_unless_done(ctx->stream, s->evIn);
  {
    int nl = 0;
    const int rc = launch_pipeline(ctx, ctx->pending.data(), n, ctx->stageTimes, &nl);
    ctx->stats.kernel_launches += nl;
    if (rc) return rc;
  }
  CK(cudaGetLastError());
  for (int i = 0; i < n; i++) {
    Slot *s = ctx->pending[i];
    s->state = SLOT_QUEUED;
    s->timed = ctx->stageTimes && i == 0;
    CK(cudaStreamWaitEvent(ctx->d2h, s->evRmd, 0));
    CK(cudaMemcpyAsync(s->hLabels, s->dLabels, (size_t)g.nctu * 16, cudaMemcpyDeviceToHost, ctx->d2h));
    if (ctx->cfg.outputs & HEVCDL_OUT_LOGITS)
      CK(cudaMemcpyAsync(s->hLogits, s->dLogits, (size_t)g.nctu * 64 * sizeof(float), cudaMemcpyDeviceToHost, ctx->d2h));
    if (ctx->cfg.rmd)
      CK(cudaMemcpyAsync(s->hCtuOff, s->dCtuOff, ((size_t)g.nctu + 1) * sizeof(int), cudaMemcpyDeviceToHost, ctx->d2h));
    CK(cudaEventRecord(s->evLabels, ctx->d2h));
  }
  ctx->pending.clear();
  return HEVCDL_OK;
}

template <class T>
int submit_impl(hevcdl_ctx *ctx, int frame, const T *y, int sy, const T *u, const T *v, int sc) {
  if (!ctx || !y || !u || !v || sy < ctx->geo.W || sc < ctx->geo.W / 2) return HEVCDL_E_INVAL;
  if (find_slot(ctx, frame)) { ctx->err = "frame id already in flight"; return HEVCDL_E_INVAL; }
  Slot *s = nullptr;
  for (auto &c : ctx->slots) if (c.state == SLOT_FREE) { s = &c; break; }
  if (!s) return HEVCDL_E_BUSY;
  const FrameGeom &g = ctx->geo;
  const int W = g.W, H = g.H, P = ctx->pitch, CP = ctx->cpitch;
  uint8_t *hy = s->hPlanes, *hu = hy + (size_t)P * H, *hv = hu + (size_t)CP * (H / 2);
  // hevcdl_cfg.pinned_input: the caller's 8-bit planes are page-locked and stay untouched until the frame has been
  // waited for, so the copy engine may read them after this call returns (no staging copy, no pointer probing)
  const bool direct = sizeof(T) == 1 && ctx->cfg.pinned_input;
  const uint8_t *src_y = reinterpret_cast<const uint8_t *>(y);
  if (direct && sy == P && sc == CP && (const uint8_t *)u == src_y + (size_t)P * H && (const uint8_t *)v == (const uint8_t *)u + (size_t)CP * (H / 2)) {
    // one contiguous pinned I420 frame whose strides equal the device pitches: a single copy
    CK(cudaMemcpyAsync(s->dY, src_y, (size_t)P * H + 2 * (size_t)CP * (H / 2), cudaMemcpyHostToDevice, ctx->h2d));
  } else if (direct) {
    CK(cudaMemcpy2DAsync(s->dY, P, src_y, sy, W, H, cudaMemcpyHostToDevice, ctx->h2d));
    CK(cudaMemcpy2DAsync(s->dU, CP, u, sc, W / 2, H / 2, cudaMemcpyHostToDevice, ctx->h2d));
    CK(cudaMemcpy2DAsync(s->dV, CP, v, sc, W / 2, H / 2, cudaMemcpyHostToDevice, ctx->h2d));
  } else {
    for (int r = 0; r < H; r++) {
      const T *row = y + (size_t)r * sy;
      if (sizeof(T) == 1) memcpy(hy + (size_t)r * P, row, W);
      else for (int x = 0; x < W; x++) hy[(size_t)r * P + x] = (uint8_t)row[x];
    }
    for (int r = 0; r < H / 2; r++) {
      const T *ru = u + (size_t)r * sc, *rv = v + (size_t)r * sc;
      if (sizeof(T) == 1) { memcpy(hu + (size_t)r * CP, ru, W / 2); memcpy(hv + (size_t)r * CP, rv, W / 2); }
      else for (int x = 0; x < W / 2; x++) { hu[(size_t)r * CP + x] = (uint8_t)ru[x]; hv[(size_t)r * CP + x] = (uint8_t)rv[x]; }
    }
    CK(cudaMemcpyAsync(s->dY, s->hPlanes, (size_t)P * H + 2 * (size_t)CP * (H / 2), cudaMemcpyHostToDevice, ctx->h2d));
  }
  CK(cudaEventRecord(s->evIn, ctx->h2d));
  s->frame = frame; s->state = SLOT_PENDING; s->pusFetched = false; s->timed = false;
  ctx->pending.push_back(s);
  if ((int)ctx->pending.size() >= ctx->batch) return flush_pending(ctx);
  return HEVCDL_OK;
}

int finish_slot(hevcdl_ctx *ctx, Slot *s) {
  if (s->state == SLOT_PENDING) {               // its batch never filled: launch what is there
    int rc = flush_pending(ctx);
    if (rc) return rc;
  }
  if (s->state == SLOT_QUEUED) {
    CK(cudaEventSynchronize(s->evLabels));
    if (s->timed) {                             // stage times of the whole launch batch, once
      float a = 0, b = 0;
      if (cudaEventSynchronize(s->evT2) == cudaSuccess) {
        if (cudaEventElapsedTime(&a, s->evT0, s->evT1) == cudaSuccess) ctx->stats.ms_cnn += a;
        if (cudaEventElapsedTime(&b, s->evT1, s->evT2) == cudaSuccess) ctx->stats.ms_rmd += b;
      }
    }
    s->state = SLOT_DONE;
    ctx->stats.frames++; ctx->stats.ctus += ctx->geo.nctu;
    if (ctx->cfg.rmd) ctx->stats.pus += s->hCtuOff[ctx->geo.nctu];
  }
  return HEVCDL_OK;
}

int fetch_pus(hevcdl_ctx *ctx, Slot *s) {
  if (s->pusFetched) return HEVCDL_OK;
  if (!ctx->cfg.rmd) { ctx->err = "context created with rmd=0"; return HEVCDL_E_INVAL; }
  const size_t n = (size_t)s->hCtuOff[ctx->geo.nctu];
  int rc = ensure_host_pu_cap(ctx, *s, n ? n : 1);
  if (rc) return rc;
  CK(cudaStreamWaitEvent(ctx->d2hPu, s->evRmd, 0));
  if (n) {
    CK(cudaMemcpyAsync(s->hPus, s->dPus, n * sizeof(hevcdl_pu), cudaMemcpyDeviceToHost, ctx->d2hPu));
    if (ctx->cfg.outputs & HEVCDL_OUT_SATD)
      CK(cudaMemcpyAsync(s->hSatd, s->dSatd, n * 35 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->d2hPu));
    CK(cudaMemcpyAsync(s->hCand, s->dCand, n * 8, cudaMemcpyDeviceToHost, ctx->d2hPu));
  }
  CK(cudaStreamSynchronize(ctx->d2hPu));
  s->pusFetched = true;
  return HEVCDL_OK;
}

}  // namespace

extern "C" {

const char *hevcdl_status_str(int st) {
  switch (st) {
    case HEVCDL_OK: return "ok";
    case HEVCDL_E_INVAL: return "invalid argument";
    case HEVCDL_E_NODEVICE: return "no usable CUDA device (sm_100 required; there is no CPU fallback)";
    case HEVCDL_E_CUDA: return "CUDA error";
    case HEVCDL_E_WEIGHTS: return "weight blob missing or malformed";
    case HEVCDL_E_NOFRAME: return "unknown frame";
    case HEVCDL_E_BUSY: return "no free frame slot";
    case HEVCDL_E_NOMEM: return "out of memory";
  }
  return "unknown status";
}

const char *hevcdl_last_error(const hevcdl_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

void *hevcdl_host_alloc(size_t bytes, int write_combined) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void hevcdl_host_free(void *p) { if (p) cudaFreeHost(p); }

// NUMA node of a CUDA device from sysfs; the calling thread is restricted to that node's CPUs, so that pinned
// allocations (first touch) and staging memcpys made by it afterwards are local to the GPU's PCIe root complex.
int hevcdl_numa_bind_thread(int device) {
  char bdf[32] = {0};
  if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, device) != cudaSuccess) { cudaGetLastError(); return HEVCDL_E_NODEVICE; }
  for (char *c = bdf; *c; c++) if (*c >= 'A' && *c <= 'F') *c += 'a' - 'A';   // sysfs names are lower case
  char path[128];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
  int node = -1;
  if (FILE *f = fopen(path, "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
  if (node < 0) return 0;                         // no NUMA information: single node
  snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
  FILE *f = fopen(path, "r");
  if (!f) return 0;
  char list[4096] = {0};
  const bool got = fgets(list, sizeof list, f) != nullptr;
  fclose(f);
  if (!got) return 0;
  cpu_set_t set, cur;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof cur, &cur) != 0) return 0;
  int nset = 0;
  for (char *p = list; *p;) {                     // "0-31,64-95"
    char *e;
    const long a = strtol(p, &e, 10);
    if (e == p) break;
    long b = a;
    if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
    for (long c = a; c <= b && c < CPU_SETSIZE; c++)
      if (CPU_ISSET(c, &cur)) { CPU_SET(c, &set); nset++; }     // never widen a mask the launcher (cgroup, taskset) gave us
    p = (*e == ',') ? e + 1 : e;
    if (*e != ',') break;
  }
  if (nset > 0) sched_setaffinity(0, sizeof set, &set);
  return node;
}

int hevcdl_create(const hevcdl_cfg *cfg, hevcdl_ctx **out) {
  if (!cfg || !out || cfg->abi_version != HEVCDL_ABI_VERSION || cfg->width <= 0 || cfg->height <= 0 ||
      (cfg->width % 8) || (cfg->height % 8) || cfg->width > 8192 || cfg->height > 8192 || !cfg->weights_path ||
      (cfg->precision != HEVCDL_PREC_FP32 && cfg->precision != HEVCDL_PREC_BF16_TC) ||
      (cfg->outputs & ~(HEVCDL_OUT_LOGITS | HEVCDL_OUT_SATD))) {
    g_create_err = "invalid hevcdl_cfg";
    return HEVCDL_E_INVAL;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) {
    cudaGetLastError();
    g_create_err = "no CUDA device: libhevcdl has no CPU fallback";
    return HEVCDL_E_NODEVICE;
  }
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10) {
    g_create_err = "device is not sm_100 (B200): kernels are built for sm_100a only";
    return HEVCDL_E_NODEVICE;
  }
  if (cfg->precision == HEVCDL_PREC_BF16_TC && !TC_BUILT) {
    g_create_err = "tensor-core (bf16 tcgen05) CNN path not built in this library";
    return HEVCDL_E_INVAL;
  }
  hevcdl_ctx *ctx = new hevcdl_ctx();
  ctx->cfg = *cfg;
  ctx->cfg.slots = cfg->slots < 1 ? 1 : cfg->slots;
  ctx->batch = cfg->batch < 1 ? 1 : (cfg->batch > MAX_BATCH ? MAX_BATCH : cfg->batch);
  if (ctx->batch > ctx->cfg.slots) ctx->batch = ctx->cfg.slots;
  if (cfg->precision != HEVCDL_PREC_BF16_TC) ctx->batch = 1;   // the fp32 parity path launches per frame
  // throughput mode: no timing events between the launches of consecutive batches unless asked for
  ctx->stageTimes = ctx->batch == 1 || getenv("HEVCDL_STAGE_TIMES") != nullptr;
  std::string wp = cfg->weights_path;
  auto fail = [&](int rc) { g_create_err = ctx->err; hevcdl_destroy(ctx); return rc; };
  if (cudaSetDevice(cfg->device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return fail(HEVCDL_E_CUDA); }
  if (cfg->numa_bind) hevcdl_numa_bind_thread(cfg->device);   // best effort: a host without NUMA topology in sysfs is not an error
  ctx->numSMs = prop.multiProcessorCount;
  FrameGeom &g = ctx->geo;
  g.W = cfg->width; g.H = cfg->height;
  g.ctu_w = (g.W + 63) / 64; g.ctu_h = (g.H + 63) / 64; g.nctu = g.ctu_w * g.ctu_h;
  ctx->pitch = ((g.W + 127) / 128) * 128; ctx->cpitch = ctx->pitch / 2;
  ctx->puCap = (size_t)g.nctu * MAX_PU_CTU;
  int rc;
  ctx->cfg.weights_path = wp.c_str();
  auto cu = [&](cudaError_t e, const char *what) {
    if (e != cudaSuccess) { ctx->err = std::string(what) + ": " + cudaGetErrorString(e); return true; }
    return false;
  };
  if (cu(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking), "stream")) return fail(HEVCDL_E_CUDA);
  if (cu(cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking), "stream")) return fail(HEVCDL_E_CUDA);
  if (cu(cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking), "stream")) return fail(HEVCDL_E_CUDA);
  if (cu(cudaStreamCreateWithFlags(&ctx->d2hPu, cudaStreamNonBlocking), "stream")) return fail(HEVCDL_E_CUDA);
  {
    const char *e = getenv("HEVCDL_RMD_STREAM");
    if (ctx->cfg.rmd && !(e && e[0] == '0') && cu(cudaStreamCreateWithFlags(&ctx->rmd, cudaStreamNonBlocking), "stream")) return fail(HEVCDL_E_CUDA);
  }
  if ((rc = load_weights(ctx))) return fail(rc);
  ctx->cfg.weights_path = nullptr;
  if (cu(cudaFuncSetAttribute(k_cnn_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, FP32_SMEM_BYTES), "smem attr") ||
      cu(cudaFuncSetAttribute(k_rmd_items, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), "carveout"))
    return fail(HEVCDL_E_CUDA);
  {
    int per_sm = 0;
    if (cu(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rmd_items, RMD_BW * 32, 0), "occupancy") || per_sm < 1) {
      if (ctx->err.empty()) ctx->err = "k_rmd_items does not fit on an SM";
      return fail(HEVCDL_E_CUDA);
    }
    ctx->rmdBlocks = per_sm * ctx->numSMs;
  }
  if ((rc = tc_configure(ctx->err))) return fail(rc);
  ctx->slots.resize(ctx->cfg.slots);
  for (auto &s : ctx->slots)
    if ((rc = alloc_slot(ctx, s))) return fail(rc);
  *out = ctx;
  return HEVCDL_OK;
}

void hevcdl_destroy(hevcdl_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
#ifdef HEVCDL_TIMELINE
  if (const char *out = getenv("HEVCDL_TIMELINE_OUT")) {
    cudaDeviceSynchronize();
    unsigned n = 0;
    cudaMemcpyFromSymbol(&n, g_tl_n, sizeof n);
    if (n > TL_CAP) n = TL_CAP;
    std::vector<unsigned long long> rec((size_t)n * 4);
    if (n && cudaMemcpyFromSymbol(rec.data(), g_tl, rec.size() * 8) == cudaSuccess) {
      if (FILE *f = fopen(out, "wb")) { fwrite(rec.data(), 8, rec.size(), f); fclose(f); }
    }
    n = 0;
    cudaMemcpyToSymbol(g_tl_n, &n, sizeof n);
  }
#endif
#ifdef HEVCDL_TRACE
  {
    cudaDeviceSynchronize();
    unsigned long long tr[64];
    if (cudaMemcpyFromSymbol(tr, tc::g_trace, sizeof tr) == cudaSuccess) {
      fprintf(stderr, "hevcdl trace (block 0 cycles, summed over all launches; epilogue sites are summed over the epilogue warps):\n"
                      " K1: total %llu | mma: w %llu empty %llu planes %llu | epi: full %llu | staging (3 warps): raw %llu free64 %llu free1 %llu\n"
                      " K2: total %llu | mma: cfree %llu cfull %llu empty %llu | epi: full %llu\n"
                      " K3: total %llu | mma: w %llu empty %llu afull %llu afree %llu | epi: full %llu\n",
              tr[16], tr[0], tr[1], tr[22], tr[2], tr[23], tr[20], tr[21], tr[17], tr[4], tr[5], tr[6], tr[7], tr[18], tr[8], tr[9], tr[10], tr[11], tr[12]);
      memset(tr, 0, sizeof tr);
      cudaMemcpyToSymbol(tc::g_trace, tr, sizeof tr);
    }
  }
#endif
  if (ctx->h2d) cudaStreamSynchronize(ctx->h2d);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->d2h) cudaStreamSynchronize(ctx->d2h);
  if (ctx->rmd) cudaStreamSynchronize(ctx->rmd);
  if (ctx->d2hPu) cudaStreamSynchronize(ctx->d2hPu);
  for (auto &s : ctx->slots) {
    cudaFree(s.dY); cudaFree(s.dLabels); cudaFree(s.dLogits); cudaFree(s.dCtuOff);
    cudaFree(s.dPus); cudaFree(s.dSatd); cudaFree(s.dCand); cudaFree(s.dCtuCnt); cudaFree(s.dCtrl); cudaFree(s.dItems);
    cudaFreeHost(s.hPlanes); cudaFreeHost(s.hLabels); cudaFreeHost(s.hLogits); cudaFreeHost(s.hCtuOff);
    cudaFreeHost(s.hPus); cudaFreeHost(s.hSatd); cudaFreeHost(s.hCand);
    if (s.evIn) cudaEventDestroy(s.evIn);
    if (s.evCnn) cudaEventDestroy(s.evCnn);
    if (s.evLabels) cudaEventDestroy(s.evLabels);
    if (s.evRmd) cudaEventDestroy(s.evRmd);
    if (s.evT0) cudaEventDestroy(s.evT0);
    if (s.evT1) cudaEventDestroy(s.evT1);
    if (s.evT2) cudaEventDestroy(s.evT2);
  }
  cudaFree(ctx->dWeights); cudaFree(ctx->dPacked); cudaFree(ctx->dExact); cudaFree(ctx->dTq); cudaFree(ctx->dRdoq); cudaFree(ctx->dDbf);
  if (ctx->hDbf) cudaFreeHost(ctx->hDbf);
  if (ctx->evAux0) cudaEventDestroy(ctx->evAux0);
  if (ctx->evAux1) cudaEventDestroy(ctx->evAux1);
  if (ctx->hExact) cudaFreeHost(ctx->hExact);
  if (ctx->hTq) cudaFreeHost(ctx->hTq);
  tc_release(&ctx->tc);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->d2h) cudaStreamDestroy(ctx->d2h);
  if (ctx->h2d) cudaStreamDestroy(ctx->h2d);
  if (ctx->d2hPu) cudaStreamDestroy(ctx->d2hPu);
  if (ctx->rmd) cudaStreamDestroy(ctx->rmd);
  cudaGetLastError();
  delete ctx;
}

int hevcdl_submit_frame_u8(hevcdl_ctx *ctx, int frame, const uint8_t *y, int sy, const uint8_t *u, const uint8_t *v, int sc) {
  if (ctx) cudaSetDevice(ctx->cfg.device);
  return submit_impl<uint8_t>(ctx, frame, y, sy, u, v, sc);
}
int hevcdl_submit_frame_pel16(hevcdl_ctx *ctx, int frame, const int16_t *y, int sy, const int16_t *u, const int16_t *v, int sc) {
  if (ctx) cudaSetDevice(ctx->cfg.device);
  return submit_impl<int16_t>(ctx, frame, y, sy, u, v, sc);
}

int hevcdl_wait_frame(hevcdl_ctx *ctx, int frame) {
  if (!ctx) return HEVCDL_E_INVAL;
  Slot *s = find_slot(ctx, frame);
  if (!s) return HEVCDL_E_NOFRAME;
  return finish_slot(ctx, s);
}

int hevcdl_ctu_labels(hevcdl_ctx *ctx, int frame, int addr, uint8_t out[16]) {
  if (!ctx || !out || addr < 0 || addr >= ctx->geo.nctu) return HEVCDL_E_INVAL;
  Slot *s = find_slot(ctx, frame);
  if (!s) return HEVCDL_E_NOFRAME;
  int rc = finish_slot(ctx, s);
  if (rc) return rc;
  memcpy(out, s->hLabels + (size_t)addr * 16, 16);
  return HEVCDL_OK;
}

int hevcdl_frame_labels(hevcdl_ctx *ctx, int frame, uint8_t *labels, float *logits) {
  if (!ctx) return HEVCDL_E_INVAL;
  Slot *s = find_slot(ctx, frame);
  if (!s) return HEVCDL_E_NOFRAME;
  int rc = finish_slot(ctx, s);
  if (rc) return rc;
  if (logits && !(ctx->cfg.outputs & HEVCDL_OUT_LOGITS)) { ctx->err = "context created without HEVCDL_OUT_LOGITS"; return HEVCDL_E_INVAL; }
  if (labels) memcpy(labels, s->hLabels, (size_t)ctx->geo.nctu * 16);
  if (logits) memcpy(logits, s->hLogits, (size_t)ctx->geo.nctu * 64 * sizeof(float));
  return HEVCDL_OK;
}

int hevcdl_frame_pu_count(hevcdl_ctx *ctx, int frame, int *npu) {
  if (!ctx || !npu) return HEVCDL_E_INVAL;
  Slot *s = find_slot(ctx, frame);
  if (!s) return HEVCDL_E_NOFRAME;
  if (!ctx->cfg.rmd) { ctx->err = "context created with rmd=0"; return HEVCDL_E_INVAL; }
  int rc = finish_slot(ctx, s);
  if (rc) return rc;
  *npu = s->hCtuOff[ctx->geo.nctu];
  return HEVCDL_OK;
}

int hevcdl_frame_pus(hevcdl_ctx *ctx, int frame, hevcdl_pu *pus, uint32_t *satd, uint8_t *cand) {
  if (!ctx) return HEVCDL_E_INVAL;
  Slot *s = find_slot(ctx, frame);
  if (!s) return HEVCDL_E_NOFRAME;
  if (satd && !(ctx->cfg.outputs & HEVCDL_OUT_SATD)) { ctx->err = "context created without HEVCDL_OUT_SATD"; return HEVCDL_E_INVAL; }
  int rc = finish_slot(ctx, s);
  if (rc) return rc;
  if ((rc = fetch_pus(ctx, s))) return rc;
  const size_t n = (size_t)s->hCtuOff[ctx->geo.nctu];
  if (pus) memcpy(pus, s->hPus, n * sizeof(hevcdl_pu));
  if (satd) memcpy(satd, s->hSatd, n * 35 * sizeof(uint32_t));
  if (cand) memcpy(cand, s->hCand, n * 8);
  return HEVCDL_OK;
}

int hevcdl_ctu_pu_range(hevcdl_ctx *ctx, int frame, int addr, int *first, int *count) {
  if (!ctx || !first || !count || addr < 0 || addr >= ctx->geo.nctu) return HEVCDL_E_INVAL;
  Slot *s = find_slot(ctx, frame);
  if (!s) return HEVCDL_E_NOFRAME;
  if (!ctx->cfg.rmd) { ctx->err = "context created with rmd=0"; return HEVCDL_E_INVAL; }
  int rc = finish_slot(ctx, s);
  if (rc) return rc;
  *first = s->hCtuOff[addr];
  *count = s->hCtuOff[addr + 1] - s->hCtuOff[addr];
  return HEVCDL_OK;
}

int hevcdl_frame_view_get(hevcdl_ctx *ctx, int frame, int want_pus, hevcdl_frame_view *out) {
  if (!ctx || !out) return HEVCDL_E_INVAL;
  Slot *s = find_slot(ctx, frame);
  if (!s) return HEVCDL_E_NOFRAME;
  int rc = finish_slot(ctx, s);
  if (rc) return rc;
  memset(out, 0, sizeof *out);
  out->labels = s->hLabels; out->logits = (ctx->cfg.outputs & HEVCDL_OUT_LOGITS) ? s->hLogits : nullptr; out->nctu = ctx->geo.nctu;
  if (ctx->cfg.rmd) {
    out->ctu_off = s->hCtuOff;
    if (want_pus) {
      if ((rc = fetch_pus(ctx, s))) return rc;
      out->npu = s->hCtuOff[ctx->geo.nctu];
      out->pus = s->hPus; out->satd = (ctx->cfg.outputs & HEVCDL_OUT_SATD) ? s->hSatd : nullptr; out->cand = s->hCand;
    }
  }
  return HEVCDL_OK;
}

int hevcdl_release_frame(hevcdl_ctx *ctx, int frame) {
  if (!ctx) return HEVCDL_E_INVAL;
  Slot *s = find_slot(ctx, frame);
  if (!s) return HEVCDL_E_NOFRAME;
  if (s->state == SLOT_QUEUED || s->state == SLOT_PENDING) {
    int rc = finish_slot(ctx, s);
    if (rc) return rc;
  }
  CK(cudaEventSynchronize(s->evRmd));
  s->state = SLOT_FREE; s->frame = -1;
  return HEVCDL_OK;
}

int hevcdl_rmd_exact(hevcdl_ctx *ctx, int n, const uint8_t *sizes, const uint8_t *org, const int16_t *lines,
                     const uint32_t *bits, const int8_t *mpm, const uint8_t *mpm_add, double sqrt_lambda,
                     uint32_t *satd, uint8_t *cand, uint8_t *ncand) {
  if (!ctx || n < 0 || (n && (!sizes || !org || !lines))) return HEVCDL_E_INVAL;
  if (n == 0) return HEVCDL_OK;
  cudaSetDevice(ctx->cfg.device);
  std::vector<int> org_off(n + 1), line_off(n + 1);
  org_off[0] = line_off[0] = 0;
  for (int i = 0; i < n; i++) {
    const int s = sizes[i];
    if (s != 4 && s != 8 && s != 16 && s != 32 && s != 64) { ctx->err = "PU size must be 4,8,16,32,64"; return HEVCDL_E_INVAL; }
    org_off[i + 1] = org_off[i] + s * s;
    line_off[i + 1] = line_off[i] + 4 * s + 1;
  }
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t b_sizes = al(n), b_org = al(org_off[n]), b_ooff = al((size_t)n * 4), b_lines = al((size_t)line_off[n] * 2),
               b_loff = al((size_t)n * 4), b_bits = al((size_t)n * 35 * 4), b_mpm = al((size_t)n * 3), b_madd = al(n),
               b_satd = al((size_t)n * 35 * 4), b_cand = al((size_t)n * 10), b_nc = al(n);
  const size_t total = b_sizes + b_org + b_ooff + b_lines + b_loff + b_bits + b_mpm + b_madd + b_satd + b_cand + b_nc;
  if (total > ctx->exactCap) {
    cudaFree(ctx->dExact); ctx->dExact = nullptr; ctx->exactCap = 0;
    if (ctx->hExact) { cudaFreeHost(ctx->hExact); ctx->hExact = nullptr; }
    const size_t cap = total < 65536 ? 65536 : 2 * total;     // per-PU callers vary the size call by call: grow rarely
    CK(cudaMalloc(&ctx->dExact, cap));
    CK(cudaMallocHost(&ctx->hExact, cap));
    ctx->exactCap = cap;
  }
  // one pinned mirror with the device layout: a single copy in, a single copy out (the call is latency-bound:
  // HM's exact mode makes one per PU)
  const size_t o_sizes = 0, o_org = o_sizes + b_sizes, o_ooff = o_org + b_org, o_lines = o_ooff + b_ooff, o_loff = o_lines + b_lines,
               o_bits = o_loff + b_loff, o_mpm = o_bits + b_bits, o_madd = o_mpm + b_mpm, o_satd = o_madd + b_madd,
               o_cand = o_satd + b_satd, o_nc = o_cand + b_cand;
  uint8_t *hp = (uint8_t *)ctx->hExact, *dp = (uint8_t *)ctx->dExact;
  memcpy(hp + o_sizes, sizes, n);
  memcpy(hp + o_org, org, org_off[n]);
  memcpy(hp + o_ooff, org_off.data(), (size_t)n * 4);
  memcpy(hp + o_lines, lines, (size_t)line_off[n] * 2);
  memcpy(hp + o_loff, line_off.data(), (size_t)n * 4);
  if (bits) memcpy(hp + o_bits, bits, (size_t)n * 35 * 4);
  if (mpm) memcpy(hp + o_mpm, mpm, (size_t)n * 3);
  if (mpm_add) memcpy(hp + o_madd, mpm_add, n);
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(dp, hp, o_satd, cudaMemcpyHostToDevice, st));
  const int grid = n < 16 * ctx->numSMs ? n : 16 * ctx->numSMs;
  k_rmd_exact<<<grid, 128, 0, st>>>(n, dp + o_sizes, dp + o_org, (const int *)(dp + o_ooff), (const int16_t *)(dp + o_lines),
                                    (const int *)(dp + o_loff), bits ? (const uint32_t *)(dp + o_bits) : nullptr,
                                    mpm ? (const int8_t *)(dp + o_mpm) : nullptr, mpm_add ? dp + o_madd : nullptr, sqrt_lambda,
                                    (uint32_t *)(dp + o_satd), dp + o_cand, dp + o_nc);
  CK(cudaGetLastError());
  ctx->stats.kernel_launches++;
  CK(cudaMemcpyAsync(hp + o_satd, dp + o_satd, total - o_satd, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (satd) memcpy(satd, hp + o_satd, (size_t)n * 35 * 4);
  if (cand) memcpy(cand, hp + o_cand, (size_t)n * 10);
  if (ncand) memcpy(ncand, hp + o_nc, n);
  return HEVCDL_OK;
}

namespace {
// Is `p` page-locked host memory (hevcdl_host_alloc, hevcdl_host_register, or the caller's own cudaHostAlloc / cudaHostRegister)?
bool host_is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// Transfers of one picture plane between the caller's strided int16 rows and a dense device plane.  Page-locked caller memory
// is copied directly (cudaMemcpy2DAsync, no host-side pass); anything else goes through the context's pinned scratch (rows
// packed before / unpacked after, `mirror` = the scratch location that mirrors the device plane).
struct PlaneIO {
  struct Pending { const uint8_t *mirror; int16_t *p; int stride, w, h; };
  cudaStream_t st;
  std::vector<Pending> later;
  cudaError_t up(uint8_t *dev, uint8_t *mirror, const int16_t *p, int stride, int w, int h) {
    if (host_is_pinned(p)) return cudaMemcpy2DAsync(dev, (size_t)w * 2, p, (size_t)stride * 2, (size_t)w * 2, h, cudaMemcpyHostToDevice, st);
    for (int r = 0; r < h; r++) memcpy(mirror + (size_t)r * w * 2, p + (size_t)r * stride, (size_t)w * 2);
    return cudaMemcpyAsync(dev, mirror, (size_t)w * h * 2, cudaMemcpyHostToDevice, st);
  }
  cudaError_t down(const uint8_t *dev, uint8_t *mirror, int16_t *p, int stride, int w, int h) {
    if (host_is_pinned(p)) return cudaMemcpy2DAsync(p, (size_t)stride * 2, dev, (size_t)w * 2, (size_t)w * 2, h, cudaMemcpyDeviceToHost, st);
    later.push_back(Pending{mirror, p, stride, w, h});
    return cudaMemcpyAsync(mirror, dev, (size_t)w * h * 2, cudaMemcpyDeviceToHost, st);
  }
  void finish() {                                 // after the stream has been synchronised
    for (const Pending &q : later)
      for (int r = 0; r < q.h; r++) memcpy(q.p + (size_t)r * q.stride, q.mirror + (size_t)r * q.w * 2, (size_t)q.w * 2);
    later.clear();
  }
};

// grow-only scratch of the in-loop entry points; keep > 0: the first `keep` bytes (the resident deblocked picture) survive a growth
int dbf_reserve(hevcdl_ctx *ctx, size_t total, size_t keep) {
  if (total <= ctx->dbfCap) return HEVCDL_OK;
  void *nd = nullptr;
  CK(cudaMalloc(&nd, total));
  if (keep && ctx->dDbf) {
    CK(cudaMemcpyAsync(nd, ctx->dDbf, keep, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  cudaFree(ctx->dDbf);
  ctx->dDbf = nd;
  if (ctx->hDbf) { cudaFreeHost(ctx->hDbf); ctx->hDbf = nullptr; }
  ctx->dbfCap = 0;
  CK(cudaMallocHost(&ctx->hDbf, total));
  ctx->dbfCap = total;
  return HEVCDL_OK;
}

// deblocking (+ optionally the SAO statistics of the deblocked picture against `org`) with one upload and one download;
// the deblocked picture stays at the start of the scratch
int inloop_impl(hevcdl_ctx *ctx, int16_t *y, int sy, int16_t *u, int16_t *v, int sc, int W, int H, const uint8_t *tu_log2, const int8_t *qp,
                int beta_off, int tc_off, int cb_off, int cr_off, const int16_t *oy, const int16_t *ou, const int16_t *ov, int osy, int osc,
                int64_t *stats) {
  if (!ctx || !y || !u || !v || !tu_log2 || !qp || W < 8 || H < 8 || (W & 7) || (H & 7) || W > 8192 || H > 8192 || sy < W || sc < W / 2 ||
      beta_off < -6 || beta_off > 6 || tc_off < -6 || tc_off > 6 || cb_off < -12 || cb_off > 12 || cr_off < -12 || cr_off > 12)
    return HEVCDL_E_INVAL;
  const bool with_stats = stats != nullptr;
  if (with_stats && (!oy || !ou || !ov || osy < W || osc < W / 2)) return HEVCDL_E_INVAL;
  const size_t nu = (size_t)(W / 4) * (H / 4);
  for (size_t i = 0; i < nu; i++)
    if (tu_log2[i] < 2 || tu_log2[i] > 5 || qp[i] < 0 || qp[i] > 51) { ctx->err = "hevcdl_deblock_frame: tu_log2 in 2..5, qp in 0..51"; return HEVCDL_E_INVAL; }
  cudaSetDevice(ctx->cfg.device);
  ctx->loopValid = false;
  const int cw = (W + 63) / 64, chh = (H + 63) / 64, nctu = cw * chh;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t b_y = al((size_t)W * H * 2), b_c = al((size_t)(W / 2) * (H / 2) * 2), b_m = al(nu), b_out = al((size_t)nctu * 3 * 5 * 64 * 8);
  const size_t o_y = 0, o_u = o_y + b_y, o_v = o_u + b_c, o_tu = o_v + b_c, o_qp = o_tu + b_m, o_oy = o_qp + b_m, o_ou = o_oy + b_y, o_ov = o_ou + b_c,
               o_out = o_ov + b_c, total = with_stats ? o_out + b_out : o_oy;
  { const int rc = dbf_reserve(ctx, total, 0); if (rc) return rc; }
  uint8_t *hp = (uint8_t *)ctx->hDbf, *dp = (uint8_t *)ctx->dDbf;
  cudaStream_t st = ctx->stream;
  PlaneIO io{st, {}};
  CK(io.up(dp + o_y, hp + o_y, y, sy, W, H)); CK(io.up(dp + o_u, hp + o_u, u, sc, W / 2, H / 2)); CK(io.up(dp + o_v, hp + o_v, v, sc, W / 2, H / 2));
  memcpy(hp + o_tu, tu_log2, nu);
  memcpy(hp + o_qp, qp, nu);
  CK(cudaMemcpyAsync(dp + o_tu, hp + o_tu, o_oy - o_tu, cudaMemcpyHostToDevice, st));
  if (with_stats) {
    CK(io.up(dp + o_oy, hp + o_oy, oy, osy, W, H)); CK(io.up(dp + o_ou, hp + o_ou, ou, osc, W / 2, H / 2)); CK(io.up(dp + o_ov, hp + o_ov, ov, osc, W / 2, H / 2));
  }
  DbfParams P{(int16_t *)(dp + o_y), (int16_t *)(dp + o_u), (int16_t *)(dp + o_v), W, W / 2, W, H, dp + o_tu, (const int8_t *)(dp + o_qp),
              beta_off, tc_off, cb_off, cr_off};
  const int nv = (W / 8 - 1) * (H / 4), nh = (W / 4) * (H / 8 - 1);
  CK(aux_begin(ctx));
  if (nv > 0) k_dbf<true><<<(nv + 255) / 256, 256, 0, st>>>(P);
  if (nh > 0) k_dbf<false><<<(nh + 255) / 256, 256, 0, st>>>(P);
  if (with_stats) {
    SaoParams S{};
    S.org[0] = (const int16_t *)(dp + o_oy); S.org[1] = (const int16_t *)(dp + o_ou); S.org[2] = (const int16_t *)(dp + o_ov);
    S.src[0] = (const int16_t *)(dp + o_y); S.src[1] = (const int16_t *)(dp + o_u); S.src[2] = (const int16_t *)(dp + o_v);
    S.W = W; S.H = H; S.ctu_w = cw; S.out = (long long *)(dp + o_out);
    k_sao_stats<<<nctu * 3, 256, 0, st>>>(S);
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->evAux1, st));
  ctx->stats.kernel_launches += (nv > 0) + (nh > 0) + (with_stats ? 1 : 0);
  CK(io.down(dp + o_y, hp + o_y, y, sy, W, H)); CK(io.down(dp + o_u, hp + o_u, u, sc, W / 2, H / 2)); CK(io.down(dp + o_v, hp + o_v, v, sc, W / 2, H / 2));
  if (with_stats) CK(cudaMemcpyAsync(hp + o_out, dp + o_out, (size_t)nctu * 3 * 5 * 64 * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  io.finish();
  if (with_stats) memcpy(stats, hp + o_out, (size_t)nctu * 3 * 5 * 64 * 8);
  ctx->loopValid = true; ctx->loopW = W; ctx->loopH = H;
  return HEVCDL_OK;
}
}  // namespace

int hevcdl_deblock_frame(hevcdl_ctx *ctx, int16_t *y, int sy, int16_t *u, int16_t *v, int sc, int W, int H, const uint8_t *tu_log2, const int8_t *qp,
                         int beta_off, int tc_off, int cb_off, int cr_off) {
  return inloop_impl(ctx, y, sy, u, v, sc, W, H, tu_log2, qp, beta_off, tc_off, cb_off, cr_off, nullptr, nullptr, nullptr, 0, 0, nullptr);
}

int hevcdl_inloop_frame(hevcdl_ctx *ctx, int16_t *y, int sy, int16_t *u, int16_t *v, int sc, int W, int H, const uint8_t *tu_log2, const int8_t *qp,
                        int beta_off, int tc_off, int cb_off, int cr_off, const int16_t *oy, const int16_t *ou, const int16_t *ov, int osy, int osc,
                        int64_t *stats) {
  if (!stats) return HEVCDL_E_INVAL;
  return inloop_impl(ctx, y, sy, u, v, sc, W, H, tu_log2, qp, beta_off, tc_off, cb_off, cr_off, oy, ou, ov, osy, osc, stats);
}

int hevcdl_sao_apply(hevcdl_ctx *ctx, const int16_t *sy, const int16_t *su, const int16_t *sv, int ssy, int ssc, int16_t *ry, int16_t *ru, int16_t *rv,
                     int rsy, int rsc, int W, int H, const hevcdl_sao_param *params) {
  const bool resident = ctx && !sy && !su && !sv;        // source = the deblocked picture hevcdl_inloop_frame left on the device
  if (!ctx || (!resident && (!sy || !su || !sv || ssy < W || ssc < W / 2)) || !ry || !ru || !rv || !params || W < 8 || H < 8 || (W & 7) || (H & 7) ||
      W > 8192 || H > 8192 || rsy < W || rsc < W / 2)
    return HEVCDL_E_INVAL;
  if (resident && !(ctx->loopValid && ctx->loopW == W && ctx->loopH == H)) {
    ctx->err = "hevcdl_sao_apply: no deblocked picture of this size is resident (hevcdl_inloop_frame must be the previous in-loop call)";
    return HEVCDL_E_NOFRAME;
  }
  const int cw = (W + 63) / 64, chh = (H + 63) / 64, nctu = cw * chh;
  for (int i = 0; i < nctu * 3; i++)
    if (params[i].type < -1 || params[i].type > 4) { ctx->err = "hevcdl_sao_apply: type in -1..4"; return HEVCDL_E_INVAL; }
  cudaSetDevice(ctx->cfg.device);
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t b_y = al((size_t)W * H * 2), b_c = al((size_t)(W / 2) * (H / 2) * 2), b_t = al((size_t)nctu * 3), b_o = al((size_t)nctu * 3 * 32);
  // the source picture sits at the start of the scratch in both cases (the layout hevcdl_inloop_frame leaves behind)
  const size_t o_sy = 0, o_su = o_sy + b_y, o_sv = o_su + b_c, o_t = o_sv + b_c, o_o = o_t + b_t, o_ry = o_o + b_o, o_ru = o_ry + b_y, o_rv = o_ru + b_c,
               total = o_rv + b_c;
  { const int rc = dbf_reserve(ctx, total, resident ? o_t : 0); if (rc) return rc; }
  uint8_t *hp = (uint8_t *)ctx->hDbf, *dp = (uint8_t *)ctx->dDbf;
  cudaStream_t st = ctx->stream;
  PlaneIO io{st, {}};
  if (!resident) {
    ctx->loopValid = false;
    CK(io.up(dp + o_sy, hp + o_sy, sy, ssy, W, H)); CK(io.up(dp + o_su, hp + o_su, su, ssc, W / 2, H / 2)); CK(io.up(dp + o_sv, hp + o_sv, sv, ssc, W / 2, H / 2));
  }
  for (int i = 0; i < nctu * 3; i++) {
    hp[o_t + i] = (uint8_t)params[i].type;
    memcpy(hp + o_o + (size_t)i * 32, params[i].offset, 32);
  }
  CK(cudaMemcpyAsync(dp + o_t, hp + o_t, o_ry - o_t, cudaMemcpyHostToDevice, st));
  SaoApplyParams P{};
  P.src[0] = (const int16_t *)(dp + o_sy); P.src[1] = (const int16_t *)(dp + o_su); P.src[2] = (const int16_t *)(dp + o_sv);
  P.res[0] = (int16_t *)(dp + o_ry); P.res[1] = (int16_t *)(dp + o_ru); P.res[2] = (int16_t *)(dp + o_rv);
  P.W = W; P.H = H; P.ctu_w = cw; P.type = (const int8_t *)(dp + o_t); P.offset = (const int8_t *)(dp + o_o);
  CK(aux_begin(ctx));
  k_sao_apply<<<nctu * 3, 256, 0, st>>>(P);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->evAux1, st));
  ctx->stats.kernel_launches++;
  CK(io.down(dp + o_ry, hp + o_ry, ry, rsy, W, H)); CK(io.down(dp + o_ru, hp + o_ru, ru, rsc, W / 2, H / 2)); CK(io.down(dp + o_rv, hp + o_rv, rv, rsc, W / 2, H / 2));
  CK(cudaStreamSynchronize(st));
  io.finish();
  return HEVCDL_OK;
}

int hevcdl_intra_pred(hevcdl_ctx *ctx, int n, const hevcdl_pred_req *reqs, const int16_t *lines, size_t nline, int16_t *pred, size_t npred) {
  static_assert(sizeof(hevcdl_pred_req) == sizeof(IntraPredReq), "hevcdl_pred_req layout");
  if (!ctx || n < 0 || (n && (!reqs || !lines || !pred))) return HEVCDL_E_INVAL;
  if (n == 0) return HEVCDL_OK;
  for (int i = 0; i < n; i++) {
    const hevcdl_pred_req &r = reqs[i];
    const size_t sz = r.log2_size >= 2 && r.log2_size <= 6 ? (size_t)1 << r.log2_size : 0;
    if (!sz || r.mode > 34 || (size_t)r.line_offset + 4 * sz + 1 > nline || (size_t)r.pred_offset + sz * sz > npred) {
      ctx->err = "hevcdl_intra_pred: log2_size 2..6, mode 0..34, line and block inside the buffers";
      return HEVCDL_E_INVAL;
    }
  }
  cudaSetDevice(ctx->cfg.device);
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t b_req = al((size_t)n * sizeof(hevcdl_pred_req)), b_line = al(nline * 2), b_pred = al(npred * 2);
  const size_t o_req = 0, o_line = o_req + b_req, o_pred = o_line + b_line, total = o_pred + b_pred;
  if (total > ctx->tqCap) {                       // shares the TU core's grow-only scratch
    cudaFree(ctx->dTq); ctx->dTq = nullptr; ctx->tqCap = 0;
    if (ctx->hTq) { cudaFreeHost(ctx->hTq); ctx->hTq = nullptr; }
    const size_t cap = total < (1u << 20) ? (1u << 20) : total + total / 2;
    CK(cudaMalloc(&ctx->dTq, cap));
    CK(cudaMallocHost(&ctx->hTq, cap));
    ctx->tqCap = cap;
  }
  uint8_t *hp = (uint8_t *)ctx->hTq, *dp = (uint8_t *)ctx->dTq;
  memcpy(hp + o_req, reqs, (size_t)n * sizeof(hevcdl_pred_req));
  memcpy(hp + o_line, lines, nline * 2);
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(dp, hp, o_pred, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(dp + o_pred, 0, b_pred, st));          // gaps between blocks come back as zeros
  const int grid = (n + PRED_WARPS - 1) / PRED_WARPS < 8 * ctx->numSMs ? (n + PRED_WARPS - 1) / PRED_WARPS : 8 * ctx->numSMs;
  CK(aux_begin(ctx));
  k_intra_pred<<<grid, PRED_WARPS * 32, 0, st>>>(n, (const IntraPredReq *)(dp + o_req), (const int16_t *)(dp + o_line), (int16_t *)(dp + o_pred));
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->evAux1, st));
  ctx->stats.kernel_launches++;
  CK(cudaMemcpyAsync(hp + o_pred, dp + o_pred, npred * 2, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  memcpy(pred, hp + o_pred, npred * 2);
  return HEVCDL_OK;
}

int hevcdl_sao_stats(hevcdl_ctx *ctx, const int16_t *oy, const int16_t *ou, const int16_t *ov, int osy, int osc, const int16_t *ry, const int16_t *ru,
                     const int16_t *rv, int rsy, int rsc, int W, int H, int64_t *stats) {
  if (!ctx || !oy || !ou || !ov || !ry || !ru || !rv || !stats || W < 8 || H < 8 || (W & 7) || (H & 7) || W > 8192 || H > 8192 || osy < W || rsy < W ||
      osc < W / 2 || rsc < W / 2)
    return HEVCDL_E_INVAL;
  cudaSetDevice(ctx->cfg.device);
  const int cw = (W + 63) / 64, chh = (H + 63) / 64, nctu = cw * chh;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t b_y = al((size_t)W * H * 2), b_c = al((size_t)(W / 2) * (H / 2) * 2), b_out = al((size_t)nctu * 3 * 5 * 64 * 8);
  const size_t o_oy = 0, o_ou = o_oy + b_y, o_ov = o_ou + b_c, o_ry = o_ov + b_c, o_ru = o_ry + b_y, o_rv = o_ru + b_c, o_out = o_rv + b_c, total = o_out + b_out;
  ctx->loopValid = false;                          // shares the in-loop entry points' grow-only scratch
  { const int rc = dbf_reserve(ctx, total, 0); if (rc) return rc; }
  uint8_t *hp = (uint8_t *)ctx->hDbf, *dp = (uint8_t *)ctx->dDbf;
  cudaStream_t st = ctx->stream;
  PlaneIO io{st, {}};
  CK(io.up(dp + o_oy, hp + o_oy, oy, osy, W, H)); CK(io.up(dp + o_ou, hp + o_ou, ou, osc, W / 2, H / 2)); CK(io.up(dp + o_ov, hp + o_ov, ov, osc, W / 2, H / 2));
  CK(io.up(dp + o_ry, hp + o_ry, ry, rsy, W, H)); CK(io.up(dp + o_ru, hp + o_ru, ru, rsc, W / 2, H / 2)); CK(io.up(dp + o_rv, hp + o_rv, rv, rsc, W / 2, H / 2));
  SaoParams P{};
  P.org[0] = (const int16_t *)(dp + o_oy); P.org[1] = (const int16_t *)(dp + o_ou); P.org[2] = (const int16_t *)(dp + o_ov);
  P.src[0] = (const int16_t *)(dp + o_ry); P.src[1] = (const int16_t *)(dp + o_ru); P.src[2] = (const int16_t *)(dp + o_rv);
  P.W = W; P.H = H; P.ctu_w = cw; P.out = (long long *)(dp + o_out);
  CK(aux_begin(ctx));
  k_sao_stats<<<nctu * 3, 256, 0, st>>>(P);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->evAux1, st));
  ctx->stats.kernel_launches++;
  CK(cudaMemcpyAsync(hp + o_out, dp + o_out, (size_t)nctu * 3 * 5 * 64 * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  memcpy(stats, hp + o_out, (size_t)nctu * 3 * 5 * 64 * 8);
  return HEVCDL_OK;
}

int hevcdl_tu_code_rdoq(hevcdl_ctx *ctx, int n, const hevcdl_tu *tus, const hevcdl_tu_rdoq *rdoq, const int32_t *est, int n_est,
                        const int16_t *resi, size_t nelem, int32_t *coeff, int16_t *level, int32_t *deq, int16_t *rec, uint32_t *abs_sum,
                        uint64_t *ssd) {
  if (!ctx || n < 0 || (n && (!tus || !resi || !level || !rec || !abs_sum)) || (rdoq && (!est || n_est <= 0))) return HEVCDL_E_INVAL;
  if (n == 0) return HEVCDL_OK;
  cudaSetDevice(ctx->cfg.device);
  for (int i = 0; i < n; i++) {
    const hevcdl_tu &t = tus[i];
    if (t.log2_size < 2 || t.log2_size > 5 || t.qp > 51 || (size_t)t.offset + ((size_t)1 << (2 * t.log2_size)) > nelem ||
        ((t.flags & HEVCDL_TU_TSKIP) && t.log2_size != 2) || (t.offset & 1) || ((t.flags & HEVCDL_TU_RDOQ) && !rdoq)) {
      ctx->err = "hevcdl_tu: log2_size 2..5, qp 0..51, even offset inside nelem, transform skip only for 4x4, HEVCDL_TU_RDOQ only with rdoq parameters";
      return HEVCDL_E_INVAL;
    }
    if (rdoq && (t.flags & HEVCDL_TU_RDOQ) &&
        (rdoq[i].est_index >= (uint32_t)n_est || rdoq[i].channel > 1 || rdoq[i].scan_type > 2 || rdoq[i].ctx_cbf >= 10 || !(rdoq[i].lambda > 0) ||
         (rdoq[i].scan_type != 0 && t.log2_size > 3))) {
      ctx->err = "hevcdl_tu_rdoq: est_index < n_est, channel 0/1, scan_type 0..2 (non-diagonal only for 4x4 / 8x8), ctx_cbf < 10, lambda > 0";
      return HEVCDL_E_INVAL;
    }
  }
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t b_tus = al((size_t)n * sizeof(hevcdl_tu)), b_resi = al(nelem * 2), b_rq = rdoq ? al((size_t)n * sizeof(hevcdl_tu_rdoq)) : 0,
               b_est = rdoq ? al((size_t)n_est * HEVCDL_EST_INTS * 4) : 0, b_coeff = coeff ? al(nelem * 4) : 0, b_level = al(nelem * 2),
               b_deq = deq ? al(nelem * 4) : 0, b_rec = al(nelem * 2), b_asum = al((size_t)n * 4), b_ssd = ssd ? al((size_t)n * 8) : 0;
  const size_t o_tus = 0, o_resi = o_tus + b_tus, o_rq = o_resi + b_resi, o_est = o_rq + b_rq, o_coeff = o_est + b_est, o_level = o_coeff + b_coeff,
               o_deq = o_level + b_level, o_rec = o_deq + b_deq, o_asum = o_rec + b_rec, o_ssd = o_asum + b_asum, total = o_ssd + b_ssd;
  if (total > ctx->tqCap) {
    cudaFree(ctx->dTq); ctx->dTq = nullptr; ctx->tqCap = 0;
    if (ctx->hTq) { cudaFreeHost(ctx->hTq); ctx->hTq = nullptr; }
    const size_t cap = total < (1u << 20) ? (1u << 20) : total + total / 2;
    CK(cudaMalloc(&ctx->dTq, cap));
    CK(cudaMallocHost(&ctx->hTq, cap));
    ctx->tqCap = cap;
  }
  int grid = (n + TQ_WARPS - 1) / TQ_WARPS < 8 * ctx->numSMs ? (n + TQ_WARPS - 1) / TQ_WARPS : 8 * ctx->numSMs;
  if (rdoq) {                                      // 48 KB of global scratch per warp: fewer, longer-lived warps
    if (grid > 2 * ctx->numSMs) grid = 2 * ctx->numSMs;
    if (grid * TQ_WARPS > ctx->rdoqWarps) {
      cudaFree(ctx->dRdoq); ctx->dRdoq = nullptr; ctx->rdoqWarps = 0;
      CK(cudaMalloc(&ctx->dRdoq, (size_t)grid * TQ_WARPS * sizeof(RdoqScratch)));
      ctx->rdoqWarps = grid * TQ_WARPS;
    }
  }
  uint8_t *hp = (uint8_t *)ctx->hTq, *dp = (uint8_t *)ctx->dTq;
  memcpy(hp + o_tus, tus, (size_t)n * sizeof(hevcdl_tu));
  memcpy(hp + o_resi, resi, nelem * 2);
  if (rdoq) {
    memcpy(hp + o_rq, rdoq, (size_t)n * sizeof(hevcdl_tu_rdoq));
    memcpy(hp + o_est, est, (size_t)n_est * HEVCDL_EST_INTS * 4);
  }
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(dp, hp, o_coeff, cudaMemcpyHostToDevice, st));
  // gaps between TU blocks (if the caller's offsets leave any) come back as zeros, not as stale bytes
  CK(cudaMemsetAsync(dp + o_coeff, 0, total - o_coeff, st));
  CK(aux_begin(ctx));
  k_tu_code<<<grid, TQ_WARPS * 32, sizeof(TqBlockS), st>>>(n, (const hevcdl_tu *)(dp + o_tus), (const int16_t *)(dp + o_resi),
                                                            coeff ? (int32_t *)(dp + o_coeff) : nullptr, (int16_t *)(dp + o_level),
                                                            deq ? (int32_t *)(dp + o_deq) : nullptr, (int16_t *)(dp + o_rec),
                                                            (uint32_t *)(dp + o_asum), ssd ? (uint64_t *)(dp + o_ssd) : nullptr,
                                                            rdoq ? (const hevcdl_tu_rdoq *)(dp + o_rq) : nullptr,
                                                            rdoq ? (const int *)(dp + o_est) : nullptr, (RdoqScratch *)ctx->dRdoq);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->evAux1, st));
  ctx->stats.kernel_launches++;
  CK(cudaMemcpyAsync(hp + o_coeff, dp + o_coeff, total - o_coeff, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (coeff) memcpy(coeff, hp + o_coeff, nelem * 4);
  memcpy(level, hp + o_level, nelem * 2);
  if (deq) memcpy(deq, hp + o_deq, nelem * 4);
  memcpy(rec, hp + o_rec, nelem * 2);
  memcpy(abs_sum, hp + o_asum, (size_t)n * 4);
  if (ssd) memcpy(ssd, hp + o_ssd, (size_t)n * 8);
  return HEVCDL_OK;
}

int hevcdl_tu_code(hevcdl_ctx *ctx, int n, const hevcdl_tu *tus, const int16_t *resi, size_t nelem, int32_t *coeff, int16_t *level,
                   int32_t *deq, int16_t *rec, uint32_t *abs_sum, uint64_t *ssd) {
  return hevcdl_tu_code_rdoq(ctx, n, tus, nullptr, nullptr, 0, resi, nelem, coeff, level, deq, rec, abs_sum, ssd);
}

int hevcdl_host_register(void *p, size_t bytes) {
  if (!p || !bytes) return HEVCDL_E_INVAL;
  if (host_is_pinned(p)) return HEVCDL_OK;
  const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) { cudaGetLastError(); return HEVCDL_E_CUDA; }
  return HEVCDL_OK;
}

int hevcdl_host_unregister(void *p) {
  if (!p) return HEVCDL_E_INVAL;
  const cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) { cudaGetLastError(); return HEVCDL_E_CUDA; }
  return HEVCDL_OK;
}

int hevcdl_last_aux_ms(hevcdl_ctx *ctx, float *ms) {
  if (!ctx || !ms) return HEVCDL_E_INVAL;
  if (!ctx->evAux0) return HEVCDL_E_NOFRAME;
  cudaSetDevice(ctx->cfg.device);
  CK(cudaEventSynchronize(ctx->evAux1));
  CK(cudaEventElapsedTime(ms, ctx->evAux0, ctx->evAux1));
  return HEVCDL_OK;
}

int hevcdl_bench_resident(hevcdl_ctx *ctx, const int *frames, int nframes, int iters, float ms[3], int *launches) {
  if (!ctx || !frames || nframes <= 0 || iters <= 0 || !ms) return HEVCDL_E_INVAL;
  cudaSetDevice(ctx->cfg.device);
  std::vector<Slot *> sl(nframes);
  for (int i = 0; i < nframes; i++) {
    sl[i] = find_slot(ctx, frames[i]);
    if (!sl[i]) return HEVCDL_E_NOFRAME;
    int rc = finish_slot(ctx, sl[i]);
    if (rc) return rc;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int nl = 0;
  // pass 1: whole pipeline, one pair of events around all iterations
  CK(cudaEventRecord(e0, ctx->stream));
  std::vector<Slot *> grp(ctx->batch);
  auto group = [&](int it) {                     // the next <= batch distinct slots of the rotation
    int n = 0;
    for (; n < ctx->batch && it + n < iters && n < nframes; n++) grp[n] = sl[(it + n) % nframes];
    return n;
  };
  Slot *last = nullptr;
  for (int it = 0; it < iters;) {
    const int n = group(it);
    const int rc = launch_pipeline(ctx, grp.data(), n, false, &nl);
    if (rc) return rc;
    it += n; last = grp[n - 1];
  }
  if (ctx->rmd && last) CK(cudaStreamWaitEvent(ctx->stream, last->evRmd, 0));   // the RMD stream's last launch is part of the pass
  CK(cudaEventRecord(e1, ctx->stream));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  CK(cudaEventElapsedTime(&ms[0], e0, e1));
  // pass 2: same work with per-stage events (stage split; not used for the headline)
  double cnn = 0, rmd = 0;
  for (int it = 0; it < iters;) {
    const int n = group(it);
    it += n;
    Slot &s = *grp[0];
    { const int rc = launch_pipeline(ctx, grp.data(), n, true, &nl); if (rc) return rc; }
    CK(cudaEventSynchronize(s.evT2));
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, s.evT0, s.evT1));
    CK(cudaEventElapsedTime(&b, s.evT1, s.evT2));
    cnn += a; rmd += b;
  }
  ms[1] = (float)cnn; ms[2] = (float)rmd;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (launches) *launches = nl / 2;
  ctx->stats.kernel_launches += nl;
  return HEVCDL_OK;
}

int hevcdl_bench_e2e(hevcdl_ctx *ctx, int first_id, int iters, int depth, int nbuf, const uint8_t *const *y,
                     const uint8_t *const *u, const uint8_t *const *v, int stride_y, int stride_c, double *seconds,
                     uint64_t *d2h_bytes, uint64_t *checksum) {
  if (!ctx || iters <= 0 || depth <= 0 || nbuf <= 0 || !y || !u || !v || !seconds) return HEVCDL_E_INVAL;
  if (depth > (int)ctx->slots.size()) depth = (int)ctx->slots.size();
  uint64_t bytes = 0, chk = 0;
  const int nctu = ctx->geo.nctu;
  auto consume = [&](int frame) -> int {
    hevcdl_frame_view fv;
    int rc = hevcdl_frame_view_get(ctx, frame, ctx->cfg.rmd, &fv);
    if (rc) return rc;
    bytes += (uint64_t)nctu * 16 + (fv.logits ? (uint64_t)nctu * 64 * sizeof(float) : 0);
    chk += fv.labels[(size_t)nctu * 16 - 1];
    if (fv.ctu_off) bytes += ((uint64_t)nctu + 1) * sizeof(int32_t);
    if (fv.npu > 0) {
      bytes += (uint64_t)fv.npu * (sizeof(hevcdl_pu) + 8 + (fv.satd ? 35 * sizeof(uint32_t) : 0));
      chk += fv.cand[(size_t)fv.npu * 8 - 8] + (fv.satd ? fv.satd[(size_t)fv.npu * 35 - 1] : 0);
    }
    return hevcdl_release_frame(ctx, frame);
  };
  const auto t0 = std::chrono::steady_clock::now();
  int done = 0;
  for (int i = 0; i < iters; i++) {
    int rc = hevcdl_submit_frame_u8(ctx, first_id + i, y[i % nbuf], stride_y, u[i % nbuf], v[i % nbuf], stride_c);
    if (rc) return rc;
    if (i + 1 - done >= depth) { rc = consume(first_id + done); if (rc) return rc; done++; }
  }
  for (; done < iters; done++) { int rc = consume(first_id + done); if (rc) return rc; }
  *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (d2h_bytes) *d2h_bytes = bytes;
  if (checksum) *checksum = chk;
  return HEVCDL_OK;
}

int hevcdl_debug_copy(hevcdl_ctx *ctx, int which, void *dst, size_t nbytes, size_t *size) {
  if (!ctx || which < 0 || which > 2) return HEVCDL_E_INVAL;
  if (ctx->cfg.precision != HEVCDL_PREC_BF16_TC) { ctx->err = "intermediates exist only on the tensor-core path"; return HEVCDL_E_INVAL; }
  cudaSetDevice(ctx->cfg.device);
  if (ctx->batch != 1) { ctx->err = "debug_copy needs a batch=1 context"; return HEVCDL_E_INVAL; }
  const size_t sz[3] = {(size_t)ctx->geo.nctu * CAT_BYTES, (size_t)ctx->geo.nctu * A2_BYTES, (size_t)ctx->tc.npad * 4096};
  const uint8_t *src[3] = {ctx->tc.cat, ctx->tc.a2, ctx->tc.feats};
  if (size) *size = sz[which];
  if (!dst) return HEVCDL_OK;
  if (nbytes < sz[which]) return HEVCDL_E_INVAL;
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(dst, src[which], sz[which], cudaMemcpyDeviceToHost));
  return HEVCDL_OK;
}

int hevcdl_debug_rerun_rmd(hevcdl_ctx *ctx, int frame, const uint8_t *labels) {
  if (!ctx || !labels) return HEVCDL_E_INVAL;
  if (!ctx->cfg.rmd) { ctx->err = "context created with rmd=0"; return HEVCDL_E_INVAL; }
  cudaSetDevice(ctx->cfg.device);
  Slot *s = find_slot(ctx, frame);
  if (!s) return HEVCDL_E_NOFRAME;
  int rc = finish_slot(ctx, s);
  if (rc) return rc;
  CK(cudaEventSynchronize(s->evRmd));
  const FrameGeom g = ctx->geo;
  cudaStream_t st = ctx->stream;
  memcpy(s->hLabels, labels, (size_t)g.nctu * 16);
  CK(cudaMemcpyAsync(s->dLabels, s->hLabels, (size_t)g.nctu * 16, cudaMemcpyHostToDevice, st));
  k_rmd_counts<<<(g.nctu + 255) / 256, 256, 0, st>>>(s->dLabels, g, s->dCtuCnt);
  RmdBatch rb{};
  rb.n = 1;
  rb.Y[0] = s->dY; rb.labels[0] = s->dLabels; rb.ctu_cnt[0] = s->dCtuCnt; rb.ctu_off[0] = s->dCtuOff;
  rb.pus[0] = s->dPus; rb.satd[0] = s->dSatd; rb.cand[0] = s->dCand;
  CK(launch_pdl(k_rmd_plan, (g.nctu + 7) / 8, 256, 0, st, rb, g, ctx->rmdBlocks, s->dItems, s->dCtrl));
  CK(launch_pdl(k_rmd_items, ctx->rmdBlocks, RMD_BW * 32, 0, st, rb, g, ctx->pitch, (const RmdItem *)s->dItems, s->dCtrl));
  CK(cudaGetLastError());
  ctx->stats.kernel_launches += 3;
  CK(cudaEventRecord(s->evRmd, st));
  CK(cudaMemcpyAsync(s->hCtuOff, s->dCtuOff, ((size_t)g.nctu + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  s->pusFetched = false;
  return HEVCDL_OK;
}

int hevcdl_get_stats(hevcdl_ctx *ctx, hevcdl_stats_t *out) {
  if (!ctx || !out) return HEVCDL_E_INVAL;
  *out = ctx->stats;
  return HEVCDL_OK;
}

void *hevcdl_stream(hevcdl_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

}  // extern "C"
