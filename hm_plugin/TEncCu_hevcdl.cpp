// hm_plugin/TEncCu_hevcdl.cpp -- the reference-side binding of libhevcdl.so.
//
// A from-scratch definition of
//     Void TEncCu::compressCtu( Int m_iFrame, TComDataCU* pCtu )
// (declared at HM_dl/source/Lib/TLibEncoder/TEncCu.h:120, reference body at TEncCu.cpp:234-287)
// that obtains the 16 CU-depth labels of the CTU from the B200 library instead of busy-polling
// ./pred/<frame>/ctu<addr>.txt (TEncCu.cpp:243-252), then runs the reference's own pruned quadtree
// search (xCompressCU, TEncCu.cpp:470) exactly as the reference does.  Nothing else in HM changes:
// same caller (TEncSlice::compressSlice, TEncSlice.cpp:879), same pre/post-conditions, same
// bitstream for the same labels.
//
// How it is linked without editing the reference: hm_plugin/Makefile compiles the reference's
// TEncCu.cpp with -DcompressCtu=compressCtu_filehandshake (its file-polling body keeps existing under
// another name) and links this translation unit's compressCtu in its place.  A maintainer applying
// the change by hand would simply replace the body at TEncCu.cpp:234-287 with the one below
// (INTEGRATION.md).
//
// Configuration comes from the environment, because the signature leaves no room for it:
//   HEVCDL_WEIGHTS    path of the HDLW weight blob (default: weights/hevc_encoder_model.hdlw next to the repo root
//                     baked in at build time as HEVCDL_DEFAULT_WEIGHTS)
//   HEVCDL_DEVICE     CUDA ordinal (default 0)
//   HEVCDL_PRECISION  fp32 (default: tightest parity with the torch sidecar) | bf16 (tcgen05 tensor cores)
//   HEVCDL_BOUNDARY_FIX 1 = raise labels of picture-edge CTUs so partial CTUs tile (default 0 = reference)
//   HEVCDL_RMD        1 = run the batched 35-mode SATD pass on the B200 and let estIntraPredLumaQT's first pass take its
//                     per-mode SATDs from it (hm_plugin/rmd_hook.h; references are ORIGINAL pixels, so mode
//                     decisions follow the +-1 % BD-rate clause, not the bit-exact one);
//                     2 = exact mode: per PU, HM's reconstructed reference samples go to hevcdl_rmd_exact (bit-exact,
//                     byte-identical bitstream, one synchronous call per PU) (default 0: HM's own pass)
// There is no fallback: any library failure aborts the encoder with the library's error text.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

#include "TLibEncoder/TEncCu.h"
#include "TLibEncoder/TEncTop.h"

#include "hevcdl.h"
#include "rmd_hook.h"

namespace {

struct HevcdlSession {
  hevcdl_ctx *ctx = nullptr;
  int width = 0, height = 0;
  int frame = -1;               // frame currently resident on the device (-1: none)
  bool gpu_rmd = false;         // HEVCDL_RMD=1: first-pass SATDs come from the device (batched, original-pixel references)
  bool exact_rmd = false;       // HEVCDL_RMD=2: ... from hevcdl_rmd_exact fed HM's reconstructed references, PU by PU
  unsigned ex_x = ~0u, ex_y = ~0u, ex_n = 0;      // PU whose 35 SATDs are cached in ex_satd
  uint32_t ex_satd[35];
  unsigned long long exact_calls = 0;
  hevcdl_frame_view view;       // results of `frame` (pinned host memory owned by the library)
  bool have_view = false;
  int ctu_first = 0, ctu_count = 0, cursor = 0;   // PU range of the CTU being compressed + last hit
  unsigned long long hook_hits = 0, hook_misses = 0;

  static void die(const char *what, int rc, hevcdl_ctx *c) {
    fprintf(stderr, "hevcdl: %s failed: %s (%s)\n", what, hevcdl_status_str(rc), hevcdl_last_error(c));
    exit(EXIT_FAILURE);
  }

  void open(int w, int h) {
    hevcdl_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = HEVCDL_ABI_VERSION;
    const char *e;
    cfg.device = (e = getenv("HEVCDL_DEVICE")) ? atoi(e) : 0;
    cfg.width = w; cfg.height = h;
    cfg.slots = 2;
    cfg.precision = ((e = getenv("HEVCDL_PRECISION")) && !strcmp(e, "bf16")) ? HEVCDL_PREC_BF16_TC : HEVCDL_PREC_FP32;
    const int rmd_mode = (e = getenv("HEVCDL_RMD")) ? atoi(e) : 0;
    cfg.rmd = rmd_mode == 1;
    gpu_rmd = rmd_mode == 1;
    exact_rmd = rmd_mode == 2;
    cfg.boundary_fix = (e = getenv("HEVCDL_BOUNDARY_FIX")) ? atoi(e) : 0;
    // weights: HEVCDL_WEIGHTS, else the path baked in at build time, else <dir of this executable>/../../weights/
    static char relpath[4096];
    cfg.weights_path = (e = getenv("HEVCDL_WEIGHTS")) ? e : HEVCDL_DEFAULT_WEIGHTS;
    if (!e && access(cfg.weights_path, R_OK) != 0) {
      const ssize_t len = readlink("/proc/self/exe", relpath, sizeof(relpath) - 64);
      if (len > 0) {
        relpath[len] = 0;
        char *slash = strrchr(relpath, '/');
        if (slash) { strcpy(slash, "/../../weights/hevc_encoder_model.hdlw"); cfg.weights_path = relpath; }
      }
    }
    const int rc = hevcdl_create(&cfg, &ctx);
    if (rc) die("hevcdl_create", rc, nullptr);
    width = w; height = h;
  }

  // Hand the picture's ORIGINAL planes (the same ones xCompressCU reads at TEncCu.cpp:484) to the device.
  // This replaces gen_frames.py:21 (ffmpeg dump) and the sidecar's whole per-frame loop (use_model.py:74-127).
  void begin_frame(int id, TComPicYuv *org) {
    const int w = org->getWidth(COMPONENT_Y), h = org->getHeight(COMPONENT_Y);
    if (!ctx) open(w, h);
    if (w != width || h != height) { fprintf(stderr, "hevcdl: picture size changed mid-sequence\n"); exit(EXIT_FAILURE); }
    if (frame >= 0) { const int rc = hevcdl_release_frame(ctx, frame); if (rc) die("hevcdl_release_frame", rc, ctx); }
    const int rc = hevcdl_submit_frame_pel16(ctx, id, org->getAddr(COMPONENT_Y), org->getStride(COMPONENT_Y),
                                             org->getAddr(COMPONENT_Cb), org->getAddr(COMPONENT_Cr), org->getStride(COMPONENT_Cb));
    if (rc) die("hevcdl_submit_frame_pel16", rc, ctx);
    frame = id;
    have_view = false;
  }

  // PU range of one CTU in the frame's device results (fetched once per frame, zero-copy)
  void begin_ctu(int addr) {
    if (!gpu_rmd) return;
    if (!have_view) {
      const int rc = hevcdl_frame_view_get(ctx, frame, 1, &view);
      if (rc) die("hevcdl_frame_view_get", rc, ctx);
      have_view = true;
    }
    ctu_first = view.ctu_off[addr];
    ctu_count = view.ctu_off[addr + 1] - ctu_first;
    cursor = 0;
  }

  // SATD of (PU at picture position x,y of the given size, mode); PUs are queried in the order the device listed them
  bool lookup(unsigned x, unsigned y, unsigned size, unsigned mode, unsigned *sad) {
    for (int k = 0; k < ctu_count; k++) {
      const int i = (cursor + k) % ctu_count;
      const hevcdl_pu &p = view.pus[ctu_first + i];
      if (p.x == x && p.y == y && p.size == size) {
        cursor = i;
        *sad += view.satd[(size_t)(ctu_first + i) * 35 + mode];
        hook_hits++;
        return true;
      }
    }
    hook_misses++;
    return false;                 // not on the device's list (cannot happen for consistent labels): HM computes it
  }

  ~HevcdlSession() {
    if (ctx) {
      if (getenv("HEVCDL_VERBOSE")) {
        hevcdl_stats_t st;
        if (!hevcdl_get_stats(ctx, &st))
          fprintf(stderr, "hevcdl: %llu frames, %llu CTUs, CNN %.3f ms, RMD %.3f ms device time, %llu kernel launches, "
                          "first-pass SATDs served %llu / missed %llu, exact PU calls %llu\n",
                  (unsigned long long)st.frames, (unsigned long long)st.ctus, st.ms_cnn, st.ms_rmd,
                  (unsigned long long)st.kernel_launches, hook_hits, hook_misses, exact_calls);
      }
      hevcdl_destroy(ctx);
    }
  }
};

HevcdlSession g_session;   // one encoder thread, one TEncCu instance (TEncTop.h:93): a process-wide session is enough

}  // namespace

Void TEncCu::compressCtu( Int m_iFrame, TComDataCU* pCtu )
{
  const UInt ctuRsAddr = pCtu->getCtuRsAddr();
  m_ppcBestCU[0]->initCtu( pCtu->getPic(), ctuRsAddr );
  m_ppcTempCU[0]->initCtu( pCtu->getPic(), ctuRsAddr );

  // First CTU of a picture the device has not seen yet: upload it; every kernel of the frame is queued
  // behind the copy and the labels of all CTUs come back in one transfer.
  if ( g_session.frame != m_iFrame )
  {
    g_session.begin_frame( m_iFrame, pCtu->getPic()->getPicYuvOrg() );
  }

  uint8_t depth8[16];
  const int rc = hevcdl_ctu_labels( g_session.ctx, m_iFrame, (int)ctuRsAddr, depth8 );   // blocks on the frame's event
  if ( rc ) HevcdlSession::die( "hevcdl_ctu_labels", rc, g_session.ctx );
  UInt label[16];                                   // same lifetime as the reference's stack array (TEncCu.cpp:247)
  for ( Int i = 0; i < 16; i++ ) label[i] = depth8[i];
  m_ppcBestCU[0]->set_pred( label );
  g_session.begin_ctu( (int)ctuRsAddr );

  DEBUG_STRING_NEW(sDebug)
  xCompressCU( m_ppcBestCU[0], m_ppcTempCU[0], 0 DEBUG_STRING_PASS_INTO(sDebug) );
  DEBUG_STRING_OUTPUT(std::cout, sDebug)

#if ADAPTIVE_QP_SELECTION
  if ( m_pcEncCfg->getUseAdaptQpSelect() && pCtu->getSlice()->getSliceType() != I_SLICE )
  {
    xCtuCollectARLStats( pCtu );
  }
#endif
}

// First-pass SATD hook (rmd_hook.h): called from the reference's estIntraPredLumaQT mode loop.
bool hevcdl_hm_rmd_satd( TComPrediction* pred, TComDataCU* pcCU, unsigned x0InCu, unsigned y0InCu, unsigned width, unsigned mode,
                         const short* org, unsigned orgStride, unsigned* sad )
{
  HevcdlSession &S = g_session;
  const unsigned x = pcCU->getCUPelX() + x0InCu, y = pcCU->getCUPelY() + y0InCu;
  if ( S.exact_rmd && S.ctx )
  {
    if ( mode == 0 || S.ex_x != x || S.ex_y != y || S.ex_n != width )   // first mode of a PU: one device call for all 35
    {
      const unsigned n = width, roiW = 2 * n + 1;
      // HM's unfiltered reference samples of this PU (TComPattern.cpp:166-176): (2n+1) x (2n+1) raster, row 0 = corner + above,
      // column 0 = corner + left.  Library layout: left column bottom-up, corner, above row left to right.
      const Pel* ext = pred->getPredictorPtr( COMPONENT_Y, false );
      int16_t line[4 * 64 + 1];
      for ( unsigned i = 0; i < 2 * n; i++ ) line[i] = (int16_t)ext[ (2 * n - i) * roiW ];
      line[2 * n] = (int16_t)ext[0];
      for ( unsigned k = 0; k < 2 * n; k++ ) line[2 * n + 1 + k] = (int16_t)ext[1 + k];
      uint8_t blk[64 * 64];
      for ( unsigned r = 0; r < n; r++ )
        for ( unsigned cidx = 0; cidx < n; cidx++ ) blk[r * n + cidx] = (uint8_t)org[r * orgStride + cidx];
      const uint8_t size8 = (uint8_t)n;
      const int rc = hevcdl_rmd_exact( S.ctx, 1, &size8, blk, line, NULL, NULL, NULL, 0.0, S.ex_satd, NULL, NULL );
      if ( rc ) HevcdlSession::die( "hevcdl_rmd_exact", rc, S.ctx );
      S.ex_x = x; S.ex_y = y; S.ex_n = n;
      S.exact_calls++;
    }
    *sad += S.ex_satd[mode];
    S.hook_hits++;
    return true;
  }
  if ( !S.gpu_rmd || !S.have_view ) return false;
  return S.lookup( x, y, width, mode, sad );
}
