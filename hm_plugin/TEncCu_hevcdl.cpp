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
//   HEVCDL_LOOKAHEAD  frames submitted ahead of the one being encoded (default 3; 0 = off).  A reader thread preads frames
//                     n+1 .. n+k of the encoder's own input file (all-intra pictures are independent and the file offset of
//                     frame f is f * W * H * 3/2) while HM encodes frame n, so only the first frame ever waits
//                     for the device.  The file is found from HEVCDL_INPUT or the encoder's own command line / -c files
//                     (InputFile, FrameSkip, FramesToBeEncoded); every prefetched frame is checked against the planes HM
//                     hands over for that frame id (hash of all three planes) -- on a mismatch the prefetch is dropped
//                     and the frame is uploaded from HM's planes as without lookahead.  8-bit 4:2:0 input without padding only.
//   HEVCDL_TQ         1 = transform + quantiser (flat, or the rate-distortion optimised one with sign-bit hiding when the encoder
//                     runs with RDOQ, its default) + dequantiser + inverse transform of every TU of xIntraCodingTUBlock on
//                     the device (hevcdl_tu_code / hevcdl_tu_code_rdoq, one synchronous call per TU): byte-identical
//                     bitstreams; TUs the core does not cover stay HM's (hm_plugin/rmd_hook.h)
//   HEVCDL_DBF        1 = the deblocking filter of every picture on the device (hevcdl_deblock_frame; hm_plugin/
//                     TComLoopFilter_hevcdl.cpp replaces TComLoopFilter::loopFilterPic): byte-identical bitstreams
//   HEVCDL_PRED       1 = every intra-predicted block (TComPrediction::predIntraAng: first pass and RD pass, luma and chroma) on
//                     the device (hevcdl_intra_pred; hm_plugin/TComPrediction_hevcdl.cpp), one synchronous call per block:
//                     byte-identical bitstreams
//   HEVCDL_SAO        1 = the two picture-wide passes of SAO on the device: the statistics of the parameter estimation
//                     (hevcdl_sao_stats; hm_plugin/TEncSAO_hevcdl.cpp replaces TEncSampleAdaptiveOffset::getStatistics) and
//                     the application of the decided offsets (hevcdl_sao_apply; hm_plugin/TComSAO_hevcdl.cpp replaces
//                     TComSampleAdaptiveOffset::offsetCTU); the RD decision between them stays HM's: byte-identical bitstreams
//                     With HEVCDL_DBF=1 as well the three passes share the picture on the device (hevcdl_inloop_frame: deblocking +
//                     statistics in one round trip, offsets applied to the resident picture); HEVCDL_INLOOP_FUSE=0 keeps them separate
//   HEVCDL_PIN        0 = do not page-lock HM's picture buffers (the in-loop calls then stage every plane through the library's buffer)
//   HEVCDL_RMD        1 = run the batched 35-mode SATD pass on the B200 and let estIntraPredLumaQT's first pass take its
//                     per-mode SATDs from it (hm_plugin/rmd_hook.h; references are ORIGINAL pixels, so mode
//                     decisions follow the +-1 % BD-rate clause, not the bit-exact one);
//                     2 = exact mode: per PU, HM's reconstructed reference samples go to hevcdl_rmd_exact (bit-exact,
//                     byte-identical bitstream, one synchronous call per PU) (default 0: HM's own pass)
//   HEVCDL_EARLY_CREATE 0 = create the device context at the first CTU instead of on a background thread at program load
// There is no fallback: any library failure aborts the encoder with the library's error text.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "TLibCommon/TComTU.h"
#include "TLibEncoder/TEncCu.h"
#include "TLibEncoder/TEncTop.h"

#include "hevcdl.h"
#include "rmd_hook.h"

namespace {

// ---- frame lookahead (replaces the overlap the reference gets from running use_model.py as a
// detached process ahead of the encoder, encmain.cpp:107-108) ---------------------------------------------------------
uint64_t hash_bytes(uint64_t h, const uint8_t *p, size_t n) {
  size_t i = 0;
  for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, p + i, 8); h = (h ^ w) * 0x9E3779B97F4A7C15ull; h ^= h >> 29; }
  for (; i < n; i++) { h = (h ^ p[i]) * 0x9E3779B97F4A7C15ull; h ^= h >> 29; }
  return h;
}
uint64_t hash_u8_plane(uint64_t h, const uint8_t *p, int w, int hgt) {          // row by row, like hash_pel_plane
  for (int y = 0; y < hgt; y++) h = hash_bytes(h, p + (size_t)y * w, w);
  return h;
}
uint64_t hash_pel_plane(uint64_t h, const Pel *p, int stride, int w, int hgt, std::vector<uint8_t> &row) {
  row.resize(w);
  for (int y = 0; y < hgt; y++) {
    const Pel *r = p + (size_t)y * stride;
    for (int x = 0; x < w; x++) row[x] = (uint8_t)r[x];
    h = hash_bytes(h, row.data(), w);
  }
  return h;
}

// What the encoder was told about its input, recovered from its own command line (/proc/self/cmdline) and -c files.
struct InputSpec {
  std::string file;
  long skip = 0, frames = -1, tsr = 1;
  int width = 0, height = 0;
  bool eight_bit = true;
  static std::string value_of(const std::string &line, const char *key) {      // "Key : value   # comment"
    size_t i = 0;
    while (i < line.size() && isspace((unsigned char)line[i])) i++;
    const size_t kl = strlen(key);
    if (line.compare(i, kl, key) != 0) return "";
    i += kl;
    while (i < line.size() && isspace((unsigned char)line[i])) i++;
    if (i >= line.size() || line[i] != ':') return "";
    std::string v = line.substr(i + 1);
    const size_t hsh = v.find('#');
    if (hsh != std::string::npos) v.erase(hsh);
    const size_t a = v.find_first_not_of(" \t\r\n"), b = v.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : v.substr(a, b - a + 1);
  }
  void take(const char *key, const std::string &v) {
    if (v.empty()) return;
    if (!strcmp(key, "InputFile")) file = v;
    else if (!strcmp(key, "FrameSkip")) skip = atol(v.c_str());
    else if (!strcmp(key, "FramesToBeEncoded")) frames = atol(v.c_str());
    else if (!strcmp(key, "TemporalSubsampleRatio")) tsr = atol(v.c_str());
    else if (!strcmp(key, "InputBitDepth")) eight_bit = atol(v.c_str()) == 8 || atol(v.c_str()) == 0;
    else if (!strcmp(key, "SourceWidth")) width = atoi(v.c_str());
    else if (!strcmp(key, "SourceHeight")) height = atoi(v.c_str());
  }
  void parse_cfg(const std::string &path) {
    std::ifstream f(path);
    std::string line;
    static const char *keys[] = {"InputFile", "FrameSkip", "FramesToBeEncoded", "TemporalSubsampleRatio", "InputBitDepth", "SourceWidth", "SourceHeight"};
    while (std::getline(f, line))
      for (const char *k : keys) take(k, value_of(line, k));
  }
  void parse_cmdline() {
    std::ifstream f("/proc/self/cmdline", std::ios::binary);
    std::vector<std::string> av;
    std::string a;
    while (std::getline(f, a, '\0')) av.push_back(a);
    static const struct { const char *sh, *lg; } opt[] = {{"-i", "InputFile"}, {"-fs", "FrameSkip"}, {"-f", "FramesToBeEncoded"},
                                                           {"-ts", "TemporalSubsampleRatio"}, {nullptr, "InputBitDepth"},
                                                           {"-wdt", "SourceWidth"}, {"-hgt", "SourceHeight"}};
    for (size_t i = 1; i < av.size(); i++) {       // later options override earlier ones, as in the reference's parser
      if (av[i] == "-c" && i + 1 < av.size()) { parse_cfg(av[++i]); continue; }
      for (const auto &o : opt) {
        const std::string lg = std::string("--") + o.lg;
        if (o.sh && av[i] == o.sh && i + 1 < av.size()) { take(o.lg, av[++i]); break; }
        if (av[i] == lg && i + 1 < av.size()) { take(o.lg, av[++i]); break; }
        if (av[i].compare(0, lg.size() + 1, lg + "=") == 0) { take(o.lg, av[i].substr(lg.size() + 1)); break; }
      }
    }
    if (const char *e = getenv("HEVCDL_INPUT")) file = e;
  }
};

// Reader thread: keeps frames (next .. next + ring) of the input file in host buffers.  Only this thread touches the
// file; only the encoder thread touches the hevcdl context.
struct FrameReader {
  int fd = -1, w = 0, h = 0;
  long skip = 0, tsr = 1, nframes = 0;             // encoder frame id f lives at file frame skip + f * tsr
  size_t fbytes = 0;
  std::vector<std::vector<uint8_t>> ring;
  std::vector<long> held;                          // encoder frame id held by each ring buffer (-1: none)
  long want_lo = 0, want_hi = -1;                  // ids the encoder thread still wants buffered: [want_lo, want_hi]
  bool stop = false;
  std::mutex mu;
  std::condition_variable cv;
  std::thread th;

  bool open(const InputSpec &in, int width, int height, int depth) {
    if (in.file.empty() || !in.eight_bit || in.tsr < 1) return false;
    fd = ::open(in.file.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    w = width; h = height; fbytes = (size_t)w * h * 3 / 2;
    if (fstat(fd, &st) != 0 || st.st_size < (off_t)fbytes || st.st_size % (off_t)fbytes) { ::close(fd); fd = -1; return false; }
    skip = in.skip; tsr = in.tsr;
    nframes = ((long)(st.st_size / (off_t)fbytes) - skip + tsr - 1) / tsr;
    if (in.frames >= 0 && in.frames < nframes) nframes = in.frames;
    ring.assign(depth + 1, std::vector<uint8_t>());
    for (auto &b : ring) b.resize(fbytes);
    held.assign(ring.size(), -1);
    th = std::thread([this] { run(); });
    return true;
  }
  void run() {
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      long f = -1;
      int slot = -1;
      cv.wait(lk, [&] {
        if (stop) return true;
        for (long c = want_lo; c <= want_hi && c < nframes; c++) {
          bool have = false;
          for (long hId : held) have |= hId == c;
          if (have) continue;
          for (size_t i = 0; i < held.size(); i++)
            if (held[i] < want_lo) { f = c; slot = (int)i; return true; }   // a buffer whose frame is no longer wanted
          return false;
        }
        return false;
      });
      if (stop) return;
      held[slot] = -2;                             // being filled
      lk.unlock();
      const off_t off = (off_t)(skip + f * tsr) * (off_t)fbytes;
      size_t got = 0;
      while (got < fbytes) {
        const ssize_t r = pread(fd, ring[slot].data() + got, fbytes - got, off + (off_t)got);
        if (r <= 0) break;
        got += (size_t)r;
      }
      lk.lock();
      held[slot] = got == fbytes ? f : -1;
      if (got != fbytes) nframes = f;              // short file: nothing beyond this frame
      cv.notify_all();
    }
  }
  // encoder thread: ask for [lo, hi]; returns the buffer of frame f if it is ready (no waiting)
  void want(long lo, long hi) { std::lock_guard<std::mutex> lk(mu); want_lo = lo; want_hi = hi; cv.notify_all(); }
  const uint8_t *ready(long f, int wait_ms = 0) {
    std::unique_lock<std::mutex> lk(mu);
    const uint8_t *buf = nullptr;
    auto have = [&] {
      for (size_t i = 0; i < held.size(); i++) if (held[i] == f) { buf = ring[i].data(); return true; }
      return f >= nframes;                         // past the end of the file: never comes
    };
    if (wait_ms > 0) cv.wait_for(lk, std::chrono::milliseconds(wait_ms), have);
    else have();
    return buf;
  }
  void close() {
    if (th.joinable()) { { std::lock_guard<std::mutex> lk(mu); stop = true; cv.notify_all(); } th.join(); }
    if (fd >= 0) ::close(fd);
    fd = -1;
  }
};

struct HevcdlSession {
  hevcdl_ctx *ctx = nullptr;
  int width = 0, height = 0;
  int frame = -1;               // frame currently resident on the device (-1: none)
  int labels_frame = -1;        // frame whose labels were last asked for (timing statistics)
  bool gpu_rmd = false;         // HEVCDL_RMD=1: first-pass SATDs come from the device (batched, original-pixel references)
  bool exact_rmd = false;       // HEVCDL_RMD=2: ... from hevcdl_rmd_exact fed HM's reconstructed references, PU by PU
  unsigned ex_x = ~0u, ex_y = ~0u, ex_n = 0;      // PU whose 35 SATDs are cached in ex_satd
  uint32_t ex_satd[35];
  unsigned long long exact_calls = 0;
  unsigned long long pred_device = 0, pred_host = 0; // intra-predicted blocks on the device / by the reference's own code
  unsigned long long sao_device = 0, sao_host = 0;
  unsigned long long pinned_buffers = 0;           // HM picture buffers page-locked for direct copies (hevcdl_host_register)
  unsigned long long inloop_resident = 0;          // pictures whose SAO passes reused the deblocking call's round trip / resident picture
  unsigned long long saoapply_device = 0, saoapply_host = 0;   // pictures whose SAO offsets were applied on the device / by the reference   // SAO statistics passes on the device / by the reference's own code
  unsigned long long dbf_device = 0, dbf_host = 0;   // pictures deblocked on the device / by the reference's own filter
  bool gpu_tq = false;          // HEVCDL_TQ=1: TU transform / quantisation / inverse transform on the device
  unsigned long long tq_calls = 0, tq_declined = 0;
  hevcdl_frame_view view;       // results of `frame` (pinned host memory owned by the library)
  bool have_view = false;
  int ctu_first = 0, ctu_count = 0, cursor = 0;   // PU range of the CTU being compressed + last hit
  unsigned long long hook_hits = 0, hook_misses = 0;
  // lookahead
  FrameReader reader;
  int lookahead = 0;            // frames submitted ahead (0: off)
  long submitted_hi = -1;       // highest frame id handed to the device from the file
  std::vector<std::pair<long, uint64_t>> ahead;   // (frame id, hash of the file's planes) of frames submitted from the file
  unsigned long long la_hits = 0, la_direct = 0, la_mismatch = 0;
  double t_create = 0, t_create_total = 0, t_wait_first = 0, t_wait_later = 0;   // seconds: encoder thread blocked for hevcdl_create / its whole duration; blocked in the first label query of frame 0 / of later frames
  static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  std::vector<uint8_t> rowbuf;

  static void die(const char *what, int rc, hevcdl_ctx *c) {
    fprintf(stderr, "hevcdl: %s failed: %s (%s)\n", what, hevcdl_status_str(rc), hevcdl_last_error(c));
    exit(EXIT_FAILURE);
  }

  // Everything hevcdl_create needs, from the environment (the signature of compressCtu leaves no room for it)
  static int env_lookahead() {
    const char *e = getenv("HEVCDL_LOOKAHEAD");
    const int k = e ? atoi(e) : 3;
    return k < 0 ? 0 : (k > 16 ? 16 : k);
  }
  static hevcdl_cfg make_cfg(int w, int h) {
    hevcdl_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = HEVCDL_ABI_VERSION;
    const char *e;
    cfg.device = (e = getenv("HEVCDL_DEVICE")) ? atoi(e) : 0;
    cfg.width = w; cfg.height = h;
    cfg.slots = 2 + env_lookahead();
    cfg.batch = 1;
    cfg.precision = ((e = getenv("HEVCDL_PRECISION")) && !strcmp(e, "bf16")) ? HEVCDL_PREC_BF16_TC : HEVCDL_PREC_FP32;
    const int rmd_mode = (e = getenv("HEVCDL_RMD")) ? atoi(e) : 0;
    cfg.rmd = rmd_mode == 1;
    cfg.outputs = rmd_mode == 1 ? HEVCDL_OUT_SATD : 0;   // the first-pass hook re-ranks with HM's own mode bits: it needs the SATD table
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
    return cfg;
  }

  // hevcdl_create costs 0.3-4 s (CUDA context creation) and compressCtu is first called only after HM has parsed its
  // configuration, allocated its pictures and read the first frame: a thread started when the program is loaded creates the
  // context meanwhile, from the picture size on the encoder's own command line / -c files (HEVCDL_EARLY_CREATE=0: off).
  struct Early {
    std::thread th;
    hevcdl_cfg cfg;
    hevcdl_ctx *ctx = nullptr;
    int rc = 0;
    double seconds = 0;
  };
  static Early *&early() { static Early *e = nullptr; return e; }
  static void start_early() {
    const char *e = getenv("HEVCDL_EARLY_CREATE");
    if (e && atoi(e) == 0) return;
    InputSpec in;
    in.parse_cmdline();
    if (in.width < 8 || in.height < 8) return;
    Early *y = new Early;
    y->cfg = make_cfg(in.width, in.height);
    y->th = std::thread([y] {
      const double t0 = now();
      y->rc = hevcdl_create(&y->cfg, &y->ctx);
      y->seconds = now() - t0;
    });
    early() = y;
  }

  void open(int w, int h) {
    const char *e;
    lookahead = env_lookahead();
    if (lookahead) {
      InputSpec in;
      in.parse_cmdline();
      if (!reader.open(in, w, h, lookahead + 1)) lookahead = 0;    // no usable input file: behave as without lookahead
    }
    const int rmd_mode = (e = getenv("HEVCDL_RMD")) ? atoi(e) : 0;
    gpu_rmd = rmd_mode == 1;
    exact_rmd = rmd_mode == 2;
    gpu_tq = (e = getenv("HEVCDL_TQ")) && atoi(e) == 1;
    const double t0 = now();
    if (Early *y = early()) {                      // created in the background since program load: wait for the rest of it
      y->th.join();
      if (y->rc == 0 && y->cfg.width == w && y->cfg.height == h) { ctx = y->ctx; t_create_total = y->seconds; }
      else if (y->ctx) hevcdl_destroy(y->ctx);     // the command line did not say what HM ended up encoding
      delete y;
      early() = nullptr;
    }
    if (!ctx) {
      hevcdl_cfg cfg = make_cfg(w, h);
      const int rc = hevcdl_create(&cfg, &ctx);
      if (rc) die("hevcdl_create", rc, nullptr);
      t_create_total = now() - t0;
    }
    t_create = now() - t0;
    width = w; height = h;
  }

  // Hand the picture's ORIGINAL planes (the same ones xCompressCU reads at TEncCu.cpp:484) to the device.
  // This replaces gen_frames.py:21 (ffmpeg dump) and the sidecar's whole per-frame loop (use_model.py:74-127).
  void begin_frame(int id, TComPicYuv *org, const TComSPS *sps) {
    const int w = org->getWidth(COMPONENT_Y), h = org->getHeight(COMPONENT_Y);
    // the CNN and the SATD pass are defined on 8-bit samples (the reference sidecar reads the 8-bit input file,
    // gen_frames.py:21); with a higher internal bit depth HM's Pel planes hold scaled values that must not be truncated
    if (sps->getBitDepth(CHANNEL_TYPE_LUMA) != 8 || sps->getBitDepth(CHANNEL_TYPE_CHROMA) != 8 || org->getChromaFormat() != CHROMA_420) {
      fprintf(stderr, "hevcdl: only 8-bit 4:2:0 encodes are supported (internal bit depth %d/%d)\n",
              sps->getBitDepth(CHANNEL_TYPE_LUMA), sps->getBitDepth(CHANNEL_TYPE_CHROMA));
      exit(EXIT_FAILURE);
    }
    if (!ctx) open(w, h);
    if (w != width || h != height) { fprintf(stderr, "hevcdl: picture size changed mid-sequence\n"); exit(EXIT_FAILURE); }
    if (frame >= 0) { const int rc = hevcdl_release_frame(ctx, frame); if (rc) die("hevcdl_release_frame", rc, ctx); }
    bool on_device = false;
    if (lookahead) {
      for (size_t i = 0; i < ahead.size(); i++) {
        if (ahead[i].first != id) continue;
        uint64_t hs = 0;
        hs = hash_pel_plane(hs, org->getAddr(COMPONENT_Y), org->getStride(COMPONENT_Y), w, h, rowbuf);
        hs = hash_pel_plane(hs, org->getAddr(COMPONENT_Cb), org->getStride(COMPONENT_Cb), w / 2, h / 2, rowbuf);
        hs = hash_pel_plane(hs, org->getAddr(COMPONENT_Cr), org->getStride(COMPONENT_Cr), w / 2, h / 2, rowbuf);
        if (hs == ahead[i].second) { on_device = true; la_hits++; ahead.erase(ahead.begin() + i); }
        else {   // the file is not what HM is encoding (pre-processing, another skip): drop everything read ahead
          la_mismatch++;
          fprintf(stderr, "hevcdl: lookahead frame %d differs from the encoder's picture; lookahead disabled\n", id);
          for (auto &a : ahead) { const int rc = hevcdl_release_frame(ctx, (int)a.first); if (rc) die("hevcdl_release_frame", rc, ctx); }
          ahead.clear();
          lookahead = 0;
          reader.close();
        }
        break;
      }
    }
    if (!on_device) {
      const int rc = hevcdl_submit_frame_pel16(ctx, id, org->getAddr(COMPONENT_Y), org->getStride(COMPONENT_Y),
                                               org->getAddr(COMPONENT_Cb), org->getAddr(COMPONENT_Cr), org->getStride(COMPONENT_Cb));
      if (rc) die("hevcdl_submit_frame_pel16", rc, ctx);
      la_direct++;
    }
    frame = id;
    have_view = false;
    if (lookahead) {
      // top up: every frame of (id, id + lookahead] the reader already holds goes to the device now (non-blocking: a staging
      // copy and queued work), the rest next time; the reader is told what to fetch meanwhile
      if (submitted_hi < id) submitted_hi = id;
      reader.want(submitted_hi + 1, (long)id + lookahead + 1);   // one frame further than the submit window: the next top-up finds it read
      while (submitted_hi < (long)id + lookahead) {
        // nothing ahead yet (first frame, or the reader fell behind): the next frame is worth a bounded wait -- the device
        // is busy with the current frame for about as long as a page-cache read takes, and HM blocks on its labels next
        const uint8_t *buf = reader.ready(submitted_hi + 1, ahead.empty() ? 25 : 0);
        if (!buf) break;
        const long f = submitted_hi + 1;
        const uint8_t *y = buf, *u = y + (size_t)w * h, *v = u + (size_t)(w / 2) * (h / 2);
        const int rc = hevcdl_submit_frame_u8(ctx, (int)f, y, w, u, v, w / 2);
        if (rc) die("hevcdl_submit_frame_u8", rc, ctx);
        uint64_t hs = hash_u8_plane(0, y, w, h);
        hs = hash_u8_plane(hs, u, w / 2, h / 2);
        hs = hash_u8_plane(hs, v, w / 2, h / 2);
        ahead.push_back(std::make_pair(f, hs));
        submitted_hi = f;
        reader.want(submitted_hi + 1, (long)id + lookahead + 1);   // one frame further than the submit window: the next top-up finds it read
      }
    }
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
    reader.close();
    if (ctx) {
      if (getenv("HEVCDL_VERBOSE")) {
        fprintf(stderr, "hevcdl: pictures deblocked on the device %llu / by the reference's filter %llu\n", dbf_device, dbf_host);
        fprintf(stderr, "hevcdl: blocks predicted on the device %llu / by the reference's code %llu\n", pred_device, pred_host);
        fprintf(stderr, "hevcdl: SAO statistics passes on the device %llu / by the reference's code %llu\n", sao_device, sao_host);
        fprintf(stderr, "hevcdl: SAO offsets applied on the device for %llu pictures / by the reference's code for %llu\n", saoapply_device, saoapply_host);
        fprintf(stderr, "hevcdl: in-loop passes of %llu pictures shared one upload (deblocked picture resident between deblocking and SAO); "
                        "%llu of HM's picture buffers page-locked for direct copies\n", inloop_resident, pinned_buffers);
        fprintf(stderr, "hevcdl: lookahead %d: %llu frames were on the device before HM asked, %llu uploaded from HM's planes, %llu mismatches\n",
                lookahead, la_hits, la_direct, la_mismatch);
        fprintf(stderr, "hevcdl: encoder thread blocked %.3f s for hevcdl_create (CUDA context + weights + buffers: %.3f s, started at program "
                        "load), %.4f s waiting for the labels of the first frame, %.4f s for all later frames together\n",
                t_create, t_create_total, t_wait_first, t_wait_later);
        hevcdl_stats_t st;
        if (!hevcdl_get_stats(ctx, &st))
          fprintf(stderr, "hevcdl: %llu frames, %llu CTUs, CNN %.3f ms, RMD %.3f ms device time, %llu kernel launches, "
                          "first-pass SATDs served %llu / missed %llu, exact PU calls %llu, TUs coded on the device %llu / left to HM %llu\n",
                  (unsigned long long)st.frames, (unsigned long long)st.ctus, st.ms_cnn, st.ms_rmd,
                  (unsigned long long)st.kernel_launches, hook_hits, hook_misses, exact_calls, tq_calls, tq_declined);
      }
      hevcdl_destroy(ctx);
    }
  }
};

HevcdlSession g_session;   // one encoder thread, one TEncCu instance (TEncTop.h:93): a process-wide session is enough
struct HevcdlEarlyStart { HevcdlEarlyStart() { HevcdlSession::start_early(); } } g_early_start;   // before main(): see HevcdlSession::Early

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
    g_session.begin_frame( m_iFrame, pCtu->getPic()->getPicYuvOrg(), pCtu->getSlice()->getSPS() );
  }

  uint8_t depth8[16];
  const bool firstQuery = g_session.labels_frame != m_iFrame;
  const double tq0 = firstQuery ? HevcdlSession::now() : 0.0;
  const int rc = hevcdl_ctu_labels( g_session.ctx, m_iFrame, (int)ctuRsAddr, depth8 );   // blocks on the frame's event
  if ( firstQuery )
  {
    ( g_session.labels_frame < 0 ? g_session.t_wait_first : g_session.t_wait_later ) += HevcdlSession::now() - tq0;
    g_session.labels_frame = m_iFrame;
  }
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

hevcdl_ctx *hevcdl_hm_context() { return g_session.ctx; }
void hevcdl_hm_count_pred( bool onDevice ) { ( onDevice ? g_session.pred_device : g_session.pred_host )++; }
// Page-lock the three component buffers of one of HM's pictures (once per buffer; HM allocates its pictures once and reuses
// them), so that the in-loop entry points copy straight between HM's strided planes and the device.  HEVCDL_PIN=0: off.
void hevcdl_hm_pin_picture( TComPicYuv *pic )
{
  static const bool on = !( getenv( "HEVCDL_PIN" ) && atoi( getenv( "HEVCDL_PIN" ) ) == 0 );
  static std::vector<const void *> done;
  if ( !on || !pic ) return;
  for ( int c = 0; c < 3; c++ )
  {
    const ComponentID id = ComponentID( c );
    Pel *buf = pic->getBuf( id );
    if ( !buf ) continue;
    bool seen = false;
    for ( const void *q : done ) seen = seen || q == buf;
    if ( seen ) continue;
    done.push_back( buf );
    if ( hevcdl_host_register( buf, (size_t)pic->getStride( id ) * pic->getTotalHeight( id ) * sizeof( Pel ) ) == 0 ) g_session.pinned_buffers++;
  }
}
void hevcdl_hm_count_inloop_resident( bool resident ) { if ( resident ) g_session.inloop_resident++; }
void hevcdl_hm_count_sao_apply( bool onDevice ) { ( onDevice ? g_session.saoapply_device : g_session.saoapply_host )++; }
void hevcdl_hm_count_sao( bool onDevice ) { ( onDevice ? g_session.sao_device : g_session.sao_host )++; }
void hevcdl_hm_count_dbf( bool onDevice ) { ( onDevice ? g_session.dbf_device : g_session.dbf_host )++; }

// TU-coding hook (rmd_hook.h): called from the reference's xIntraCodingTUBlock in place of transformNxN + invTransformNxN.
static_assert( sizeof(estBitsSbacStruct) == HEVCDL_EST_INTS * sizeof(int32_t), "hevcdl_tu_code_rdoq takes the reference's estBitsSbacStruct as is" );

bool hevcdl_hm_tu_code( TComDataCU* pcCU, TComTU& rTu, int compIDi, short* piResi, unsigned uiStride, int* pcCoeff, int* puiAbsSum,
                        int qp, bool useTransformSkip, bool rdoqOn, const void* estBits )
{
  HevcdlSession &S = g_session;
  if ( !S.gpu_tq || !S.ctx ) return false;
  const ComponentID compID = ComponentID( compIDi );
  const UInt uiAbsPartIdx = rTu.GetAbsPartIdxTU();
  const TComRectangle &rect = rTu.getRect( compID );
  const UInt n = rect.width;
  // what the device core covers (csrc/tq.cuh): square 4..32 TUs, flat quantiser, no bypass, no scaling lists, 8-bit
  // (the flat quantiser's own sign-bit hiding, signBitHidingHDQ, is not on the device: flat + SDH stays HM's)
  const bool sdh = pcCU->getSlice()->getPPS()->getSignDataHidingEnabledFlag();
  if ( rect.width != rect.height || n < 4 || n > 32 || pcCU->getCUTransquantBypass( uiAbsPartIdx ) || ( !rdoqOn && sdh ) ||
       pcCU->getSlice()->getSPS()->getScalingListFlag() ||   /* RDOQ_CHROMA is 1 in the reference (TComTrQuant.cpp:64): chroma TUs take the same quantiser */
       qp < 0 || qp > 51 || ( useTransformSkip && n != 4 ) )
  {
    S.tq_declined++;
    return false;
  }
  int16_t resi[32 * 32], level[32 * 32], rec[32 * 32];
  for ( UInt y = 0; y < n; y++ )
    for ( UInt x = 0; x < n; x++ ) resi[y * n + x] = piResi[y * uiStride + x];
  hevcdl_tu tu;
  tu.log2_size = (uint8_t)( n == 4 ? 2 : n == 8 ? 3 : n == 16 ? 4 : 5 );
  tu.qp = (uint8_t)qp;
  tu.flags = (uint8_t)( ( useTransformSkip ? HEVCDL_TU_TSKIP : ( rTu.useDST( compID ) ? HEVCDL_TU_DST : 0 ) ) |
                        ( pcCU->getSlice()->getSliceType() == I_SLICE ? 0 : HEVCDL_TU_INTER ) );
  tu.reserved = 0;
  tu.offset = 0;
  uint32_t absSum = 0;
  int rc;
  if ( rdoqOn )
  {
    // what xRateDistOptQuant reads from encoder state (TComTrQuant.cpp:2130-2205, 2446-2460)
    hevcdl_tu_rdoq rq;
    rq.lambda = pcCU->getSlice()->getLambdas()[compID];      // = TComTrQuant::m_dLambda after selectLambda(compID) (TEncSlice.cpp:133,139)
    rq.est_index = 0;
    rq.channel = (uint8_t)toChannelType( compID );
    rq.scan_type = (uint8_t)pcCU->getCoefScanIdx( uiAbsPartIdx, n, n, compID );
    rq.ctx_cbf = (uint8_t)( pcCU->getCtxQtCbf( rTu, toChannelType( compID ) ) + getCBFContextOffset( compID ) );
    rq.flags = (uint8_t)( ( sdh ? 1 : 0 ) | ( pcCU->isIntra( uiAbsPartIdx ) ? 2 : 0 ) | ( pcCU->getTransformIdx( uiAbsPartIdx ) == 0 ? 4 : 0 ) );
    tu.flags |= HEVCDL_TU_RDOQ;
    rc = hevcdl_tu_code_rdoq( S.ctx, 1, &tu, &rq, (const int32_t*)estBits, 1, resi, (size_t)n * n, NULL, level, NULL, rec, &absSum, NULL );
  }
  else rc = hevcdl_tu_code( S.ctx, 1, &tu, resi, (size_t)n * n, NULL, level, NULL, rec, &absSum, NULL );
  if ( rc ) HevcdlSession::die( "hevcdl_tu_code", rc, S.ctx );
  S.tq_calls++;
  // exactly what transformNxN (TComTrQuant.cpp:1450-1534) and the inverse-transform if/else (TEncSearch.cpp:1310-1328) leave behind
  *puiAbsSum = (int)absSum;
  pcCU->setCbfPartRange( ( ( absSum > 0 ? 1 : 0 ) << rTu.GetTransformDepthRel() ), compID, uiAbsPartIdx, rTu.GetAbsPartIdxNumParts( compID ) );
  for ( UInt i = 0; i < n * n; i++ ) pcCoeff[i] = absSum > 0 ? (int)level[i] : 0;
  for ( UInt y = 0; y < n; y++ )
    for ( UInt x = 0; x < n; x++ ) piResi[y * uiStride + x] = absSum > 0 ? rec[y * n + x] : 0;
  return true;
}
