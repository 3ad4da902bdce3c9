// tools/tc_probe.cu -- hardware probe for the tcgen05 conventions the CNN kernels rely on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_probe tools/tc_probe.cu
// Checks (1) canonical no-swizzle K-major descriptors, (2) sliding-window descriptors over a
// padded plane (SBO = row pitch, LBO = 16 B), and measures (3) cycles per MMA for several N with
// A/B in shared memory and (4) tcgen05.ld throughput.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../hevc-deep-learning-pipeline_b200/csrc/tc_ptx.cuh"

using namespace hevcdl::tc;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } \
  } while (0)

struct Args {
  const __nv_bfloat16 *A, *B;   // A: raw bytes copied to smem at offset 0; B likewise at offset a_bytes
  int a_bytes, b_bytes;
  uint32_t a_start, a_lbo, a_sbo, a_kstep;   // descriptor parameters (bytes); a_kstep: start advance per K-step
  uint32_t b_start, b_lbo, b_sbo, b_kstep;
  int M, N, ksteps, reps, nacc;
  int bg;                       // background activity of warps 1-3 while warp 0 issues: 0 none, 1 tcgen05.ld, 2 shared-memory traffic, 3 bulk copies
  const uint8_t *gsrc;          // source of the background bulk copies
  float *D;                     // [128][N]
  long long *cycles;
};

__global__ void __launch_bounds__(128, 1) k_probe(Args a) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < a.a_bytes / 4; i += 128) reinterpret_cast<uint32_t *>(sm)[i] = reinterpret_cast<const uint32_t *>(a.A)[i];
  uint8_t *smb = sm + ((a.a_bytes + 1023) & ~1023);
  for (int i = tid; i < a.b_bytes / 4; i += 128) reinterpret_cast<uint32_t *>(smb)[i] = reinterpret_cast<const uint32_t *>(a.B)[i];
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_slot;
  const uint32_t idesc = idesc_bf16(a.M, a.N);
  long long t0 = 0, t1 = 0;
  __shared__ volatile int stop_flag;
  __shared__ __align__(8) uint64_t bgbar;
  if (tid == 0) { stop_flag = 0; mbar_init(&bgbar, 1); mbar_init_fence(); }
  __syncthreads();
  if (warp != 0 && a.bg) {
    float acc = 0.f;
    uint32_t ph = 0;
    while (!stop_flag) {
      if (a.bg == 1) {                       // TMEM reads of columns the MMAs do not touch
        float v[32];
        tmem_ld32(tmem_addr(tbase, warp * 32, 256 + 32 * (warp & 1)), v);
        tmem_ld_wait();
        acc += v[lane & 31];
      } else if (a.bg == 2) {                // shared-memory loads/stores in a scratch region behind the operands
        volatile float *scr = reinterpret_cast<volatile float *>(sm + 96 * 1024);
        for (int i = 0; i < 16; i++) { acc += scr[(tid + 32 * i) & 1023]; scr[(tid * 3 + i) & 1023] = acc; }
      } else if (a.bg == 3 && warp == 1) {   // bulk copies global -> shared (16 KB each)
        if (lane == 0) {
          mbar_expect_tx(&bgbar, 16384);
          bulk_g2s(sm + 96 * 1024, a.gsrc, 16384, &bgbar);
          mbar_wait(&bgbar, ph & 1);
          ph++;
        }
        __syncwarp();
      }
    }
    if (acc == 123.f) a.cycles[3] = 1;
  }
  if (warp == 0 && elect_one()) {
    const uint32_t sa = smem_u32(sm) + a.a_start, sb = smem_u32(smb) + a.b_start;
    const uint64_t da = smem_desc(sa, a.a_lbo, a.a_sbo), db = smem_desc(sb, a.b_lbo, a.b_sbo);
    const uint64_t dka = (uint64_t)(a.a_kstep >> 4), dkb = (uint64_t)(a.b_kstep >> 4);
    long long ti = 0;
    if (a.nacc == 1) {
      t0 = clock64();
      for (int r = 0; r < a.reps; r++)
        for (int k = 0; k < a.ksteps; k++) mma_bf16_ss(tbase, da + k * dka, db + k * dkb, idesc, (r | k) ? 1u : 0u);
      ti = clock64();
    } else {   // 4 accumulators, round-robin (N <= 128)
      t0 = clock64();
      for (int r = 0; r < a.reps; r += 4) {
        const uint32_t acc = r ? 1u : 0u;
        const uint64_t oa = (uint64_t)((r >> 2) & 7) * dka, ob = (uint64_t)((r >> 2) & 7) * dkb;   // distinct operands when kstep != 0
        mma_bf16_ss(tbase, da + oa, db + ob, idesc, acc);
        mma_bf16_ss(tbase + a.N, da + oa, db + ob, idesc, acc);
        mma_bf16_ss(tbase + 2 * a.N, da + oa, db + ob, idesc, acc);
        mma_bf16_ss(tbase + 3 * a.N, da + oa, db + ob, idesc, acc);
      }
      ti = clock64();
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (a.cycles) a.cycles[1] = ti - t0;
    if (a.cycles) a.cycles[0] = t1 - t0;
    stop_flag = 1;
  }
  __syncthreads();
  fence_after_sync();
  if (a.D) {
    for (int c0 = 0; c0 < a.N; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_addr(tbase, warp * 32, c0), v);
      tmem_ld_wait();
      for (int j = 0; j < 16; j++) a.D[(size_t)(warp * 32 + lane) * a.N + c0 + j] = v[j];
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// tcgen05.ld throughput: every warp reads `cols` columns of its lane quarter `iters` times
__global__ void __launch_bounds__(256, 1) k_ldprobe(int iters, int cols, long long *cycles, float *sink) {
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_slot;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  const int cbase = (warp >> 2) * 256;
  for (int it = 0; it < iters; it++) {
    for (int c = 0; c < cols; c += 32) {
      float v[32];
      tmem_ld32(tmem_addr(tbase, (warp & 3) * 32, cbase + c), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j++) acc += v[j];
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  if (acc == 123.456f) sink[tid] = acc;
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

static float frand_int(int lim) { return (float)((rand() % (2 * lim + 1)) - lim); }

int main() {
  srand(1);
  int fails = 0;
  long long *d_cyc;
  CK(cudaMalloc(&d_cyc, 64));
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

  // ---- test 1: canonical layout, M=128, N=64, K=64 --------------------------------------------
  {
    const int M = 128, N = 64, K = 64;
    std::vector<float> A(M * K), B(N * K);
    for (auto &v : A) v = frand_int(4);
    for (auto &v : B) v = frand_int(4);
    // canonical: elem (r,k) at (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2 ; LBO = 128, SBO = (K/8)*128
    auto pack = [&](const std::vector<float> &X, int R) {
      std::vector<__nv_bfloat16> o(R * K);
      for (int r = 0; r < R; r++)
        for (int k = 0; k < K; k++) o[((r / 8) * (K / 8) * 128 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2) / 2] = __float2bfloat16(X[r * K + k]);
      return o;
    };
    auto pa = pack(A, M), pb = pack(B, N);
    __nv_bfloat16 *dA, *dB;
    float *dD;
    CK(cudaMalloc(&dA, pa.size() * 2)); CK(cudaMalloc(&dB, pb.size() * 2)); CK(cudaMalloc(&dD, M * N * 4));
    CK(cudaMemcpy(dA, pa.data(), pa.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, pb.data(), pb.size() * 2, cudaMemcpyHostToDevice));
    Args a{dA, dB, (int)pa.size() * 2, (int)pb.size() * 2, 0, 128, (K / 8) * 128, 256, 0, 128, (K / 8) * 128, 256, M, N, K / 16, 1, 1, 0, nullptr, dD, d_cyc};
    k_probe<<<1, 128, 64 * 1024>>>(a);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(M * N);
    CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < M; m++)
      for (int n = 0; n < N; n++) {
        float ref = 0;
        for (int k = 0; k < K; k++) ref += A[m * K + k] * B[n * K + k];
        if (ref != D[m * N + n]) { if (bad < 5) printf("  t1 mismatch m=%d n=%d ref=%g got=%g\n", m, n, ref, D[m * N + n]); bad++; }
      }
    printf("test1 canonical K-major no-swizzle GEMM 128x64x64: %s (%d mismatches)\n", bad ? "FAIL" : "PASS", bad);
    fails += bad != 0;
  }

  // ---- test 2: sliding windows over a padded plane --------------------------------------------
  // plane[rows][pitch_units] of 16-byte units (8 bf16).  A(m,k): m = g*8+i -> unit (row0+g, col0+i) for k<8, next unit for k>=8.
  {
    const int rows = 20, pitch_u = 12, N = 32;
    std::vector<float> P(rows * pitch_u * 8), B(N * 16);
    for (auto &v : P) v = frand_int(4);
    for (auto &v : B) v = frand_int(4);
    std::vector<__nv_bfloat16> pp(P.size()), pb(N * 16);
    for (size_t i = 0; i < P.size(); i++) pp[i] = __float2bfloat16(P[i]);
    for (int n = 0; n < N; n++)
      for (int k = 0; k < 16; k++) pb[((n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) / 2] = __float2bfloat16(B[n * 16 + k]);
    __nv_bfloat16 *dA, *dB;
    float *dD;
    CK(cudaMalloc(&dA, pp.size() * 2)); CK(cudaMalloc(&dB, pb.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, pp.data(), pp.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, pb.data(), pb.size() * 2, cudaMemcpyHostToDevice));
    const int row0 = 2, col0 = 3;   // window origin: NOT 128-byte aligned
    Args a{dA, dB, (int)pp.size() * 2, (int)pb.size() * 2, (uint32_t)((row0 * pitch_u + col0) * 16), 16, pitch_u * 16, 0,
           0, 128, 256, 0, 128, N, 1, 1, 1, 0, nullptr, dD, d_cyc};
    k_probe<<<1, 128, 64 * 1024>>>(a);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * N);
    CK(cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < 128; m++)
      for (int n = 0; n < N; n++) {
        const int g = m / 8, i = m % 8;
        float ref = 0;
        for (int k = 0; k < 16; k++) ref += P[((row0 + g) * pitch_u + col0 + i) * 8 + k] * B[n * 16 + k];
        if (ref != D[m * N + n]) { if (bad < 5) printf("  t2 mismatch m=%d n=%d ref=%g got=%g\n", m, n, ref, D[m * N + n]); bad++; }
      }
    printf("test2 sliding-window descriptors (SBO=row pitch, LBO=16B, unaligned origin): %s (%d mismatches)\n", bad ? "FAIL" : "PASS", bad);
    fails += bad != 0;
  }

  // ---- test 3: cycles per MMA (A, B in smem), window-style A -----------------------------------
  {
    __nv_bfloat16 *dA, *dB;
    CK(cudaMalloc(&dA, 64 * 1024)); CK(cudaMalloc(&dB, 64 * 1024));
    CK(cudaMemset(dA, 0, 64 * 1024)); CK(cudaMemset(dB, 0, 64 * 1024));
    const int Ns[] = {16, 64, 128, 256};
    for (int M : {128, 64})
      for (int N : Ns)
        for (int nacc : {1, 4}) {
          if (nacc == 4 && N > 128) continue;
          Args a{dA, dB, 32 * 1024, 32 * 1024, 0, 16u, 192u, 0, 0, 128, 256, 0, M, N, 1, 512, nacc, 0, nullptr, nullptr, d_cyc};
          k_probe<<<1, 128, 80 * 1024>>>(a);
          CK(cudaDeviceSynchronize());
          long long c[2];
          CK(cudaMemcpy(c, d_cyc, 16, cudaMemcpyDeviceToHost));
          printf("test3 M=%3d N=%3d K=16 nacc=%d: %.1f cycles/MMA total, %.1f issue (ideal math %.0f)\n", M, N, nacc, c[0] / 512.0, c[1] / 512.0,
                 128.0 * N / 256);
        }
  }

  // ---- test 5: cycles per MMA with the operand layouts the CNN kernels actually use, distinct operands per MMA ------
  {
    __nv_bfloat16 *dA, *dB;
    CK(cudaMalloc(&dA, 64 * 1024)); CK(cudaMalloc(&dB, 64 * 1024));
    CK(cudaMemset(dA, 0, 64 * 1024)); CK(cudaMemset(dB, 0, 64 * 1024));
    struct Cfg { const char *name; uint32_t alb, asb, aks, blb, bsb, bks; int N, nacc; };
    const Cfg cfgs[] = {
        {"canonical A + canonical B, N=64", 128, 256, 4096, 128, 256, 2048, 64, 4},
        {"canonical A + canonical B, N=128", 128, 256, 4096, 128, 256, 4096, 128, 4},
        {"canonical A + canonical B, N=32", 128, 256, 4096, 128, 256, 1024, 32, 4},
        {"K2-like: window A (LBO 8192, SBO 288, +16B/step) + canonical B, N=64", 8192, 288, 16, 128, 256, 2048, 64, 4},
        {"K1-like: window A (LBO 8192, SBO 544, +16B/step) + canonical B, N=128", 8192, 544, 16, 128, 256, 4096, 128, 4},
        {"fc-like: A (LBO 128, SBO 1024, +256B/step) + B (LBO 128, SBO 1024), N=32", 128, 1024, 256, 128, 1024, 256, 32, 4},
        {"K3-like: canonical A + window B (LBO 6400, SBO 160, +16B/step), N=256", 128, 256, 4096, 6400, 160, 16, 256, 1},
        {"canonical A + canonical B, N=256", 128, 256, 4096, 128, 256, 0, 256, 1},
    };
    for (const Cfg &c : cfgs) {
      Args a{dA, dB, 48 * 1024, 48 * 1024, 0, c.alb, c.asb, c.aks, 0, c.blb, c.bsb, c.bks, 128, c.N, c.nacc == 1 ? 8 : 1, c.nacc == 1 ? 64 : 512, c.nacc, 0, nullptr, nullptr, d_cyc};
      k_probe<<<1, 128, 100 * 1024>>>(a);
      CK(cudaDeviceSynchronize());
      long long cy[2];
      CK(cudaMemcpy(cy, d_cyc, 16, cudaMemcpyDeviceToHost));
      const double bytes = 128 * 32.0 + c.N * 32.0;
      printf("test5 %-78s nacc=%d: %.1f cycles/MMA (%.0f B of operands -> %.0f B/cycle; math floor %.0f)\n", c.name, c.nacc, cy[0] / 512.0, bytes,
             bytes / (cy[0] / 512.0), 128.0 * c.N / 256);
    }
  }

  // ---- test 6: the same MMAs while the other warps of the CTA keep TMEM / shared memory / the bulk-copy engine busy ----
  {
    __nv_bfloat16 *dA, *dB;
    uint8_t *gsrc;
    CK(cudaMalloc(&dA, 64 * 1024)); CK(cudaMalloc(&dB, 64 * 1024)); CK(cudaMalloc(&gsrc, 64 * 1024));
    CK(cudaMemset(dA, 0, 64 * 1024)); CK(cudaMemset(dB, 0, 64 * 1024)); CK(cudaMemset(gsrc, 0, 64 * 1024));
    const char *bgname[4] = {"idle", "tcgen05.ld", "shared-memory traffic", "bulk copies"};
    for (int N : {64, 128})
      for (int bg = 0; bg < 4; bg++) {
        Args a{dA, dB, 48 * 1024, 48 * 1024, 0, 8192u, 288u, 16u, 0, 128, 256, (uint32_t)(N * 32), 128, N, 1, 512, 4, bg, gsrc, nullptr, d_cyc};
        k_probe<<<1, 128, 120 * 1024>>>(a);
        CK(cudaDeviceSynchronize());
        long long cy[2];
        CK(cudaMemcpy(cy, d_cyc, 16, cudaMemcpyDeviceToHost));
        printf("test6 window A + canonical B, N=%3d, 4 accumulators, other warps: %-22s %.1f cycles/MMA\n", N, bgname[bg], cy[0] / 512.0);
      }
  }

  // ---- test 4: tcgen05.ld throughput ------------------------------------------------------------
  {
    float *sink;
    CK(cudaMalloc(&sink, 4096));
    for (int warps : {4, 8}) {
      k_ldprobe<<<1, warps * 32>>>(64, 256, d_cyc, sink);
      CK(cudaDeviceSynchronize());
      long long c;
      CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
      const double bytes = 64.0 * 256 * 4 * 32 * warps;
      printf("test4 tcgen05.ld 32x32b.x32, %d warps: %.1f B/cycle/SM (%lld cycles)\n", warps, bytes / c, c);
    }
  }
  printf("tc_probe: %s\n", fails ? "FAILED" : "ALL PASS");
  return fails ? 1 : 0;
}
