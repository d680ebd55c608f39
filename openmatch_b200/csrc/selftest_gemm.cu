// Standalone bring-up / regression probe for the tcgen05 GEMM core (not part of the shipped library).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/selftest_gemm selftest_gemm.cu
//   run  : build/selftest_gemm [dump_dir]
// Exact check: small-integer bf16 operands make every product and partial sum exactly representable
// in fp32, so the tensor-core result must equal the CPU result bit for bit.
#include <algorithm>
#include <execinfo.h>
#include <math.h>
#include <signal.h>
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "gemm.cuh"
#include "gemm2sm.cuh"
#include "scan_epilogue.cuh"

using namespace om;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static uint32_t rng_state = 12345u;
static inline uint32_t rnd() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 8;
}

struct EpiCount {  // perf probe: counts accumulators above a threshold (mimics the search filter)
  unsigned long long* counter;
  float thr;
  int M, N;
  static constexpr int kPasses = 1;
  static constexpr bool kPrefetch = false;
  __host__ __device__ static constexpr int smem_bytes(int) { return 0; }
  struct State {
    int cnt;
  };
  __device__ __forceinline__ void begin(State& s, int, int, int) const { s.cnt = 0; }
  __device__ __forceinline__ void chunk(State& s, int row, int col0, const float (&v)[32]) const {
#pragma unroll
    for (int i = 0; i < 32; ++i) s.cnt += (v[i] > thr) ? 1 : 0;
  }
  __device__ __forceinline__ void end(State& s, int row) const {
    if (s.cnt) atomicAdd(counter, (unsigned long long)s.cnt);
  }
};

template <int BN, int STAGES, bool MF, int EW = 4>
static int check_case(int M, int N, int K, int num_sms, const char* dump_dir, bool dyn = false) {
  printf("[case] BN=%d STAGES=%d M_FASTEST=%d EW=%d dyn=%d  M=%d N=%d K=%d ... ", BN, STAGES, (int)MF, EW, (int)dyn, M, N, K);
  fflush(stdout);
  std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);
  std::vector<float> fA((size_t)M * K), fB((size_t)N * K);
  for (size_t i = 0; i < hA.size(); ++i) {
    fA[i] = (float)((int)(rnd() % 7) - 3);
    hA[i] = __float2bfloat16(fA[i]);
  }
  for (size_t i = 0; i < hB.size(); ++i) {
    fB[i] = (float)((int)(rnd() % 7) - 3);
    hB[i] = __float2bfloat16(fB[i]);
  }
  __nv_bfloat16 *dA, *dB;
  float* dC;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dC, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, (size_t)M * N * 4));
  EpiStoreF32 epi{dC, N, nullptr, nullptr, 0, M, N};
  cudaError_t e = launch_gemm<BN, STAGES, MF, EW>(dA, K, dB, K, M, N, K, epi, num_sms, 0, dyn);
  if (e != cudaSuccess) {
    printf("LAUNCH FAILED: %s\n", cudaGetErrorString(e));
    return 1;
  }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("KERNEL FAILED: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  unsigned int fault = read_clear_dev_fault();
  if (fault) printf("DEVICE FAULT word=0x%08x (site %u, block %u) ", fault, (fault >> 16) & 0x7fff, fault & 0xffff);
  std::vector<float> hC((size_t)M * N);
  CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  double maxerr = 0;
  int printed = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < K; ++k) ref += fA[(size_t)m * K + k] * fB[(size_t)n * K + k];
      float got = hC[(size_t)m * N + n];
      if (!(got == ref)) {
        ++bad;
        double d = fabs((double)got - ref);
        if (d > maxerr || d != d) maxerr = d;
        if (printed < 6) {
          printf("\n   mismatch C[%d,%d] got %g want %g", m, n, got, ref);
          ++printed;
        }
      }
    }
  printf("%s  (%zu / %zu mismatches, max |err| %g)\n", bad ? "\n   FAIL" : "ok", bad, hC.size(), maxerr);
  if (bad && dump_dir) {
    char path[512];
    snprintf(path, sizeof path, "%s/gemm_dump_BN%d_M%d_N%d_K%d.bin", dump_dir, BN, M, N, K);
    FILE* f = fopen(path, "wb");
    if (f) {
      int hdr[4] = {M, N, K, BN};
      fwrite(hdr, 4, 4, f);
      fwrite(fA.data(), 4, fA.size(), f);
      fwrite(fB.data(), 4, fB.size(), f);
      fwrite(hC.data(), 4, hC.size(), f);
      fclose(f);
      printf("   dumped operands + result to %s\n", path);
    }
  }
  cudaFree(dA), cudaFree(dB), cudaFree(dC);
  return bad ? 1 : 0;
}

template <int BN, int STAGES, bool MF, int EW = 4>
static void perf_case(const char* name, int M, int N, int K, int num_sms, int iters);

// 2-CTA (cta_group::2) core: same exact check
template <int STAGES, bool MF, int EW>
static int check_case2(int M, int N, int K, int num_sms, bool dyn = false) {
  printf("[case 2sm] STAGES=%d M_FASTEST=%d EW=%d dyn=%d  M=%d N=%d K=%d ... ", STAGES, (int)MF, EW, (int)dyn, M, N, K);
  fflush(stdout);
  std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);
  std::vector<float> fA((size_t)M * K), fB((size_t)N * K);
  for (size_t i = 0; i < hA.size(); ++i) {
    fA[i] = (float)((int)(rnd() % 7) - 3);
    hA[i] = __float2bfloat16(fA[i]);
  }
  for (size_t i = 0; i < hB.size(); ++i) {
    fB[i] = (float)((int)(rnd() % 7) - 3);
    hB[i] = __float2bfloat16(fB[i]);
  }
  __nv_bfloat16 *dA, *dB;
  float* dC;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dC, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, (size_t)M * N * 4));
  EpiStoreF32 epi{dC, N, nullptr, nullptr, 0, M, N};
  cudaError_t e = launch_gemm2<STAGES, MF, EW>(dA, K, dB, K, M, N, K, epi, num_sms, 0, dyn);
  if (e != cudaSuccess) {
    printf("LAUNCH FAILED: %s\n", cudaGetErrorString(e));
    return 1;
  }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("KERNEL FAILED: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  unsigned int fault = read_clear_dev_fault();
  if (fault) printf("DEVICE FAULT word=0x%08x (site %u, block %u) ", fault, (fault >> 16) & 0x7fff, fault & 0xffff);
  std::vector<float> hC((size_t)M * N);
  CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  int printed = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < K; ++k) ref += fA[(size_t)m * K + k] * fB[(size_t)n * K + k];
      const float got = hC[(size_t)m * N + n];
      if (!(got == ref)) {
        ++bad;
        if (printed < 8) {
          printf("\n   mismatch C[%d,%d] got %g want %g", m, n, got, ref);
          ++printed;
        }
      }
    }
  printf("%s  (%zu / %zu mismatches)\n", bad ? "\n   FAIL" : "ok", bad, hC.size());
  cudaFree(dA), cudaFree(dB), cudaFree(dC);
  return (bad || fault) ? 1 : 0;
}

template <int STAGES, bool MF, int EW, int MODE = 0>
static void perf_case2(const char* name, int M, int N, int K, int num_sms, int iters, bool dyn = false) {
  __nv_bfloat16 *dA, *dB;
  unsigned long long* dcnt;
  CK(cudaMalloc(&dA, (size_t)M * K * 2));
  CK(cudaMalloc(&dB, (size_t)N * K * 2));
  CK(cudaMalloc(&dcnt, 8));
  CK(cudaMemset(dcnt, 0, 8));
  {
    std::vector<uint16_t> h((size_t)1 << 22);
    for (auto& x : h) x = (uint16_t)(0x3c00 + (rnd() % 0x400)) | (uint16_t)((rnd() & 1) << 15);
    for (size_t off = 0; off < (size_t)M * K; off += h.size())
      CK(cudaMemcpy(dA + off, h.data(), std::min(h.size(), (size_t)M * K - off) * 2, cudaMemcpyHostToDevice));
    for (size_t off = 0; off < (size_t)N * K; off += h.size())
      CK(cudaMemcpy(dB + off, h.data(), std::min(h.size(), (size_t)N * K - off) * 2, cudaMemcpyHostToDevice));
  }
  EpiCount epi{dcnt, 1.0e30f, M, N};
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int pairs = 0;
  for (int i = 0; i < 2; ++i) CK((launch_gemm2<STAGES, MF, EW, false, MODE>(dA, K, dB, K, M, N, K, epi, num_sms, 0, dyn, &pairs)));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) CK((launch_gemm2<STAGES, MF, EW, false, MODE>(dA, K, dB, K, M, N, K, epi, num_sms, 0, dyn)));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  unsigned int fault = read_clear_dev_fault();
  double tf = 2.0 * M * N * (double)K / (ms * 1e-3) / 1e12;
  printf("[perf 2sm mode=%d dyn=%d] %-28s ST=%d MF=%d EW=%d pairs=%d  M=%d N=%d K=%d : %.3f ms  %.1f TFLOP/s  fault=0x%x\n", MODE,
         (int)dyn, name, STAGES, (int)MF, EW, pairs, M, N, K, ms, tf, fault);
  fflush(stdout);
  cudaFree(dA), cudaFree(dB), cudaFree(dcnt);
}

template <int EW, bool F16 = false, bool PAIR = false>
static void perf_scan(const char* name, int M, int N, int K, int num_sms, int iters, float thr_value, bool dyn = false,
                      std::vector<unsigned long long>* keys_out = nullptr);

static int run_2sm(int sms) {
  int fails = 0;
  fails += check_case2<6, false, 4>(256, 256, 64, sms);     // one pair tile, one k-block
  if (fails) return fails;                                  // nothing else can work
  fails += check_case2<6, false, 8>(300, 520, 192, sms);    // ragged edges, several tiles
  fails += check_case2<6, true, 8>(1000, 3000, 768, sms);   // both accumulator buffers, M fastest
  fails += check_case2<6, false, 4>(256, 256, 512, sms);    // ring wraps
  fails += check_case2<5, true, 8>(3000, 9000, 256, sms, true);   // dynamic pair scheduler
  fails += check_case2<5, false, 4>(300, 520, 192, sms, true);    // fewer tiles than pairs
  fails += check_case2<5, true, 8>(2048, 2304, 768, sms, true);
  // the product's scan epilogue on both cores: identical survivor sets
  {
    std::vector<unsigned long long> k1, k2;
    perf_scan<8, false, false>("scan 1M single-CTA", 6980, 1 << 20, 768, sms, 3, 38.f, true, &k1);
    perf_scan<8, false, true>("scan 1M pair", 6980, 1 << 20, 768, sms, 3, 38.f, true, &k2);
    const bool same = k1 == k2;
    printf("[scan pair vs single] survivor sets %s (%zu words)\n", same ? "IDENTICAL" : "DIFFER", k1.size());
    fails += same ? 0 : 1;
  }
  // rate probes (garbage results): the pair-wide MMA alone, the 2-CTA TMA path alone
  perf_case2<6, false, 8, 1>("8192^3 MMA only", 8192, 8192, 8192, sms, 5);
  perf_case2<6, false, 8, 2>("8192^3 loads only", 8192, 8192, 8192, sms, 5);
  perf_case2<6, false, 8>("encoder FFN1 shape", 32768, 3072, 768, sms, 10);
  perf_case<256, 4, false, 8>("encoder FFN1 shape (1-CTA)", 32768, 3072, 768, sms, 10);
  perf_case2<6, false, 8>("cublas-peak shape 8192^3", 8192, 8192, 8192, sms, 5);
  perf_case<256, 4, false, 8>("cublas-peak shape 8192^3 (1-CTA)", 8192, 8192, 8192, sms, 5);
  perf_case2<5, true, 8>("search 6980 x 4M sustained", 6980, 1 << 22, 768, sms, 12, true);
  perf_case<256, 4, true, 8>("search 6980 x 4M sustained (1-CTA)", 6980, 1 << 22, 768, sms, 12);
  for (int rep = 0; rep < 2; ++rep) {
    perf_scan<8, false, true>("scan 4M pair", 6980, 1 << 22, 768, sms, 12, 42.f, true);
    perf_scan<8, false, false>("scan 4M single-CTA", 6980, 1 << 22, 768, sms, 12, 42.f, true);
  }
  perf_scan<8, false, true>("scan 4M pair static", 6980, 1 << 22, 768, sms, 12, 42.f, false);
  fflush(stdout);
  return fails;
}

template <int BN, int STAGES, bool MF, int EW>
static void perf_case(const char* name, int M, int N, int K, int num_sms, int iters) {
  __nv_bfloat16 *dA, *dB;
  unsigned long long* dcnt;
  CK(cudaMalloc(&dA, (size_t)M * K * 2));
  CK(cudaMalloc(&dB, (size_t)N * K * 2));
  CK(cudaMalloc(&dcnt, 8));
  CK(cudaMemset(dcnt, 0, 8));
  // pseudo-random bf16 bit patterns in a sane exponent range: 0x3c00..0x3fff (|x| in [0.0078, 2)) with sign
  {
    std::vector<uint16_t> h((size_t)1 << 22);
    for (auto& x : h) x = (uint16_t)(0x3c00 + (rnd() % 0x400)) | (uint16_t)((rnd() & 1) << 15);
    for (size_t off = 0; off < (size_t)M * K; off += h.size())
      CK(cudaMemcpy(dA + off, h.data(), std::min(h.size(), (size_t)M * K - off) * 2, cudaMemcpyHostToDevice));
    for (size_t off = 0; off < (size_t)N * K; off += h.size())
      CK(cudaMemcpy(dB + off, h.data(), std::min(h.size(), (size_t)N * K - off) * 2, cudaMemcpyHostToDevice));
  }
  EpiCount epi{dcnt, 1.0e30f, M, N};
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) CK((launch_gemm<BN, STAGES, MF, EW>(dA, K, dB, K, M, N, K, epi, num_sms, 0)));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) CK((launch_gemm<BN, STAGES, MF, EW>(dA, K, dB, K, M, N, K, epi, num_sms, 0)));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  unsigned int fault = read_clear_dev_fault();
  double tf = 2.0 * M * N * (double)K / (ms * 1e-3) / 1e12;
  printf("[perf] %-28s BN=%d ST=%d MF=%d  M=%d N=%d K=%d : %.3f ms  %.1f TFLOP/s  fault=0x%x\n", name, BN, STAGES,
         (int)MF, M, N, K, ms, tf, fault);
  cudaFree(dA), cudaFree(dB), cudaFree(dcnt);
}

// the product's scan epilogue on the search shape, thresholds set so that nothing survives
template <int EW, bool F16, bool PAIR>
static void perf_scan(const char* name, int M, int N, int K, int num_sms, int iters, float thr_value, bool dyn,
                      std::vector<unsigned long long>* keys_out) {
  if (keys_out) rng_state = 777u;  // survivor sets are compared across calls: same operands every time
  auto launch = [&](const EpiScan<false, EW * 32>& epi, __nv_bfloat16* dA, __nv_bfloat16* dB) {
    if constexpr (PAIR)
      return launch_gemm2<5, true, EW, F16>(dA, K, dB, K, M, N, K, epi, num_sms, 0, dyn);
    else
      return launch_gemm<256, 4, true, EW, EpiScan<false, EW * 32>, F16>(dA, K, dB, K, M, N, K, epi, num_sms, 0, dyn);
  };
  __nv_bfloat16 *dA, *dB;
  CK(cudaMalloc(&dA, (size_t)M * K * 2));
  CK(cudaMalloc(&dB, (size_t)N * K * 2));
  {
    std::vector<uint16_t> h((size_t)1 << 22);
    for (auto& x : h) x = (uint16_t)(0x3c00 + (rnd() % 0x400)) | (uint16_t)((rnd() & 1) << 15);
    for (size_t off = 0; off < (size_t)M * K; off += h.size())
      CK(cudaMemcpy(dA + off, h.data(), std::min(h.size(), (size_t)M * K - off) * 2, cudaMemcpyHostToDevice));
    // different values for B (A == B rows would make |x|^2-sized scores), and a period that is not a multiple
    // of the row length
    std::vector<uint16_t> hb(((size_t)1 << 22) + 4099);
    for (auto& x : hb) x = (uint16_t)(0x3c00 + ((rnd() >> 3) % 0x400)) | (uint16_t)(((rnd() >> 5) & 1) << 15);
    for (size_t off = 0; off < (size_t)N * K; off += hb.size())
      CK(cudaMemcpy(dB + off, hb.data(), std::min(hb.size(), (size_t)N * K - off) * 2, cudaMemcpyHostToDevice));
  }
  const int C = 1 << 16;
  float* thr;
  unsigned long long* cand;
  int *count, *ovf;
  CK(cudaMalloc(&thr, M * 4));
  CK(cudaMalloc(&cand, (size_t)M * C * 8));
  CK(cudaMalloc(&count, M * 4));
  CK(cudaMalloc(&ovf, 4));
  std::vector<float> ht(M, thr_value);
  CK(cudaMemcpy(thr, ht.data(), M * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(count, 0, M * 4));
  CK(cudaMemset(ovf, 0, 4));
  EpiScan<false, EW * 32> epi{thr, cand, count, ovf, M, N, C, 0u};
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) CK(launch(epi, dA, dB));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) CK(launch(epi, dA, dB));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  int hovf = 0;
  CK(cudaMemcpy(&hovf, ovf, 4, cudaMemcpyDeviceToHost));
  // survivors per launch (count[] accumulated over warm-up + timed launches)
  std::vector<int> hc(M);
  CK(cudaMemcpy(hc.data(), count, M * 4, cudaMemcpyDeviceToHost));
  double surv = 0;
  for (int c : hc) surv += c;
  surv /= (iters + 2);
  if (keys_out) {  // survivors of ONE launch, sorted per query: identical across filter variants by construction
    CK(cudaMemset(count, 0, M * 4));
    CK(launch(epi, dA, dB));
    CK(cudaDeviceSynchronize());
    std::vector<int> c1(M);
    CK(cudaMemcpy(c1.data(), count, M * 4, cudaMemcpyDeviceToHost));
    keys_out->clear();
    std::vector<unsigned long long> row(C);
    for (int q = 0; q < M; ++q) {
      const int n = std::min(c1[q], C);
      CK(cudaMemcpy(row.data(), cand + (size_t)q * C, (size_t)n * 8, cudaMemcpyDeviceToHost));
      std::sort(row.begin(), row.begin() + n);
      keys_out->push_back((unsigned long long)n);
      keys_out->insert(keys_out->end(), row.begin(), row.begin() + n);
    }
  }
  printf("[perf] %-24s EpiScan pair=%d EW=%d dyn=%d thr=%g N=%d : %.3f ms %.1f TFLOP/s fault=0x%x ovf=%d survivors/query=%.0f (%.2f per warp-tile)\n",
         name, (int)PAIR, EW, (int)dyn, thr_value, N, ms, 2.0 * M * N * (double)K / (ms * 1e-3) / 1e12, read_clear_dev_fault(), hovf,
         surv / M, surv / M * 32.0 / (N / 256.0) / (EW / 4));
  cudaFree(dA), cudaFree(dB), cudaFree(thr), cudaFree(cand), cudaFree(count), cudaFree(ovf);
}

static void on_segv(int sig) {
  void* frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "\n*** fatal signal, backtrace:\n";
  (void)!write(1, msg, sizeof msg - 1);
  backtrace_symbols_fd(frames, n, 1);
  _exit(128 + sig);
}

// filter variants of scan_epilogue.cuh: same survivor sets (exact), cost per survivor compared at several rates
int main(int argc, char** argv) {
  signal(SIGSEGV, on_segv);
  signal(SIGABRT, on_segv);
  signal(SIGBUS, on_segv);
  const char* dump_dir = argc > 1 ? argv[1] : nullptr;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s  sm_%d%d  SMs=%d  smem/block optin=%zu\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  const int sms = prop.multiProcessorCount;
  int fails = 0;
  if (argc > 1 && strcmp(argv[1], "--f16cmp") == 0) {
    // fp16 vs bf16 operands on the scan shape, long enough to sit in the power-capped (sustained) regime
    for (int rep = 0; rep < 2; ++rep) {
      perf_scan<8, false>("bf16 sustained 4M", 6980, 1 << 22, 768, sms, 12, 42.f, true);
      perf_scan<8, true>("fp16 sustained 4M", 6980, 1 << 22, 768, sms, 12, 42.f, true);
    }
    return 0;
  }
  if (argc > 1 && strcmp(argv[1], "--2sm") == 0) {
    const int f2 = run_2sm(sms);
    printf("2sm: %d failing case(s)\n", f2);
    return f2 ? 1 : 0;
  }
  fails += check_case<256, 4, false>(128, 256, 64, sms, dump_dir);   // one tile, one k-block
  fails += check_case<256, 4, false>(128, 256, 256, sms, dump_dir);  // one tile, ring wraps once
  fails += check_case<256, 4, false>(300, 520, 192, sms, dump_dir);  // ragged edges, several tiles
  fails += check_case<128, 4, false>(300, 520, 192, sms, dump_dir);
  fails += check_case<64, 4, false>(200, 200, 128, sms, dump_dir);
  fails += check_case<256, 4, true>(1000, 3000, 768, sms, dump_dir);   // > 1 tile per CTA, both acc buffers
  fails += check_case<256, 4, false>(2048, 2304, 768, sms, dump_dir);  // many tiles per CTA
  fails += check_case<256, 4, false, 8>(2048, 2304, 768, sms, dump_dir);  // 8 epilogue warps
  fails += check_case<64, 4, false, 8>(300, 200, 128, sms, dump_dir);
  fails += check_case<192, 5, false, 8>(1000, 768, 768, sms, dump_dir);  // 192-wide tiles, odd chunk tail
  fails += check_case<192, 5, false, 4>(300, 500, 192, sms, dump_dir);
  fails += check_case<256, 4, true, 8>(3000, 9000, 256, sms, dump_dir, true);   // dynamic tile scheduler
  fails += check_case<256, 4, false, 4>(300, 520, 192, sms, dump_dir, true);    // fewer tiles than CTAs
  fails += check_case<192, 5, false, 8>(2048, 768, 768, sms, dump_dir, true);
  if (fails) {
    printf("SELFTEST FAILED (%d cases)\n", fails);
    return 1;
  }
  perf_case<256, 4, false>("encoder FFN1 shape", 32768, 3072, 768, sms, 10);
  perf_case<256, 4, false>("encoder FFN2 shape", 32768, 768, 3072, sms, 10);
  perf_case<256, 4, false>("encoder QKV shape", 32768, 2304, 768, sms, 10);
  perf_case<128, 6, false>("encoder FFN1 shape", 32768, 3072, 768, sms, 10);
  perf_case<192, 5, false, 8>("encoder FFN2 shape BN192", 32768, 768, 3072, sms, 10);
  perf_case<256, 4, true>("search 6980 x 1M", 6980, 1 << 20, 768, sms, 3);
  perf_case<256, 4, true, 8>("search 6980 x 1M, 8 epi warps", 6980, 1 << 20, 768, sms, 3);
  perf_scan<8>("scan epilogue, no survivors", 6980, 1 << 20, 768, sms, 3, 1e30f);
  perf_scan<4>("scan epilogue, no survivors", 6980, 1 << 20, 768, sms, 3, 1e30f);
  for (float t : {48.f, 42.f, 38.f, 35.f, 32.f}) perf_scan<8>("scan 1M", 6980, 1 << 20, 768, sms, 3, t, true);
  perf_scan<8>("scan 1M static", 6980, 1 << 20, 768, sms, 3, 38.f, false);
  perf_scan<8>("scan epilogue, 4M rows", 6980, 1 << 22, 768, sms, 2, 1e30f);
  perf_scan<8>("scan epilogue, 4M rows", 6980, 1 << 22, 768, sms, 2, 1e30f, true);
  perf_scan<8>("scan epilogue, thr 45", 6980, 1 << 22, 768, sms, 2, 45.0f, true);
  perf_case<256, 4, false>("search 6980 x 1M (n fastest)", 6980, 1 << 20, 768, sms, 3);
  perf_case<256, 4, false>("cublas-peak shape 8192^3", 8192, 8192, 8192, sms, 5);
  printf("SELFTEST OK\n");
  return 0;
}
