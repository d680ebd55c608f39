// Exact inner-product top-k over an HBM-resident flat index — the B200 replacement of
// faiss.IndexFlatIP.{add,search,reset} as the reference uses it
// (src/openmatch/retriever/dense_retriever.py:38-41,105,133-137,180) and of the IndexShards merge behind
// index_cpu_to_gpu_multiple(shard=True) (:43-58).
//
// Storage per index shard: fp32 master rows [n, d] (what index.add received; used for exact re-scoring)
// plus a bf16 scan copy [n, dpad] (tensor-core operand).
//
// search(q, k):
//   1. SCAN   bf16 Q * X^T on tcgen05 (gemm.cuh mainloop) with the top-k filter fused into the epilogue:
//             scores never leave TMEM/registers; a thread owns one query row, compares its 32-column chunk
//             against that query's running threshold and appends the rare survivors
//             (key = orderable(score) << 32 | ~row) to the query's candidate list in HBM.
//             The corpus is swept in rounds of geometrically growing size; after each round
//   2. SELECT a per-query bitonic sort in shared memory keeps the best kp = k + slack candidates and
//             publishes the kp-th score as the next round's (strict) threshold.  Expected survivors per
//             round ~ kp, so the list capacity C >= 2.5 kp + 512 is ample for exchangeable data; an overflow (e.g.
//             adversarially sorted corpus) is detected and the query chunk is redone with an overflow-proof
//             fixed-size round schedule, so the result is always exact w.r.t. the bf16 stage.
//   3. FINAL  the kp candidates are re-scored against the fp32 master rows (fp32 FMA), sorted by
//             (score desc, row asc) and the top k emitted as (D fp32, I int64) — faiss's output contract.
#include <float.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "common.h"
#include "gemm.cuh"
#include "scan_epilogue.cuh"

#ifndef OM_SCAN_VARIANT
#define OM_SCAN_VARIANT 0  // filter code shape, see scan_epilogue.cuh (1 / 2: unmeasured candidates)
#endif

namespace om {

// ---------------------------------------------------------------------------------------------------
// shared-memory bitonic sort (descending) of P = 2^m keys by nthreads threads
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* s, int P, int tid, int nthreads) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (P >> 1); t += nthreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // insert a 0 bit at position log2(j)
        const int p = i | j;
        const unsigned long long a = s[i], b = s[p];
        const bool desc = (i & k) == 0;
        if ((a < b) == desc) {
          s[i] = b;
          s[p] = a;
        }
      }
      __syncthreads();
    }
  }
}

// SELECT: one CTA per query.  MSB-first radix select of the kp-th largest key among the cnt candidates, then
// compaction of the keys >= it to the head of the list (unordered) and publication of its score as the new
// strict threshold.  The 32 score bits are resolved with three 11/11/10-bit passes; the 32 row bits are only
// walked (four 8-bit passes) when the kp-th score is tied with more candidates than there are slots left.
// (A full shared-memory sort here costs ~35 GB of smem traffic per round at nq = 6980 — it was 43 % of the
// search time.)
__global__ void __launch_bounds__(256) select_kernel(unsigned long long* cand, int* count, float* thr, int C, int kp,
                                                     int cnt_override) {
  extern __shared__ unsigned long long skeys[];
  __shared__ int hist[2048];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining, s_out, s_bin_count, s_wsum[8];
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  unsigned long long* mine = cand + static_cast<size_t>(q) * C;
  int cnt = cnt_override >= 0 ? cnt_override : count[q];
  cnt = cnt < C ? cnt : C;
  if (cnt < kp) {  // nothing to drop yet: no threshold
    if (tid == 0) {
      count[q] = cnt;
      thr[q] = __int_as_float(0xff800000);
    }
    return;
  }
  if (cnt == kp) {
    // exactly kp entries (a round without survivors, or a shard of exactly kp rows): the list stays as it is and
    // the threshold is its smallest key.  (Resetting it to -inf here let the next round accept every row.)
    unsigned long long m = ~0ull;
    for (int i = tid; i < cnt; i += blockDim.x) m = min(m, mine[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) skeys[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i < static_cast<int>(blockDim.x >> 5); ++i) m = min(m, skeys[i]);
      count[q] = cnt;
      thr[q] = key_score(m);
    }
    return;
  }
  for (int i = tid; i < cnt; i += blockDim.x) skeys[i] = mine[i];
  if (tid == 0) {
    s_prefix = 0ull;
    s_remaining = kp;
    s_out = 0;
  }
  __syncthreads();
  unsigned long long mask = 0ull;
  // digit schedule: (shift, bits)
  const int shifts[7] = {53, 42, 32, 24, 16, 8, 0};
  const int widths[7] = {11, 11, 10, 8, 8, 8, 8};
  for (int pass = 0; pass < 7; ++pass) {
    const int shift = shifts[pass], nbins = 1 << widths[pass];
    for (int i = tid; i < nbins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    for (int i0 = 0; i0 < cnt; i0 += blockDim.x) {
      const int i = i0 + tid;
      const bool live = i < cnt && (skeys[i] & mask) == prefix;
      const int digit = live ? static_cast<int>((skeys[i] >> shift) & static_cast<unsigned long long>(nbins - 1)) : -1;
      // warp-aggregate equal digits (top bits of similar scores collide heavily)
      const unsigned peers = __match_any_sync(0xffffffffu, digit);
      if (live && lane == (__ffs(peers) - 1)) atomicAdd(&hist[digit], __popc(peers));
    }
    __syncthreads();
    {
      // Block-wide scan from the top bin down: thread t owns the `per` consecutive bins at positions
      // [t * per, (t + 1) * per) counted from the top (read in a per-lane rotated order so the 32 lanes hit 32
      // banks); find the bin where the running count reaches `remaining`.
      const int per = nbins >> 8;  // 8, 4 or 1
      const int remaining = s_remaining;
      const int rot = per > 1 ? lane / (32 / per) : 0;
      int sum = 0;
      for (int j = 0; j < per; ++j) sum += hist[nbins - 1 - (tid * per + ((j + rot) & (per - 1)))];
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_wsum[tid >> 5] = incl;
      __syncthreads();
      int base = 0;
      for (int w = 0; w < (tid >> 5); ++w) base += s_wsum[w];
      const int excl = base + incl - sum;
      if (excl < remaining && remaining <= excl + sum) {  // exactly one thread
        int run = excl;
        for (int j = 0; j < per; ++j) {
          const int bin = nbins - 1 - (tid * per + j), c = hist[bin];
          if (run < remaining && remaining <= run + c) {
            s_prefix = prefix | (static_cast<unsigned long long>(bin) << shift);
            s_remaining = remaining - run;
            s_bin_count = c;
          }
          run += c;
        }
      }
    }
    mask |= static_cast<unsigned long long>(nbins - 1) << shift;
    __syncthreads();
    // after the score bits: if every candidate sharing the kp-th score fits, no need to look at row bits
    if (pass == 2 && s_bin_count == s_remaining) break;
  }
  // keys >= kth under `mask` (keys are unique, so with all 64 bits resolved exactly kp keys qualify; with only
  // the score bits resolved the whole tie group qualifies and it fits by the check above)
  const unsigned long long kth = s_prefix;
  for (int i0 = 0; i0 < cnt; i0 += blockDim.x) {  // compaction, one shared-memory atomic per warp
    const int i = i0 + tid;
    const bool keep = i < cnt && (skeys[i] & mask) >= kth;
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    int base = 0;
    if (lane == 0 && b) base = atomicAdd(&s_out, __popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) mine[base + __popc(b & ((1u << lane) - 1u))] = skeys[i];
  }
  if (tid == 0) {
    count[q] = kp;
    thr[q] = key_score(kth);
  }
}

// Equal-width score bins shared by every shard of a sharded search (same inputs -> same bin on every rank).
constexpr int kFloorBins = 256;
struct FloorBins {
  float lo, scale;  // bin = floor((s - lo) * scale), clamped to [0, kFloorBins - 1]; s < lo -> -1
  __device__ static FloorBins make(float lo, float hi) {
    FloorBins f;
    f.lo = lo;
    const float w = hi - lo;
    f.scale = (w > 0.f && w < 3.0e38f) ? static_cast<float>(kFloorBins) / w : 0.f;  // lo = -inf / hi <= lo: one bin
    return f;
  }
  __device__ int bin(float s) const {
    if (s < lo) return -1;
    if (!(scale > 0.f)) return 0;
    const float t = (s - lo) * scale;
    return t >= static_cast<float>(kFloorBins - 1) ? kFloorBins - 1 : static_cast<int>(t);
  }
};

// FINAL: one CTA per query: exact fp32 re-score of the surviving candidates against the master rows,
// sort by (score desc, row asc), emit top-k.
__global__ void __launch_bounds__(256) finalize_kernel(const unsigned long long* cand, const int* count, int C,
                                                       const float* __restrict__ qf, const float* __restrict__ xf,
                                                       int d, int k, float* D, int64_t* I, int64_t id_offset,
                                                       const float* __restrict__ range, const int* __restrict__ ghist,
                                                       int nq_total, int kp, int* kept_max) {
  extern __shared__ unsigned long long fsm[];
  const int q = blockIdx.x;
  int cnt = count[q];
  int P = 2;
  while (P < cnt) P <<= 1;
  unsigned long long* skeys = fsm;
  float* sq = reinterpret_cast<float*>(fsm + P);
  for (int i = threadIdx.x; i < d; i += blockDim.x) sq[i] = qf[static_cast<size_t>(q) * d + i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const unsigned long long* mine = cand + static_cast<size_t>(q) * C;
  // sharded search: only candidates in or above the histogram bin that holds the global kp-th bf16-stage score
  // can be in the global top-k; the others are not re-scored
  __shared__ int s_minbin;
  __shared__ int s_n;
  FloorBins fb{__int_as_float(0xff800000), 0.f};
  if (range) fb = FloorBins::make(range[q], range[nq_total + q]);
  if (threadIdx.x == 0) {
    s_minbin = 0;
    s_n = 0;
  }
  __syncthreads();
  if (range && ghist && warp == 0) {
    // lowest bin b with count(bins >= b) >= kp: lane l owns bins [l * per, (l + 1) * per); suffix sums over lanes
    constexpr int per = kFloorBins / 32;
    int h[per], sum = 0;
#pragma unroll
    for (int j = 0; j < per; ++j) {
      h[j] = ghist[static_cast<size_t>(q) * kFloorBins + lane * per + j];
      sum += h[j];
    }
    int suffix = sum;  // sum over lanes >= this one
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_down_sync(0xffffffffu, suffix, o);
      if (lane + o < 32) suffix += v;
    }
    int run = suffix - sum;  // count in the bins above this lane's
    if (run < kp && kp <= suffix) {  // at most one lane; none if the lists hold fewer than kp rows (no pruning)
      int b = per - 1;
#pragma unroll
      for (int j = per - 1; j >= 0; --j) {
        if (run < kp) b = j;
        run += h[j];
      }
      s_minbin = lane * per + b;
    }
  }
  __syncthreads();
  const int minbin = s_minbin;
  for (int j = warp; j < cnt; j += nw) {
    if (range && fb.bin(key_score(mine[j])) < minbin) continue;
    const uint32_t row = key_row(mine[j]);
    const float* x = xf + static_cast<size_t>(row) * d;
    float acc = 0.f;
    if ((d & 3) == 0) {
      const float4* x4 = reinterpret_cast<const float4*>(x);
      const float4* q4 = reinterpret_cast<const float4*>(sq);
      for (int i = lane; i < (d >> 2); i += 32) {
        const float4 a = __ldg(x4 + i), b = q4[i];
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
        acc = fmaf(a.z, b.z, acc);
        acc = fmaf(a.w, b.w, acc);
      }
    } else {
      for (int i = lane; i < d; i += 32) acc = fmaf(__ldg(x + i), sq[i], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) skeys[atomicAdd(&s_n, 1)] = make_key(acc, row);  // compacted: the sort covers survivors only
  }
  __syncthreads();
  const int n = s_n;
  int P2 = 2;
  while (P2 < n) P2 <<= 1;
  for (int i = n + threadIdx.x; i < P2; i += blockDim.x) skeys[i] = 0ull;
  __syncthreads();
  bitonic_sort_desc(skeys, P2, threadIdx.x, blockDim.x);
  for (int r = threadIdx.x; r < k; r += blockDim.x) {
    float s = -FLT_MAX;
    int64_t id = -1;
    if (r < n) {
      s = key_score(skeys[r]);
      id = id_offset + static_cast<int64_t>(key_row(skeys[r]));
    }
    D[static_cast<size_t>(q) * k + r] = s;
    I[static_cast<size_t>(q) * k + r] = id;
  }
  // longest valid prefix over the queries: lets the caller exchange [nq, kept] instead of [nq, k]
  if (kept_max && threadIdx.x == 0 && n > 0) atomicMax(kept_max, n < k ? n : k);
}

// COUNT (sharded search): histogram of this shard's surviving candidates over kFloorBins equal-width score bins
// spanning [range[0][q], range[1][q]] = (best local floor, best local score) over the shards.
__global__ void __launch_bounds__(256) floor_hist_kernel(const unsigned long long* cand, const int* count, int C,
                                                         const float* __restrict__ range, int nq, int* hist) {
  __shared__ int sh[kFloorBins];
  const int q = blockIdx.x;
  for (int i = threadIdx.x; i < kFloorBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const FloorBins fb = FloorBins::make(range[q], range[nq + q]);
  const int cnt = count[q];
  const unsigned long long* mine = cand + static_cast<size_t>(q) * C;
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
    const int b = fb.bin(key_score(mine[j]));
    if (b >= 0) atomicAdd(&sh[b], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kFloorBins; i += blockDim.x) hist[static_cast<size_t>(q) * kFloorBins + i] = sh[i];
}

// per query: {kp-th (floor) and best bf16-stage score} of this shard's candidate list -> range[0][q], range[1][q]
__global__ void __launch_bounds__(256) local_range_kernel(const unsigned long long* cand, const int* count,
                                                          const float* thr, int C, int nq, int has_floor,
                                                          float* range) {
  __shared__ float smax[8];
  const int q = blockIdx.x;
  const int cnt = count[q];
  const unsigned long long* mine = cand + static_cast<size_t>(q) * C;
  float m = __int_as_float(0xff800000);
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) m = fmaxf(m, key_score(mine[j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, smax[i]);
    range[q] = has_floor ? thr[q] : __int_as_float(0xff800000);  // a floor needs k + slack local rows
    range[nq + q] = m;
  }
}

// fp32 [n, d] -> bf16 [n, dpad] (pad columns zeroed)
__global__ void f32_to_bf16_rows(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n, int d,
                                 int dpad) {
  const int64_t total = n * static_cast<int64_t>(dpad);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / dpad;
    const int c = static_cast<int>(i - r * dpad);
    dst[i] = __float2bfloat16(c < d ? src[r * d + c] : 0.f);
  }
}
template <typename T>
__global__ void to_f32(const T* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = static_cast<float>(src[i]);
}
__global__ void fill_i32(int* p, int v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// MERGE (sharded search exchange step): one CTA per query over nparts * k candidates.
__global__ void __launch_bounds__(256) merge_kernel(const float* Dp, const int64_t* Ip, int nparts, int nq, int k,
                                                    int k_out, float* D, int64_t* I) {
  // keys: orderable(score) << 32 | ~slot, with ties broken by id through a second pass on equal scores
  extern __shared__ unsigned long long msm[];
  const int q = blockIdx.x;
  const int total = nparts * k;
  int P = 2;
  while (P < total) P <<= 1;
  unsigned long long* skeys = msm;           // [P] (score, slot)
  int64_t* sid = reinterpret_cast<int64_t*>(msm + P);  // [total]
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    unsigned long long key = 0ull;
    if (i < total) {
      const int part = i / k, r = i - part * k;
      const size_t off = (static_cast<size_t>(part) * nq + q) * k + r;
      const int64_t id = Ip[off];
      sid[i] = id;
      if (id >= 0) key = (static_cast<unsigned long long>(f32_orderable(Dp[off])) << 32) | (0xffffffffu - i);
    }
    skeys[i] = key;
  }
  __syncthreads();
  bitonic_sort_desc(skeys, P, threadIdx.x, blockDim.x);
  // Shards hold disjoint, increasing id ranges and each shard list is already (score desc, id asc), so slot
  // order == id order among equal scores: the (score, slot) sort is the (score, id) sort.
  for (int r = threadIdx.x; r < k_out; r += blockDim.x) {
    const unsigned long long key = r < P ? skeys[r] : 0ull;
    float s = -FLT_MAX;
    int64_t id = -1;
    if (key != 0ull) {
      s = f32_from_orderable(static_cast<uint32_t>(key >> 32));
      id = sid[0xffffffffu - static_cast<uint32_t>(key & 0xffffffffull)];
    }
    D[static_cast<size_t>(q) * k_out + r] = s;
    I[static_cast<size_t>(q) * k_out + r] = id;
  }
}

}  // namespace om

using namespace om;

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct om_index {
  int d = 0, dpad = 0;
  int64_t n = 0, cap = 0;
  float* xf = nullptr;
  __nv_bfloat16* xb = nullptr;
  int64_t rescore_slack = -1;
  int force_safe = 0;
  int dynamic_sched = 1;  // claim scan tiles from a global counter (keeps CTAs on neighbouring corpus tiles)
  int growth = 2;  // each round scans (growth - 1) x the rows seen so far (2 measured best on B200)
  int64_t st_rounds = 0, st_retries = 0, st_capacity = 0, st_launches = 0;
  // optional per-phase device timing (CUDA events on the launching stream), enabled by set_param("profile", 1)
  int profile = 0;
  double st_scan_us = 0, st_select_us = 0, st_final_us = 0;
  std::vector<cudaEvent_t> ev;  // pool: [2i] start, [2i+1] stop
  std::vector<int> ev_kind;     // 0 scan, 1 select, 2 finalize
  size_t ev_used = 0;
  // workspace
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // state between om_index_search_begin and om_index_search_finish
  struct Plan {
    bool valid = false;
    int nq = 0, k = 0, kp = 0, kp_target = 0, C = 0, growth = 2;  // kp = min(kp_target = k + slack, rows)
    size_t o_qf = 0, o_qb = 0, o_cand = 0, o_count = 0, o_thr = 0, o_ovf = 0, o_D = 0, o_I = 0;
  } plan;
};

static int index_grow(om_index* ix, int64_t need) {
  if (need <= ix->cap) return 0;
  int64_t ncap = std::max<int64_t>(need, ix->cap + ix->cap / 2);
  ncap = round_up(std::max<int64_t>(ncap, 1024), 256);
  float* nxf = nullptr;
  __nv_bfloat16* nxb = nullptr;
  OM_CUDA(cudaMalloc(&nxf, static_cast<size_t>(ncap) * ix->d * sizeof(float)));
  cudaError_t e = cudaMalloc(&nxb, static_cast<size_t>(ncap) * ix->dpad * sizeof(__nv_bfloat16));
  if (e != cudaSuccess) {
    cudaFree(nxf);
    cudaGetLastError();
    return fail(OM_ENOMEM, "index: cannot allocate bf16 scan copy for %lld rows", (long long)ncap);
  }
  if (ix->n > 0) {
    OM_CUDA(cudaMemcpy(nxf, ix->xf, static_cast<size_t>(ix->n) * ix->d * sizeof(float), cudaMemcpyDeviceToDevice));
    OM_CUDA(cudaMemcpy(nxb, ix->xb, static_cast<size_t>(ix->n) * ix->dpad * sizeof(__nv_bfloat16),
                       cudaMemcpyDeviceToDevice));
  }
  cudaFree(ix->xf);
  cudaFree(ix->xb);
  ix->xf = nxf;
  ix->xb = nxb;
  ix->cap = ncap;
  return 0;
}

static int ws_reserve(om_index* ix, size_t bytes) {
  if (bytes <= ix->ws_bytes) return 0;
  if (ix->ws) cudaFree(ix->ws);
  ix->ws = nullptr;
  ix->ws_bytes = 0;
  OM_CUDA(cudaMalloc(&ix->ws, bytes));
  ix->ws_bytes = bytes;
  return 0;
}

static inline int grid_for(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  return static_cast<int>(std::min<int64_t>(std::max<int64_t>(g, 1), 148 * 16));
}

extern "C" {

int om_index_create(int d, om_index** out) {
  if (!out || d <= 0) return fail(OM_EINVAL, "om_index_create: d must be positive");
  OM_TRY(device_sm_count());
  om_index* ix = new (std::nothrow) om_index();
  if (!ix) return fail(OM_ENOMEM, "om_index_create: out of host memory");
  ix->d = d;
  ix->dpad = static_cast<int>(round_up(d, 8));  // 16-byte row pitch for TMA
  *out = ix;
  return 0;
}

void om_index_destroy(om_index* ix) {
  if (!ix) return;
  cudaFree(ix->xf);
  cudaFree(ix->xb);
  cudaFree(ix->ws);
  for (cudaEvent_t e : ix->ev) cudaEventDestroy(e);
  delete ix;
}

int64_t om_index_ntotal(const om_index* ix) { return ix ? ix->n : 0; }
int om_index_dim(const om_index* ix) { return ix ? ix->d : 0; }

int om_index_reset(om_index* ix) {
  if (!ix) return fail(OM_EINVAL, "om_index_reset: null index");
  ix->n = 0;
  ix->plan.valid = false;  // a search begun on the old contents cannot be finished
  return 0;
}

int om_index_reserve(om_index* ix, int64_t n, float** dev_rows) {
  if (!ix || n < 0 || !dev_rows) return fail(OM_EINVAL, "om_index_reserve: bad arguments");
  if (ix->n + n > 0xfffffff0ll) return fail(OM_EINVAL, "index shard limited to 2^32-16 rows; shard the corpus");
  OM_TRY(index_grow(ix, ix->n + n));
  *dev_rows = ix->xf + static_cast<size_t>(ix->n) * ix->d;
  return 0;
}

int om_index_commit(om_index* ix, int64_t n, void* stream) {
  if (!ix || n < 0 || ix->n + n > ix->cap) return fail(OM_EINVAL, "om_index_commit: more rows than reserved");
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  f32_to_bf16_rows<<<grid_for(n * ix->dpad, 256), 256, 0, st>>>(ix->xf + static_cast<size_t>(ix->n) * ix->d,
                                                               ix->xb + static_cast<size_t>(ix->n) * ix->dpad, n,
                                                               ix->d, ix->dpad);
  OM_CUDA(cudaGetLastError());
  ix->n += n;
  return 0;
}

int om_index_add(om_index* ix, const void* x, om_memkind kind, om_dtype dtype, int64_t n, void* stream) {
  if (!ix || (!x && n > 0) || n < 0) return fail(OM_EINVAL, "om_index_add: bad arguments");
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* dst = nullptr;
  OM_TRY(om_index_reserve(ix, n, &dst));
  const size_t elems = static_cast<size_t>(n) * ix->d;
  if (dtype == OM_F32) {
    OM_CUDA(cudaMemcpyAsync(dst, x, elems * 4, kind == OM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                            st));
  } else if (dtype == OM_BF16 || dtype == OM_F16) {
    const void* src = x;
    void* tmp = nullptr;
    if (kind == OM_HOST) {
      OM_CUDA(cudaMalloc(&tmp, elems * 2));
      OM_CUDA(cudaMemcpyAsync(tmp, x, elems * 2, cudaMemcpyHostToDevice, st));
      src = tmp;
    }
    if (dtype == OM_BF16)
      to_f32<<<grid_for(elems, 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), dst, (int64_t)elems);
    else
      to_f32<<<grid_for(elems, 256), 256, 0, st>>>(static_cast<const __half*>(src), dst, (int64_t)elems);
    OM_CUDA(cudaGetLastError());
    if (tmp) {
      OM_CUDA(cudaStreamSynchronize(st));
      cudaFree(tmp);
    }
  } else {
    return fail(OM_EINVAL, "om_index_add: unsupported dtype %d", (int)dtype);
  }
  OM_TRY(om_index_commit(ix, n, stream));
  if (kind == OM_HOST) OM_CUDA(cudaStreamSynchronize(st));  // caller may free / reuse its host buffer
  return 0;
}

int om_index_set_param(om_index* ix, const char* name, int64_t value) {
  if (!ix || !name) return fail(OM_EINVAL, "om_index_set_param: bad arguments");
  if (!strcmp(name, "rescore_slack")) {
    ix->rescore_slack = value;
  } else if (!strcmp(name, "force_safe_rounds")) {
    ix->force_safe = value != 0;
  } else if (!strcmp(name, "round_growth")) {
    if (value < 2 || value > 8) return fail(OM_EINVAL, "round_growth must be in [2, 8]");
    ix->growth = static_cast<int>(value);
  } else if (!strcmp(name, "dynamic_sched")) {
    ix->dynamic_sched = value != 0;
  } else if (!strcmp(name, "profile")) {
    ix->profile = static_cast<int>(value);
  } else {
    return fail(OM_EINVAL, "om_index_set_param: unknown parameter '%s'", name);
  }
  return 0;
}

int64_t om_index_get_stat(const om_index* ix, const char* name) {
  if (!ix || !name) return -1;
  if (!strcmp(name, "rounds")) return ix->st_rounds;
  if (!strcmp(name, "overflow_retries")) return ix->st_retries;
  if (!strcmp(name, "candidates")) return ix->st_capacity;
  if (!strcmp(name, "launches")) return ix->st_launches;
  if (!strcmp(name, "scan_ns")) return static_cast<int64_t>(ix->st_scan_us * 1e3);
  if (!strcmp(name, "select_ns")) return static_cast<int64_t>(ix->st_select_us * 1e3);
  if (!strcmp(name, "finalize_ns")) return static_cast<int64_t>(ix->st_final_us * 1e3);
  return -1;
}

}  // extern "C"

namespace {

// profiling helpers: bracket a launch with events from the index's pool
struct Timed {
  om_index* ix;
  cudaStream_t st;
  bool on;
  Timed(om_index* ix_, cudaStream_t st_, int kind) : ix(ix_), st(st_), on(ix_->profile != 0) {
    if (!on) return;
    if (ix->ev_used + 2 > ix->ev.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
        on = false;
        return;
      }
      ix->ev.push_back(a);
      ix->ev.push_back(b);
      ix->ev_kind.push_back(0);
    }
    ix->ev_kind[ix->ev_used / 2] = kind;
    cudaEventRecord(ix->ev[ix->ev_used], st);
  }
  ~Timed() {
    if (!on) return;
    cudaEventRecord(ix->ev[ix->ev_used + 1], st);
    ix->ev_used += 2;
  }
};

void collect_profile(om_index* ix) {
  for (size_t i = 0; i + 1 < ix->ev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ix->ev[i], ix->ev[i + 1]) != cudaSuccess) continue;
    const int kind = ix->ev_kind[i / 2];
    (kind == 0 ? ix->st_scan_us : kind == 1 ? ix->st_select_us : ix->st_final_us) += ms * 1e3;
    if (ix->profile >= 2) fprintf(stderr, "[om profile] launch %zu %s %.3f ms\n", i / 2, kind == 0 ? "scan" : kind == 1 ? "select" : "finalize", ms);
  }
  ix->ev_used = 0;
}

struct ChunkWs {
  unsigned long long* cand;
  int* count;
  float* thr;
  int* overflow;
};

// One sweep of the corpus for a chunk of queries.  safe=false: doubling rounds; safe=true: fixed rounds of
// C - kp rows, which cannot overflow.
int sweep(om_index* ix, const __nv_bfloat16* qb, int nq, int kp, int C, int growth_, const ChunkWs& w, bool safe,
          int sms, cudaStream_t st) {
  const int64_t growth = growth_;
  const int64_t N = ix->n;
  OM_CUDA(cudaMemsetAsync(w.overflow, 0, sizeof(int), st));
  int64_t pos = 0;
  const size_t sel_smem = static_cast<size_t>(C) * 8;
  static bool sel_attr = false;
  if (!sel_attr) {
    OM_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    sel_attr = true;
  }
  bool first = true;
  while (pos < N) {
    int64_t step;
    if (first)
      step = std::min<int64_t>(N, C);
    else if (safe)
      step = std::min<int64_t>(N - pos, std::max<int64_t>(256, ((C - kp) / 256) * 256));
    else
      step = std::min<int64_t>(N - pos, (growth - 1) * pos);
    {
      Timed t(ix, st, 0);
      const __nv_bfloat16* xrows = ix->xb + static_cast<size_t>(pos) * ix->dpad;
      cudaError_t e;
      if (first) {
        EpiScan<true> epi{w.thr, w.cand, w.count, w.overflow, nq, static_cast<int>(step), C, static_cast<uint32_t>(pos)};
        e = launch_gemm<256, 4, true, 8>(qb, ix->dpad, xrows, ix->dpad, nq, static_cast<int>(step), ix->d, epi, sms, st,
                                         ix->dynamic_sched != 0);
      } else {
        EpiScan<false, 256, OM_SCAN_VARIANT> epi{w.thr, w.cand, w.count, w.overflow, nq, static_cast<int>(step), C,
                                                 static_cast<uint32_t>(pos)};
        e = launch_gemm<256, 4, true, 8>(qb, ix->dpad, xrows, ix->dpad, nq, static_cast<int>(step), ix->d, epi, sms, st,
                                         ix->dynamic_sched != 0);
      }
      if (e != cudaSuccess) return fail(OM_ECUDA, "scan kernel launch failed: %s", cudaGetErrorString(e));
    }
    {
      Timed t(ix, st, 1);
      select_kernel<<<nq, 256, sel_smem, st>>>(w.cand, w.count, w.thr, C, kp, first ? static_cast<int>(step) : -1);
    }
    OM_CUDA(cudaGetLastError());
    ix->st_launches += 2;
    pos += step;
    first = false;
    ix->st_rounds++;
  }
  return 0;
}

}  // namespace

namespace {

constexpr int kQueryChunk = 16384;

ChunkWs plan_ws(om_index* ix) {
  uint8_t* base = static_cast<uint8_t*>(ix->ws);
  const om_index::Plan& p = ix->plan;
  return ChunkWs{reinterpret_cast<unsigned long long*>(base + p.o_cand), reinterpret_cast<int*>(base + p.o_count),
                 reinterpret_cast<float*>(base + p.o_thr), reinterpret_cast<int*>(base + p.o_ovf)};
}

// sizes the candidate lists, carves the workspace and uploads / converts the queries
int search_prepare(om_index* ix, const void* q, om_memkind q_kind, int nq, int k, bool host_out, cudaStream_t st) {
  const int d = ix->d, dpad = ix->dpad;
  if (d > 16384) return fail(OM_EINVAL, "om_index_search: d > 16384 unsupported");
  int64_t slack = ix->rescore_slack >= 0 ? ix->rescore_slack : std::max<int64_t>(64, k / 8);
  const int64_t kp64 = std::min<int64_t>(static_cast<int64_t>(k) + slack, std::max<int64_t>(ix->n, 1));
  if (kp64 > 4096) return fail(OM_EINVAL, "om_index_search: k + slack = %lld exceeds 4096", (long long)kp64);
  om_index::Plan& p = ix->plan;
  p.valid = false;
  p.nq = nq;
  p.k = k;
  p.kp = static_cast<int>(kp64);
  p.kp_target = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(k) + slack, 4096));
  // Expected list length after a round that multiplies the rows seen by g is ~g kp (kp kept + ~(g-1) kp new
  // survivors); 25 % + 512 entries of head-room cover its spread for exchangeable row order.
  for (p.growth = ix->growth;; --p.growth) {  // large k: slower-growing schedule that fits the 16384-entry select
    p.C = 1024;
    while (p.C < (5 * p.growth * p.kp) / 4 + 512) p.C <<= 1;
    if (p.C <= 16384 || p.growth == 2) break;
  }
  if (p.C > 16384) return fail(OM_EINVAL, "om_index_search: candidate list of %d entries exceeds 16384; lower k", p.C);
  ix->st_capacity = p.C;
  ix->st_rounds = 0;
  ix->st_retries = 0;
  ix->st_launches = 1;  // the query fp32 -> bf16 conversion below
  ix->st_scan_us = ix->st_select_us = ix->st_final_us = 0;
  ix->ev_used = 0;
  const int nqc_max = std::min(nq, kQueryChunk);
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    size_t o = off;
    off += round_up(bytes, 256);
    return o;
  };
  p.o_qf = carve(static_cast<size_t>(nq) * d * 4);
  p.o_qb = carve(static_cast<size_t>(nq) * dpad * 2);
  p.o_cand = carve(static_cast<size_t>(nqc_max) * p.C * 8);
  p.o_count = carve(static_cast<size_t>(nqc_max) * 4);
  p.o_thr = carve(static_cast<size_t>(nqc_max) * 4);
  p.o_ovf = carve(256);
  p.o_D = carve(host_out ? static_cast<size_t>(nq) * k * 4 : 0);
  p.o_I = carve(host_out ? static_cast<size_t>(nq) * k * 8 : 0);
  OM_TRY(ws_reserve(ix, off));
  uint8_t* base = static_cast<uint8_t*>(ix->ws);
  float* qf = reinterpret_cast<float*>(base + p.o_qf);
  __nv_bfloat16* qb = reinterpret_cast<__nv_bfloat16*>(base + p.o_qb);
  OM_CUDA(cudaMemcpyAsync(qf, q, static_cast<size_t>(nq) * d * 4,
                          q_kind == OM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
  f32_to_bf16_rows<<<grid_for(static_cast<int64_t>(nq) * dpad, 256), 256, 0, st>>>(qf, qb, nq, d, dpad);
  OM_CUDA(cudaGetLastError());
  static bool fin_attr = false;
  if (!fin_attr) {
    OM_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 8 + 65536));
    fin_attr = true;
  }
  return 0;
}

// bf16 scan sweep of the shard for queries [q0, q0 + nqc), with the overflow check and the safe-schedule retry
int search_sweep_checked(om_index* ix, int q0, int nqc, int sms, cudaStream_t st) {
  const om_index::Plan& p = ix->plan;
  const ChunkWs w = plan_ws(ix);
  const __nv_bfloat16* qb = reinterpret_cast<const __nv_bfloat16*>(static_cast<uint8_t*>(ix->ws) + p.o_qb);
  if (ix->n == 0) {
    fill_i32<<<(nqc + 255) / 256, 256, 0, st>>>(w.count, 0, nqc);
    OM_CUDA(cudaGetLastError());
    return 0;
  }
  bool safe = ix->force_safe != 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    OM_TRY(sweep(ix, qb + static_cast<size_t>(q0) * ix->dpad, nqc, p.kp, p.C, p.growth, w, safe, sms, st));
    int ovf = 0;
    OM_CUDA(cudaMemcpyAsync(&ovf, w.overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    OM_CUDA(cudaStreamSynchronize(st));
    const unsigned int fault = read_clear_dev_fault();
    if (fault) return fail(OM_EFAULT, "scan kernel pipeline fault 0x%08x", fault);
    if (!ovf) return 0;
    if (safe) return fail(OM_EFAULT, "candidate list overflow in the overflow-proof schedule (bug)");
    safe = true;
    ix->st_retries++;
  }
  return 0;
}

int search_finalize(om_index* ix, int q0, int nqc, float* dD, int64_t* dI, int64_t id_offset, const float* range,
                    const int* ghist, int* kept_max, cudaStream_t st) {
  const om_index::Plan& p = ix->plan;
  const ChunkWs w = plan_ws(ix);
  const float* qf = reinterpret_cast<const float*>(static_cast<uint8_t*>(ix->ws) + p.o_qf);
  int P2 = 2;
  while (P2 < p.kp) P2 <<= 1;
  const size_t fin_smem = static_cast<size_t>(P2) * 8 + static_cast<size_t>(ix->d) * 4;
  {
    Timed t(ix, st, 2);
    finalize_kernel<<<nqc, 256, fin_smem, st>>>(w.cand, w.count, p.C, qf + static_cast<size_t>(q0) * ix->d, ix->xf, ix->d,
                                                p.k, dD + static_cast<size_t>(q0) * p.k, dI + static_cast<size_t>(q0) * p.k,
                                                id_offset, range, ghist, p.nq, p.kp_target, kept_max);
  }
  OM_CUDA(cudaGetLastError());
  ix->st_launches += 1;
  return 0;
}

int search_emit(om_index* ix, float* D, int64_t* I, float* dD, int64_t* dI, om_memkind out_kind, cudaStream_t st) {
  const om_index::Plan& p = ix->plan;
  if (out_kind == OM_HOST) {
    OM_CUDA(cudaMemcpyAsync(D, dD, static_cast<size_t>(p.nq) * p.k * 4, cudaMemcpyDeviceToHost, st));
    OM_CUDA(cudaMemcpyAsync(I, dI, static_cast<size_t>(p.nq) * p.k * 8, cudaMemcpyDeviceToHost, st));
  }
  OM_CUDA(cudaStreamSynchronize(st));
  if (ix->profile) collect_profile(ix);
  return 0;
}

}  // namespace

extern "C" int om_index_search(om_index* ix, const void* q, om_memkind q_kind, int nq, int k, float* D, int64_t* I,
                               om_memkind out_kind, int64_t id_offset, void* stream) {
  if (!ix || (nq > 0 && (!q || !D || !I)) || nq < 0 || k <= 0)
    return fail(OM_EINVAL, "om_index_search: bad arguments (nq=%d k=%d)", nq, k);
  if (nq == 0) return 0;
  const int sms = device_sm_count();
  if (sms < 0) return sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  OM_TRY(search_prepare(ix, q, q_kind, nq, k, out_kind == OM_HOST, st));
  uint8_t* base = static_cast<uint8_t*>(ix->ws);
  float* dD = out_kind == OM_HOST ? reinterpret_cast<float*>(base + ix->plan.o_D) : D;
  int64_t* dI = out_kind == OM_HOST ? reinterpret_cast<int64_t*>(base + ix->plan.o_I) : I;
  for (int q0 = 0; q0 < nq; q0 += kQueryChunk) {
    const int nqc = std::min(kQueryChunk, nq - q0);
    OM_TRY(search_sweep_checked(ix, q0, nqc, sms, st));
    OM_TRY(search_finalize(ix, q0, nqc, dD, dI, id_offset, nullptr, nullptr, nullptr, st));
  }
  return search_emit(ix, D, I, dD, dI, out_kind, st);
}

// Three-phase search for row-sharded indexes (one shard per process); see include/openmatch_b200.h.
extern "C" int om_index_search_begin(om_index* ix, const void* q, om_memkind q_kind, int nq, int k, float* local_range,
                                     void* stream) {
  if (!ix || nq <= 0 || !q || !local_range || k <= 0) return fail(OM_EINVAL, "om_index_search_begin: bad arguments");
  if (nq > kQueryChunk) return fail(OM_EINVAL, "om_index_search_begin: at most %d queries per call", kQueryChunk);
  const int sms = device_sm_count();
  if (sms < 0) return sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  OM_TRY(search_prepare(ix, q, q_kind, nq, k, false, st));
  OM_TRY(search_sweep_checked(ix, 0, nq, sms, st));
  const ChunkWs w = plan_ws(ix);
  if (ix->n == 0) {
    fill_i32<<<(nq + 255) / 256, 256, 0, st>>>(reinterpret_cast<int*>(w.thr), static_cast<int>(0xff800000), nq);
    OM_CUDA(cudaGetLastError());
  }
  local_range_kernel<<<nq, 256, 0, st>>>(w.cand, w.count, w.thr, ix->plan.C, nq, ix->n >= ix->plan.kp_target ? 1 : 0,
                                         local_range);
  OM_CUDA(cudaGetLastError());
  ix->st_launches += 1;
  ix->plan.valid = true;
  return 0;
}

extern "C" int om_index_search_count(om_index* ix, const float* global_range, int* local_hist, void* stream) {
  if (!ix || !global_range || !local_hist) return fail(OM_EINVAL, "om_index_search_count: bad arguments");
  if (!ix->plan.valid) return fail(OM_ESTATE, "om_index_search_count: no search in progress (call om_index_search_begin)");
  const ChunkWs w = plan_ws(ix);
  floor_hist_kernel<<<ix->plan.nq, 256, 0, static_cast<cudaStream_t>(stream)>>>(w.cand, w.count, ix->plan.C, global_range,
                                                                             ix->plan.nq, local_hist);
  OM_CUDA(cudaGetLastError());
  ix->st_launches += 1;
  return 0;
}

extern "C" int om_index_search_finish(om_index* ix, const float* global_range, const int* global_hist, float* D,
                                      int64_t* I, int64_t id_offset, int* kept_max, void* stream) {
  if (!ix || !D || !I || (global_hist && !global_range)) return fail(OM_EINVAL, "om_index_search_finish: bad arguments");
  if (!ix->plan.valid) return fail(OM_ESTATE, "om_index_search_finish: no search in progress (call om_index_search_begin)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ix->plan.valid = false;
  if (kept_max) OM_CUDA(cudaMemsetAsync(kept_max, 0, sizeof(int), st));
  OM_TRY(search_finalize(ix, 0, ix->plan.nq, D, I, id_offset, global_range, global_hist, kept_max, st));
  return search_emit(ix, D, I, D, I, OM_DEVICE, st);
}

extern "C" int om_search_floor_bins(void) { return kFloorBins; }

extern "C" int om_topk_merge_n(const float* D_parts, const int64_t* I_parts, int nparts, int nq, int k_in, int k_out,
                               float* D, int64_t* I, void* stream) {
  if (nparts <= 0 || nq < 0 || k_in <= 0 || k_out <= 0 || !D_parts || !I_parts || !D || !I)
    return fail(OM_EINVAL, "om_topk_merge: bad arguments");
  if (nq == 0) return 0;
  OM_TRY(device_sm_count());
  const int64_t total = static_cast<int64_t>(nparts) * k_in;
  if (total > 8192) return fail(OM_EINVAL, "om_topk_merge: nparts * k = %lld exceeds 8192", (long long)total);
  int P = 2;
  while (P < total) P <<= 1;
  const size_t smem = static_cast<size_t>(P) * 8 + static_cast<size_t>(total) * 8;
  static bool attr = false;
  if (!attr) {
    OM_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16));
    attr = true;
  }
  merge_kernel<<<nq, 256, smem, static_cast<cudaStream_t>(stream)>>>(D_parts, I_parts, nparts, nq, k_in, k_out, D, I);
  OM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int om_topk_merge(const float* D_parts, const int64_t* I_parts, int nparts, int nq, int k, float* D,
                             int64_t* I, void* stream) {
  return om_topk_merge_n(D_parts, I_parts, nparts, nq, k, k, D, I, stream);
}
