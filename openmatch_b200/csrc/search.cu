// Exact inner-product top-k over an HBM-resident flat index — the B200 replacement of
// faiss.IndexFlatIP.{add,search,reset} as the reference uses it
// (src/openmatch/retriever/dense_retriever.py:38-41,105,133-137,180) and of the IndexShards merge behind
// index_cpu_to_gpu_multiple(shard=True) (:43-58).
//
// Storage per index shard: fp32 master rows [n, d] (what index.add received; used for exact re-scoring)
// plus an fp16 scan copy [n, dpad] (tensor-core operand; IEEE half keeps 3 more significand bits than bf16 at the
// same tcgen05 rate, which is what makes the exactness certificate below affordable).
//
// search(q, k):
//   1. SCAN    fp16 Q * X^T on tcgen05 (gemm.cuh mainloop) with the top-k filter fused into the epilogue:
//              scores never leave TMEM/registers; a thread owns one query row, compares its 32-column chunk
//              against that query's running threshold and appends the rare survivors
//              (key = orderable(score) << 32 | ~row) to the query's candidate list in HBM.
//              The corpus is swept in rounds of geometrically growing size; after each round
//   2. SELECT  a per-query radix select keeps the best kp = k + slack candidates and publishes the kp-th score
//              as the next round's (strict) threshold.  An overflowing list (adversarially sorted corpus) is
//              detected and the level is redone with an overflow-proof fixed-size round schedule.
//   3. FINAL   the kp candidates are re-scored against the fp32 master rows (fp32 FMA, fixed summation order),
//              sorted by (score desc, row asc) and the top k emitted as (D fp32, I int64) — faiss's output contract.
//   4. CERTIFY per query, a rigorous a-posteriori bound E(q) on |stage score - fp32 score| over ALL rows
//              (measured quantisation-error norms of the corpus and of the query + an fp32 accumulation term)
//              proves that no row outside the candidate list can reach the k-th fp32 score:
//                  s_k(fp32) - tau(stage, kp-th) > E(q).
//              Queries that fail it are re-run with the widest candidate list (k + slack = 4096) and, if still
//              uncertified (e.g. thousands of near-duplicate rows that collide in half precision), by an exact
//              fp32 scan on the CUDA cores that uses the same summation order as FINAL.  The result is therefore
//              always the exact top-k by fp32 inner product, ties by row id — never "top-k up to fp16 noise".
#include <cuda_fp16.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "common.h"
#include "gemm.cuh"
#include "gemm2sm.cuh"
#include "nccl_dyn.h"
#include "scan_epilogue.cuh"

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

// Lowest bin b with count(bins >= b) >= kp (0 when the lists hold fewer than kp rows: no pruning), computed by one
// warp from a query's reduced histogram: lane l owns bins [l * per, (l + 1) * per); suffix sums over the lanes.
__device__ __forceinline__ int warp_min_bin(const int* __restrict__ hrow, int kp, int lane) {
  constexpr int per = kFloorBins / 32;
  int h[per], sum = 0;
#pragma unroll
  for (int j = 0; j < per; ++j) {
    h[j] = hrow[lane * per + j];
    sum += h[j];
  }
  int suffix = sum;  // sum over lanes >= this one
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_down_sync(0xffffffffu, suffix, o);
    if (lane + o < 32) suffix += v;
  }
  int run = suffix - sum;  // count in the bins above this lane's
  int mine = 0;
  const bool owner = run < kp && kp <= suffix;  // at most one lane
  if (owner) {
    int b = per - 1;
#pragma unroll
    for (int j = per - 1; j >= 0; --j) {
      if (run < kp) b = j;
      run += h[j];
    }
    mine = lane * per + b;
  }
  const unsigned who = __ballot_sync(0xffffffffu, owner);
  return who ? __shfl_sync(0xffffffffu, mine, __ffs(who) - 1) : 0;
}

// FINAL: one CTA per query: exact fp32 re-score of the surviving candidates against the master rows,
// sort by (score desc, row asc), emit top-k.
__global__ void __launch_bounds__(256) finalize_kernel(const unsigned long long* cand, const int* count, int C,
                                                       const float* __restrict__ qf, const float* __restrict__ xf,
                                                       int d, int k, float* D, int64_t* I, int64_t id_offset,
                                                       const float* __restrict__ range, const int* __restrict__ ghist,
                                                       int nq_total, int kp, int* kept_max, int k_out, int* exceed,
                                                       int stage_scores) {
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
    const int mb = warp_min_bin(ghist + static_cast<size_t>(q) * kFloorBins, kp, lane);
    if (lane == 0) s_minbin = mb;
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
    if (stage_scores) acc = key_score(mine[j]);  // debug: report the candidate-stage score instead
    if (lane == 0) skeys[atomicAdd(&s_n, 1)] = make_key(acc, row);  // compacted: the sort covers survivors only
  }
  __syncthreads();
  const int n = s_n;
  int P2 = 2;
  while (P2 < n) P2 <<= 1;
  for (int i = n + threadIdx.x; i < P2; i += blockDim.x) skeys[i] = 0ull;
  __syncthreads();
  bitonic_sort_desc(skeys, P2, threadIdx.x, blockDim.x);
  // k_out <= k entries are written per query (row pitch k_out): the sharded exchange ships a fixed-width prefix
  for (int r = threadIdx.x; r < k_out; r += blockDim.x) {
    float s = -FLT_MAX;
    int64_t id = -1;
    if (r < n) {
      s = key_score(skeys[r]);
      id = id_offset + static_cast<int64_t>(key_row(skeys[r]));
    }
    D[static_cast<size_t>(q) * k_out + r] = s;
    I[static_cast<size_t>(q) * k_out + r] = id;
  }
  // longest valid prefix over the queries: lets the caller exchange [nq, kept] instead of [nq, k]
  if (kept_max && threadIdx.x == 0 && n > 0) atomicMax(kept_max, n < k ? n : k);
  // a list longer than the shipped prefix: the caller must redo the exchange at full width
  if (exceed && threadIdx.x == 0 && (n < k ? n : k) > k_out) *exceed = 1;
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
                                                          float* range, const float* gstats) {
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
    if (q == 0 && gstats) {  // MAX-reduced together with the ranges: every shard certifies against the global maxima
      range[2 * nq] = gstats[0];
      range[2 * nq + 1] = gstats[1];
    }
  }
}

// fp32 [n, d] -> fp16 [n, dpad] (pad columns zeroed; finite values beyond the half range saturate — the measured
// error norm below then makes the certificate fail and the exact path takes over), one warp per row.
// Optional outputs: per-row ||x_h|| and ||x - x_h|| (queries) and running maxima of ||x|| and ||x - x_h|| over all
// rows ever committed (index): the inputs of the exactness certificate.  Non-negative floats order like ints, so
// the maxima are kept with atomicMax on the bit patterns.
__global__ void __launch_bounds__(256) rows_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst,
                                                          int64_t n, int d, int dpad, float* __restrict__ hnorm,
                                                          float* __restrict__ enorm, float* gstats) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  float mx = 0.f, me = 0.f;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += nwarps) {
    const float* x = src + r * d;
    __half* y = dst + r * dpad;
    float sx = 0.f, sh = 0.f, se = 0.f;
    for (int c = 2 * lane; c < dpad; c += 64) {  // dpad is a multiple of 8
      const float a = c < d ? x[c] : 0.f, b = c + 1 < d ? x[c + 1] : 0.f;
      const float ac = fminf(fmaxf(a, -65504.f), 65504.f), bc = fminf(fmaxf(b, -65504.f), 65504.f);
      const __half2 h = __floats2half2_rn(ac, bc);  // NaN stays NaN (fmin/fmax return the other operand: guard below)
      *reinterpret_cast<__half2*>(y + c) = h;
      const float ha = __low2float(h), hb = __high2float(h);
      const float ea = a - ha, eb = b - hb;
      sx = fmaf(a, a, fmaf(b, b, sx));
      sh = fmaf(ha, ha, fmaf(hb, hb, sh));
      se = fmaf(ea, ea, fmaf(eb, eb, se));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sh += __shfl_xor_sync(0xffffffffu, sh, o);
      se += __shfl_xor_sync(0xffffffffu, se, o);
    }
    // NaN / inf inputs poison the norms on purpose: the certificate then never holds and the exact scan answers
    const float nx = sqrtf(sx), nh = sqrtf(sh), ne = sqrtf(se);
    if (lane == 0) {
      if (hnorm) hnorm[r] = nh;
      if (enorm) enorm[r] = ne;
    }
    mx = (nx > mx || nx != nx) ? nx : mx;
    me = (ne > me || ne != ne) ? ne : me;
  }
  if (gstats && lane == 0) {
    atomicMax(reinterpret_cast<int*>(gstats), __float_as_int(mx != mx ? __int_as_float(0x7fc00000) : mx));
    atomicMax(reinterpret_cast<int*>(gstats) + 1, __float_as_int(me != me ? __int_as_float(0x7fc00000) : me));
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

// MERGE (sharded search exchange step): one CTA per query over nparts * k candidates.  Part p's lists start at
// Dp + p * stride_d / Ip + p * stride_i (elements), [nq, k] row-major each.
__global__ void __launch_bounds__(256) merge_kernel(const float* Dp, const int64_t* Ip, int64_t stride_d,
                                                    int64_t stride_i, int nparts, int nq, int k, int k_out, float* D,
                                                    int64_t* I) {
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
      const size_t off = static_cast<size_t>(q) * k + r;
      const int64_t id = Ip[part * stride_i + off];
      sid[i] = id;
      if (id >= 0) key = (static_cast<unsigned long long>(f32_orderable(Dp[part * stride_d + off])) << 32) | (0xffffffffu - i);
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

// CERTIFY: one warp per query.  Every row outside the re-scored candidate set has a stage score <= tau, and
// |stage - fp32| <= E(q) for every row of the corpus, so s_k - tau > E(q) proves that the emitted top-k is the
// exact fp32 top-k.  With x_h / q_h the half-precision operands, B the exact product sum of the rounded operands:
//   |fp32 - exact|   <= 30 * 2^-24 * |q||x|                      (FINAL's 24-FMA chain + 5-level tree at d = 768)
//   |exact - B|      <= ||q_h|| * ||x - x_h|| + ||q - q_h|| * ||x||          (Cauchy-Schwarz on the two error terms)
//   |B - stage|      <= d * 2^-22 * ||q_h|| * ||x_h||            (fp32 accumulation of exact products in the tensor
//                       core: d adds, each off by at most 2^-23 of the running magnitude if the hardware truncates
//                       instead of rounding; x2 head-room.  tests/test_search_gpu.py measures the real value.)
// with the corpus norms replaced by their maxima over the index (kept by rows_to_f16_kernel; sharded search: over all
// shards).  tau is the kp-th stage score of the candidate list — sharded search: the largest of the shards' floors.
// A NaN anywhere makes the comparison false: the query is flagged and answered by the exact path.
__global__ void __launch_bounds__(256) certify_kernel(const float* __restrict__ D, const int64_t* __restrict__ I, int k,
                                                      const float* __restrict__ floors, int64_t floor_stride, int nparts,
                                                      const float* __restrict__ gstats, int64_t stats_stride,
                                                      const float* __restrict__ hn, const float* __restrict__ en, int d,
                                                      int nq, int q_base, int* flags, int* nflag) {
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= nq) return;
  // tau: no row outside the re-scored candidate lists has a stage score above the largest per-shard floor (kp-th stage
  // score of the shard's list; -inf when the shard's list holds every row of the shard); error-norm maxima likewise
  float tau = __int_as_float(0xff800000), xmax = 0.f, exmax = 0.f;
  for (int p = 0; p < nparts; ++p) {
    tau = fmaxf(tau, floors[p * floor_stride + q]);
    const float a = gstats[p * stats_stride], b = gstats[p * stats_stride + 1];
    xmax = (a > xmax || a != a) ? a : xmax;  // NaN sticks: the comparison below then fails
    exmax = (b > exmax || b != b) ? b : exmax;
  }
  const float a = hn[q], b = en[q];
  const float E = 1.001f * (a * exmax + b * xmax + static_cast<float>(d + 16) * 2.384185791015625e-07f * (a + b) * (xmax + exmax));
  // tau = -inf: every list holds every row of its shard, nothing was left out.  Otherwise k results are needed to compare.
  const bool full = I[static_cast<size_t>(q) * k + (k - 1)] >= 0;
  const bool ok = !(tau > __int_as_float(0xff800000)) ? true : (full && D[static_cast<size_t>(q) * k + (k - 1)] - tau > E);
  // flags, not an appended list: the order of the uncertified queries must be the same on every rank of a sharded search
  if (lane == 0) {
    flags[q_base + q] = ok ? 0 : 1;
    if (!ok) atomicAdd(nflag, 1);
  }
}

// list[0 .. count) = ascending indices i with flags[i] != 0 (one block; chunked block-wide scan)
__global__ void __launch_bounds__(1024) compact_flags_kernel(const int* __restrict__ flags, int n, int* __restrict__ list) {
  __shared__ int warp_sums[32];
  __shared__ int base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    const bool f = i < n && flags[i] != 0;
    const unsigned b = __ballot_sync(0xffffffffu, f);
    if (lane == 0) warp_sums[warp] = __popc(b);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_sums[w];
    if (f) list[off + __popc(b & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += warp_sums[w];
      base += t;
    }
    __syncthreads();
  }
}

// EXACT fp32 scan on the CUDA cores (the certificate's last resort and the "exact_only" test mode): one warp per
// row group, queries of the tile in shared memory.  Per (query, row) the summation order is EXACTLY the one of
// finalize_kernel (lane-strided float4 FMA chain, then the xor-butterfly), so both produce bit-identical scores.
// Survivors (score > the query's strict threshold) are appended to the same candidate lists the tensor-core scan
// uses; dense = 1: first round, every score stored at position = column.
template <int NQT, int ROWS>
__global__ void __launch_bounds__(256) exact_scan_kernel(const float* __restrict__ xf, int64_t n_rows, uint32_t row_base,
                                                         const float* __restrict__ qf, int nq, int d, int nqt,
                                                         const float* __restrict__ thr, unsigned long long* cand,
                                                         int* count, int* overflow, int C, int dense) {
  extern __shared__ float sq[];
  const int q0 = blockIdx.y * nqt;
  const int nact = min(nqt, nq - q0);
  for (int i = threadIdx.x; i < nact * d; i += blockDim.x) sq[i] = qf[static_cast<size_t>(q0) * d + i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float t = __int_as_float(0x7f800000);
  if (lane < nact && !dense) t = thr[q0 + lane];
  const int64_t wg = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5), nw = static_cast<int64_t>(gridDim.x) * 8;
  for (int64_t r0 = wg * ROWS; r0 < n_rows; r0 += nw * ROWS) {
    float acc[ROWS][NQT];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
      for (int j = 0; j < NQT; ++j) acc[rr][j] = 0.f;
    if ((d & 3) == 0) {
      const int d4 = d >> 2;
      for (int i = lane; i < d4; i += 32) {
        float4 a[ROWS];
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr)
          a[rr] = r0 + rr < n_rows ? __ldg(reinterpret_cast<const float4*>(xf + static_cast<size_t>(r0 + rr) * d) + i)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NQT; ++j) {
          if (j < nact) {
            const float4 b = reinterpret_cast<const float4*>(sq + static_cast<size_t>(j) * d)[i];
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) {
              float s = acc[rr][j];
              s = fmaf(a[rr].x, b.x, s);
              s = fmaf(a[rr].y, b.y, s);
              s = fmaf(a[rr].z, b.z, s);
              s = fmaf(a[rr].w, b.w, s);
              acc[rr][j] = s;
            }
          }
        }
      }
    } else {
      for (int i = lane; i < d; i += 32) {
        float a[ROWS];
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr) a[rr] = r0 + rr < n_rows ? __ldg(xf + static_cast<size_t>(r0 + rr) * d + i) : 0.f;
#pragma unroll
        for (int j = 0; j < NQT; ++j)
          if (j < nact) {
            const float b = sq[static_cast<size_t>(j) * d + i];
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) acc[rr][j] = fmaf(a[rr], b, acc[rr][j]);
          }
      }
    }
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
      for (int j = 0; j < NQT; ++j)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[rr][j] += __shfl_xor_sync(0xffffffffu, acc[rr][j], o);
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
      float mine = acc[rr][0];  // lane j keeps query j's score (all lanes hold identical sums)
#pragma unroll
      for (int j = 1; j < NQT; ++j)
        if (lane == j) mine = acc[rr][j];
      const int64_t col = r0 + rr;
      if (lane < nact && col < n_rows) {
        const int q = q0 + lane;
        const unsigned long long key = make_key(mine, row_base + static_cast<uint32_t>(col));
        if (dense) {
          cand[static_cast<size_t>(q) * C + col] = key;
        } else if (mine > t) {
          const int pos = atomicAdd(count + q, 1);
          if (pos < C)
            cand[static_cast<size_t>(q) * C + pos] = key;
          else
            *overflow = 1;
        }
      }
    }
  }
}

// dst[i] = src[list[i]] (rows of d floats): the flagged queries of an escalation level
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ list, int n, int d,
                                   float* __restrict__ dst) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < static_cast<int64_t>(n) * d;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / d);
    dst[i] = src[static_cast<size_t>(list[r]) * d + (i - static_cast<int64_t>(r) * d)];
  }
}
// D[list[i]] = Dsub[i], I[list[i]] = Isub[i] (rows of k)
__global__ void scatter_results_kernel(const float* __restrict__ Dsub, const int64_t* __restrict__ Isub,
                                       const int* __restrict__ list, int n, int k, float* __restrict__ D,
                                       int64_t* __restrict__ I) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < static_cast<int64_t>(n) * k;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / k);
    const size_t o = static_cast<size_t>(list[r]) * k + (i - static_cast<int64_t>(r) * k);
    D[o] = Dsub[i];
    I[o] = Isub[i];
  }
}
// out[i] = outer[inner[i]]: flagged-within-flagged -> indices into the full query set
__global__ void compose_list_kernel(const int* __restrict__ outer, const int* __restrict__ inner, int n, int* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = outer[inner[i]];
}

}  // namespace om

using namespace om;

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct om_comm {
  void* nccl = nullptr;
  int rank = 0, world = 1;
};

namespace {

constexpr int kQueryChunk = 16384;
constexpr int kMaxCandidates = 4096;  // k + slack ceiling (finalize sorts the list in shared memory)

struct DevBuf {  // grow-only device scratch
  void* p = nullptr;
  size_t bytes = 0;
  int reserve(size_t need) {
    if (need <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    OM_CUDA(cudaMalloc(&p, need));
    bytes = need;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

// One pass of the pipeline (scan -> select -> re-score [-> exchange -> merge] -> certify) over nq device-resident
// queries; all pointers are carved from the index's level workspace.
struct Level {
  bool valid = false;  // three-phase API: a search is in progress
  int nq = 0, k = 0, kp = 0, kp_target = 0, C = 0, growth = 2, mode = 0, world = 1, kc = 0, nqc_max = 0;
  const float* qf = nullptr;  // [nq, d] fp32 (not owned by the workspace)
  __half* qh = nullptr;       // [nq, dpad] scan operand
  float *hn = nullptr, *en = nullptr;  // per query ||q_h||, ||q - q_h||
  unsigned long long* cand = nullptr;
  int* count = nullptr;
  float* thr = nullptr;
  int* status = nullptr;  // [0] list overflow, [2] uncertified queries
  uint8_t *send = nullptr, *recv = nullptr;
};

}  // namespace

struct om_index {
  int d = 0, dpad = 0;
  int64_t n = 0, cap = 0;
  float* xf = nullptr;
  __half* xh = nullptr;
  float* gstats = nullptr;  // device [2]: max ||x||, max ||x - x_h|| over the committed rows (float bit patterns)
  int64_t rescore_slack = -1;
  int force_safe = 0;
  int dynamic_sched = 1;  // claim scan tiles from a global counter (keeps CTAs on neighbouring corpus tiles)
  int pair_scan = 1;      // scan GEMM on CTA pairs (cta_group::2, gemm2sm.cuh); 0 = the single-CTA core of gemm.cuh
  int growth = 0;         // each round scans (growth - 1) x the rows seen so far; 0 = auto: 2 for query batches (measured
                          // best at nq = 6 980: fewest filter survivors), 8 for <= 256 queries (HBM-bound streaming
                          // regime: 5 instead of 13 dependent scan + select launch pairs over 8.8 M rows)
  int certify = 1;        // 0: legacy behaviour (top-k of the half-precision candidate stage, no proof)
  int exact_only = 0;     // 1: answer every query with the exact fp32 scan (testing / reference timing)
  int stage_scores = 0;   // 1: emit candidate-stage scores instead of fp32 re-scores (measuring the error model)
  int64_t st_rounds = 0, st_retries = 0, st_capacity = 0, st_launches = 0;
  int64_t st_flagged = 0, st_flagged_wide = 0, st_exact = 0, st_wide_exchange = 0;
  // optional per-phase device timing (CUDA events on the launching stream), enabled by set_param("profile", 1)
  int profile = 0;
  double st_scan_us = 0, st_select_us = 0, st_final_us = 0, st_other_us = 0;
  std::vector<cudaEvent_t> ev;  // pool: [2i] start, [2i+1] stop
  std::vector<int> ev_kind;     // 0 scan, 1 select, 2 finalize, 3 exchange / merge / certify
  size_t ev_used = 0;
  DevBuf ws, ows, sws;  // level workspace / whole-search staging / escalation sub-batch
  int* h_status = nullptr;  // pinned host mirror of Level::status
  Level plan;               // state between om_index_search_begin and om_index_search_finish
};

static int index_grow(om_index* ix, int64_t need) {
  if (need <= ix->cap) return 0;
  int64_t ncap = std::max<int64_t>(need, ix->cap + ix->cap / 2);
  ncap = round_up(std::max<int64_t>(ncap, 1024), 256);
  float* nxf = nullptr;
  __half* nxh = nullptr;
  // rows may still be in flight on the caller's stream(s) (encoder writing reserved rows, a pending commit)
  OM_CUDA(cudaDeviceSynchronize());
  OM_CUDA(cudaMalloc(&nxf, static_cast<size_t>(ncap) * ix->d * sizeof(float)));
  cudaError_t e = cudaMalloc(&nxh, static_cast<size_t>(ncap) * ix->dpad * sizeof(__half));
  if (e != cudaSuccess) {
    cudaFree(nxf);
    cudaGetLastError();
    return fail(OM_ENOMEM, "index: cannot allocate the fp16 scan copy for %lld rows", (long long)ncap);
  }
  if (ix->n > 0) {
    OM_CUDA(cudaMemcpy(nxf, ix->xf, static_cast<size_t>(ix->n) * ix->d * sizeof(float), cudaMemcpyDeviceToDevice));
    OM_CUDA(cudaMemcpy(nxh, ix->xh, static_cast<size_t>(ix->n) * ix->dpad * sizeof(__half), cudaMemcpyDeviceToDevice));
  }
  OM_CUDA(cudaDeviceSynchronize());
  cudaFree(ix->xf);
  cudaFree(ix->xh);
  ix->xf = nxf;
  ix->xh = nxh;
  ix->cap = ncap;
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
  if (cudaMalloc(&ix->gstats, 2 * sizeof(float)) != cudaSuccess || cudaMemset(ix->gstats, 0, 2 * sizeof(float)) != cudaSuccess ||
      cudaHostAlloc(&ix->h_status, 8 * sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(ix->gstats);
    delete ix;
    return fail(OM_ENOMEM, "om_index_create: cannot allocate index state");
  }
  *out = ix;
  return 0;
}

void om_index_destroy(om_index* ix) {
  if (!ix) return;
  cudaFree(ix->xf);
  cudaFree(ix->xh);
  cudaFree(ix->gstats);
  if (ix->h_status) cudaFreeHost(ix->h_status);
  ix->ws.release();
  ix->ows.release();
  ix->sws.release();
  for (cudaEvent_t e : ix->ev) cudaEventDestroy(e);
  delete ix;
}

int64_t om_index_ntotal(const om_index* ix) { return ix ? ix->n : 0; }
int om_index_dim(const om_index* ix) { return ix ? ix->d : 0; }

int om_index_reset(om_index* ix) {
  if (!ix) return fail(OM_EINVAL, "om_index_reset: null index");
  ix->n = 0;
  ix->plan.valid = false;  // a search begun on the old contents cannot be finished
  OM_CUDA(cudaMemset(ix->gstats, 0, 2 * sizeof(float)));
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
  rows_to_f16_kernel<<<grid_for(n, 8), 256, 0, st>>>(ix->xf + static_cast<size_t>(ix->n) * ix->d,
                                                     ix->xh + static_cast<size_t>(ix->n) * ix->dpad, n, ix->d, ix->dpad,
                                                     nullptr, nullptr, ix->gstats);
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
    if (value != 0 && (value < 2 || value > 8)) return fail(OM_EINVAL, "round_growth must be 0 (auto) or in [2, 8]");
    ix->growth = static_cast<int>(value);
  } else if (!strcmp(name, "dynamic_sched")) {
    ix->dynamic_sched = value != 0;
  } else if (!strcmp(name, "pair_scan")) {
    ix->pair_scan = value != 0;
  } else if (!strcmp(name, "profile")) {
    ix->profile = static_cast<int>(value);
  } else if (!strcmp(name, "certify")) {
    ix->certify = value != 0;
  } else if (!strcmp(name, "exact_only")) {
    ix->exact_only = value != 0;
  } else if (!strcmp(name, "debug_stage_scores")) {
    ix->stage_scores = value != 0;
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
  if (!strcmp(name, "uncertified")) return ix->st_flagged;
  if (!strcmp(name, "uncertified_wide")) return ix->st_flagged_wide;
  if (!strcmp(name, "exact_queries")) return ix->st_exact;
  if (!strcmp(name, "wide_exchanges")) return ix->st_wide_exchange;
  if (!strcmp(name, "scan_ns")) return static_cast<int64_t>(ix->st_scan_us * 1e3);
  if (!strcmp(name, "select_ns")) return static_cast<int64_t>(ix->st_select_us * 1e3);
  if (!strcmp(name, "finalize_ns")) return static_cast<int64_t>(ix->st_final_us * 1e3);
  if (!strcmp(name, "other_ns")) return static_cast<int64_t>(ix->st_other_us * 1e3);
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
    static const char* nvtx_names[4] = {"om.search.scan", "om.search.select", "om.search.rescore", "om.search.exchange_certify"};
    nvtxRangePushA(nvtx_names[kind & 3]);
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
    nvtxRangePop();
    if (!on) return;
    cudaEventRecord(ix->ev[ix->ev_used + 1], st);
    ix->ev_used += 2;
  }
};

void collect_profile(om_index* ix) {
  static const char* names[4] = {"scan", "select", "finalize", "exchange+certify"};
  for (size_t i = 0; i + 1 < ix->ev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ix->ev[i], ix->ev[i + 1]) != cudaSuccess) continue;
    const int kind = ix->ev_kind[i / 2];
    (kind == 0 ? ix->st_scan_us : kind == 1 ? ix->st_select_us : kind == 2 ? ix->st_final_us : ix->st_other_us) += ms * 1e3;
    if (ix->profile >= 2) fprintf(stderr, "[om profile] launch %zu %s %.3f ms\n", i / 2, names[kind & 3], ms);
  }
  ix->ev_used = 0;
}

int once_attrs() {
  static bool done = false;  // one device per process (enforced by device_sm_count)
  if (done) return 0;
  OM_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
  OM_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxCandidates * 8 + 65536));
  OM_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16));
  OM_CUDA(cudaFuncSetAttribute(exact_scan_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  done = true;
  return 0;
}

// Packed per-shard block of the sharded exchange for nqc queries shipping kc entries each:
//   [D f32 nqc x kc | I i64 nqc x kc | floor f32 nqc (kp-th stage score of the shard's list) | error-norm maxima f32 x 2]
struct ExchangeBlock {
  size_t off_i, off_floor, off_stats, bytes;
};
inline ExchangeBlock exchange_block(size_t nqc, size_t kc) {
  ExchangeBlock b;
  b.off_i = round_up(nqc * kc * 4, 256);
  b.off_floor = b.off_i + round_up(nqc * kc * 8, 256);
  b.off_stats = b.off_floor + round_up(nqc * 4, 256);
  b.bytes = b.off_stats + 256;
  return b;
}
inline size_t exchange_block_bytes(size_t nqc, size_t kc) { return exchange_block(nqc, kc).bytes; }

// sizes the candidate lists of a level, carves the level workspace and converts the queries
int level_prepare(om_index* ix, Level& L, const float* qf, int nq, int k, int kp_target, int mode, int world,
                  cudaStream_t st) {
  const int d = ix->d, dpad = ix->dpad;
  if (d > 16384) return fail(OM_EINVAL, "om_index_search: d > 16384 unsupported");
  if (k > kMaxCandidates) return fail(OM_EINVAL, "om_index_search: k = %d exceeds %d", k, kMaxCandidates);
  OM_TRY(once_attrs());
  L.valid = false;
  L.qf = qf;
  L.nq = nq;
  L.k = k;
  L.mode = mode;
  L.world = world;
  L.kp_target = std::min(std::max(kp_target, k), kMaxCandidates);
  // Row-sharded search: the global top-(k + slack) draws ~(k + slack) / W rows from every shard (binomial: mean m = kp / W,
  // deviation sqrt(m)), so each shard keeps a list of m + 6 sqrt(m) + 32 rows instead of k + slack: scan survivors, select
  // and re-score work all shrink W-fold and the whole list is shipped (no agreement phase).  A shard that holds more of the answer than its list (skewed shards) shows up
  // as a high floor: the certificate fails and the query is escalated to the 4096-wide level, which keeps full lists.
  if (world > 1 && mode == 0 && L.kp_target < kMaxCandidates) {
    const double m = static_cast<double>(L.kp_target) / world;
    L.kp_target = static_cast<int>(std::min<int64_t>(L.kp_target, round_up(static_cast<int64_t>(m + 6.0 * sqrt(m) + 32.0), 32)));
  }
  L.kp = static_cast<int>(std::min<int64_t>(L.kp_target, std::max<int64_t>(ix->n, 1)));
  // Expected list length after a round that multiplies the rows seen by g is ~g kp (kp kept + ~(g-1) kp new
  // survivors); 25 % + 512 entries of head-room cover its spread for exchangeable row order.
  for (L.growth = ix->growth > 0 ? ix->growth : (nq <= 256 ? 8 : 2);; --L.growth) {  // large k: slower-growing schedule that fits the 16384-entry select
    L.C = 1024;
    while (L.C < (5 * L.growth * L.kp) / 4 + 512) L.C <<= 1;
    if (L.C <= 16384 || L.growth == 2) break;
  }
  if (L.C > 16384) return fail(OM_EINVAL, "om_index_search: candidate list of %d entries exceeds 16384; lower k", L.C);
  L.kc = std::min(k, L.kp_target);  // entries a shard ships per query (it cannot contribute more than k)
  L.nqc_max = std::min(nq, kQueryChunk);
  ix->st_capacity = L.C;
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    size_t o = off;
    off += round_up(bytes, 256);
    return o;
  };
  const size_t nqc = L.nqc_max;
  const size_t o_qh = carve(static_cast<size_t>(nq) * dpad * 2), o_hn = carve(static_cast<size_t>(nq) * 4),
               o_en = carve(static_cast<size_t>(nq) * 4), o_cand = carve(nqc * L.C * 8), o_count = carve(nqc * 4),
               o_thr = carve(nqc * 4), o_status = carve(256);
  size_t o_send = 0, o_recv = 0;
  if (world > 1) {
    const size_t blk = exchange_block_bytes(nqc, L.kc);
    o_send = carve(blk);
    o_recv = carve(blk * world);
  }
  OM_TRY(ix->ws.reserve(off));
  uint8_t* base = static_cast<uint8_t*>(ix->ws.p);
  L.qh = reinterpret_cast<__half*>(base + o_qh);
  L.hn = reinterpret_cast<float*>(base + o_hn);
  L.en = reinterpret_cast<float*>(base + o_en);
  L.cand = reinterpret_cast<unsigned long long*>(base + o_cand);
  L.count = reinterpret_cast<int*>(base + o_count);
  L.thr = reinterpret_cast<float*>(base + o_thr);
  L.status = reinterpret_cast<int*>(base + o_status);
  L.send = world > 1 ? base + o_send : nullptr;
  L.recv = world > 1 ? base + o_recv : nullptr;
  if (mode == 0) {
    rows_to_f16_kernel<<<grid_for(nq, 8), 256, 0, st>>>(qf, L.qh, nq, d, dpad, L.hn, L.en, nullptr);
    OM_CUDA(cudaGetLastError());
    ix->st_launches += 1;
  }
  return 0;
}

// One sweep of the shard for queries [q0, q0 + nqc) of the level.  safe = false: doubling rounds; safe = true: fixed
// rounds of C - kp rows, which cannot overflow.  mode 0: fp16 tensor-core scan, 1: exact fp32 scan.
int sweep_chunk(om_index* ix, const Level& L, int q0, int nqc, bool safe, int sms, cudaStream_t st) {
  const int64_t growth = L.growth;
  const int64_t N = ix->n;
  const int C = L.C, kp = L.kp;
  if (N == 0) {
    fill_i32<<<(nqc + 255) / 256, 256, 0, st>>>(L.count, 0, nqc);
    fill_i32<<<(nqc + 255) / 256, 256, 0, st>>>(reinterpret_cast<int*>(L.thr), static_cast<int>(0xff800000), nqc);
    OM_CUDA(cudaGetLastError());
    return 0;
  }
  const __half* qh = L.qh + static_cast<size_t>(q0) * ix->dpad;
  const float* qf = L.qf + static_cast<size_t>(q0) * ix->d;
  int* overflow = L.status;
  int64_t pos = 0;
  const size_t sel_smem = static_cast<size_t>(C) * 8;
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
      if (L.mode == 0) {
        const __half* xrows = ix->xh + static_cast<size_t>(pos) * ix->dpad;
        cudaError_t e = cudaErrorNotSupported;
        const bool dynsched = ix->dynamic_sched != 0;
        const int ncols = static_cast<int>(step);
        // CTA pairs own 256 query rows per tile: with <= 128 queries the peer's half would be padding (and the sweep is
        // HBM-bound there: the single-CTA ring keeps more corpus bytes in flight per SM)
        const bool pair = ix->pair_scan != 0 && nqc > kBlockM;
        if (first) {
          EpiScan<true> epi{L.thr, L.cand, L.count, overflow, nqc, ncols, C, static_cast<uint32_t>(pos)};
          if (pair) e = launch_gemm2<5, true, 8, true>(qh, ix->dpad, xrows, ix->dpad, nqc, ncols, ix->d, epi, sms, st, dynsched);
          if (e == cudaErrorNotSupported)  // no CTA pair fits (a device without two free SMs per TPC)
            e = launch_gemm<256, 4, true, 8, EpiScan<true>, true>(qh, ix->dpad, xrows, ix->dpad, nqc, ncols, ix->d, epi, sms, st, dynsched);
        } else {
          EpiScan<false> epi{L.thr, L.cand, L.count, overflow, nqc, ncols, C, static_cast<uint32_t>(pos)};
          if (pair) e = launch_gemm2<5, true, 8, true>(qh, ix->dpad, xrows, ix->dpad, nqc, ncols, ix->d, epi, sms, st, dynsched);
          if (e == cudaErrorNotSupported)
            e = launch_gemm<256, 4, true, 8, EpiScan<false>, true>(qh, ix->dpad, xrows, ix->dpad, nqc, ncols, ix->d, epi, sms, st, dynsched);
        }
        if (e != cudaSuccess) return fail(OM_ECUDA, "scan kernel launch failed: %s", cudaGetErrorString(e));
      } else {
        const int nqt = std::max(1, std::min(8, (96 * 1024) / (ix->d * 4)));
        dim3 grid(static_cast<unsigned>(std::min<int64_t>((step + 15) / 16, static_cast<int64_t>(sms) * 4)),
                  static_cast<unsigned>((nqc + nqt - 1) / nqt));
        exact_scan_kernel<8, 2><<<grid, 256, static_cast<size_t>(nqt) * ix->d * 4, st>>>(
            ix->xf + static_cast<size_t>(pos) * ix->d, step, static_cast<uint32_t>(pos), qf, nqc, ix->d, nqt, L.thr, L.cand,
            L.count, overflow, C, first ? 1 : 0);
        OM_CUDA(cudaGetLastError());
      }
    }
    {
      Timed t(ix, st, 1);
      select_kernel<<<nqc, 256, sel_smem, st>>>(L.cand, L.count, L.thr, C, kp, first ? static_cast<int>(step) : -1);
    }
    OM_CUDA(cudaGetLastError());
    ix->st_launches += 2;
    pos += step;
    first = false;
    ix->st_rounds++;
  }
  return 0;
}

int finalize_chunk(om_index* ix, const Level& L, int q0, int nqc, float* D, int64_t* I, int k_out, int64_t id_offset,
                   const float* range, const int* ghist, int* kept_max, int* exceed, cudaStream_t st) {
  int P2 = 2;
  while (P2 < L.kp) P2 <<= 1;
  const size_t fin_smem = static_cast<size_t>(P2) * 8 + static_cast<size_t>(ix->d) * 4;
  {
    Timed t(ix, st, 2);
    finalize_kernel<<<nqc, 256, fin_smem, st>>>(L.cand, L.count, L.C, L.qf + static_cast<size_t>(q0) * ix->d, ix->xf, ix->d,
                                                L.k, D, I, id_offset, range, ghist, nqc, L.kp_target, kept_max, k_out,
                                                exceed, ix->stage_scores);
  }
  OM_CUDA(cudaGetLastError());
  ix->st_launches += 1;
  return 0;
}

#define OM_NCCL(expr)                                                                                   \
  do {                                                                                                  \
    int r__ = (expr);                                                                                   \
    if (r__ != 0) return fail(OM_ECUDA, "%s failed: %s", #expr, nccl_api().GetErrorString(r__));        \
  } while (0)

// Merge of nparts [nq, k_in] lists -> [nq, k_out]; more than 8192 entries per query are merged hierarchically.
int merge_parts(const float* Dp, const int64_t* Ip, int64_t stride_d, int64_t stride_i, int nparts, int nq, int k_in,
                int k_out, float* D, int64_t* I, cudaStream_t st) {
  OM_TRY(once_attrs());
  if (k_in > 8192) return fail(OM_EINVAL, "om_topk_merge: k_in = %d exceeds 8192", k_in);
  if (static_cast<int64_t>(nparts) * k_in <= 8192) {
    int P = 2;
    while (P < nparts * k_in) P <<= 1;
    const size_t smem = static_cast<size_t>(P) * 8 + static_cast<size_t>(nparts) * k_in * 8;
    merge_kernel<<<nq, 256, smem, st>>>(Dp, Ip, stride_d, stride_i, nparts, nq, k_in, k_out, D, I);
    OM_CUDA(cudaGetLastError());
    return 0;
  }
  // groups of g parts -> intermediate lists of width k_mid, then recurse on the groups
  const int g = std::max(2, 8192 / k_in);
  const int ngroups = (nparts + g - 1) / g;
  const int k_mid = static_cast<int>(std::min<int64_t>(k_out, static_cast<int64_t>(g) * k_in));
  float* Dm = nullptr;
  int64_t* Im = nullptr;
  const size_t per = static_cast<size_t>(nq) * k_mid;
  OM_CUDA(cudaMallocAsync(&Dm, per * ngroups * 4, st));
  OM_CUDA(cudaMallocAsync(&Im, per * ngroups * 8, st));
  int rc = 0;
  for (int gi = 0; gi < ngroups && rc == 0; ++gi) {
    const int p0 = gi * g, np = std::min(g, nparts - p0);
    rc = merge_parts(Dp + p0 * stride_d, Ip + p0 * stride_i, stride_d, stride_i, np, nq, k_in, k_mid, Dm + gi * per,
                     Im + gi * per, st);
  }
  if (rc == 0) rc = merge_parts(Dm, Im, static_cast<int64_t>(per), static_cast<int64_t>(per), ngroups, nq, k_mid, k_out, D, I, st);
  cudaFreeAsync(Dm, st);
  cudaFreeAsync(Im, st);
  return rc;
}

// Exchange of one query chunk of a row-sharded level: every shard re-scores its (shard-sized) candidate list in fp32 and
// ships it whole — sorted scores, global ids, the list's stage-score floor and the shard's error-norm maxima — in ONE
// packed all-gather; every rank then merges the W lists.  (The certificate is run by the caller on the merged result.)
int exchange_chunk(om_index* ix, om_comm* comm, const Level& L, int q0, int nqc, float* dD, int64_t* dI, int64_t id_offset,
                   cudaStream_t st) {
  NcclApi& nc = nccl_api();
  const int W = comm->world;
  const ExchangeBlock b = exchange_block(nqc, L.kc);
  OM_TRY(finalize_chunk(ix, L, q0, nqc, reinterpret_cast<float*>(L.send), reinterpret_cast<int64_t*>(L.send + b.off_i), L.kc,
                        id_offset, nullptr, nullptr, nullptr, nullptr, st));
  {
    Timed t(ix, st, 3);
    OM_CUDA(cudaMemcpyAsync(L.send + b.off_floor, L.thr, static_cast<size_t>(nqc) * 4, cudaMemcpyDeviceToDevice, st));
    OM_CUDA(cudaMemcpyAsync(L.send + b.off_stats, ix->gstats, 8, cudaMemcpyDeviceToDevice, st));
    OM_NCCL(nc.AllGather(L.send, L.recv, b.bytes, kNcclInt8, comm->nccl, st));
    OM_TRY(merge_parts(reinterpret_cast<const float*>(L.recv), reinterpret_cast<const int64_t*>(L.recv + b.off_i),
                       static_cast<int64_t>(b.bytes / 4), static_cast<int64_t>(b.bytes / 8), W, nqc, L.kc, L.k,
                       dD + static_cast<size_t>(q0) * L.k, dI + static_cast<size_t>(q0) * L.k, st));
  }
  ix->st_launches += 2;
  return 0;
}

// Runs one level over all its queries and returns the number of uncertified ones (their indices in flag_list).
// One host synchronisation at the end (status word); a list overflow redoes the level with the safe schedule.
int run_level(om_index* ix, om_comm* comm, Level& L, float* dD, int64_t* dI, int64_t id_offset, int* flags,
              int* flag_list, int* nflag_out, cudaStream_t st) {
  const int sms = device_sm_count();
  if (sms < 0) return sms;
  NvtxRange nvtx(L.mode == 1 ? "om.search.level_exact" : (L.kp_target >= kMaxCandidates ? "om.search.level_wide" : "om.search.level0"));
  const bool sharded = comm && comm->world > 1;
  bool safe = ix->force_safe != 0;
  const bool certify = ix->certify && L.mode == 0 && flags != nullptr && !ix->stage_scores;
  for (int attempt = 0; attempt < 4; ++attempt) {
    OM_CUDA(cudaMemsetAsync(L.status, 0, 32, st));
    for (int q0 = 0; q0 < L.nq; q0 += kQueryChunk) {
      const int nqc = std::min(kQueryChunk, L.nq - q0);
      OM_TRY(sweep_chunk(ix, L, q0, nqc, safe, sms, st));
      if (sharded)
        OM_TRY(exchange_chunk(ix, comm, L, q0, nqc, dD, dI, id_offset, st));
      else
        OM_TRY(finalize_chunk(ix, L, q0, nqc, dD + static_cast<size_t>(q0) * L.k, dI + static_cast<size_t>(q0) * L.k, L.k,
                              id_offset, nullptr, nullptr, nullptr, nullptr, st));
      if (certify) {
        Timed t(ix, st, 3);
        const ExchangeBlock b = exchange_block(nqc, L.kc);
        const float* floors = sharded ? reinterpret_cast<const float*>(L.recv + b.off_floor) : L.thr;
        const float* gst = sharded ? reinterpret_cast<const float*>(L.recv + b.off_stats) : ix->gstats;
        certify_kernel<<<(nqc + 7) / 8, 256, 0, st>>>(dD + static_cast<size_t>(q0) * L.k, dI + static_cast<size_t>(q0) * L.k,
                                                      L.k, floors, static_cast<int64_t>(b.bytes / 4), sharded ? comm->world : 1,
                                                      gst, static_cast<int64_t>(b.bytes / 4), L.hn + q0, L.en + q0, ix->d, nqc,
                                                      q0, flags, L.status + 2);
        OM_CUDA(cudaGetLastError());
        ix->st_launches += 1;
      }
    }
    if (sharded)  // every rank must take the same retry decision (list overflow on any shard)
      OM_NCCL(nccl_api().AllReduce(L.status, L.status, 1, kNcclInt32, kNcclMax, comm->nccl, st));
    OM_CUDA(cudaMemcpyAsync(ix->h_status, L.status, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    OM_CUDA(cudaStreamSynchronize(st));
    const unsigned int fault = read_clear_dev_fault();
    if (fault) return fail(OM_EFAULT, "scan kernel pipeline fault 0x%08x", fault);
    if (ix->h_status[0]) {
      if (safe) return fail(OM_EFAULT, "candidate list overflow in the overflow-proof schedule (bug)");
      safe = true;
      ix->st_retries++;
      continue;
    }
    *nflag_out = certify ? ix->h_status[2] : 0;
    if (*nflag_out > 0) {  // ascending list of the uncertified queries (identical on every rank)
      compact_flags_kernel<<<1, 1024, 0, st>>>(flags, L.nq, flag_list);
      OM_CUDA(cudaGetLastError());
      ix->st_launches += 1;
    }
    return 0;
  }
  return fail(OM_EFAULT, "search level did not converge (bug)");
}

// The whole search: level 0 (all queries, k + slack candidates) -> level 1 (uncertified queries, widest list) ->
// level 2 (still uncertified: exact fp32 scan).  comm == nullptr / world 1: single shard.
int search_impl(om_index* ix, om_comm* comm, const void* q, om_memkind q_kind, int nq, int k, float* D, int64_t* I,
                om_memkind out_kind, int64_t id_offset, cudaStream_t st) {
  NvtxRange nvtx("om.search");
  const int d = ix->d;
  const int world = comm ? comm->world : 1;
  ix->plan.valid = false;  // the level workspace is about to be reused
  ix->st_rounds = ix->st_retries = ix->st_launches = 0;
  ix->st_flagged = ix->st_flagged_wide = ix->st_exact = ix->st_wide_exchange = 0;
  ix->st_scan_us = ix->st_select_us = ix->st_final_us = ix->st_other_us = 0;
  ix->ev_used = 0;
  // whole-search staging: queries (if they arrive from the host), results (if they leave to the host), flag list
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    size_t o = off;
    off += round_up(bytes, 256);
    return o;
  };
  const size_t o_q = carve(q_kind == OM_HOST ? static_cast<size_t>(nq) * d * 4 : 0);
  const size_t o_D = carve(out_kind == OM_HOST ? static_cast<size_t>(nq) * k * 4 : 0);
  const size_t o_I = carve(out_kind == OM_HOST ? static_cast<size_t>(nq) * k * 8 : 0);
  const size_t o_flags = carve(static_cast<size_t>(nq) * 4);  // per-query 0 / 1 written by the certificate of a level
  const size_t o_flag = carve(static_cast<size_t>(nq) * 4);   // level-0 uncertified queries (ascending indices)
  const size_t o_sub = carve(static_cast<size_t>(nq) * 4);    // uncertified within an escalation sub-batch
  const size_t o_flag2 = carve(static_cast<size_t>(nq) * 4);  // ... composed back to indices into the full set
  OM_TRY(ix->ows.reserve(off));
  uint8_t* ob = static_cast<uint8_t*>(ix->ows.p);
  const float* qf = static_cast<const float*>(q);
  if (q_kind == OM_HOST) {
    OM_CUDA(cudaMemcpyAsync(ob + o_q, q, static_cast<size_t>(nq) * d * 4, cudaMemcpyHostToDevice, st));
    qf = reinterpret_cast<const float*>(ob + o_q);
  }
  float* dD = out_kind == OM_HOST ? reinterpret_cast<float*>(ob + o_D) : D;
  int64_t* dI = out_kind == OM_HOST ? reinterpret_cast<int64_t*>(ob + o_I) : I;
  int* flags = reinterpret_cast<int*>(ob + o_flags);
  int* flag_list = reinterpret_cast<int*>(ob + o_flag);
  int* sub_flags = reinterpret_cast<int*>(ob + o_sub);
  int* flag_list2 = reinterpret_cast<int*>(ob + o_flag2);

  const int64_t slack = ix->rescore_slack >= 0 ? ix->rescore_slack : std::max<int64_t>(128, k / 5);
  const int kp0 = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(k) + slack, kMaxCandidates));
  int nf = 0;
  if (!ix->exact_only) {
    Level L;
    OM_TRY(level_prepare(ix, L, qf, nq, k, kp0, 0, world, st));
    OM_TRY(run_level(ix, comm, L, dD, dI, id_offset, flags, flag_list, &nf, st));
    ix->st_flagged = nf;
  }
  // Escalation: the queries listed in `list` (indices into the full set; nullptr = all of them) are gathered into a
  // compact sub-batch, answered by one more level and scattered back over their rows of (dD, dI).
  auto run_sub = [&](const int* list, int n_sub, int kp_target, int mode, int* nf_out) -> int {
    Level Ls;
    if (!list) {
      OM_TRY(level_prepare(ix, Ls, qf, n_sub, k, kp_target, mode, world, st));
      return run_level(ix, comm, Ls, dD, dI, id_offset, flags, sub_flags, nf_out, st);
    }
    size_t so = 0;
    auto scarve = [&](size_t bytes) {
      size_t o = so;
      so += round_up(bytes, 256);
      return o;
    };
    const size_t s_q = scarve(static_cast<size_t>(n_sub) * d * 4), s_D = scarve(static_cast<size_t>(n_sub) * k * 4),
                 s_I = scarve(static_cast<size_t>(n_sub) * k * 8);
    OM_TRY(ix->sws.reserve(so));
    uint8_t* sb = static_cast<uint8_t*>(ix->sws.p);
    float* qsub = reinterpret_cast<float*>(sb + s_q);
    float* Ds = reinterpret_cast<float*>(sb + s_D);
    int64_t* Is = reinterpret_cast<int64_t*>(sb + s_I);
    gather_rows_kernel<<<grid_for(static_cast<int64_t>(n_sub) * d, 256), 256, 0, st>>>(qf, list, n_sub, d, qsub);
    OM_CUDA(cudaGetLastError());
    OM_TRY(level_prepare(ix, Ls, qsub, n_sub, k, kp_target, mode, world, st));
    OM_TRY(run_level(ix, comm, Ls, Ds, Is, id_offset, flags, sub_flags, nf_out, st));
    scatter_results_kernel<<<grid_for(static_cast<int64_t>(n_sub) * k, 256), 256, 0, st>>>(Ds, Is, list, n_sub, k, dD, dI);
    OM_CUDA(cudaGetLastError());
    ix->st_launches += 2;
    return 0;
  };
  if (ix->exact_only) {
    int dummy = 0;
    ix->st_exact = nq;
    OM_TRY(run_sub(nullptr, nq, k, 1, &dummy));
  } else if (nf > 0) {
    const int* list = flag_list;
    // level 1: widest candidate list the select / sort kernels take.  (Sharded: the decision must not depend on this
    // rank's row count — every rank runs the same levels.)
    if (kp0 < kMaxCandidates && (world > 1 || ix->n > kp0)) {
      int nf1 = 0;
      OM_TRY(run_sub(list, nf, kMaxCandidates, 0, &nf1));
      if (nf1 > 0) {
        compose_list_kernel<<<(nf1 + 255) / 256, 256, 0, st>>>(list, sub_flags, nf1, flag_list2);
        OM_CUDA(cudaGetLastError());
        list = flag_list2;
      }
      nf = nf1;
    }
    ix->st_flagged_wide = nf;
    if (nf > 0) {  // level 2: exact fp32 scan
      int dummy = 0;
      ix->st_exact = nf;
      OM_TRY(run_sub(list, nf, k, 1, &dummy));
    }
  }
  if (out_kind == OM_HOST) {
    OM_CUDA(cudaMemcpyAsync(D, dD, static_cast<size_t>(nq) * k * 4, cudaMemcpyDeviceToHost, st));
    OM_CUDA(cudaMemcpyAsync(I, dI, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, st));
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
  OM_TRY(device_sm_count());
  return search_impl(ix, nullptr, q, q_kind, nq, k, D, I, out_kind, id_offset, static_cast<cudaStream_t>(stream));
}

// ---- row-sharded search with the exchange inside the library (NCCL over NVLink) -----------------------------------
extern "C" int om_comm_unique_id(char* out128) {
  if (!out128) return fail(OM_EINVAL, "om_comm_unique_id: null buffer");
  NcclApi& nc = nccl_api();
  if (!nc.handle) return fail(OM_ESTATE, "NCCL unavailable: %s", nc.why ? nc.why : "?");
  NcclUid id;
  OM_NCCL(nc.GetUniqueId(&id));
  memcpy(out128, id.internal, 128);
  return 0;
}

extern "C" int om_comm_init(const char* unique_id, int rank, int world, om_comm** out) {
  if (!unique_id || !out || world <= 0 || rank < 0 || rank >= world) return fail(OM_EINVAL, "om_comm_init: bad arguments");
  OM_TRY(device_sm_count());
  NcclApi& nc = nccl_api();
  if (!nc.handle) return fail(OM_ESTATE, "NCCL unavailable: %s", nc.why ? nc.why : "?");
  om_comm* c = new (std::nothrow) om_comm();
  if (!c) return fail(OM_ENOMEM, "om_comm_init: out of host memory");
  c->rank = rank;
  c->world = world;
  NcclUid id;
  memcpy(id.internal, unique_id, 128);
  const int r = nc.CommInitRank(&c->nccl, world, id, rank);
  if (r != 0) {
    delete c;
    return fail(OM_ECUDA, "ncclCommInitRank failed: %s", nc.GetErrorString(r));
  }
  *out = c;
  return 0;
}

extern "C" void om_comm_destroy(om_comm* c) {
  if (!c) return;
  if (c->nccl) nccl_api().CommDestroy(c->nccl);
  delete c;
}

extern "C" int om_index_search_sharded(om_index* ix, om_comm* comm, const void* q, om_memkind q_kind, int nq, int k,
                                       float* D, int64_t* I, om_memkind out_kind, int64_t id_offset, void* stream) {
  if (!ix || !comm || (nq > 0 && (!q || !D || !I)) || nq < 0 || k <= 0)
    return fail(OM_EINVAL, "om_index_search_sharded: bad arguments (nq=%d k=%d)", nq, k);
  if (nq == 0) return 0;
  OM_TRY(device_sm_count());
  return search_impl(ix, comm, q, q_kind, nq, k, D, I, out_kind, id_offset, static_cast<cudaStream_t>(stream));
}

// ---- three-phase building blocks (one shard per call; the caller reduces between the phases) ----------------------
extern "C" int om_index_search_begin(om_index* ix, const void* q, om_memkind q_kind, int nq, int k, float* local_range,
                                     void* stream) {
  if (!ix || nq <= 0 || !q || !local_range || k <= 0) return fail(OM_EINVAL, "om_index_search_begin: bad arguments");
  if (nq > kQueryChunk) return fail(OM_EINVAL, "om_index_search_begin: at most %d queries per call", kQueryChunk);
  const int sms = device_sm_count();
  if (sms < 0) return sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the queries must outlive this call (finish re-scores against them): keep a copy in the staging buffer
  OM_TRY(ix->ows.reserve(round_up(static_cast<size_t>(nq) * ix->d * 4, 256)));
  OM_CUDA(cudaMemcpyAsync(ix->ows.p, q, static_cast<size_t>(nq) * ix->d * 4,
                          q_kind == OM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
  ix->st_rounds = ix->st_retries = ix->st_launches = 0;
  const int64_t slack = ix->rescore_slack >= 0 ? ix->rescore_slack : std::max<int64_t>(128, k / 5);
  Level& L = ix->plan;
  OM_TRY(level_prepare(ix, L, static_cast<const float*>(ix->ows.p), nq, k,
                       static_cast<int>(std::min<int64_t>(static_cast<int64_t>(k) + slack, kMaxCandidates)), 0, 1, st));
  bool safe = ix->force_safe != 0;
  for (int attempt = 0;; ++attempt) {
    OM_CUDA(cudaMemsetAsync(L.status, 0, 32, st));
    OM_TRY(sweep_chunk(ix, L, 0, nq, safe, sms, st));
    OM_CUDA(cudaMemcpyAsync(ix->h_status, L.status, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    OM_CUDA(cudaStreamSynchronize(st));
    const unsigned int fault = read_clear_dev_fault();
    if (fault) return fail(OM_EFAULT, "scan kernel pipeline fault 0x%08x", fault);
    if (!ix->h_status[0]) break;
    if (safe || attempt > 0) return fail(OM_EFAULT, "candidate list overflow in the overflow-proof schedule (bug)");
    safe = true;
    ix->st_retries++;
  }
  local_range_kernel<<<nq, 256, 0, st>>>(L.cand, L.count, L.thr, L.C, nq, ix->n >= L.kp_target ? 1 : 0, local_range, nullptr);
  OM_CUDA(cudaGetLastError());
  ix->st_launches += 1;
  L.valid = true;
  return 0;
}

extern "C" int om_index_search_count(om_index* ix, const float* global_range, int* local_hist, void* stream) {
  if (!ix || !global_range || !local_hist) return fail(OM_EINVAL, "om_index_search_count: bad arguments");
  if (!ix->plan.valid) return fail(OM_ESTATE, "om_index_search_count: no search in progress (call om_index_search_begin)");
  const Level& L = ix->plan;
  floor_hist_kernel<<<L.nq, 256, 0, static_cast<cudaStream_t>(stream)>>>(L.cand, L.count, L.C, global_range, L.nq, local_hist);
  OM_CUDA(cudaGetLastError());
  ix->st_launches += 1;
  return 0;
}

extern "C" int om_index_search_finish(om_index* ix, const float* global_range, const int* global_hist, float* D,
                                      int64_t* I, int64_t id_offset, int* kept_max, void* stream) {
  if (!ix || !D || !I || (global_hist && !global_range)) return fail(OM_EINVAL, "om_index_search_finish: bad arguments");
  if (!ix->plan.valid) return fail(OM_ESTATE, "om_index_search_finish: no search in progress (call om_index_search_begin)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Level& L = ix->plan;
  L.valid = false;
  if (kept_max) OM_CUDA(cudaMemsetAsync(kept_max, 0, sizeof(int), st));
  OM_TRY(finalize_chunk(ix, L, 0, L.nq, D, I, L.k, id_offset, global_range, global_hist, kept_max, nullptr, st));
  OM_CUDA(cudaStreamSynchronize(st));
  if (ix->profile) collect_profile(ix);
  return 0;
}

extern "C" int om_search_floor_bins(void) { return kFloorBins; }

extern "C" int om_topk_merge_n(const float* D_parts, const int64_t* I_parts, int nparts, int nq, int k_in, int k_out,
                               float* D, int64_t* I, void* stream) {
  if (nparts <= 0 || nq < 0 || k_in <= 0 || k_out <= 0 || !D_parts || !I_parts || !D || !I)
    return fail(OM_EINVAL, "om_topk_merge: bad arguments");
  if (nq == 0) return 0;
  OM_TRY(device_sm_count());
  const int64_t stride = static_cast<int64_t>(nq) * k_in;
  return merge_parts(D_parts, I_parts, stride, stride, nparts, nq, k_in, k_out, D, I, static_cast<cudaStream_t>(stream));
}

extern "C" int om_topk_merge(const float* D_parts, const int64_t* I_parts, int nparts, int nq, int k, float* D,
                             int64_t* I, void* stream) {
  return om_topk_merge_n(D_parts, I_parts, nparts, nq, k, k, D, I, stream);
}
