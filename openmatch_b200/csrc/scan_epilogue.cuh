// Fused top-k filter epilogue of the search scan GEMM + candidate key encoding (shared by search.cu and the
// bring-up probe selftest_gemm.cu).
#pragma once
#include <string.h>

#include "gemm.cuh"

namespace om {

// ---------------------------------------------------------------------------------------------------
// candidate keys: descending unsigned order == (score descending, row ascending)
// ---------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_orderable(float s) {
  s = s + 0.0f;  // -0 -> +0
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(s);
#else
  uint32_t u;
  memcpy(&u, &s, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_from_orderable(uint32_t u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float s;
  memcpy(&s, &u, 4);
  return s;
#endif
}
__host__ __device__ __forceinline__ unsigned long long make_key(float s, uint32_t row) {
  return (static_cast<unsigned long long>(f32_orderable(s)) << 32) | static_cast<unsigned long long>(0xffffffffu - row);
}
__host__ __device__ __forceinline__ uint32_t key_row(unsigned long long k) {
  return 0xffffffffu - static_cast<uint32_t>(k & 0xffffffffull);
}
__host__ __device__ __forceinline__ float key_score(unsigned long long k) {
  return f32_from_orderable(static_cast<uint32_t>(k >> 32));
}

// ---------------------------------------------------------------------------------------------------
// fused scan epilogue
// ---------------------------------------------------------------------------------------------------
// VARIANT selects the filter's code shape for the sparse rounds (same results, different cost per survivor):
//   0  shipped: one 32-column max, then a 32-bit mask and a 31-SEL select tree per survivor
//   1  candidate for the next round (unmeasured): 4 group maxima of 8 columns; only groups that beat the
//      threshold are expanded (8-bit mask, 7-SEL tree), the group's 8 values chosen by a select on the group id
//   2  candidate (unmeasured): like 1 with the four group bodies unrolled (no group select, more cold code)
//   3  candidate (unmeasured): warp-cooperative extraction - a lane whose chunk beats its threshold broadcasts its
//      32 values with 32 shuffles so that lane i holds column i; one compare + ballot finds every survivor of the
//      chunk at once and the surviving lanes store their own keys (no mask loop, no select tree, no single-lane
//      dependent chain; cost independent of the number of survivors in the chunk)
// Measured with selftest perf_scan (variant 0): 1 survivor per warp-tile costs ~15 % of the scan rate, i.e. about
// a thousand cycles of a single-lane dependent chain per survivor - the term to shrink.
template <bool DENSE, int EPI_THREADS = 256, int VARIANT = 0>
struct EpiScan {
  const float* thr;          // [nq] strict lower bound per query
  unsigned long long* cand;  // [nq, C]
  int* count;                // [nq]
  int* overflow;             // single flag
  int nq, n_cols, C;
  uint32_t row_base;  // corpus row of column 0 of this round
  // DENSE: first round, every score is stored at position = column (no threshold yet)
  // Pass 0: a thread compares its 32-column chunks against its query's threshold and parks the rare
  // survivors (up to kStash per tile) in a private shared-memory stash.  The accumulator buffer is then
  // released to the MMA warp, and end() ISSUES one atomicAdd that reserves the survivors' slots in the
  // query's candidate list; the atomic's result is consumed only at the end of the thread's NEXT tile
  // (double-buffered stash), so its L2 round trip is hidden behind a whole tile of work.
  // Overflow path (dense early rounds): if some lane of the warp found more than kStash survivors, the warp
  // sweeps the tile a second time (pass 1) and those lanes append the excess synchronously.
  // Everything is force-inlined and State never has its address taken, so it lives in registers.
  static constexpr int kPasses = 2;
  static constexpr bool kPrefetch = false;
  static constexpr int kStash = 8;                      // survivors a thread can park per tile
  static constexpr int kEpiThreads = EPI_THREADS;       // 8 epilogue warps in the product
  __host__ __device__ static constexpr int smem_bytes(int) { return DENSE ? 0 : 2 * kStash * kEpiThreads * 8; }
  struct State {
    float t;
    int k, n, skip, pos2, tid, buf;
    unsigned long long* stash;     // [2][kStash][kEpiThreads], this thread owns column `tid` of buffer `buf`
    int p_n, p_pos, p_row, p_buf;  // reservation in flight: p_n keys of buffer p_buf go to row p_row at p_pos
  };
  __device__ __forceinline__ void bind(State& s, uint8_t* smem, int epi_tid) const {
    s.stash = reinterpret_cast<unsigned long long*>(smem);
    s.tid = epi_tid;
    s.buf = 0;
    s.p_n = 0;
  }
  __device__ __forceinline__ void finish(State& s) const { drain(s); }
  __device__ __forceinline__ void begin(State& s, int row, int, int) const {
    s.t = (row < nq && !DENSE) ? thr[row] : __int_as_float(0x7f800000);
    s.k = 0;
    s.n = 0;
    if constexpr (DENSE) {  // no stash / reservation in the dense round
      s.p_n = 0;
      s.buf = 0;
    }
  }
  __device__ __forceinline__ unsigned long long* slot(const State& s, int buf, int j) const {
    return s.stash + (static_cast<size_t>(buf) * kStash + j) * kEpiThreads + s.tid;
  }
  __device__ __forceinline__ void drain(State& s) const {
    if (s.p_n > 0) {
      unsigned long long* mine = cand + static_cast<size_t>(s.p_row) * C;
      for (int j = 0; j < s.p_n; ++j)
        if (s.p_pos + j < C) mine[s.p_pos + j] = *slot(s, s.p_buf, j);
      if (s.p_pos + s.p_n > C) *overflow = 1;
      s.p_n = 0;
    }
  }
  __device__ __forceinline__ bool need_pass(const State& s, int) const {
    if constexpr (DENSE) return false;
    return __any_sync(0xffffffffu, s.n > kStash);
  }
  __device__ __forceinline__ void between(State& s, int row) const {
    s.skip = kStash;
    s.pos2 = C;
    if (s.n > kStash) {
      s.pos2 = atomicAdd(count + row, s.n - kStash);
      if (s.pos2 + (s.n - kStash) > C) *overflow = 1;
    }
  }
  __device__ __forceinline__ void end(State& s, int row) const {
    if constexpr (DENSE) return;
    drain(s);  // last tile's survivors: their atomic was issued a whole tile ago
    if (s.k > 0) {
      s.p_pos = atomicAdd(count + row, s.k);  // result first used by the next drain()
      s.p_n = s.k;
      s.p_row = row;
      s.p_buf = s.buf;
      s.buf ^= 1;
    }
  }
  __device__ __forceinline__ void chunk(State& s, int row, int col0, const float (&v)[32], int pass) const {
    if constexpr (VARIANT == 3 && !DENSE) {
      chunk_coop(s, row, col0, v, pass);  // warp-collective: every lane stays in
      return;
    }
    if (row >= nq || col0 >= n_cols) return;
    unsigned long long* mine = cand + static_cast<size_t>(row) * C;
    if constexpr (DENSE) {
      if (col0 + 32 <= n_cols) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          ulonglong2 kk;
          kk.x = make_key(v[i], row_base + col0 + i);
          kk.y = make_key(v[i + 1], row_base + col0 + i + 1);
          *reinterpret_cast<ulonglong2*>(mine + col0 + i) = kk;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col0 + i < n_cols) mine[col0 + i] = make_key(v[i], row_base + col0 + i);
      }
      return;
    }
    const float t = s.t;
    const int lim = n_cols - col0;  // columns >= lim are out of range (only in the last tile)
    if (pass == 1 && s.n <= kStash) return;
    if constexpr (VARIANT != 0) {
      chunk_grouped(s, row, col0, v, pass, t, lim, mine);
      return;
    }
    float mx = v[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) mx = fmaxf(mx, v[i]);
    if (!(mx > t)) return;  // common case: nothing in this chunk beats the threshold
    // Survivor path.  Kept deliberately COMPACT (a bit mask + a short loop with a select tree instead of 32
    // unrolled predicated blocks): it is executed rarely per warp, so its instructions are cold in the
    // instruction cache and every extra cache line costs hundreds of cycles (measured ~1000 cycles per
    // survivor with the unrolled form).
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) mask |= (v[i] > t ? 1u : 0u) << i;
    if (lim < 32) mask &= (1u << lim) - 1u;
#pragma unroll 1
    while (mask) {
      const int i = __ffs(mask) - 1;
      mask &= mask - 1;
      const unsigned long long key = make_key(pick32(v, i), row_base + col0 + i);
      if (pass == 0) {
        if (s.k < kStash) {
          *slot(s, s.buf, s.k) = key;
          ++s.k;
        }
        ++s.n;
      } else if (s.skip > 0) {
        --s.skip;
      } else {
        if (s.pos2 < C) mine[s.pos2] = key;
        ++s.pos2;
      }
    }
  }
  // one survivor: column col (absolute within the round) with score val
  __device__ __forceinline__ void keep(State& s, float val, int col, int pass, unsigned long long* mine) const {
    const unsigned long long key = make_key(val, row_base + col);
    if (pass == 0) {
      if (s.k < kStash) {
        *slot(s, s.buf, s.k) = key;
        ++s.k;
      }
      ++s.n;
    } else if (s.skip > 0) {
      --s.skip;
    } else {
      if (s.pos2 < C) mine[s.pos2] = key;
      ++s.pos2;
    }
  }
  // w[j] for a run-time j in [0, 8): 3-level select tree (7 SEL)
  __device__ __forceinline__ static float pick8(const float (&w)[8], int j) {
    float a[4], b[2];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = (j & 1) ? w[2 * u + 1] : w[2 * u];
#pragma unroll
    for (int u = 0; u < 2; ++u) b[u] = (j & 2) ? a[2 * u + 1] : a[2 * u];
    return (j & 4) ? b[1] : b[0];
  }
  __device__ __forceinline__ void expand_group(State& s, const float (&w)[8], int col_g, int lim_g, float t, int pass,
                                               unsigned long long* mine) const {
    uint32_t m8 = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) m8 |= (w[j] > t ? 1u : 0u) << j;
    if (lim_g < 8) m8 &= lim_g > 0 ? (1u << lim_g) - 1u : 0u;
#pragma unroll 1
    while (m8) {
      const int j = __ffs(m8) - 1;
      m8 &= m8 - 1;
      keep(s, pick8(w, j), col_g + j, pass, mine);
    }
  }
  // VARIANT 1 / 2 (see the struct comment); visits survivors in increasing column order like variant 0
  __device__ __forceinline__ void chunk_grouped(State& s, int row, int col0, const float (&v)[32], int pass, float t,
                                                int lim, unsigned long long* mine) const {
    float g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float m1 = fmaxf(fmaxf(v[8 * q], v[8 * q + 1]), v[8 * q + 2]);
      const float m2 = fmaxf(fmaxf(v[8 * q + 3], v[8 * q + 4]), v[8 * q + 5]);
      g[q] = fmaxf(fmaxf(fmaxf(m1, m2), v[8 * q + 6]), v[8 * q + 7]);
    }
    const float mx = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
    if (!(mx > t)) return;  // common case: nothing in this chunk beats the threshold
    if constexpr (VARIANT == 2) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (g[q] > t) {
          float w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = v[8 * q + j];
          expand_group(s, w, col0 + 8 * q, lim - 8 * q, t, pass, mine);
        }
      }
    } else {
      uint32_t gm = (g[0] > t ? 1u : 0u) | (g[1] > t ? 2u : 0u) | (g[2] > t ? 4u : 0u) | (g[3] > t ? 8u : 0u);
#pragma unroll 1
      while (gm) {
        const int q = __ffs(gm) - 1;
        gm &= gm - 1;
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float lo = (q & 1) ? v[8 + j] : v[j], hi = (q & 1) ? v[24 + j] : v[16 + j];
          w[j] = (q & 2) ? hi : lo;
        }
        expand_group(s, w, col0 + 8 * q, lim - 8 * q, t, pass, mine);
      }
    }
  }
  // VARIANT 3 (see the struct comment).  Must be called by all 32 lanes of the warp (the GEMM epilogue loop is
  // warp-uniform).  Survivors are appended in increasing column order, exactly like variant 0.
  __device__ __forceinline__ void chunk_coop(State& s, int row, int col0, const float (&v)[32], int pass) const {
    const unsigned full = 0xffffffffu;
    const int lane = static_cast<int>(threadIdx.x & 31u);
    if (col0 >= n_cols) return;  // warp-uniform
    const bool active = row < nq && !(pass == 1 && s.n <= kStash);
    float mx = v[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) mx = fmaxf(mx, v[i]);
    unsigned hot = __ballot_sync(full, active && mx > s.t);
    if (hot == 0u) return;  // common case: no lane of the warp has a survivor in this chunk
    const int lim = n_cols - col0;  // columns >= lim are out of range (only in the last tile)
#pragma unroll 1
    while (hot) {
      const int L = __ffs(hot) - 1;  // the lane (query row) being expanded
      hot &= hot - 1;
      float x = 0.f;  // lane i receives column i of lane L's chunk
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float y = __shfl_sync(full, v[i], L);
        if (lane == i) x = y;
      }
      const float tL = __shfl_sync(full, s.t, L);
      const int rowL = __shfl_sync(full, row, L);
      const bool sv = x > tL && lane < lim;
      const unsigned m = __ballot_sync(full, sv);
      const int cnt = __popc(m), rank = __popc(m & ((1u << lane) - 1u));
      const unsigned long long key = make_key(x, row_base + col0 + lane);
      if (pass == 0) {
        const int kL = __shfl_sync(full, s.k, L), bufL = __shfl_sync(full, s.buf, L);
        const int j = kL + rank;
        if (sv && j < kStash)
          s.stash[(static_cast<size_t>(bufL) * kStash + j) * kEpiThreads + (s.tid - lane + L)] = key;
        if (lane == L) {
          s.k = kL + cnt < kStash ? kL + cnt : kStash;
          s.n += cnt;
        }
      } else {
        const int skipL = __shfl_sync(full, s.skip, L), posL = __shfl_sync(full, s.pos2, L);
        if (sv && rank >= skipL) {
          const int p = posL + (rank - skipL);
          if (p < C) cand[static_cast<size_t>(rowL) * C + p] = key;
        }
        if (lane == L) {
          const int used = cnt < skipL ? cnt : skipL;
          s.skip = skipL - used;
          s.pos2 = posL + (cnt - used);
        }
      }
    }
  }
  // v[i] for a run-time i without local memory: 5-level select tree (31 SEL)
  __device__ __forceinline__ static float pick32(const float (&v)[32], int i) {
    float a[16], b[8], c[4], d[2];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = (i & 1) ? v[2 * j + 1] : v[2 * j];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = (i & 2) ? a[2 * j + 1] : a[2 * j];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = (i & 4) ? b[2 * j + 1] : b[2 * j];
#pragma unroll
    for (int j = 0; j < 2; ++j) d[j] = (i & 8) ? c[2 * j + 1] : c[2 * j];
    return (i & 16) ? d[1] : d[0];
  }
};

}  // namespace om
