// Persistent, warp-specialised tcgen05 GEMM core for sm_100a:   C[m, n] = sum_k A[m, k] * B[n, k]
// (both operands K-major bf16, fp32 accumulation in tensor memory).  This one mainloop serves the
// three hot steps: encoder linear layers (A = activations [T, H], B = nn.Linear weight [out, in]),
// brute-force search (A = queries [nq, d], B = corpus rows [N, d]) and contrastive logits (Q * P^T).
//
//   warp 0 (one elected lane)  TMA producer : global -> STAGES-deep smem ring (128B-swizzled boxes)
//   warp 1 (one elected lane)  MMA issuer   : tcgen05.mma 128 x BN x 16, accumulators in TMEM,
//                                             double-buffered (2 x BN columns)
//   warp 2                     TMEM allocator
//   warps 4..4+EPI_WARPS       epilogue     : tcgen05.ld -> registers -> Epi functor (fused op).
//                                             EPI_WARPS = 4: one thread per accumulator row; 8 / 16: two / four
//                                             threads per row, each owning a column group of the tile (math-
//                                             heavy epilogues such as bias + erf-GELU need the extra warps to
//                                             hide TMEM-load and MUFU latency)
//
// Pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue); static persistent
// tile schedule (tile = blockIdx.x + i * gridDim.x).
#pragma once
#include <type_traits>

#include "ptx.cuh"
#include "tmap.cuh"

namespace om {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmProducerThreads = 128;  // warps 0..3: TMA, MMA, TMEM alloc, idle

template <int BN, int STAGES>
struct GemmCfg {
  static_assert(BN == 64 || BN == 128 || BN == 192 || BN == 256, "BN must be 64, 128, 192 or 256");
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kEpiOffset = kBarOffset + 1024;         // epilogue scratch starts 1024-B aligned (TMA store)
  static constexpr int kSmemBytes = kEpiOffset + 1024;         // + slack for 1024-B alignment of the base
  static constexpr int kTmemCols = BN <= 64 ? 128 : (BN <= 128 ? 256 : 512);  // power of two >= 2 * BN
};

// Epilogue functor contract (all methods __device__, called by the epilogue threads; thread <-> row (half)):
//   struct State;                                             per-thread, per-tile scratch
//   void begin(State&, int row, int m_blk, int n_blk) const;  once per tile
//   void chunk(State&, int row, int col0, const float (&v)[32]) const;   v = C[row, col0 .. col0+31]
//   void end(State&, int row) const;                           once per tile
//   static constexpr bool kPrefetch = false;                   true: prefetch(State&, row, col0) is called for the
//        thread's first chunk BEFORE waiting on the accumulator, and chunk(..., int next_col0) receives the
//        column of the thread's next chunk (-1: none) so that global operands (residual) are always one
//        chunk ahead of the math
//   __host__ __device__ static constexpr int smem_bytes(int epi_warps);            > 0: that much dynamic smem is reserved for the functor
//        and handed over through bind(State&, uint8_t* smem, int epilogue_thread_index) once per thread; the
//        State object persists across the thread's tiles and finish(State&) is called after the last one
//   end() runs AFTER the thread's warp has released the accumulator buffer: long-latency tails (atomics,
//        global stores) placed there overlap the next tile's MMAs
//   static constexpr int kPasses = 1;                          2: the accumulator tile may be read twice:
//        chunk(..., int pass) runs for pass 0; if need_pass(State&, 1) (warp-uniform) is true, between(State&,
//        row) and a second sweep with pass 1 follow (TMEM re-reads are cheap; the search filter uses this as
//        its overflow path when a thread finds more survivors than its stash holds)
// Rows >= M and columns >= N contain zeros (TMA out-of-bounds fill) and must be masked by the functor.

// Epi::kRolled (optional, default false): see the epilogue loop
template <class Epi, class = void>
struct epi_rolled : std::false_type {};
template <class Epi>
struct epi_rolled<Epi, std::void_t<decltype(Epi::kRolled)>> : std::bool_constant<Epi::kRolled> {};

template <int BN, int STAGES, bool M_FASTEST, int EPI_WARPS, class Epi, bool F16 = false>
__global__ void __launch_bounds__(kGemmProducerThreads + 32 * EPI_WARPS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                    int K, const __grid_constant__ Epi epi, int* tile_counter) {
  using Cfg = GemmCfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* epi_smem = smem + Cfg::kEpiOffset;  // Epi::smem_bytes() of scratch owned by the epilogue functor
  // dynamic tile scheduler (tile_counter != nullptr): the producer claims tile indices from a global counter
  // and publishes them to the MMA and epilogue roles through a 4-deep smem ring
  constexpr int kSched = 4;
  uint64_t* sfull_bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);
  uint64_t* sempty_bar = sfull_bar + kSched;
  volatile int* tile_ring = reinterpret_cast<volatile int*>(sempty_bar + kSched);
  const bool dyn = tile_counter != nullptr;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(threadIdx.x & 31);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_WARPS);  // one arrive per epilogue warp
    }
    for (int i = 0; i < kSched; ++i) {
      mbar_init(&sfull_bar[i], 1);
      mbar_init(&sempty_bar[i], 1 + EPI_WARPS);  // MMA thread + one lane per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int num_m = (M + kBlockM - 1) / kBlockM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + kBlockK - 1) / kBlockK;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------ TMA producer ------------------------------
      uint32_t stage = 0, phase = 0, sslot = 0, sphase = 0;
      int tile = blockIdx.x;
      if (dyn) {
        tile = atomicAdd(tile_counter, 1);
        if (tile >= num_tiles) tile = -1;
      }
      while (true) {
        if (dyn) {  // publish (also the -1 sentinel) to the consumer roles
          mbar_wait(&sempty_bar[sslot], sphase ^ 1u, 5);
          tile_ring[sslot] = tile;
          mbar_arrive(&sfull_bar[sslot]);
          if (++sslot == kSched) {
            sslot = 0;
            sphase ^= 1u;
          }
          if (tile < 0) break;
        } else if (tile >= num_tiles) {
          break;
        }
        // claim the next tile now; the atomic's round trip overlaps this tile's loads
        int next = dyn ? atomicAdd(tile_counter, 1) : tile + static_cast<int>(gridDim.x);
        const int m_blk = M_FASTEST ? tile % num_m : tile / num_n;
        const int n_blk = M_FASTEST ? tile / num_m : tile % num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * kBlockK, m_blk * kBlockM);
          tma_load_2d(sa + Cfg::kABytes, &tmB, &full_bar[stage], kb * kBlockK, n_blk * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tile = (dyn && next >= num_tiles) ? -1 : next;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------ MMA issuer ------------------------------
      // F16: both operands IEEE half instead of bf16 (same instruction, same rate, 3 more significand bits)
      constexpr uint32_t idesc = F16 ? umma_idesc_f16(kBlockM, BN) : umma_idesc_bf16(kBlockM, BN);
      uint32_t stage = 0, phase = 0, sslot = 0, sphase = 0;
      int it = 0;
      for (int tile = blockIdx.x;; tile += gridDim.x, ++it) {
        if (dyn) {
          mbar_wait(&sfull_bar[sslot], sphase, 6);
          const int t = tile_ring[sslot];
          mbar_arrive(&sempty_bar[sslot]);
          if (++sslot == kSched) {
            sslot = 0;
            sphase ^= 1u;
          }
          if (t < 0) break;
        } else if (tile >= num_tiles) {
          break;
        }
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1u, 2);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * kUmmaK * 2, kDescKMajorSW128);
            const uint64_t db = umma_smem_desc(b_addr + k * kUmmaK * 2, kDescKMajorSW128);
            umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    static_assert(EPI_WARPS == 4 || EPI_WARPS == 8 || EPI_WARPS == 12 || EPI_WARPS == 16, "EPI_WARPS: 4, 8, 12 or 16");
    const int ew = (warp - 4) & 3;      // == warp % 4: the TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;   // column group (of EPI_WARPS / 4) owned by this warp
    static_assert((BN / 32) % (EPI_WARPS / 4) == 0, "column chunks must split evenly over the epilogue warps");
    constexpr int kChunks = BN / 32 / (EPI_WARPS / 4);
    int it = 0;
    typename Epi::State st;  // lives across tiles: functors may keep work in flight from one tile to the next
    if constexpr (Epi::smem_bytes(EPI_WARPS) > 0)
      epi.bind(st, epi_smem, static_cast<int>(threadIdx.x) - kGemmProducerThreads);
    uint32_t sslot = 0, sphase = 0;
    for (int tile = blockIdx.x;; tile += gridDim.x, ++it) {
      if (dyn) {
        mbar_wait_warp(&sfull_bar[sslot], sphase, 7);
        tile = tile_ring[sslot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&sempty_bar[sslot]);
        if (++sslot == kSched) {
          sslot = 0;
          sphase ^= 1u;
        }
        if (tile < 0) break;
      } else if (tile >= num_tiles) {
        break;
      }
      const int m_blk = M_FASTEST ? tile % num_m : tile / num_n;
      const int n_blk = M_FASTEST ? tile / num_m : tile % num_n;
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const int row = m_blk * kBlockM + ew * 32 + lane;
      epi.begin(st, row, m_blk, n_blk);
      if constexpr (Epi::kPrefetch) epi.prefetch(st, row, n_blk * BN + half * kChunks * 32);  // before the wait
      mbar_wait_warp(&tfull_bar[as], aphase, 4);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
      for (int pass = 0; pass < Epi::kPasses; ++pass) {
        if constexpr (Epi::kPasses > 1) {
          if (pass > 0) {
            if (!epi.need_pass(st, pass)) break;  // warp-uniform decision
            epi.between(st, row);
          }
        }
        // Software-pipelined TMEM reads: the load of chunk c+1 is in flight while chunk c is processed.  The
        // loop stays ROLLED over chunk pairs (two register buffers): fully unrolling it made the scan kernel
        // 30 K SASS instructions and instruction-fetch bound (2.3x slower).
        const int c0 = half * kChunks;
        auto run = [&](const uint32_t (&rb)[32], int c, bool has_next) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rb[i]);
          if constexpr (Epi::kPasses > 1)
            epi.chunk(st, row, n_blk * BN + c * 32, v, pass);
          else if constexpr (Epi::kPrefetch)
            epi.chunk(st, row, n_blk * BN + c * 32, v, has_next ? n_blk * BN + (c + 1) * 32 : -1);
          else
            epi.chunk(st, row, n_blk * BN + c * 32, v);
        };
        if constexpr (epi_rolled<Epi>::value) {
          // heavy epilogues (hundreds of instructions per chunk): ONE copy of the chunk code, chunks processed in a rolled
          // loop; the exposed TMEM-load latency (~100 cycles per chunk) is noise next to the chunk's own work, while
          // three inlined copies (double-buffered loads + tail) made the kernel 4.7 K instructions and fetch-bound
          uint32_t ra[32];
#pragma unroll 1
          for (int c = c0; c < c0 + kChunks; ++c) {
            tmem_ld_32x32b_x32(taddr + c * 32, ra);
            tmem_ld_wait();
            run(ra, c, c + 1 < c0 + kChunks);
          }
          continue;
        }
        uint32_t ra[32], rb[32];
        tmem_ld_32x32b_x32(taddr + c0 * 32, ra);
        if constexpr (kChunks == 1) {
          tmem_ld_wait();
          run(ra, c0, false);
        } else {
#pragma unroll 1
          for (int c = c0; c + 1 < c0 + kChunks; c += 2) {
            tmem_ld_wait();
            tmem_ld_32x32b_x32(taddr + (c + 1) * 32, rb);
            run(ra, c, true);
            tmem_ld_wait();
            const bool more = c + 2 < c0 + kChunks;
            if (more) tmem_ld_32x32b_x32(taddr + (c + 2) * 32, ra);
            run(rb, c + 1, more);
          }
          if constexpr (kChunks % 2 == 1) {  // odd tail (e.g. BN = 192 with 8 epilogue warps: 3 chunks each)
            tmem_ld_wait();
            run(ra, c0 + kChunks - 1, false);
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      epi.end(st, row);
    }
    if constexpr (Epi::smem_bytes(EPI_WARPS) > 0) epi.finish(st);  // drain whatever the functor still has in flight
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// Host launcher.  A: [M, K] bf16 row pitch lda elements; B: [N, K] bf16 row pitch ldb elements.
// Returns cudaSuccess / a CUDA error; tensor-map failures map to cudaErrorInvalidValue.
// Pool of tile counters for the dynamic scheduler: launch i uses counter i % 64 (zeroed on the launching stream
// just before the kernel), so up to 64 dynamically scheduled GEMMs may be in flight.
static inline int* next_tile_counter(cudaStream_t stream) {
  static int* pool = nullptr;
  static unsigned next = 0;
  if (!pool && cudaMalloc(&pool, 64 * sizeof(int)) != cudaSuccess) return nullptr;
  int* c = pool + (next++ & 63u);
  if (cudaMemsetAsync(c, 0, sizeof(int), stream) != cudaSuccess) return nullptr;
  return c;
}

// dynamic_sched: claim tiles from a global counter instead of the static blockIdx + i * gridDim order.  With
// the static order CTAs drift apart over thousands of tiles and stop sharing operand tiles in L2 (measured on
// the search scan: 143 GB of DRAM reads for a 6.4 GB shard); the dynamic order keeps all CTAs on neighbouring
// tiles.
template <int BN, int STAGES, bool M_FASTEST, int EPI_WARPS, class Epi, bool F16 = false>
static inline cudaError_t launch_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K,
                                      const Epi& epi, int num_sms, cudaStream_t stream, bool dynamic_sched = false) {
  using Cfg = GemmCfg<BN, STAGES>;
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  CUtensorMap tmA, tmB;
  if (make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, kBlockK, kBlockM) != 0)
    return cudaErrorInvalidValue;
  if (make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, kBlockK, BN) != 0)
    return cudaErrorInvalidValue;
  auto kern = gemm_bf16_tn_kernel<BN, STAGES, M_FASTEST, EPI_WARPS, Epi, F16>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes + Epi::smem_bytes(EPI_WARPS));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int num_tiles = ((M + kBlockM - 1) / kBlockM) * ((N + BN - 1) / BN);
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  int* counter = nullptr;
  if (dynamic_sched) {
    counter = next_tile_counter(stream);
    if (!counter) return cudaErrorMemoryAllocation;
  }
  kern<<<grid, kGemmProducerThreads + 32 * EPI_WARPS, Cfg::kSmemBytes + Epi::smem_bytes(EPI_WARPS), stream>>>(tmA, tmB, M, N, K,
                                                                                                     epi, counter);
  return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------------
// Generic store epilogues
// --------------------------------------------------------------------------------------------------
struct EpiStoreF32 {  // C fp32 = acc (+ bias[n]) (+ resid[m, n])
  float* C;
  int64_t ldc;
  const float* bias;   // nullable, [N]
  const float* resid;  // nullable, [M, ldr]; may alias C (in-place residual add)
  int64_t ldr;
  int M, N;
  static constexpr int kPasses = 1;
  static constexpr bool kPrefetch = true;
  __host__ __device__ static constexpr int smem_bytes(int) { return 0; }
  struct State {
    float4 pre[8];  // residual of the chunk about to be processed
  };
  __device__ __forceinline__ void begin(State&, int, int, int) const {}
  __device__ __forceinline__ void end(State&, int) const {}
  __device__ __forceinline__ bool fast(int row, int col0) const {
    return row < M && col0 + 32 <= N && (((ldc | ldr) & 3) == 0);
  }
  __device__ __forceinline__ void prefetch(State& s, int row, int col0) const {
    if (resid && col0 >= 0 && fast(row, col0)) {
      const float4* r = reinterpret_cast<const float4*>(resid + (int64_t)row * ldr + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) s.pre[j] = r[j];
    }
  }
  __device__ __forceinline__ void chunk(State& s, int row, int col0, const float (&v)[32], int next_col0) const {
    if (row >= M || col0 >= N) return;
    float* out = C + (int64_t)row * ldc + col0;
    if (fast(row, col0)) {
      float4 cur[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = resid ? s.pre[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      // all loads of the NEXT chunk are issued before any store of this one: with resid aliasing C the
      // compiler must otherwise order every load behind the previous store (one HBM round trip each)
      prefetch(s, row, next_col0);
      float4 b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        b[j] = bias ? __ldg(reinterpret_cast<const float4*>(bias + col0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 o;
        o.x = v[4 * j] + b[j].x + cur[j].x;
        o.y = v[4 * j + 1] + b[j].y + cur[j].y;
        o.z = v[4 * j + 2] + b[j].z + cur[j].z;
        o.w = v[4 * j + 3] + b[j].w + cur[j].w;
        reinterpret_cast<float4*>(out)[j] = o;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < N) {
          float o = v[i];
          if (bias) o += bias[col0 + i];
          if (resid) o += resid[(int64_t)row * ldr + col0 + i];
          out[i] = o;
        }
      prefetch(s, row, next_col0);
    }
  }
};

}  // namespace om
