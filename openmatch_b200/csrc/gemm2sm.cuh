// EXPERIMENTAL (selftest only, not linked into libopenmatch_b200.so): 2-CTA (tcgen05 cta_group::2) variant of the
// GEMM core in gemm.cuh:   C[m, n] = sum_k A[m, k] * B[n, k].
// Measured on B200 (build/selftest_gemm --2sm): bit-exact, but ~770 TFLOP/s on every shape (the single-CTA core:
// 1 440 - 1 710), independent of ring depth and wait flavour - see DESIGN.md section 7.
//
// Idea: the single-CTA 128 x 256 tile moves (128 + 256) x 64 x 2 B of operands into shared memory per
// 2 x 128 x 256 x 64 FLOP (85 FLOP/B).  Here a CTA PAIR (cluster of 2, the two SMs of a TPC) owns a
// 256 x 256 tile: CTA r loads A rows [128 r, 128 r + 128) and B rows [128 r, 128 r + 128) of the tile, the
// leader (rank 0) issues tcgen05.mma.cta_group::2 with M = 256, N = 256, and each SM's tensor core reads the
// B half it does not hold from its peer's shared memory.  Per SM: 32 KB of operands per 4.2 MFLOP = 128 FLOP/B.
//
//   both CTAs, warp 0 lane 0   TMA producer : own A tile + own B half -> own smem ring; the transaction bytes
//                                             of BOTH CTAs complete on the LEADER's full barrier (.cta_group::2)
//   leader,    warp 1 lane 0   MMA issuer   : waits the leader's full barrier, issues the pair-wide MMAs;
//                                             tcgen05.commit ... multicast frees the smem slot / publishes the
//                                             accumulator in BOTH CTAs
//   both CTAs, warp 2          TMEM allocator (cta_group::2: same warp index in both CTAs)
//   both CTAs, warps 4..       epilogue     : own 128 accumulator rows (TMEM lanes) -> Epi functor; the
//                                             "accumulator drained" arrivals of both CTAs go to the leader
//
// Static persistent schedule over pairs (tile = pair + i * num_pairs); same Epi functor contract as gemm.cuh
// (kPasses == 1 functors).
#pragma once
#include <algorithm>

#include "gemm.cuh"

namespace om {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_count_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a pointer into this CTA's shared memory) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_rank(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr),
               "r"(bytes)
               : "memory");
}
// TMA load into THIS CTA's shared memory whose transaction bytes complete on an mbarrier of the pair's leader
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint32_t leader_bar_cluster_addr,
                                                int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the 9-operand form CUTLASS emits (explicit all-zero disable-output-lane mask)
__device__ __forceinline__ void umma_bf16_ss_2sm_masked(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                        uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// arrive (count 1) on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so
// far have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}

// SPIN = 1: poll with mbarrier.test_wait (never suspends the thread) and cluster-scope acquire.  Bring-up switch for
// the question "does a suspended try_wait wake up promptly when the completing arrival comes from the peer SM?"
__device__ __forceinline__ bool mbar_test_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
template <int SPIN>
__device__ __forceinline__ void wait2(uint64_t* bar, uint32_t parity, uint32_t site) {
  if constexpr (SPIN == 0) {
    mbar_wait(bar, parity, site);
  } else {
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_test_wait_cluster(bar, parity)) {
      if ((++spins & 4095u) != 0) continue;
      if (*reinterpret_cast<volatile unsigned int*>(&om_dev_fault) != 0u) return;
      if (clock64() - t0 > OM_WAIT_TIMEOUT_CYCLES) {
        atomicCAS(&om_dev_fault, 0u, (site << 16) | (blockIdx.x & 0xffffu) | 0x80000000u);
        return;
      }
    }
  }
}
template <int SPIN>
__device__ __forceinline__ void wait2_warp(uint64_t* bar, uint32_t parity, uint32_t site) {
  if constexpr (SPIN == 0) {
    mbar_wait_warp(bar, parity, site);
  } else {
    if ((threadIdx.x & 31u) == 0) {
      const long long t0 = clock64();
      uint32_t spins = 0;
      while (!mbar_test_wait_cluster(bar, parity)) {
        __nanosleep(32);
        if ((++spins & 4095u) != 0) continue;
        if (*reinterpret_cast<volatile unsigned int*>(&om_dev_fault) != 0u) break;
        if (clock64() - t0 > OM_WAIT_TIMEOUT_CYCLES) {
          atomicCAS(&om_dev_fault, 0u, (site << 16) | (blockIdx.x & 0xffffu) | 0x80000000u);
          break;
        }
      }
    }
    __syncwarp();
  }
}

template <int STAGES, int TILE_N = 256>
struct Gemm2Cfg {
  static_assert(TILE_N == 256 || TILE_N == 128, "pair tile N: 256 or 128");
  static constexpr int kTileM = 256, kTileN = TILE_N;       // per CTA pair
  static constexpr int kABytes = kBlockM * kBlockK * 2;       // this CTA's 128 A rows
  static constexpr int kBBytes = (kTileN / 2) * kBlockK * 2;  // this CTA's half of B
  static constexpr int kStageBytes = kABytes + kBBytes;       // 32 KB
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kEpiOffset = kBarOffset + 1024;
  static constexpr int kSmemBytes = kEpiOffset + 1024;
  static constexpr int kTmemCols = 2 * TILE_N;  // double-buffered accumulator columns per CTA
};

// SPIN: 0 suspending try_wait, 1 spinning test_wait.  TILE_N: pair tile width.  MASKED: 9-operand MMA form.
// MODE (rate probes, results are garbage): 1 = MMA only (no TMA, the issuer never waits for operands),
// 2 = loads only (the issuer waits and commits but issues no MMA).
// RELAY: the peer's TMA loads complete on a barrier in the peer's OWN shared memory and one relay thread forwards a
// single arrival per stage to the leader (instead of every complete_tx of the peer's TMA crossing to the leader SM).
// TRACE (measurement only): pair 0 records %globaltimer per k-block: [rank][role][kb], roles: 0 producer passed the
// empty wait, 1 producer issued its loads, 2 MMA thread passed the full wait, 3 MMA thread issued the commit,
// 4 relay saw the local stage land
__device__ unsigned long long om_2sm_trace[2][5][64];
__device__ __forceinline__ void trace2(bool on, uint32_t rank, int role, int i) {
  if (on && i < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    om_2sm_trace[rank][role][i] = t;
  }
}
template <int STAGES, bool M_FASTEST, int EPI_WARPS, class Epi, int SPIN = 0, int TILE_N = 256, bool MASKED = false,
          int MODE = 0, bool RELAY = false, bool TRACE = false>
__global__ void __launch_bounds__(kGemmProducerThreads + 32 * EPI_WARPS, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                     int K, const __grid_constant__ Epi epi) {
  using Cfg = Gemm2Cfg<STAGES, TILE_N>;
  static_assert(Epi::kPasses == 1, "2-CTA core supports single-pass epilogues");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOffset);  // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;                                    // one per CTA (multicast commit)
  uint64_t* tfull_bar = empty_bar + STAGES;                                   // one per CTA (multicast commit)
  uint64_t* tempty_bar = tfull_bar + 2;                                       // used in the leader only
  uint64_t* lfull_bar = tempty_bar + 2;                                       // RELAY: peer-local "my loads landed"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lfull_bar + STAGES);
  uint8_t* epi_smem = smem + Cfg::kEpiOffset;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(threadIdx.x & 31);
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = static_cast<int>(cluster_id_x());
  const int num_pairs = static_cast<int>(cluster_count_x());

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      // RELAY: the leader's arrive.expect_tx + the relay's arrival; otherwise ONE arrive.expect_tx by the leader that
      // covers the bytes of both CTAs (the peer only issues its loads: a remote mbarrier arrive with cluster-scope
      // release stalls the issuing thread for ~0.6 - 0.8 us, measured, which throttled the peer to one stage per that)
      mbar_init(&full_bar[i], RELAY ? 2 : 1);
      mbar_init(&empty_bar[i], 1);
      mbar_init(&lfull_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * EPI_WARPS);  // every epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before_sync();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before anyone signals across the pair
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int num_m = (M + Cfg::kTileM - 1) / Cfg::kTileM;
  const int num_n = (N + Cfg::kTileN - 1) / Cfg::kTileN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + kBlockK - 1) / kBlockK;

  if (warp == 0) {
    if (lane == 0 && MODE != 1) {
      // ------------------------------ TMA producer (both CTAs) ------------------------------
      uint32_t stage = 0, phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = M_FASTEST ? tile % num_m : tile / num_n;
        const int n_blk = M_FASTEST ? tile / num_m : tile % num_n;
        const int row_a = m_blk * Cfg::kTileM + static_cast<int>(rank) * kBlockM;
        const int row_b = n_blk * Cfg::kTileN + static_cast<int>(rank) * (Cfg::kTileN / 2);
        for (int kb = 0; kb < num_k; ++kb) {
          wait2<SPIN>(&empty_bar[stage], phase ^ 1u, 1);
          trace2(TRACE && tile == 0, rank, 0, kb);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          if constexpr (RELAY) {
            uint64_t* bar = rank == 0 ? &full_bar[stage] : &lfull_bar[stage];
            mbar_arrive_expect_tx(bar, Cfg::kStageBytes);
            tma_load_2d(sa, &tmA, bar, kb * kBlockK, row_a);
            tma_load_2d(sa + Cfg::kABytes, &tmB, bar, kb * kBlockK, row_b);
          } else {
            const uint32_t leader_full = mapa_rank(&full_bar[stage], 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            tma_load_2d_2sm(sa, &tmA, leader_full, kb * kBlockK, row_a);
            tma_load_2d_2sm(sa + Cfg::kABytes, &tmB, leader_full, kb * kBlockK, row_b);
          }
          trace2(TRACE && tile == 0, rank, 1, kb);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ------------------------------ MMA issuer (leader only) ------------------------------
      constexpr uint32_t idesc = umma_idesc_bf16(Cfg::kTileM, Cfg::kTileN);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        wait2<SPIN>(&tempty_bar[as], aphase ^ 1u, 2);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * Cfg::kTileN;
        for (int kb = 0; kb < num_k; ++kb) {
          if constexpr (MODE != 1) wait2<SPIN>(&full_bar[stage], phase, 3);
          tc_fence_after_sync();
          trace2(TRACE && tile == 0, 0, 2, kb);
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * kUmmaK * 2, kDescKMajorSW128);
            const uint64_t db = umma_smem_desc(b_addr + k * kUmmaK * 2, kDescKMajorSW128);
            if constexpr (MODE == 2) {
              (void)da;
              (void)db;
            } else if constexpr (MASKED) {
              umma_bf16_ss_2sm_masked(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            } else {
              umma_bf16_ss_2sm(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit_2sm(&empty_bar[stage]);  // frees this stage in both CTAs
          trace2(TRACE && tile == 0, 0, 3, kb);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_2sm(&tfull_bar[as]);  // accumulator complete -> both epilogues
      }
    }
  } else if (warp == 3) {
    if (RELAY && lane == 0 && rank == 1 && MODE != 1) {
      // ------------------------------ relay (peer only): stage landed here -> one arrival at the leader ----------
      uint32_t stage = 0, phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs)
        for (int kb = 0; kb < num_k; ++kb) {
          wait2<SPIN>(&lfull_bar[stage], phase, 5);
          trace2(TRACE && tile == 0, 1, 4, kb);
          mbar_arrive_cluster(mapa_rank(&full_bar[stage], 0));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue (both CTAs, own 128 rows) ------------------------------
    static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "EPI_WARPS: 4 or 8");
    const int ew = (warp - 4) & 3;
    const int half = (warp - 4) >> 2;
    constexpr int kChunks = Cfg::kTileN / 32 / (EPI_WARPS / 4);
    int it = 0;
    typename Epi::State st;
    if constexpr (Epi::smem_bytes(EPI_WARPS) > 0)
      epi.bind(st, epi_smem, static_cast<int>(threadIdx.x) - kGemmProducerThreads);
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int m_blk = M_FASTEST ? tile % num_m : tile / num_n;
      const int n_blk = M_FASTEST ? tile / num_m : tile % num_n;
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const int m_blk128 = m_blk * 2 + static_cast<int>(rank);  // in units of 128 rows, as the functors expect
      const int row = m_blk128 * kBlockM + ew * 32 + lane;
      const int col_base = n_blk * Cfg::kTileN;
      epi.begin(st, row, m_blk128, n_blk);
      if constexpr (Epi::kPrefetch) epi.prefetch(st, row, col_base + half * kChunks * 32);
      wait2_warp<SPIN>(&tfull_bar[as], aphase, 4);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + as * Cfg::kTileN + (static_cast<uint32_t>(ew * 32) << 16);
      const int c0 = half * kChunks;
      auto run = [&](const uint32_t (&rb)[32], int c, bool has_next) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rb[i]);
        if constexpr (Epi::kPrefetch)
          epi.chunk(st, row, col_base + c * 32, v, has_next ? col_base + (c + 1) * 32 : -1);
        else
          epi.chunk(st, row, col_base + c * 32, v);
      };
      uint32_t ra[32], rb[32];
      tmem_ld_32x32b_x32(taddr + c0 * 32, ra);
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
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0)
          mbar_arrive(&tempty_bar[as]);
        else
          mbar_arrive_cluster(mapa_rank(&tempty_bar[as], 0));
      }
      epi.end(st, row);
    }
    if constexpr (Epi::smem_bytes(EPI_WARPS) > 0) epi.finish(st);
  }

  tc_fence_before_sync();
  cluster_sync_all();  // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

// Which SMs host the two CTAs of each cluster?  out[2 * cluster + rank] = %smid.  (cta_group::2 needs the two SMs of
// one TPC; a cluster of 2 is the only placement control there is.)
__global__ void cluster_smid_kernel(unsigned* out) {
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    out[2 * cluster_id_x() + cluster_ctarank()] = smid;
  }
}

// Host launcher (cluster of 2 CTAs along x).  A: [M, K] bf16 row pitch lda; B: [N, K] bf16 row pitch ldb.
template <int STAGES, bool M_FASTEST, int EPI_WARPS, int SPIN = 0, int TILE_N = 256, bool MASKED = false, int MODE = 0,
          bool RELAY = false, bool TRACE = false, class Epi>
static inline cudaError_t launch_gemm2(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K,
                                       const Epi& epi, int num_sms, cudaStream_t stream, int* pairs_out = nullptr) {
  using Cfg = Gemm2Cfg<STAGES, TILE_N>;
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  CUtensorMap tmA, tmB;
  if (make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, kBlockK, kBlockM) != 0)
    return cudaErrorInvalidValue;
  if (make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, kBlockK, Cfg::kTileN / 2) != 0)
    return cudaErrorInvalidValue;
  auto kern = gemm2_bf16_tn_kernel<STAGES, M_FASTEST, EPI_WARPS, Epi, SPIN, TILE_N, MASKED, MODE, RELAY, TRACE>;
  const int smem_bytes = Cfg::kSmemBytes + Epi::smem_bytes(EPI_WARPS);
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int num_tiles = ((M + Cfg::kTileM - 1) / Cfg::kTileM) * ((N + Cfg::kTileN - 1) / Cfg::kTileN);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kGemmProducerThreads + 32 * EPI_WARPS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // The persistent schedule needs every cluster co-resident: a GPC with an odd SM count leaves one SM without a
  // partner, so fewer than num_sms / 2 pairs fit (a second wave of clusters would double the run time).
  static int max_pairs = 0;  // per instantiation
  if (!max_pairs) {
    cfg.gridDim = dim3(2 * (num_sms / 2));
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess) return e;
    max_pairs = std::max(1, std::min(n, num_sms / 2));
  }
  const int pairs = std::max(1, std::min(num_tiles, max_pairs));
  cfg.gridDim = dim3(2 * pairs);
  if (pairs_out) *pairs_out = pairs;
  return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, M, N, K, epi);
}

}  // namespace om
