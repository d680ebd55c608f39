// 2-CTA (tcgen05 cta_group::2) variant of the GEMM core in gemm.cuh:   C[m, n] = sum_k A[m, k] * B[n, k].
//
// The single-CTA 128 x 256 tile moves (128 + 256) x 64 x 2 B = 48 KB of operands into an SM per 4.2 MFLOP, and the
// SM's ingest path (~64 B/clk) is what bounds that mainloop.  Here a CTA PAIR (cluster of 2 = the two SMs of a TPC)
// owns a 256 x 256 tile: CTA r loads A rows [128 r, +128) and B rows [128 r, +128) of the tile, the leader (rank 0)
// issues tcgen05.mma.cta_group::2 with M = 256, N = 256, and each SM's tensor core reads the B half it does not hold
// from its peer's shared memory.  Per SM: 32 KB of operands per 4.2 MFLOP - the mainloop becomes MMA-bound.
// Measured (build/selftest_gemm --2sm, profiles/r02_2sm_leadertx.log): 1 711 TFLOP/s on the encoder FFN1 shape
// (single-CTA core 1 600), 1 614 on 8192^3 (1 373), 1 515 on the sustained search sweep (1 437); bit-exact.
//
//   both CTAs, warp 0 lane 0   TMA producer : own A tile + own B half -> own smem ring; the transaction bytes of
//                                             BOTH CTAs complete on the LEADER's full barrier (.cta_group::2), armed
//                                             by ONE arrive.expect_tx of the leader for the bytes of both.  (A remote
//                                             mbarrier arrive from the peer - the first version of this file - stalls
//                                             the issuing thread 0.6 - 0.8 us and throttled the pair to 770 TFLOP/s:
//                                             profiles/r02_2sm_trace_remote_arrive.log.)
//   leader,    warp 1 lane 0   MMA issuer   : waits the leader's full barrier, issues the pair-wide MMAs;
//                                             tcgen05.commit ... multicast frees the smem slot / publishes the
//                                             accumulator in BOTH CTAs
//   both CTAs, warp 2          TMEM allocator (cta_group::2: same warp index in both CTAs)
//   both CTAs, warps 4..       epilogue     : own 128 accumulator rows (TMEM lanes) -> Epi functor (same contract as
//                                             gemm.cuh, including multi-pass functors); the "accumulator drained"
//                                             arrivals of both CTAs go to the leader
//
// Tile schedule: static (tile = pair + i * num_pairs) or dynamic (tile_counter != nullptr): the leader's producer
// claims pair tiles from a global counter, pushes each claim into the peer's shared memory with one 8-byte remote
// atomic max (sequence number | tile; the peer's producer polls its own shared memory), and both producers publish the tile
// to the roles of their CTA through the same 4-deep ring as gemm.cuh.
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
// Relaxed: the only user is "accumulator drained" (the tcgen05.ld results are already in registers, ordered by
// tcgen05.wait::ld + fence::before_thread_sync); a cluster-scope RELEASE arrive stalls the issuing thread for
// 0.6 - 0.8 us (profiles/r02_2sm_trace_remote_arrive.log).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// fire-and-forget remote reduction (no result: the issuing thread does not wait for the peer SM)
__device__ __forceinline__ void red_max_cluster_u64(uint32_t cluster_addr, unsigned long long v) {
  asm volatile("red.relaxed.cluster.shared::cluster.max.u64 [%0], %1;" ::"r"(cluster_addr), "l"(v) : "memory");
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
__device__ __forceinline__ void umma_f16_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
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
// arrive (count 1) on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so
// far have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}

template <int STAGES>
struct Gemm2Cfg {
  static constexpr int kTileM = 256, kTileN = 256;            // per CTA pair
  static constexpr int kABytes = kBlockM * kBlockK * 2;       // this CTA's 128 A rows
  static constexpr int kBBytes = (kTileN / 2) * kBlockK * 2;  // this CTA's half of B
  static constexpr int kStageBytes = kABytes + kBBytes;       // 32 KB
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kEpiOffset = kBarOffset + 1024;
  static constexpr int kSmemBytes = kEpiOffset + 1024;
  static constexpr int kTmemCols = 2 * kTileN;  // double-buffered accumulator columns per CTA
};

// MODE (rate probes of the selftest, results are garbage): 1 = MMA only (no TMA, the issuer never waits for
// operands), 2 = loads only (the issuer waits and commits but issues no MMA).  0 in the product.
template <int STAGES, bool M_FASTEST, int EPI_WARPS, class Epi, bool F16 = false, int MODE = 0>
__global__ void __launch_bounds__(kGemmProducerThreads + 32 * EPI_WARPS, 1)
gemm2_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                const __grid_constant__ Epi epi, int* tile_counter) {
  using Cfg = Gemm2Cfg<STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOffset);  // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;                                    // one per CTA (multicast commit)
  uint64_t* tfull_bar = empty_bar + STAGES;                                   // one per CTA (multicast commit)
  uint64_t* tempty_bar = tfull_bar + 2;                                       // used in the leader only
  constexpr int kSched = 4, kPush = 8;
  uint64_t* sfull_bar = tempty_bar + 2;  // tile ring of this CTA: producer -> MMA thread (leader) + epilogue warps
  uint64_t* sempty_bar = sfull_bar + kSched;
  volatile unsigned long long* push_ring = reinterpret_cast<volatile unsigned long long*>(sempty_bar + kSched);  // leader -> peer
  volatile int* tile_ring = reinterpret_cast<volatile int*>(push_ring + kPush);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_cast<int*>(tile_ring) + kSched);
  uint8_t* epi_smem = smem + Cfg::kEpiOffset;
  static_assert((2 * STAGES + 4 + 2 * kSched + kPush) * 8 + kSched * 4 + 8 <= 1024, "barrier block overflows its 1 KB");

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(threadIdx.x & 31);
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = static_cast<int>(cluster_id_x());
  const int num_pairs = static_cast<int>(cluster_count_x());
  const bool dyn = tile_counter != nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);  // ONE arrive.expect_tx by the leader covering the bytes of both CTAs
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * EPI_WARPS);  // every epilogue warp of both CTAs
    }
    for (int i = 0; i < kSched; ++i) {
      mbar_init(&sfull_bar[i], 1);
      mbar_init(&sempty_bar[i], (rank == 0 ? 1 : 0) + EPI_WARPS);  // MMA thread (leader) + one lane per epilogue warp
    }
    for (int i = 0; i < kPush; ++i) push_ring[i] = 0ull;
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before_sync();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before anyone signals across the pair
  __syncthreads();     // (redundant with the cluster barrier; it is the one compute-sanitizer's racecheck models)
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int num_m = (M + Cfg::kTileM - 1) / Cfg::kTileM;
  const int num_n = (N + Cfg::kTileN - 1) / Cfg::kTileN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + kBlockK - 1) / kBlockK;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------ TMA producer (both CTAs) ------------------------------
      uint32_t stage = 0, phase = 0, sslot = 0, sphase = 0;
      unsigned seq = 1;  // claims so far + 1 (0 = "nothing pushed yet" in the peer's ring)
      int tile = pair, next = 0;
      if (dyn && rank == 0) {
        tile = atomicAdd(tile_counter, 1);
        if (tile >= num_tiles) tile = -1;
      }
      while (true) {
        if (dyn) {
          if (rank == 0) {
            // push the claim to the peer: one 8-byte remote atomic max (sequence | tile only grows in a slot), nobody
            // waits for it here
            red_max_cluster_u64(mapa_rank(const_cast<unsigned long long*>(&push_ring[seq % kPush]), 1),
                                (static_cast<unsigned long long>(seq) << 32) | static_cast<uint32_t>(tile));
          } else {  // the peer polls its own shared memory (atomically) for claim number `seq`
            const long long t0 = clock64();
            unsigned long long v;
            while (static_cast<unsigned>((v = atomicMax(const_cast<unsigned long long*>(&push_ring[seq % kPush]), 0ull)) >> 32) != seq) {
              if (clock64() - t0 > OM_WAIT_TIMEOUT_CYCLES) {
                atomicCAS(&om_dev_fault, 0u, (9u << 16) | (blockIdx.x & 0xffffu) | 0x80000000u);
                v = 0xffffffffull;  // -1: stop
                break;
              }
            }
            tile = static_cast<int>(static_cast<uint32_t>(v));
          }
          ++seq;
          // publish (also the -1 sentinel) to the consumer roles of this CTA
          mbar_wait(&sempty_bar[sslot], sphase ^ 1u, 5);
          tile_ring[sslot] = tile;
          mbar_arrive(&sfull_bar[sslot]);
          if (++sslot == kSched) {
            sslot = 0;
            sphase ^= 1u;
          }
          if (tile < 0) break;
          // claim the next tile now; the atomic's round trip overlaps this tile's loads
          if (rank == 0) next = atomicAdd(tile_counter, 1);
        } else {
          if (tile >= num_tiles) break;
          next = tile + num_pairs;
        }
        if constexpr (MODE != 1) {
          const int m_blk = M_FASTEST ? tile % num_m : tile / num_n;
          const int n_blk = M_FASTEST ? tile / num_m : tile % num_n;
          const int row_a = m_blk * Cfg::kTileM + static_cast<int>(rank) * kBlockM;
          const int row_b = n_blk * Cfg::kTileN + static_cast<int>(rank) * (Cfg::kTileN / 2);
          for (int kb = 0; kb < num_k; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            const uint32_t leader_full = mapa_rank(&full_bar[stage], 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            tma_load_2d_2sm(sa, &tmA, leader_full, kb * kBlockK, row_a);
            tma_load_2d_2sm(sa + Cfg::kABytes, &tmB, leader_full, kb * kBlockK, row_b);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        tile = (dyn && next >= num_tiles) ? -1 : next;
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ------------------------------ MMA issuer (leader only) ------------------------------
      constexpr uint32_t idesc = F16 ? umma_idesc_f16(Cfg::kTileM, Cfg::kTileN) : umma_idesc_bf16(Cfg::kTileM, Cfg::kTileN);
      uint32_t stage = 0, phase = 0, sslot = 0, sphase = 0;
      int it = 0;
      for (int tile = pair;; tile += num_pairs, ++it) {
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
        const uint32_t d_tmem = tmem_base + as * Cfg::kTileN;
        for (int kb = 0; kb < num_k; ++kb) {
          if constexpr (MODE != 1) mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * kUmmaK * 2, kDescKMajorSW128);
            const uint64_t db = umma_smem_desc(b_addr + k * kUmmaK * 2, kDescKMajorSW128);
            if constexpr (MODE == 2) {
              (void)da;
              (void)db;
            } else {
              umma_f16_ss_2sm(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit_2sm(&empty_bar[stage]);  // frees this stage in both CTAs
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_2sm(&tfull_bar[as]);  // accumulator complete -> both epilogues
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
    uint32_t sslot = 0, sphase = 0;
    const uint32_t leader_tempty0 = mapa_rank(&tempty_bar[0], 0);
    for (int tile = pair;; tile += num_pairs, ++it) {
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
      const int m_blk128 = m_blk * 2 + static_cast<int>(rank);  // in units of 128 rows, as the functors expect
      const int row = m_blk128 * kBlockM + ew * 32 + lane;
      const int col_base = n_blk * Cfg::kTileN;
      epi.begin(st, row, m_blk128, n_blk);
      if constexpr (Epi::kPrefetch) epi.prefetch(st, row, col_base + half * kChunks * 32);
      mbar_wait_warp(&tfull_bar[as], aphase, 4);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + as * Cfg::kTileN + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
      for (int pass = 0; pass < Epi::kPasses; ++pass) {
        if constexpr (Epi::kPasses > 1) {
          if (pass > 0) {
            if (!epi.need_pass(st, pass)) break;  // warp-uniform decision
            epi.between(st, row);
          }
        }
        const int c0 = half * kChunks;
        auto run = [&](const uint32_t (&rb)[32], int c, bool has_next) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rb[i]);
          if constexpr (Epi::kPasses > 1)
            epi.chunk(st, row, col_base + c * 32, v, pass);
          else if constexpr (Epi::kPrefetch)
            epi.chunk(st, row, col_base + c * 32, v, has_next ? col_base + (c + 1) * 32 : -1);
          else
            epi.chunk(st, row, col_base + c * 32, v);
        };
        if constexpr (epi_rolled<Epi>::value) {  // heavy functors: one copy of the chunk code (see gemm.cuh)
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
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0)
          mbar_arrive(&tempty_bar[as]);
        else
          mbar_arrive_cluster(leader_tempty0 + as * 8u);
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

// Host launcher (cluster of 2 CTAs along x).  A: [M, K] row pitch lda; B: [N, K] row pitch ldb; 2-byte elements (bf16,
// or IEEE half with F16).  dynamic_sched: see gemm.cuh.  Returns cudaErrorNotSupported if no cluster of this kernel
// fits on the device.
template <int STAGES, bool M_FASTEST, int EPI_WARPS, bool F16 = false, int MODE = 0, class Epi>
static inline cudaError_t launch_gemm2(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K,
                                       const Epi& epi, int num_sms, cudaStream_t stream, bool dynamic_sched = false,
                                       int* pairs_out = nullptr) {
  using Cfg = Gemm2Cfg<STAGES>;
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  CUtensorMap tmA, tmB;
  if (make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, kBlockK, kBlockM) != 0)
    return cudaErrorInvalidValue;
  if (make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, kBlockK, Cfg::kTileN / 2) != 0)
    return cudaErrorInvalidValue;
  auto kern = gemm2_tn_kernel<STAGES, M_FASTEST, EPI_WARPS, Epi, F16, MODE>;
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
  // partner, so fewer than num_sms / 2 pairs may fit (a second wave of clusters would double the run time).
  static int max_pairs = 0;  // per instantiation
  if (!max_pairs) {
    cfg.gridDim = dim3(2 * (num_sms / 2));
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess) return e;
    if (n < 1) return cudaErrorNotSupported;
    max_pairs = std::min(n, num_sms / 2);
  }
  const int pairs = std::max(1, std::min(num_tiles, max_pairs));
  cfg.gridDim = dim3(2 * pairs);
  if (pairs_out) *pairs_out = pairs;
  int* counter = nullptr;
  if (dynamic_sched) {
    counter = next_tile_counter(stream);
    if (!counter) return cudaErrorMemoryAllocation;
  }
  return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, M, N, K, epi, counter);
}

}  // namespace om
