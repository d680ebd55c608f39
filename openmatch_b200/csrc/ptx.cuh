// Inline-PTX primitives for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld), UMMA shared-memory + instruction descriptors.  Hand-written; no CUTLASS dependency.
//
// Every blocking wait is bounded: a wait that exceeds OM_WAIT_TIMEOUT_CYCLES records a diagnostic
// code in the translation unit's fault word and returns; every other waiter then bails out as soon as
// it sees the word set.  A pipeline bug therefore ends the kernel with garbage results plus a non-zero
// fault word that the host turns into an error, instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef OM_WAIT_TIMEOUT_CYCLES
#define OM_WAIT_TIMEOUT_CYCLES (1000000000ll)  // ~0.5 s at 1.9 GHz
#endif

namespace om {

// One fault word per translation unit (static): 0 = ok, else (site << 16) | (blockIdx.x & 0xffff).
static __device__ unsigned int om_dev_fault;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Make generic-proxy writes to shared memory visible to the async proxy (TMA / UMMA reads).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Same for global memory: generic-proxy stores (any CTA) that a later TMA load will read.  Executed by the writer
// before the release / barrier and by the TMA-issuing thread after the acquire.
__device__ __forceinline__ void fence_proxy_async_global() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, uint32_t site) {
  // try_wait suspends the thread in hardware until the phase completes or a time limit expires, so this
  // loop is not a busy spin.  The fault word (a global load, ~1 us) and the clock are consulted only every
  // 1024 wake-ups: polling them on every iteration added a DRAM round trip to every pipeline hand-off.
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
#if defined(OM_SPIN_SLEEP_NS) && OM_SPIN_SLEEP_NS > 0
    __nanosleep(OM_SPIN_SLEEP_NS);
#endif
    if ((++spins & 1023u) != 0) continue;
    if (*reinterpret_cast<volatile unsigned int*>(&om_dev_fault) != 0u) return;
    if (clock64() - t0 > OM_WAIT_TIMEOUT_CYCLES) {
      atomicCAS(&om_dev_fault, 0u, (site << 16) | (blockIdx.x & 0xffffu) | 0x80000000u);
      return;
    }
  }
}
// Bounded wait on phase `parity`. `site` identifies the call site in the fault word.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t site) {
  if (mbar_try_wait(bar, parity)) return;
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, site);
}
// Warp-collective wait for consumer warps: lane 0 polls (with a short sleep between probes), the rest of the
// warp parks at __syncwarp.  Hundreds of threads spinning on try_wait compete with the TMA/MMA threads for
// the barrier unit: measured 2x slowdown of the scan kernel when all 256 epilogue threads polled.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, uint32_t site) {
#ifdef OM_EPI_ALL_LANES_WAIT
  if (!mbar_try_wait(bar, parity)) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(OM_EPI_ALL_LANES_WAIT);
  }
  return;
#endif
  if ((threadIdx.x & 31u) == 0) {
    if (!mbar_try_wait(bar, parity)) {
      const long long t0 = clock64();
      uint32_t spins = 0;
      while (!mbar_try_wait(bar, parity)) {
#ifndef OM_EPI_POLL_SLEEP_NS
#define OM_EPI_POLL_SLEEP_NS 64
#endif
        __nanosleep(OM_EPI_POLL_SLEEP_NS);
        if ((++spins & 1023u) != 0) continue;
        if (*reinterpret_cast<volatile unsigned int*>(&om_dev_fault) != 0u) break;
        if (clock64() - t0 > OM_WAIT_TIMEOUT_CYCLES) {
          atomicCAS(&om_dev_fault, 0u, (site << 16) | (blockIdx.x & 0xffffu) | 0x80000000u);
          break;
        }
      }
    }
  }
  __syncwarp();
}
// Host helpers: read-and-clear this translation unit's fault word (call after a stream sync).
static inline unsigned int read_clear_dev_fault() {
  unsigned int v = 0, z = 0;
  if (cudaMemcpyFromSymbol(&v, om_dev_fault, sizeof(v)) != cudaSuccess) return 0xffffffffu;
  if (v) cudaMemcpyToSymbol(om_dev_fault, &z, sizeof(z));
  return v;
}

// ----------------------------------------------------------------------------------------------
// TMA: 2-D tiled bulk tensor load, global -> shared, completion on an mbarrier (complete_tx::bytes)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// c0 = coordinate along the contiguous (inner) dimension in elements, c1 = row coordinate.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0,
                                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0,
                                                 int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "l"(policy)
      : "memory");
}
// L2 prefetch of one box (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* tmap, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(c0), "r"(c1)
               : "memory");
}
// TMA store: shared (128B-swizzled box) -> global, tracked by the thread's bulk async-group.
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk groups have finished READING their shared-memory sources
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ----------------------------------------------------------------------------------------------
// tcgen05: tensor memory allocation, MMA issue, commit, TMEM loads
// ----------------------------------------------------------------------------------------------
// Whole-warp collective. ncols: power of two in [32, 512]. The base TMEM address lands in *dst_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; single-thread issue. accumulate==0 overwrites D.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (count 1) on `bar` once all previously issued tcgen05.mma of this thread have completed.
// Implies tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// Warp-collective: lane t receives 32 consecutive fp32 columns of TMEM lane (lane_base + t).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout per PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor")
// ----------------------------------------------------------------------------------------------
// Shared-memory operand descriptor.
//   [0,14)  start address >> 4        [16,30) leading byte offset >> 4     [32,46) stride byte offset >> 4
//   [46,48) descriptor version (1 on sm_100)   [49,52) base offset   [61,64) swizzle (2 = 128 B)
__host__ __device__ constexpr uint64_t umma_smem_desc_base(uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                           uint32_t layout_type) {
  return (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) |
         (static_cast<uint64_t>(layout_type & 7u) << 61);
}
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint64_t base) {
  return base | static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
}
constexpr uint32_t kSwizzle128B = 2;  // UMMA layout_type encoding
// K-major operand, rows of 64 bf16 (=128 B) swizzled in 8-row / 1024-B atoms (what a TMA box of
// {64 elems, R rows} with CU_TENSOR_MAP_SWIZZLE_128B produces): SBO = 1024 B between 8-row groups.
constexpr uint64_t kDescKMajorSW128 = umma_smem_desc_base(16, 1024, kSwizzle128B);

// Instruction descriptor for kind::f16 with BF16 A/B, FP32 accumulate.
//   [4,6) D fmt (1=f32)  [7,10) A fmt (1=bf16)  [10,13) B fmt  [15] A major (0=K)  [16] B major
//   [17,23) N>>3   [24,29) M>>4
// A and B in IEEE fp16 (format code 0), fp32 accumulate, K-major operands
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major = 0,
                                                       uint32_t b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// Packed fp32x2 arithmetic (sm_100: FFMA2 — two fp32 FMAs per issue slot); used by issue-bound epilogues.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(r)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(r)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(r)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 splat2(float x) { return make_float2(x, x); }

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace om
