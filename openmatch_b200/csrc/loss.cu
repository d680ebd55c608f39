// In-batch-negatives contrastive loss, forward + backward, as ONE persistent tcgen05 kernel, replacing
//   logits = x @ y.T ; F.cross_entropy(logits, target)          src/openmatch/loss.py:7-15
//   scores = q_reps @ p_reps.T ; CrossEntropyLoss(mean)          src/openmatch/modeling/dense_retrieval_model.py:113-122
// and their autograd backward (~8 PyTorch launches forward, as many backward).
//
// loss_fused_kernel: cooperative grid (<= 1 CTA per SM), four phases separated by grid barriers; every GEMM runs
// on the tcgen05 pipeline of gemm.cuh's design (TMA -> 4-stage smem ring -> tcgen05.mma 128x128x16 -> TMEM ->
// tcgen05.ld epilogue), with the pipeline state carried from phase to phase:
//   PREP    Q, P -> bf16 row-major copies and bf16 transposes (operands of the backward GEMMs)
//   LOGITS  S = Q P^T, fp32 [nq, np]
//   SOFTMAX one warp per query row: log-sum-exp (fp32), loss_i = lse_i - s_i,t_i, G = w (softmax - onehot)
//           -> bf16 G, G^T; after the barrier the last CTA reduces the row losses in a fixed order (deterministic)
//   GRADS   dQ = G P   (A = G   [nq, np], B = P^T [d, np])
//           dP = G^T Q (A = G^T [np, nq], B = Q^T [d, nq])
#include <algorithm>

#include "common.h"
#include "gemm.cuh"

namespace om {

constexpr int kLossBN = 128, kLossStages = 4, kLossThreads = 256;
using LossCfg = GemmCfg<kLossBN, kLossStages>;

struct LossMaps {  // [0] logits (Q, P)   [1] dQ (G, P^T)   [2] dP (G^T, Q^T); boxes {64, 128}, 128-B swizzle
  CUtensorMap a[3], b[3];
};

struct LossArgs {
  const void* Q;
  const void* P;
  int is_bf16;
  int nq, np, d, dpad, nqp, npp;
  const int64_t* target;
  float w, loss_scale;
  __nv_bfloat16 *qb, *pb, *qt, *pt, *G, *GT;
  float *S, *row_loss, *loss_out, *dQ, *dP;
  int* bad_target;
  unsigned* grid_bar;
};

// Self-resetting grid barrier (the cooperative-groups scheme): CTA 0 adds 0x80000000 - (G - 1), the others 1, so
// the top bit flips exactly when all G have arrived and the low bits return to zero.  Bounded spin: a lost CTA
// raises the fault word instead of hanging the GPU.
__device__ __forceinline__ void grid_sync(unsigned* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned inc = blockIdx.x == 0 ? 0x80000000u - (gridDim.x - 1u) : 1u;
    __threadfence();
    const unsigned old = atomicAdd(bar, inc);
    const long long t0 = clock64();
    while (((old ^ *reinterpret_cast<volatile unsigned*>(bar)) & 0x80000000u) == 0u) {
      if (clock64() - t0 > OM_WAIT_TIMEOUT_CYCLES) {
        atomicCAS(&om_dev_fault, 0u, 0x80ee0000u | (blockIdx.x & 0xffffu));
        break;
      }
    }
    __threadfence();
  }
  __syncthreads();
}

struct LossSmem {
  uint8_t* ring;
  uint64_t *full_bar, *empty_bar, *tfull_bar, *tempty_bar;
  uint32_t tmem_base;
};
struct Pipe {  // per-thread pipeline position, carried across the GEMM phases (every role advances identically)
  uint32_t stage = 0, phase = 0;
  int it = 0;
};

// C[M, N] fp32 (row pitch ldc) = A[M, K] B[N, K]^T over this CTA's share of the 128 x 128 tiles; `rot` rotates the
// tile -> CTA assignment so that back-to-back GEMMs start on different CTAs.
__device__ __forceinline__ void gemm_phase(const CUtensorMap* tmA, const CUtensorMap* tmB, int M, int N, int K, float* C,
                                           int ldc, int rot, const LossSmem& sm, Pipe& pipe, int warp, int lane) {
  const int num_n = (N + kLossBN - 1) / kLossBN;
  const int num_tiles = ((M + kBlockM - 1) / kBlockM) * num_n;
  const int num_k = (K + kBlockK - 1) / kBlockK;
  const int G = static_cast<int>(gridDim.x);
  const int first = (static_cast<int>(blockIdx.x) + G - rot % G) % G;
  if (warp == 0) {
    if (lane == 0) {  // TMA producer
      fence_proxy_async_global();  // operands were written with ordinary stores by other CTAs before the barrier
      for (int tile = first; tile < num_tiles; tile += G) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&sm.empty_bar[pipe.stage], pipe.phase ^ 1u, 1);
          uint8_t* sa = sm.ring + pipe.stage * LossCfg::kStageBytes;
          mbar_arrive_expect_tx(&sm.full_bar[pipe.stage], LossCfg::kStageBytes);
          tma_load_2d(sa, tmA, &sm.full_bar[pipe.stage], kb * kBlockK, m_blk * kBlockM);
          tma_load_2d(sa + LossCfg::kABytes, tmB, &sm.full_bar[pipe.stage], kb * kBlockK, n_blk * kLossBN);
          if (++pipe.stage == kLossStages) {
            pipe.stage = 0;
            pipe.phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // MMA issuer
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, kLossBN);
      for (int tile = first; tile < num_tiles; tile += G, ++pipe.it) {
        const uint32_t as = pipe.it & 1, aphase = (pipe.it >> 1) & 1;
        mbar_wait(&sm.tempty_bar[as], aphase ^ 1u, 2);
        tc_fence_after_sync();
        const uint32_t d_tmem = sm.tmem_base + as * kLossBN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&sm.full_bar[pipe.stage], pipe.phase, 3);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sm.ring + pipe.stage * LossCfg::kStageBytes);
          const uint32_t b_addr = a_addr + LossCfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * kUmmaK * 2, kDescKMajorSW128);
            const uint64_t db = umma_smem_desc(b_addr + k * kUmmaK * 2, kDescKMajorSW128);
            umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&sm.empty_bar[pipe.stage]);
          if (++pipe.stage == kLossStages) {
            pipe.stage = 0;
            pipe.phase ^= 1u;
          }
        }
        umma_commit(&sm.tfull_bar[as]);
      }
    }
  } else if (warp >= 4) {  // epilogue: warp w owns TMEM lanes [32 (w % 4), +32) = rows of the tile
    const int ew = warp & 3;
    for (int tile = first; tile < num_tiles; tile += G, ++pipe.it) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const uint32_t as = pipe.it & 1, aphase = (pipe.it >> 1) & 1;
      const int row = m_blk * kBlockM + ew * 32 + lane;
      mbar_wait_warp(&sm.tfull_bar[as], aphase, 4);
      tc_fence_after_sync();
      const uint32_t taddr = sm.tmem_base + as * kLossBN + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < kLossBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c * 32, r);
        tmem_ld_wait();
        const int col0 = n_blk * kLossBN + c * 32;
        if (row < M && col0 < N) {
          float* out = C + static_cast<int64_t>(row) * ldc + col0;
          if (col0 + 32 <= N && (ldc & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(out)[j] =
                  make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                              __uint_as_float(r[4 * j + 3]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < N) out[i] = __uint_as_float(r[i]);
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tempty_bar[as]);
    }
  }
}

// src [rows, cols] (fp32 or bf16) -> dst bf16 [rows, ldd] and dstT bf16 [cols, ldt], 64 x 64 tiles starting at this
// CTA's index, 16 independent loads per thread in flight (pads untouched: TMA never reads beyond the logical
// extent)
constexpr int kPrepTile = 64;
template <typename T>
__device__ __forceinline__ void prep_tiles(const T* __restrict__ src, int rows, int cols, __nv_bfloat16* dst, int ldd,
                                           __nv_bfloat16* dstT, int ldt, __nv_bfloat16 (*tile)[kPrepTile + 1],
                                           int t_begin, int t_step) {
  constexpr int kPer = kPrepTile * kPrepTile / kLossThreads;  // 16
  const int tiles_c = (cols + kPrepTile - 1) / kPrepTile, tiles = tiles_c * ((rows + kPrepTile - 1) / kPrepTile);
  for (int t = t_begin; t < tiles; t += t_step) {
    const int c0 = (t % tiles_c) * kPrepTile, r0 = (t / tiles_c) * kPrepTile;
    float v[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int idx = i * kLossThreads + static_cast<int>(threadIdx.x);
      const int r = r0 + idx / kPrepTile, c = c0 + idx % kPrepTile;
      v[i] = (r < rows && c < cols) ? static_cast<float>(src[static_cast<int64_t>(r) * cols + c]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int idx = i * kLossThreads + static_cast<int>(threadIdx.x);
      const int dr = idx / kPrepTile, dc = idx % kPrepTile;
      const __nv_bfloat16 b = __float2bfloat16(v[i]);
      if (r0 + dr < rows && c0 + dc < cols) dst[static_cast<int64_t>(r0 + dr) * ldd + c0 + dc] = b;
      tile[dr][dc] = b;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int idx = i * kLossThreads + static_cast<int>(threadIdx.x);
      const int dc = idx / kPrepTile, dr = idx % kPrepTile;  // consecutive threads: consecutive rows of src
      if (c0 + dc < cols && r0 + dr < rows) dstT[static_cast<int64_t>(c0 + dc) * ldt + r0 + dr] = tile[dr][dc];
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kLossThreads, 1)
loss_fused_kernel(const __grid_constant__ LossMaps maps, const __grid_constant__ LossArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __nv_bfloat16 tile[kPrepTile][kPrepTile + 1];
  __shared__ double red[kLossThreads];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  LossSmem sm;
  sm.ring = smem;
  sm.full_bar = reinterpret_cast<uint64_t*>(smem + LossCfg::kBarOffset);
  sm.empty_bar = sm.full_bar + kLossStages;
  sm.tfull_bar = sm.empty_bar + kLossStages;
  sm.tempty_bar = sm.tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm.tempty_bar + 2);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(threadIdx.x & 31);
  const int G = static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 3; ++i) {
      tma_prefetch_desc(&maps.a[i]);
      tma_prefetch_desc(&maps.b[i]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kLossStages; ++i) {
      mbar_init(&sm.full_bar[i], 1);
      mbar_init(&sm.empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.tfull_bar[i], 1);
      mbar_init(&sm.tempty_bar[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, LossCfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  sm.tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  Pipe pipe;

  // ------------------------------ PREP ------------------------------
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.bad_target = 0;
  {
    // P tiles start at CTA 0, Q tiles where the P tiles end, so the short Q list lands on otherwise idle CTAs
    const int p_tiles = ((a.d + kPrepTile - 1) / kPrepTile) * ((a.np + kPrepTile - 1) / kPrepTile);
    const int q_begin = (static_cast<int>(blockIdx.x) + G - p_tiles % G) % G;
    if (a.is_bf16) {
      prep_tiles(static_cast<const __nv_bfloat16*>(a.P), a.np, a.d, a.pb, a.dpad, a.pt, a.npp, tile, blockIdx.x, G);
      prep_tiles(static_cast<const __nv_bfloat16*>(a.Q), a.nq, a.d, a.qb, a.dpad, a.qt, a.nqp, tile, q_begin, G);
    } else {
      prep_tiles(static_cast<const float*>(a.P), a.np, a.d, a.pb, a.dpad, a.pt, a.npp, tile, blockIdx.x, G);
      prep_tiles(static_cast<const float*>(a.Q), a.nq, a.d, a.qb, a.dpad, a.qt, a.nqp, tile, q_begin, G);
    }
  }
  fence_proxy_async_global();
  grid_sync(a.grid_bar);

  // ------------------------------ LOGITS ------------------------------
  gemm_phase(&maps.a[0], &maps.b[0], a.nq, a.np, a.d, a.S, a.np, 0, sm, pipe, warp, lane);
  grid_sync(a.grid_bar);

  // ------------------------------ SOFTMAX + GRAD OF THE LOGITS ------------------------------
  {
    const int tpq = a.np / a.nq;
    const float wl = a.w * a.loss_scale;
    for (int q = warp * G + static_cast<int>(blockIdx.x); q < a.nq; q += (kLossThreads / 32) * G) {
      const float* s = a.S + static_cast<int64_t>(q) * a.np;
      float m = -INFINITY;
      for (int j = lane; j < a.np; j += 32) m = fmaxf(m, s[j]);
      m = warp_max(m);
      float z = 0.f;
      for (int j = lane; j < a.np; j += 32) z += expf(s[j] - m);
      z = warp_sum(z);
      int64_t t = a.target ? a.target[q] : static_cast<int64_t>(q) * tpq;
      if (t < 0 || t >= a.np) {
        if (lane == 0) *a.bad_target = 1;
        t = 0;
      }
      if (lane == 0) a.row_loss[q] = (m + logf(z)) - s[t];
      if (a.dQ || a.dP) {
        const float inv = 1.0f / z;
        for (int j = lane; j < a.np; j += 32) {
          float g = expf(s[j] - m) * inv;
          if (j == t) g -= 1.0f;
          const __nv_bfloat16 gb = __float2bfloat16(g * wl);
          if (a.dQ) a.G[static_cast<int64_t>(q) * a.npp + j] = gb;
          if (a.dP) a.GT[static_cast<int64_t>(j) * a.nqp + q] = gb;
        }
      }
    }
  }
  fence_proxy_async_global();
  grid_sync(a.grid_bar);

  // ------------------------------ LOSS (last CTA, fixed summation order) ------------------------------
  if (static_cast<int>(blockIdx.x) == G - 1) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < a.nq; i += kLossThreads) acc += a.row_loss[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kLossThreads / 2; s > 0; s >>= 1) {
      if (static_cast<int>(threadIdx.x) < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    // out-of-range targets (PyTorch would device-assert) poison the loss with NaN instead of being ignored
    if (threadIdx.x == 0)
      *a.loss_out = *a.bad_target ? __int_as_float(0x7fc00000) : static_cast<float>(red[0] * a.w * a.loss_scale);
  }

  // ------------------------------ GRADS ------------------------------
  int rot = 0;
  if (a.dQ) {
    gemm_phase(&maps.a[1], &maps.b[1], a.nq, a.d, a.np, a.dQ, a.d, rot, sm, pipe, warp, lane);
    rot = ((a.nq + kBlockM - 1) / kBlockM) * ((a.d + kLossBN - 1) / kLossBN);
  }
  if (a.dP) gemm_phase(&maps.a[2], &maps.b[2], a.np, a.d, a.nq, a.dP, a.d, rot, sm, pipe, warp, lane);

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(sm.tmem_base, LossCfg::kTmemCols);
  }
}

struct LossWs {
  void* p = nullptr;
  size_t bytes = 0;
  unsigned* grid_bar = nullptr;  // persistent (self-resetting), zeroed once
  // tensor maps are rebuilt only when the problem or the workspace changes
  LossMaps maps;
  const void* maps_base = nullptr;
  int maps_nq = 0, maps_np = 0, maps_d = 0;
};
static LossWs g_loss_ws;  // grown on demand; one process drives one GPU (see header)

}  // namespace om

using namespace om;

extern "C" int om_contrastive_loss_fwd_bwd(const void* Q, const void* P, om_dtype dtype, int nq, int np, int d,
                                           const int64_t* target, int reduction, float loss_scale, float* loss_out,
                                           float* dQ, float* dP, float* scores_out, void* stream) {
  if (!Q || !P || !loss_out || nq <= 0 || np <= 0 || d <= 0)
    return fail(OM_EINVAL, "om_contrastive_loss_fwd_bwd: bad arguments (nq=%d np=%d d=%d)", nq, np, d);
  if (dtype != OM_F32 && dtype != OM_BF16) return fail(OM_EINVAL, "loss: dtype must be f32 or bf16");
  if (reduction != OM_REDUCE_MEAN && reduction != OM_REDUCE_SUM) return fail(OM_EINVAL, "loss: bad reduction");
  if (!target && np < nq) return fail(OM_EINVAL, "loss: default target needs np >= nq");
  const int sms = device_sm_count();
  if (sms < 0) return sms;
  NvtxRange nvtx("om.loss");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LossWs& ws = g_loss_ws;
  const int dpad = (int)round_up(d, 8), nqp = (int)round_up(nq, 8), npp = (int)round_up(np, 8);
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    size_t o = off;
    off += round_up(bytes, 256);
    return o;
  };
  const size_t o_qb = carve((size_t)nq * dpad * 2), o_pb = carve((size_t)np * dpad * 2);
  const size_t o_qt = carve((size_t)d * nqp * 2), o_pt = carve((size_t)d * npp * 2);
  const size_t o_s = carve(scores_out ? 0 : (size_t)nq * np * 4);
  const size_t o_g = carve((size_t)nq * npp * 2), o_gt = carve((size_t)np * nqp * 2);
  const size_t o_rl = carve((size_t)nq * 4), o_flag = carve(256);
  if (off > ws.bytes) {
    if (ws.p) {
      OM_CUDA(cudaStreamSynchronize(st));
      cudaFree(ws.p);
    }
    ws.p = nullptr;
    ws.bytes = 0;
    ws.maps_base = nullptr;
    OM_CUDA(cudaMalloc(&ws.p, off));
    ws.bytes = off;
  }
  if (!ws.grid_bar) {
    OM_CUDA(cudaMalloc(&ws.grid_bar, 256));
    OM_CUDA(cudaMemset(ws.grid_bar, 0, 256));
  }
  uint8_t* base = static_cast<uint8_t*>(ws.p);
  LossArgs a;
  a.Q = Q;
  a.P = P;
  a.is_bf16 = dtype == OM_BF16;
  a.nq = nq;
  a.np = np;
  a.d = d;
  a.dpad = dpad;
  a.nqp = nqp;
  a.npp = npp;
  a.target = target;
  a.w = reduction == OM_REDUCE_MEAN ? 1.0f / nq : 1.0f;
  a.loss_scale = loss_scale;
  a.qb = reinterpret_cast<__nv_bfloat16*>(base + o_qb);
  a.pb = reinterpret_cast<__nv_bfloat16*>(base + o_pb);
  a.qt = reinterpret_cast<__nv_bfloat16*>(base + o_qt);
  a.pt = reinterpret_cast<__nv_bfloat16*>(base + o_pt);
  a.G = reinterpret_cast<__nv_bfloat16*>(base + o_g);
  a.GT = reinterpret_cast<__nv_bfloat16*>(base + o_gt);
  a.S = scores_out ? scores_out : reinterpret_cast<float*>(base + o_s);
  a.row_loss = reinterpret_cast<float*>(base + o_rl);
  a.loss_out = loss_out;
  a.dQ = dQ;
  a.dP = dP;
  a.bad_target = reinterpret_cast<int*>(base + o_flag);
  a.grid_bar = ws.grid_bar;

  if (ws.maps_base != ws.p || ws.maps_nq != nq || ws.maps_np != np || ws.maps_d != d) {
    int rc = 0;
    rc |= make_tmap_bf16_2d(&ws.maps.a[0], a.qb, (uint64_t)d, (uint64_t)nq, (uint64_t)dpad * 2, kBlockK, kBlockM);
    rc |= make_tmap_bf16_2d(&ws.maps.b[0], a.pb, (uint64_t)d, (uint64_t)np, (uint64_t)dpad * 2, kBlockK, kLossBN);
    rc |= make_tmap_bf16_2d(&ws.maps.a[1], a.G, (uint64_t)np, (uint64_t)nq, (uint64_t)npp * 2, kBlockK, kBlockM);
    rc |= make_tmap_bf16_2d(&ws.maps.b[1], a.pt, (uint64_t)np, (uint64_t)d, (uint64_t)npp * 2, kBlockK, kLossBN);
    rc |= make_tmap_bf16_2d(&ws.maps.a[2], a.GT, (uint64_t)nq, (uint64_t)np, (uint64_t)nqp * 2, kBlockK, kBlockM);
    rc |= make_tmap_bf16_2d(&ws.maps.b[2], a.qt, (uint64_t)nq, (uint64_t)d, (uint64_t)nqp * 2, kBlockK, kLossBN);
    if (rc != 0) return fail(OM_ECUDA, "loss: tensor-map encode failed (%d)", rc);
    ws.maps_base = ws.p;
    ws.maps_nq = nq;
    ws.maps_np = np;
    ws.maps_d = d;
  }

  static int max_ctas = 0;
  if (!max_ctas) {
    OM_CUDA(cudaFuncSetAttribute(loss_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LossCfg::kSmemBytes));
    int per_sm = 0;
    OM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, loss_fused_kernel, kLossThreads, LossCfg::kSmemBytes));
    if (per_sm < 1) return fail(OM_ECUDA, "loss: fused kernel does not fit on an SM");
    max_ctas = sms;  // one CTA per SM: the whole grid is co-resident (required by the grid barrier)
  }
  auto tiles = [](int m, int n) { return ((m + kBlockM - 1) / kBlockM) * ((n + kLossBN - 1) / kLossBN); };
  const int gemm_tiles = std::max(tiles(nq, np), (dQ ? tiles(nq, d) : 0) + (dP ? tiles(np, d) : 0));
  const int prep_tiles_n = ((d + kPrepTile - 1) / kPrepTile) * ((np + kPrepTile - 1) / kPrepTile + (nq + kPrepTile - 1) / kPrepTile);
  int grid = std::max(std::max(gemm_tiles, (nq + 7) / 8), prep_tiles_n);
  grid = std::max(1, std::min(grid, max_ctas));
  void* params[] = {const_cast<LossMaps*>(&ws.maps), &a};
  OM_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(loss_fused_kernel), dim3(grid), dim3(kLossThreads), params,
                                      LossCfg::kSmemBytes, st));
  return 0;
}
