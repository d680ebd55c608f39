// In-batch-negatives contrastive loss, forward + backward, replacing
//   logits = x @ y.T ; F.cross_entropy(logits, target)          src/openmatch/loss.py:7-15
//   scores = q_reps @ p_reps.T ; CrossEntropyLoss(mean)          src/openmatch/modeling/dense_retrieval_model.py:113-122
// and their autograd backward.
//
//   PREP    Q, P -> bf16 row-major copies and bf16 transposes (operands of the backward GEMMs)
//   LOGITS  S = Q P^T on tcgen05 (gemm.cuh), fp32 [nq, np]
//   SOFTMAX per query row: log-sum-exp (fp32), loss_i = lse_i - s_i,t_i, G = w (softmax - onehot) -> bf16 G, G^T
//   GRADS   dQ = G P   (tcgen05: A = G   [nq, np], B = P^T [d, np])
//           dP = G^T Q (tcgen05: A = G^T [np, nq], B = Q^T [d, nq])
//   REDUCE  loss = scale * sum_i loss_i * w   (single block, fixed order: deterministic)
#include <algorithm>

#include "common.h"
#include "gemm.cuh"

namespace om {

// src [rows, cols] (fp32 or bf16) -> dst bf16 [rows, ldd] and dstT bf16 [cols, ldt] (pads untouched: TMA
// never reads beyond the logical extent)
template <typename T>
__global__ void prep_kernel(const T* __restrict__ src, int rows, int cols, __nv_bfloat16* dst, int ldd,
                            __nv_bfloat16* dstT, int ldt) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int r = r0 + dy, c = c0 + threadIdx.x;
    __nv_bfloat16 v = __float2bfloat16(0.f);
    if (r < rows && c < cols) {
      v = __float2bfloat16(static_cast<float>(src[static_cast<int64_t>(r) * cols + c]));
      dst[static_cast<int64_t>(r) * ldd + c] = v;
    }
    tile[dy][threadIdx.x] = v;
  }
  __syncthreads();
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int c = c0 + dy, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dstT[static_cast<int64_t>(c) * ldt + r] = tile[threadIdx.x][dy];
  }
}

// One CTA per query row.  S fp32 [nq, np] -> row loss, G bf16 [nq, ldg], G^T bf16 [np, ldgt]
__global__ void __launch_bounds__(256) softmax_grad_kernel(const float* __restrict__ S, int nq, int np,
                                                           const int64_t* __restrict__ target, int tpq, float w,
                                                           float* row_loss, __nv_bfloat16* G, int ldg,
                                                           __nv_bfloat16* GT, int ldgt, int* bad_target) {
  const int q = blockIdx.x;
  const float* s = S + static_cast<int64_t>(q) * np;
  __shared__ float red[8];
  __shared__ float bcast;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < np; j += blockDim.x) m = fmaxf(m, s[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = red[0];
    for (int i = 1; i < 8; ++i) mm = fmaxf(mm, red[i]);
    bcast = mm;
  }
  __syncthreads();
  m = bcast;
  float z = 0.f;
  for (int j = threadIdx.x; j < np; j += blockDim.x) z += expf(s[j] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = z;
  __syncthreads();
  if (threadIdx.x == 0) {
    float zz = 0.f;
    for (int i = 0; i < 8; ++i) zz += red[i];
    bcast = zz;
  }
  __syncthreads();
  z = bcast;
  int64_t t = target ? target[q] : static_cast<int64_t>(q) * tpq;
  if (t < 0 || t >= np) {
    if (threadIdx.x == 0) *bad_target = 1;
    t = 0;
  }
  if (threadIdx.x == 0) row_loss[q] = (m + logf(z)) - s[t];
  const float inv = 1.0f / z;
  for (int j = threadIdx.x; j < np; j += blockDim.x) {
    float g = expf(s[j] - m) * inv;
    if (j == t) g -= 1.0f;
    const __nv_bfloat16 gb = __float2bfloat16(g * w);
    if (G) G[static_cast<int64_t>(q) * ldg + j] = gb;
    if (GT) GT[static_cast<int64_t>(j) * ldgt + q] = gb;
  }
}

// out-of-range targets (PyTorch would device-assert) poison the loss with NaN instead of being ignored
__global__ void loss_reduce_kernel(const float* row_loss, int nq, float w, float scale, const int* bad_target,
                                   float* loss_out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nq; i += blockDim.x) acc += row_loss[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = *bad_target ? __int_as_float(0x7fc00000) : static_cast<float>(red[0] * w * scale);
}

struct LossWs {
  void* p = nullptr;
  size_t bytes = 0;
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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
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
  if (off > g_loss_ws.bytes) {
    if (g_loss_ws.p) {
      OM_CUDA(cudaStreamSynchronize(st));
      cudaFree(g_loss_ws.p);
    }
    g_loss_ws.p = nullptr;
    g_loss_ws.bytes = 0;
    OM_CUDA(cudaMalloc(&g_loss_ws.p, off));
    g_loss_ws.bytes = off;
  }
  uint8_t* base = static_cast<uint8_t*>(g_loss_ws.p);
  auto* qb = reinterpret_cast<__nv_bfloat16*>(base + o_qb);
  auto* pb = reinterpret_cast<__nv_bfloat16*>(base + o_pb);
  auto* qt = reinterpret_cast<__nv_bfloat16*>(base + o_qt);
  auto* pt = reinterpret_cast<__nv_bfloat16*>(base + o_pt);
  float* S = scores_out ? scores_out : reinterpret_cast<float*>(base + o_s);
  auto* G = reinterpret_cast<__nv_bfloat16*>(base + o_g);
  auto* GT = reinterpret_cast<__nv_bfloat16*>(base + o_gt);
  float* row_loss = reinterpret_cast<float*>(base + o_rl);
  int* flag = reinterpret_cast<int*>(base + o_flag);

  const dim3 tb(32, 8);
  if (dtype == OM_F32) {
    prep_kernel<<<dim3((d + 31) / 32, (nq + 31) / 32), tb, 0, st>>>(static_cast<const float*>(Q), nq, d, qb, dpad, qt, nqp);
    prep_kernel<<<dim3((d + 31) / 32, (np + 31) / 32), tb, 0, st>>>(static_cast<const float*>(P), np, d, pb, dpad, pt, npp);
  } else {
    prep_kernel<<<dim3((d + 31) / 32, (nq + 31) / 32), tb, 0, st>>>(static_cast<const __nv_bfloat16*>(Q), nq, d, qb, dpad, qt, nqp);
    prep_kernel<<<dim3((d + 31) / 32, (np + 31) / 32), tb, 0, st>>>(static_cast<const __nv_bfloat16*>(P), np, d, pb, dpad, pt, npp);
  }
  OM_CUDA(cudaGetLastError());
  {
    EpiStoreF32 epi{S, np, nullptr, nullptr, 0, nq, np};
    cudaError_t e = launch_gemm<128, 4, false, 4>(qb, dpad, pb, dpad, nq, np, d, epi, sms, st);
    if (e != cudaSuccess) return fail(OM_ECUDA, "loss logits GEMM launch failed: %s", cudaGetErrorString(e));
  }
  const float w = reduction == OM_REDUCE_MEAN ? 1.0f / nq : 1.0f;
  const bool need_grad = dQ || dP;
  OM_CUDA(cudaMemsetAsync(flag, 0, 4, st));
  softmax_grad_kernel<<<nq, 256, 0, st>>>(S, nq, np, target, np / nq, w * loss_scale, row_loss,
                                          need_grad && dQ ? G : nullptr, npp, need_grad && dP ? GT : nullptr, nqp, flag);
  OM_CUDA(cudaGetLastError());
  loss_reduce_kernel<<<1, 256, 0, st>>>(row_loss, nq, w, loss_scale, flag, loss_out);
  OM_CUDA(cudaGetLastError());
  if (dQ) {
    EpiStoreF32 epi{dQ, d, nullptr, nullptr, 0, nq, d};
    cudaError_t e = launch_gemm<128, 4, false, 4>(G, npp, pt, npp, nq, d, np, epi, sms, st);
    if (e != cudaSuccess) return fail(OM_ECUDA, "loss dQ GEMM launch failed: %s", cudaGetErrorString(e));
  }
  if (dP) {
    EpiStoreF32 epi{dP, d, nullptr, nullptr, 0, np, d};
    cudaError_t e = launch_gemm<128, 4, false, 4>(GT, nqp, qt, nqp, np, d, nq, epi, sms, st);
    if (e != cudaSuccess) return fail(OM_ECUDA, "loss dP GEMM launch failed: %s", cudaGetErrorString(e));
  }
  return 0;
}
