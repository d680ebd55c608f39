// B200-native encoder forward behind DRModel.encode (src/openmatch/modeling/dense_retrieval_model.py:133-155):
// HF BertModel (post-LN, GELU-erf) or T5EncoderModel (pre-RMSNorm, ReLU, shared relative position bias)
// -> 'first' / 'mean' pooling (:145-150, src/openmatch/utils.py:233-235) -> bias-free LinearHead
// (src/openmatch/modeling/linear.py:19,22-23) -> F.normalize (:153-154).
//
// Per layer (T = B*L tokens, H hidden, I = heads*64 attention width, F ffn).  LayerNorm / RMSNorm never runs as a
// kernel of its own: the residual stream is kept UN-normalised (s, fp32 + a bf16 copy) together with per-row (sum,
// sum of squares); the normalisation is folded algebraically into the GEMM that consumes it,
//        LN(s) W^T = rstd * (s Wf^T) + (W beta + b),   Wf = W diag(gamma) with every row centred (sum_i Wf[j, i] = 0,
//        which makes the "- rstd * mean * rowsum" term vanish; RMSNorm has no mean, rows stay as they are),
// and into the residual read of the GEMM that produces the next s:
//   QKV    tcgen05 GEMM [T,H]x[3I,H]^T on bf16(s) and the folded weights; epilogue: rstd[row] * acc + folded bias
//          -> bf16 Q|K [T,2I] and V transposed [I, T]
//   ATTN   one CTA per (128-row tile, head): S = Q K^T (tcgen05, TMEM) -> masked softmax in registers
//          (thread = query row) -> P (bf16, 128B-swizzled smem) -> O = P V (tcgen05) -> ctx bf16 [T,I]
//   OPROJ  tcgen05 GEMM [T,I]x[H,I]^T, epilogue (EpiResidNorm): s' = acc + bias + LN(s) (BERT) / + s (T5), written in
//          place as fp32 (TMA load + TMA store of the residual tile) and as bf16, row statistics of s' accumulated
//   FFN1   tcgen05 GEMM [T,H]x[F,H]^T on bf16(s') and folded weights; epilogue: rstd[row] * acc + folded bias, GELU(erf) /
//          ReLU -> bf16 [T,F]
//   FFN2   tcgen05 GEMM [T,F]x[H,F]^T, epilogue (EpiResidNorm) like OPROJ -> next layer's s
// One norm kernel runs after the last layer (last_hidden_state for pooling).  Activations feeding tensor cores are
// bf16; the residual stream, normalisation statistics, softmax, pooling, head and L2-normalisation are fp32.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "common.h"
#include "gemm.cuh"
#include "gemm2sm.cuh"

namespace om {

constexpr int kHeadDim = 64;
constexpr int kMaxL = 128;       // one attention tile; sequences of at most kMaxL tokens take attn_kernel
constexpr int kMaxLongL = 512;   // longer sequences (multiples of 128 tokens) take attn_long_kernel
constexpr float kLog2e = 1.4426950408889634f;

// ===================================================================================================
// GEMM epilogues
// ===================================================================================================
// GELU(x) = x/2 (1 + erf(x / sqrt 2)) for two elements per instruction (FFMA2): the FFN1 epilogue is ISSUE-bound
// (with erff() / a 13-term rational it needed more issue cycles per tile than the tensor cores need for the
// tile's MMAs), so erf is the cheapest approximation that is invisible after bf16 rounding of the output:
// z P(z^2) / Q(z^2) on [-4, 4], P cubic, Q cubic with Q >= 1, fitted against math.erf: max |err| 2.1e-5
// (bf16 output rounding is 2e-3 relative); 6 FFMA2 + 1 MUFU.RCP per element pair half.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  float2 z = mul2(x, splat2(0.70710678118654752440f));
  z.x = fminf(fmaxf(z.x, -4.f), 4.f);
  z.y = fminf(fmaxf(z.y, -4.f), 4.f);
  const float2 z2 = mul2(z, z);
  float2 p = fma2(splat2(0.0006061712047085166f), z2, splat2(0.041214898228645325f));
  p = fma2(p, z2, splat2(0.1745881289243698f));
  p = fma2(p, z2, splat2(1.1282498836517334f));
  float2 q = fma2(splat2(0.008151348680257797f), z2, splat2(0.10013707727193832f));
  q = fma2(q, z2, splat2(0.48736122250556946f));
  q = fma2(q, z2, splat2(1.0f));
  const float2 erf = mul2(mul2(p, z), make_float2(rcp_approx(q.x), rcp_approx(q.y)));
  const float2 hx = mul2(x, splat2(0.5f));
  return fma2(hx, erf, hx);
}

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

// Normalisation of one row of the residual stream from its (sum, sum of squares): rstd and rstd * mean.
// The statistics arrive as kStatParts partial (sum, sumsq) pairs per row — one per (column tile, epilogue column group)
// of the GEMM that produced the row, summed here in a fixed order (deterministic, no atomics); unused slots hold zeros.
// LayerNorm: var = E[x^2] - mean^2 (fp32; clamped at 0), RMSNorm: mean = 0.  stats == nullptr: identity (rstd 1, rm 0).
constexpr int kStatParts = 8;
struct RowNorm {
  const float* stats;  // [T, kStatParts, 2] or nullptr
  float inv_h, eps;
  int rms;
  __device__ __forceinline__ void get(int row, int M, float& rstd, float& rm) const {
    rstd = 1.f;
    rm = 0.f;
    if (stats && row < M) {
      const float4* p = reinterpret_cast<const float4*>(stats + static_cast<int64_t>(row) * (2 * kStatParts));
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int j = 0; j < kStatParts / 2; ++j) {
        const float4 t = __ldg(p + j);
        sum += t.x;
        sq += t.y;
        sum += t.z;
        sq += t.w;
      }
      const float mean = rms ? 0.f : sum * inv_h;
      const float var = fmaxf(sq * inv_h - mean * mean, 0.f);
      rstd = rsqrtf(var + eps);
      rm = rstd * mean;
    }
  }
};

// Coalescing stage for bf16 epilogue outputs.  A thread owns one accumulator ROW, so a direct 16-byte store
// per lane touches 32 different cache lines per warp instruction and the SM's load/store unit — not the tensor
// core — bounds the GEMM (measured: tensor pipe ~50 % with direct stores, 92 % with no stores).  Instead two
// consecutive 32-column chunks (= 128 bytes per row) are written into a warp-private 4 KB shared-memory tile
// in the TMA SWIZZLE_128B layout (16-byte pieces XOR-ed with row & 7) and ONE lane issues a TMA store of the
// 32-row x 64-column box: no LSU work for the global write, rows beyond M are clipped by the tensor map.
struct StagedBf16 {
  static constexpr int kBytesPerWarp = 4096;
  uint32_t held[16];  // first chunk of the pair, packed bf16x2
  int have, held_col, in_flight;
  uint8_t* tile;
  __device__ __forceinline__ void bind(uint8_t* smem, int epi_tid) {
    tile = smem + (epi_tid >> 5) * kBytesPerWarp;  // 1024-byte aligned (swizzle atom)
    have = 0;
    in_flight = 0;
  }
  // direct (uncoalesced) store of one 32-column chunk: tail of an odd chunk count
  __device__ __forceinline__ static void store_direct(const uint32_t (&pk)[16], __nv_bfloat16* out, int64_t ldo, int row,
                                                      int col0, int M) {
    if (row >= M) return;
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ldo + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
  }
  // pk = packed chunk [col0, col0+32) of this lane's row; every lane of the warp must call this
  __device__ __forceinline__ void push(const uint32_t (&pk)[16], const CUtensorMap* tm_out, int row, int col0) {
    if (!have) {
#pragma unroll
      for (int i = 0; i < 16; ++i) held[i] = pk[i];
      have = 1;
      held_col = col0;
      return;
    }
    have = 0;
    const int lane = threadIdx.x & 31;
    if (in_flight) {  // the previous TMA store must have finished reading the tile
      if (lane == 0) bulk_wait_group_read0();
      __syncwarp();
    }
    uint8_t* mine = tile + lane * 128;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      *reinterpret_cast<uint4*>(mine + ((p ^ (lane & 7)) << 4)) =
          make_uint4(held[4 * p], held[4 * p + 1], held[4 * p + 2], held[4 * p + 3]);
      *reinterpret_cast<uint4*>(mine + (((p + 4) ^ (lane & 7)) << 4)) =
          make_uint4(pk[4 * p], pk[4 * p + 1], pk[4 * p + 2], pk[4 * p + 3]);
    }
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA (async proxy)
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tm_out, tile, held_col, row);  // lane 0's row is the first row of the warp's 32-row slab
      bulk_commit_group();
    }
    in_flight = 1;
  }
  __device__ __forceinline__ void flush_tail(__nv_bfloat16* out, int64_t ldo, int row, int M) {
    if (have) {
      store_direct(held, out, ldo, row, held_col, M);
      have = 0;
    }
  }
  __device__ __forceinline__ void finish() {
    if (in_flight && (threadIdx.x & 31) == 0) bulk_wait_group0();
  }
};

// out bf16 [M, ldo] = act(rstd[row] * acc + bias[col])   (norm.stats == nullptr: rstd = 1, plain act(acc + bias)).
// The mean term of a folded LayerNorm needs no work here: the folded weight rows are centred (fold_norm_kernel), so
// sum_i s_i W''[j, i] already equals sum_i (s_i - mean) W'[j, i].
template <int ACT>
struct EpiBiasActBf16 {
  CUtensorMap tm_out;  // box {64 cols, 32 rows} over out, SWIZZLE_128B (TMA store)
  __nv_bfloat16* out;
  int64_t ldo;
  const float* bias;  // nullable (folded bias W beta + b when the input is normalised)
  int M, N;
  RowNorm norm;
  static constexpr int kPasses = 1;
  static constexpr bool kPrefetch = false;
  __host__ __device__ static constexpr int smem_bytes(int epi_warps) { return epi_warps * StagedBf16::kBytesPerWarp; }
  struct State {
    StagedBf16 stage;
    float rstd, rm;
  };
  __device__ __forceinline__ void bind(State& s, uint8_t* smem, int epi_tid) const { s.stage.bind(smem, epi_tid); }
  __device__ __forceinline__ void finish(State& s) const { s.stage.finish(); }
  __device__ __forceinline__ void begin(State& s, int row, int, int) const {
    s.stage.have = 0;
    norm.get(row, M, s.rstd, s.rm);
  }
  __device__ __forceinline__ void end(State& s, int row) const { s.stage.flush_tail(out, ldo, row, M); }
  __device__ __forceinline__ void chunk(State& s, int row, int col0, const float (&v)[32]) const {
    if (col0 >= N) return;  // warp-uniform; N is a multiple of 32 for every encoder GEMM (checked on the host)
    uint32_t packed[16];
    float bv[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = bias ? __ldg(reinterpret_cast<const float4*>(bias + col0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      bv[4 * j] = t.x, bv[4 * j + 1] = t.y, bv[4 * j + 2] = t.z, bv[4 * j + 3] = t.w;
    }
    const float2 rs = splat2(s.rstd);
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float2 ab = fma2(rs, make_float2(v[i], v[i + 1]), make_float2(bv[i], bv[i + 1]));
      if (ACT == ACT_GELU) {
        ab = gelu_erf2(ab);
      } else if (ACT == ACT_RELU) {
        ab.x = fmaxf(ab.x, 0.f);
        ab.y = fmaxf(ab.y, 0.f);
      }
      packed[i >> 1] = pack_bf16x2(ab.x, ab.y);
    }
    s.stage.push(packed, &tm_out, row, col0);
  }
};

// QKV projection: columns [0, 2I) -> qk bf16 [T, 2I]; columns [2I, 3I) -> vt bf16 [I, ldv], V transposed
// (the P*V GEMM wants V^T K-major, i.e. token-contiguous) in the attention kernel's tile-local column order:
// token m of attention tile m / valid_rows sits at column tile * 128 + m % valid_rows, so every tile's
// V^T box starts at a 256-byte aligned column.
struct EpiQKV {
  CUtensorMap tm_qk;  // box {64 cols, 32 rows} over qk, SWIZZLE_128B (TMA store)
  __nv_bfloat16* qk;
  __nv_bfloat16* vt;
  int64_t ldv;
  const float* bias;  // nullable, [3I] (folded: W beta + b)
  int M, I2;          // I2 = 2*I
  int valid_rows;     // tokens per attention tile (spt * L)
  RowNorm norm;       // normalisation of the input rows, applied here (see the file header)
  static constexpr int kPasses = 1;
  static constexpr bool kPrefetch = false;
  __host__ __device__ static constexpr int smem_bytes(int epi_warps) { return epi_warps * StagedBf16::kBytesPerWarp; }
  struct State {
    int vcol;
    StagedBf16 stage;
    float rstd, rm;
  };
  __device__ __forceinline__ void bind(State& s, uint8_t* smem, int epi_tid) const { s.stage.bind(smem, epi_tid); }
  __device__ __forceinline__ void finish(State& s) const { s.stage.finish(); }
  __device__ __forceinline__ void begin(State& s, int row, int, int) const {
    s.vcol = (row / valid_rows) * 128 + row % valid_rows;
    s.stage.have = 0;
    norm.get(row, M, s.rstd, s.rm);
  }
  __device__ __forceinline__ void end(State& s, int row) const { s.stage.flush_tail(qk, I2, row, M); }
  __device__ __forceinline__ void chunk(State& s, int row, int col0, const float (&v)[32]) const {
    if (col0 >= I2 + (I2 >> 1)) return;  // warp-uniform
    float bv[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = bias ? __ldg(reinterpret_cast<const float4*>(bias + col0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      bv[4 * j] = t.x, bv[4 * j + 1] = t.y, bv[4 * j + 2] = t.z, bv[4 * j + 3] = t.w;
    }
    const float rstd = s.rstd;
    if (col0 < I2) {
      uint32_t packed[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2)
        packed[i >> 1] = pack_bf16x2(fmaf(rstd, v[i], bv[i]), fmaf(rstd, v[i + 1], bv[i + 1]));
      s.stage.push(packed, &tm_qk, row, col0);
    } else if (row < M) {
      __nv_bfloat16* dst = vt + static_cast<int64_t>(col0 - I2) * ldv + s.vcol;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        dst[static_cast<int64_t>(i) * ldv] = __float2bfloat16(fmaf(rstd, v[i], bv[i]));  // lanes = consecutive tokens
    }
  }
};

// Residual epilogue of the O-proj and FFN2 GEMMs:   s'[m, n] = acc + bias[n] + R(s[m, n])
//   R = LayerNorm of the residual stream with the producer's (gamma, beta) and the row's (mean, rstd)   (BERT, post-LN)
//   R = identity                                                                                        (T5, pre-norm)
// s lives in fp32 [T, H] and is updated IN PLACE; bf16(s') goes to xb (the next GEMM's A operand) and every (thread, tile)
// writes its partial (sum, sum of squares) of s' into its own slot of stats_out for whoever normalises s' next.
// A thread owns an accumulator ROW, so touching global memory directly would cost one cache line per lane and
// instruction; instead each warp moves its 32-row x 32-column chunks through shared memory with TMA:
//   TMA load  s[32 rows, 32 cols] fp32 -> 4 KB tile (SWIZZLE_128B; the first chunk of a tile is requested before the
//             accumulator wait, the residual of the NEXT tile is pulled into L2 a tile ahead; one tile per warp: the
//             shared memory a second one would take buys the mainloop its 4th ring stage, worth more — measured)
//   in place  thread r rewrites row r of the tile (16-byte pieces XOR-ed with r & 7: conflict-free) and writes bf16(s')
//             into a 2 KB tile (SWIZZLE_64B: 64-byte rows, also conflict-free)
//   TMA store both tiles; rows >= M are clipped by the tensor maps.
struct EpiResidNorm {
  CUtensorMap tm_s;   // fp32 s [T, H], box {32 cols, 32 rows}, SWIZZLE_128B (load and store)
  CUtensorMap tm_xb;  // bf16 xb [T, H], box {32 cols, 32 rows}, SWIZZLE_64B (store)
  const float* bias;  // nullable [N]
  RowNorm norm;       // statistics of s (norm.stats == nullptr: R = identity)
  const float* gamma;  // [N] LayerNorm weight / bias applied to the residual (BERT); unused when norm.stats == nullptr
  const float* beta;
  float* stats_out;  // [T, kStatParts, 2]: slot n_blk * parts + column group <- this thread's (sum, sumsq) of s'
  int parts;         // column groups per tile (epilogue warps / 4); (N / BN) * parts <= kStatParts
  int M, N;
  int bn;            // tile width (BN of the GEMM): geometry of the static tile schedule, for the L2 prefetch below
  static constexpr int kPasses = 1;
  static constexpr bool kPrefetch = true;
  static constexpr bool kRolled = true;  // one copy of chunk() in the kernel (gemm.cuh)
  static constexpr int kF32Tile = 4096, kBf16Tile = 2048, kPerWarp = kF32Tile + kBf16Tile;
  static constexpr int kMaxN = 1024;     // per-column constants staged in shared memory: 2 * kMaxN floats
  __host__ __device__ static constexpr int smem_bytes(int epi_warps) { return epi_warps * kPerWarp + 1024 + 2 * kMaxN * 4; }
  struct State {
    uint8_t* f32tile;   // [4096]
    uint8_t* bf16tile;  // [2048]
    uint64_t* bars;     // [1]
    const float* gsm;   // [N] gamma of the residual's LayerNorm (1 when there is none)
    const float* bsm;   // [N] bias + beta
    uint32_t parity;    // phase parity of bars[0]
    int slot;           // statistics slot of this (tile, column group)
    int group;          // column group of this warp
    float rstd, rm, rsum, rsq;
  };
  __device__ __forceinline__ void bind(State& s, uint8_t* smem, int epi_tid) const {
    const int w = epi_tid >> 5, nw = blockDim.x / 32 - kGemmProducerThreads / 32;
    s.f32tile = smem + w * kF32Tile;                        // 1024-byte aligned (swizzle atoms)
    s.bf16tile = smem + nw * kF32Tile + w * kBf16Tile;       // 512-byte aligned is enough for SWIZZLE_64B
    s.bars = reinterpret_cast<uint64_t*>(smem + nw * kPerWarp) + w;
    s.parity = 0;
    s.group = w >> 2;
    if ((epi_tid & 31) == 0) {
      mbar_init(&s.bars[0], 1);
      fence_barrier_init();
    }
    // per-column constants, once per CTA: every lane of a chunk reads the same column's gamma / (bias + beta), so they
    // come from shared memory as broadcasts instead of 24 dependent global loads per chunk and thread
    float* gs = reinterpret_cast<float*>(smem + nw * kPerWarp + 1024);
    float* bs = gs + kMaxN;
    const bool ln = norm.stats != nullptr;
    for (int c = epi_tid; c < N; c += nw * 32) {
      gs[c] = ln ? gamma[c] : 1.f;
      bs[c] = (bias ? bias[c] : 0.f) + (ln ? beta[c] : 0.f);
    }
    s.gsm = gs;
    s.bsm = bs;
    named_bar_sync(1, nw * 32);  // the epilogue warps only (the producer / MMA warps never join barrier 1)
  }
  __device__ __forceinline__ void finish(State&) const {
    if ((threadIdx.x & 31) == 0) bulk_wait_group0();
  }
  __device__ __forceinline__ void begin(State& s, int row, int m_blk, int n_blk) const {
    norm.get(row, M, s.rstd, s.rm);
    s.rsum = 0.f;
    s.rsq = 0.f;
    s.slot = n_blk * parts + s.group;
    // The residual of the tile this CTA processes NEXT (static schedule: tile + gridDim.x, n fastest) is pulled into L2
    // now, a whole tile ahead: its TMA loads then cost an L2 hit instead of an exposed HBM round trip per chunk.
    if (bn > 0 && (threadIdx.x & 31) == 0) {  // bn == 0: CTA-pair schedule, no look-ahead
      const int num_n = (N + bn - 1) / bn, num_m = (M + kBlockM - 1) / kBlockM;
      const int next = m_blk * num_n + n_blk + static_cast<int>(gridDim.x);
      if (next < num_m * num_n) {
        const int nm = next / num_n, nn = next - nm * num_n;
        const int row0 = row + (nm - m_blk) * kBlockM;  // lane 0's row = first row of the warp's slab
        const int cols = bn / parts;                     // columns per column group
#pragma unroll 1
        for (int c = 0; c < cols; c += 32) tma_prefetch_l2_2d(&tm_s, nn * bn + s.group * cols + c, row0);
      }
    }
  }
  // request the residual of chunk [row0 .. row0+32) x [col0 .. col0+32) into tile `b` (lane 0; the tile's last TMA store
  // must have finished READING it: the caller waits for that)
  __device__ __forceinline__ void request(State& s, int row0, int col0) const {
    bulk_wait_group_read0();  // the tile's last TMA store has finished reading it
    mbar_arrive_expect_tx(&s.bars[0], kF32Tile);
    tma_load_2d(s.f32tile, &tm_s, &s.bars[0], col0, row0);
  }
  __device__ __forceinline__ void prefetch(State& s, int row, int col0) const {
    if (col0 >= N) return;  // warp-uniform
    if ((threadIdx.x & 31) == 0) request(s, row, col0);  // lane 0's row = first row of the warp's 32-row slab
  }
  __device__ __forceinline__ void end(State& s, int row) const {
    if (row < M)
      *reinterpret_cast<float2*>(stats_out + (static_cast<int64_t>(row) * kStatParts + s.slot) * 2) = make_float2(s.rsum, s.rsq);
  }
  __device__ __forceinline__ void chunk(State& s, int row, int col0, const float (&v)[32], int next_col0) const {
    if (col0 >= N) return;  // warp-uniform
    const int lane = threadIdx.x & 31;
    mbar_wait_warp(&s.bars[0], s.parity, 20);
    s.parity ^= 1u;
    uint8_t* mine = s.f32tile + lane * 128;
    uint8_t* mineb = s.bf16tile + lane * 64;
    // R = (s - mean) * rstd * gamma + beta = s * (rstd * gamma) + (beta - rstd * mean * gamma); without a LayerNorm on the
    // residual (T5) gamma = 1, rstd = 1, mean = 0:  s' = acc + s * (rstd g) + ((bias + beta) - rstd mean g)
    const float rstd = s.rstd, nrm = -s.rm;
    const float4* gp = reinterpret_cast<const float4*>(s.gsm + col0);
    const float4* bp = reinterpret_cast<const float4*>(s.bsm + col0);
    float rsum = 0.f, rsq = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      float4* slot = reinterpret_cast<float4*>(mine + ((p ^ (lane & 7)) << 4));
      const float4 r = *slot;
      const float4 g = gp[p], bb = bp[p];
      float4 o;
      o.x = fmaf(r.x, rstd * g.x, fmaf(nrm, g.x, bb.x)) + v[4 * p];
      o.y = fmaf(r.y, rstd * g.y, fmaf(nrm, g.y, bb.y)) + v[4 * p + 1];
      o.z = fmaf(r.z, rstd * g.z, fmaf(nrm, g.z, bb.z)) + v[4 * p + 2];
      o.w = fmaf(r.w, rstd * g.w, fmaf(nrm, g.w, bb.w)) + v[4 * p + 3];
      *slot = o;
      rsum += (o.x + o.y) + (o.z + o.w);
      rsq = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, rsq))));
      // bf16 tile: 64-byte rows, 16-byte piece (p >> 1) XOR-ed with (row >> 1) & 3 (SWIZZLE_64B)
      uint2* hb = reinterpret_cast<uint2*>(mineb + ((((p >> 1) ^ ((lane >> 1) & 3)) << 4) | ((p & 1) << 3)));
      *hb = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
    if (row < M) {  // rows beyond M hold zero-filled residuals + garbage-free accumulators, but must not count
      s.rsum += rsum;
      s.rsq += rsq;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&tm_s, s.f32tile, col0, row);
      tma_store_2d(&tm_xb, s.bf16tile, col0, row);
      bulk_commit_group();
      if (next_col0 >= 0 && next_col0 < N) request(s, row, next_col0);  // waits for the stores above to release the tile
    }
  }
};

// ===================================================================================================
// row-wise kernels: one warp per token row, H % 128 == 0, H <= 1024
// ===================================================================================================
constexpr int kMaxVec = 8;  // float4 per lane: 8 * 4 * 32 = 1024 columns

template <bool RMS>
__device__ __forceinline__ void norm_row(float4 (&v)[kMaxVec], int nvec, int H, float eps) {
  float s = 0.f;
  if (!RMS) {
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j)
      if (j < nvec) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  const float mean = RMS ? 0.f : s / H;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nvec) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / H + eps);
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nvec) {
      v[j].x = (v[j].x - mean) * rstd;
      v[j].y = (v[j].y - mean) * rstd;
      v[j].z = (v[j].z - mean) * rstd;
      v[j].w = (v[j].w - mean) * rstd;
    }
}

__device__ __forceinline__ void affine_store(const float4 (&v)[kMaxVec], int nvec, int lane, const float* gamma,
                                             const float* beta, float* out_f32, __nv_bfloat16* out_bf16) {
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nvec) {
      const int c = (j * 32 + lane) * 4;
      const float4 g = *reinterpret_cast<const float4*>(gamma + c);
      float4 y = make_float4(v[j].x * g.x, v[j].y * g.y, v[j].z * g.z, v[j].w * g.w);
      if (beta) {
        const float4 b = *reinterpret_cast<const float4*>(beta + c);
        y.x += b.x, y.y += b.y, y.z += b.z, y.w += b.w;
      }
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + c) = y;
      if (out_bf16) *reinterpret_cast<uint2*>(out_bf16 + c) = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
    }
}

// sum = h (+ add); y = Norm(sum) * gamma (+ beta).  Writes sum back to h when STORE_SUM (T5's pre-norm residual
// stream), y as fp32 (nullable, may alias h: BERT's post-LN stream / T5's final norm) and as bf16 (nullable: the
// next GEMM's A operand).  `add` is the bf16 output of the preceding O-proj / FFN2 GEMM: doing the residual add
// here keeps those GEMM epilogues store-only — with the fp32 residual read in the epilogue they were bound by
// exposed DRAM latency (tensor pipe 20 % on O-proj) — and this kernel streams at HBM speed anyway.
template <bool RMS, bool STORE_SUM>
__global__ void __launch_bounds__(128) norm_kernel(float* h, const __nv_bfloat16* __restrict__ add, const float* gamma,
                                                   const float* beta, float eps, int T, int H, float* out_f32,
                                                   __nv_bfloat16* out_bf16) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  const int nvec = H >> 7;
  float4 v[kMaxVec];
  float* src = h + static_cast<int64_t>(row) * H;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nvec) v[j] = *reinterpret_cast<const float4*>(src + (j * 32 + lane) * 4);
  if (add) {
    const __nv_bfloat16* a = add + static_cast<int64_t>(row) * H;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j)
      if (j < nvec) {
        const uint2 raw = *reinterpret_cast<const uint2*>(a + (j * 32 + lane) * 4);
        const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
        const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
        v[j].x += lo.x, v[j].y += lo.y, v[j].z += hi.x, v[j].w += hi.y;
        if (STORE_SUM) *reinterpret_cast<float4*>(src + (j * 32 + lane) * 4) = v[j];
      }
  }
  norm_row<RMS>(v, nvec, H, eps);
  affine_store(v, nvec, lane, gamma, beta, out_f32 ? out_f32 + static_cast<int64_t>(row) * H : nullptr,
               out_bf16 ? out_bf16 + static_cast<int64_t>(row) * H : nullptr);
}

// Row statistics of a freshly embedded row: slot 0 of the row's kStatParts partials holds (sum, sum of squares), the
// other slots zeros (the GEMM epilogues that normalise the row sum all slots, see RowNorm).
__device__ __forceinline__ void store_row_stats(const float4 (&v)[kMaxVec], int nvec, int lane, float* stats_row) {
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nvec) {
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      q = fmaf(v[j].x, v[j].x, fmaf(v[j].y, v[j].y, fmaf(v[j].z, v[j].z, fmaf(v[j].w, v[j].w, q))));
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane < kStatParts) reinterpret_cast<float2*>(stats_row)[lane] = lane == 0 ? make_float2(s, q) : make_float2(0.f, 0.f);
}

__device__ __forceinline__ void raw_store(const float4 (&v)[kMaxVec], int nvec, int lane, float* out_f32,
                                          __nv_bfloat16* out_bf16) {
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nvec) {
      const int c = (j * 32 + lane) * 4;
      *reinterpret_cast<float4*>(out_f32 + c) = v[j];
      *reinterpret_cast<uint2*>(out_bf16 + c) = make_uint2(pack_bf16x2(v[j].x, v[j].y), pack_bf16x2(v[j].z, v[j].w));
    }
}

// BERT embeddings (modeling_bert.py:53-112): s = word[id] + type[tt] + pos[l], UN-normalised (fp32 + bf16) with its row
// statistics; the embedding LayerNorm is applied by the first layer's QKV / O-proj epilogues like every other LayerNorm.
__global__ void __launch_bounds__(128) bert_embed_kernel(const int64_t* ids, const int64_t* tts, const float* word,
                                                         const float* type, const float* pos, int T, int L, int H, int vocab,
                                                         int type_vocab, float* out_f32, __nv_bfloat16* out_bf16,
                                                         float* stats) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  const int nvec = H >> 7;
  int64_t id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  int64_t tt = tts ? tts[row] : 0;
  tt = tt < 0 ? 0 : (tt >= type_vocab ? type_vocab - 1 : tt);
  const int l = row % L;
  float4 v[kMaxVec];
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nvec) {
      const int c = (j * 32 + lane) * 4;
      const float4 a = *reinterpret_cast<const float4*>(word + id * H + c);
      const float4 b = *reinterpret_cast<const float4*>(type + tt * H + c);
      const float4 p = *reinterpret_cast<const float4*>(pos + static_cast<int64_t>(l) * H + c);
      v[j] = make_float4((a.x + b.x) + p.x, (a.y + b.y) + p.y, (a.z + b.z) + p.z, (a.w + b.w) + p.w);
    }
  raw_store(v, nvec, lane, out_f32 + static_cast<int64_t>(row) * H, out_bf16 + static_cast<int64_t>(row) * H);
  store_row_stats(v, nvec, lane, stats + static_cast<int64_t>(row) * (2 * kStatParts));
}

// T5: h = embed_tokens[id] (no position embedding, no scaling; modeling_t5.py:682,734), fp32 + bf16 + row statistics
__global__ void __launch_bounds__(128) t5_embed_kernel(const int64_t* ids, const float* emb, int T, int H, int vocab,
                                                       float* out_f32, __nv_bfloat16* out_bf16, float* stats) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  const int nvec = H >> 7;
  int64_t id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  float4 v[kMaxVec];
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nvec) v[j] = *reinterpret_cast<const float4*>(emb + id * H + (j * 32 + lane) * 4);
  raw_store(v, nvec, lane, out_f32 + static_cast<int64_t>(row) * H, out_bf16 + static_cast<int64_t>(row) * H);
  store_row_stats(v, nvec, lane, stats + static_cast<int64_t>(row) * (2 * kStatParts));
}

// Weight folding (once, at om_encoder_finalize):  Wf[j, i] = bf16(W[j, i] * gamma[i] - centre * mean_i(W[j, i] * gamma[i]))
// and bfold[j] = b[j] + sum_i W[j, i] * beta[i] (fp32, un-centred W).  With centred rows (LayerNorm) the mean term of the
// folded normalisation vanishes identically:  sum_i s_i Wf[j, i] = sum_i (s_i - mean(s)) W[j, i] gamma[i]  (up to the bf16
// rounding of Wf, ~1e-3 of a weight: below the bf16 rounding of the GEMM's output).  RMSNorm (T5): centre = 0, beta = none.
// One warp per output row j.
__global__ void __launch_bounds__(256) fold_norm_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ bias,
                                                        int N, int K, int centre, __nv_bfloat16* __restrict__ Wf,
                                                        float* bfold) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= N) return;
  const float* w = W + static_cast<int64_t>(j) * K;
  float c = 0.f, b = 0.f;
  for (int i = lane; i < K; i += 32) {
    const float wi = w[i];
    c = fmaf(wi, gamma[i], c);
    if (beta) b = fmaf(wi, beta[i], b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c += __shfl_xor_sync(0xffffffffu, c, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const float shift = centre ? c / static_cast<float>(K) : 0.f;
  for (int i = lane; i < K; i += 32) Wf[static_cast<int64_t>(j) * K + i] = __float2bfloat16(fmaf(w[i], gamma[i], -shift));
  if (lane == 0 && bfold) bfold[j] = b + (bias ? bias[j] : 0.f);
}

__global__ void keymask_kernel(const int64_t* attn_mask, float* kmask, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T) kmask[i] = attn_mask[i] != 0 ? 0.f : __int_as_float(0xff800000);
}

// ===================================================================================================
// attention: one CTA per (tile of 128 token rows = spt whole sequences, head)
// ===================================================================================================
struct AttnParams {
  int T, L, spt, I, Tvalid_rows;  // Tvalid_rows = spt * L: rows of the tile that belong to it
  float scale_log2;               // softmax scale * log2(e)
  const float* kmask;             // [T] 0 / -inf
  const float* relbias_log2;      // nullable [heads, 2*kMaxL-1], already multiplied by log2(e)
  __nv_bfloat16* ctx;             // [T, I]
};

// P (32 KB) overwrites the Q and K tiles, which are dead once S = Q K^T has completed; O overwrites the first
// 64 TMEM columns of S, dead once the softmax has read it: 52 KB smem + 128 TMEM columns per CTA -> 4 CTAs/SM.
constexpr int kAttnSmemQ = 0, kAttnSmemK = 16384, kAttnSmemV = 32768, kAttnSmemP = 0;
constexpr int kAttnSmemMisc = 49152;  // kb[128] f32, rel[256] f32, barriers, tmem slot
constexpr int kAttnTmemCols = 128;
constexpr int kAttnSmemBytes = kAttnSmemMisc + 512 + 1024 + 64 + 1024;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(128, 4)
attn_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmVt, AttnParams p, int n_tiles,
            int n_heads) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_kb = reinterpret_cast<float*>(smem + kAttnSmemMisc);
  float* s_rel = s_kb + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rel + 256);  // [0] load, [1] S ready, [2] O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmVt);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, kAttnTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // Persistent CTA (grid = 4 per SM): barriers, tensor memory and descriptors are set up once; work items are
  // (tile, head) pairs taken head-fastest, so that the CTAs running at the same time read the same token rows of qk
  // (neighbouring 128-byte slices of one DRAM page instead of one slice from each of many pages).
  const int n_items = n_tiles * n_heads;
  uint32_t par = 0;
#pragma unroll 1
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, par ^= 1u) {
  const int tile = item / n_heads, head = item - tile * n_heads;
  const int row0 = tile * p.Tvalid_rows;  // first token of this tile
  if (tid == 0) {  // shared memory and tensor memory of the previous item are free (trailing barrier): loads go out first
    mbar_arrive_expect_tx(&bars[0], 3 * 16384);
    tma_load_2d(smem + kAttnSmemQ, &tmQK, &bars[0], head * kHeadDim, row0);
    tma_load_2d(smem + kAttnSmemK, &tmQK, &bars[0], p.I + head * kHeadDim, row0);
    tma_load_2d(smem + kAttnSmemV, &tmVt, &bars[0], tile * 128, head * kHeadDim);
    tma_load_2d(smem + kAttnSmemV + 8192, &tmVt, &bars[0], tile * 128 + 64, head * kHeadDim);
  }
  {
    const int tok = row0 + tid;
    // key validity as 4 x 32-bit words (bit c%32 of word c/32): one ballot per warp
    const bool key_ok = (tid < p.Tvalid_rows && tok < p.T) && p.kmask[tok] == 0.f;
    const unsigned bits = __ballot_sync(0xffffffffu, key_ok);
    if ((tid & 31) == 0) reinterpret_cast<uint32_t*>(s_kb)[warp] = bits;
    if (p.relbias_log2) {
      for (int i = tid; i < 2 * kMaxL - 1; i += 128) s_rel[i] = p.relbias_log2[head * (2 * kMaxL - 1) + i];
    }
  }
  __syncthreads();  // key bits / relative bias of this item are in place

  if (tid == 0) {
    mbar_wait(&bars[0], par, 10);
    tc_fence_after_sync();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    const uint32_t qa = smem_u32(smem + kAttnSmemQ), ka = smem_u32(smem + kAttnSmemK);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16_ss(tmem_base, umma_smem_desc(qa + k * 32, kDescKMajorSW128),
                   umma_smem_desc(ka + k * 32, kDescKMajorSW128), idesc_s, k != 0 ? 1u : 0u);
    umma_commit(&bars[1]);
  }
  mbar_wait_warp(&bars[1], par, 11);
  tc_fence_after_sync();

  // ---- softmax: thread r owns query row r of the tile ----
  const int r = tid;
  const bool row_valid = r < p.Tvalid_rows && row0 + r < p.T;
  const int seq = r / p.L;
  const int c_lo = row_valid ? seq * p.L : 0, c_hi = row_valid ? c_lo + p.L : 0;
  // this row may attend key c iff the key is valid AND belongs to the row's own sequence [c_lo, c_hi)
  uint32_t allow[4];
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int lo = c_lo - 32 * w, hi = c_hi - 32 * w;  // range relative to word w
    const uint32_t ge = lo <= 0 ? 0xffffffffu : (lo >= 32 ? 0u : (0xffffffffu << lo));
    const uint32_t lt = hi >= 32 ? 0xffffffffu : (hi <= 0 ? 0u : (0xffffffffu >> (32 - hi)));
    allow[w] = reinterpret_cast<const uint32_t*>(s_kb)[w] & ge & lt;
  }
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const bool has_rel = p.relbias_log2 != nullptr;
  float m = __int_as_float(0xff800000);
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    uint32_t raw[32];
    tmem_ld_32x32b_x32(taddr + c4 * 32, raw);
    tmem_ld_wait();
    const uint32_t aw = allow[c4];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = c4 * 32 + i;
      float s = __uint_as_float(raw[i]) * p.scale_log2;
      if (has_rel) s += s_rel[c - r + (kMaxL - 1)];
      if (!(aw & (1u << i))) s = __int_as_float(0xff800000);
      m = fmaxf(m, s);
    }
  }
  const bool dead = !(m > __int_as_float(0xff800000));  // every key masked (padding row)
  const float mm = dead ? 0.f : m;
  float sum = 0.f;
  uint8_t* sP = smem + kAttnSmemP;
  const float neg_mm = -mm;
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    uint32_t raw[32];
    tmem_ld_32x32b_x32(taddr + c4 * 32, raw);
    tmem_ld_wait();
    const uint32_t aw = dead ? 0u : allow[c4];
    uint32_t packed[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float pv[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = c4 * 32 + i + u;
        float s = fmaf(__uint_as_float(raw[i + u]), p.scale_log2, neg_mm);
        if (has_rel) s += s_rel[c - r + (kMaxL - 1)];
        pv[u] = (aw & (1u << (i + u))) ? ex2_approx(s) : 0.f;
      }
      sum += pv[0] + pv[1];
      packed[i >> 1] = pack_bf16x2(pv[0], pv[1]);
    }
    // P[r, c4*32 .. +32) -> K-major SWIZZLE_128B layout: block = c / 64, 16-B chunk j XOR (r & 7)
    uint8_t* blk = sP + (c4 >> 1) * 16384 + r * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int chunk = ((c4 & 1) * 4 + j) ^ (r & 7);
      *reinterpret_cast<uint4*>(blk + chunk * 16) =
          make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
    }
  }
  const float inv = sum > 0.f ? 1.0f / sum : 0.f;
  fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before_sync();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after_sync();
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64);
    const uint32_t pa = smem_u32(sP), va = smem_u32(smem + kAttnSmemV);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      umma_bf16_ss(tmem_base, umma_smem_desc(pa + (k >> 2) * 16384 + (k & 3) * 32, kDescKMajorSW128),
                   umma_smem_desc(va + (k >> 2) * 8192 + (k & 3) * 32, kDescKMajorSW128), idesc_o, k != 0 ? 1u : 0u);
    umma_commit(&bars[2]);
  }
  mbar_wait_warp(&bars[2], par, 12);
  tc_fence_after_sync();
#pragma unroll 1
  for (int c2 = 0; c2 < 2; ++c2) {
    uint32_t raw[32];
    tmem_ld_32x32b_x32(taddr + c2 * 32, raw);
    tmem_ld_wait();
    if (row_valid) {
      uint4* dst = reinterpret_cast<uint4*>(p.ctx + static_cast<int64_t>(row0 + r) * p.I + head * kHeadDim + c2 * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          w[u] = pack_bf16x2(__uint_as_float(raw[8 * j + 2 * u]) * inv, __uint_as_float(raw[8 * j + 2 * u + 1]) * inv);
        dst[j] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  // the next item's TMA loads overwrite Q / K / V and its first MMA overwrites S: every thread must be done with O
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  }  // item loop
  if (warp == 0) tmem_dealloc(tmem_base, kAttnTmemCols);
}

// ---------------------------------------------------------------------------------------------------
// Sequences longer than one tile (L = 256 / 384 / 512, multiples of 128): one CTA per (128-row query tile, head)
// loops over the sequence's 128-key tiles with an online softmax:  S_j = Q K_j^T (tcgen05 -> TMEM) -> running
// max / sum in registers (thread = query row) -> P_j (bf16, swizzled smem) -> O_j = P_j V_j (tcgen05 -> TMEM) ->
// acc = acc * alpha + O_j in registers.  Q stays in smem; K_j / V_j are re-loaded by TMA per iteration.
// ---------------------------------------------------------------------------------------------------
constexpr int kAttnLongSmemQ = 0, kAttnLongSmemK = 16384, kAttnLongSmemV = 32768, kAttnLongSmemP = 49152;
constexpr int kAttnLongSmemMisc = 81920;  // rel[1024] f32, key bits [4 tiles x 4 words], barriers, tmem slot
constexpr int kAttnLongTmemCols = 256;    // S_j: columns [0, 128); O_j: columns [128, 192)
constexpr int kAttnLongSmemBytes = kAttnLongSmemMisc + 4096 + 64 + 64 + 64 + 1024;

__global__ void __launch_bounds__(128, 2)
attn_long_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmVt, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_rel = reinterpret_cast<float*>(smem + kAttnLongSmemMisc);
  uint32_t* s_kb = reinterpret_cast<uint32_t*>(s_rel + 1024);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_kb + 16);  // [0] loads, [1] S ready, [2] O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = blockIdx.x, head = blockIdx.y;
  const int nk = p.L / 128;                 // key tiles per sequence
  const int row0 = qt * 128;                // first token of this query tile
  const int kt0 = (qt / nk) * nk;           // first tile of the sequence this query tile belongs to
  const int qpos0 = (qt - kt0) * 128;       // position of the tile's first query inside its sequence

  if (tid == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmVt);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, kAttnLongTmemCols);
    tmem_relinquish();
  }
  for (int j = 0; j < nk; ++j) {  // key validity of every key tile of the sequence: one ballot per warp and tile
    const int tok = (kt0 + j) * 128 + tid;
    const bool key_ok = tok < p.T && p.kmask[tok] == 0.f;
    const unsigned bits = __ballot_sync(0xffffffffu, key_ok);
    if ((tid & 31) == 0) s_kb[j * 4 + warp] = bits;
  }
  if (p.relbias_log2) {
    for (int i = tid; i < 2 * kMaxLongL - 1; i += 128) s_rel[i] = p.relbias_log2[head * (2 * kMaxLongL - 1) + i];
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  const int r = tid;
  const bool row_valid = row0 + r < p.T;
  const bool has_rel = p.relbias_log2 != nullptr;
  const int rel0 = (kMaxLongL - 1) - (qpos0 + r);  // + key position = index into s_rel
  float m_run = __int_as_float(0xff800000), sum = 0.f;
  float acc[kHeadDim];
#pragma unroll
  for (int i = 0; i < kHeadDim; ++i) acc[i] = 0.f;
  uint8_t* sP = smem + kAttnLongSmemP;

#pragma unroll 1
  for (int j = 0; j < nk; ++j) {
    const uint32_t par = static_cast<uint32_t>(j & 1);
    if (tid == 0) {
      // K / V of the previous iteration are dead: S_{j-1} and O_{j-1} have completed (every thread waited on them)
      mbar_arrive_expect_tx(&bars[0], (j == 0 ? 3 : 2) * 16384);
      if (j == 0) tma_load_2d(smem + kAttnLongSmemQ, &tmQK, &bars[0], head * kHeadDim, row0);
      tma_load_2d(smem + kAttnLongSmemK, &tmQK, &bars[0], p.I + head * kHeadDim, (kt0 + j) * 128);
      tma_load_2d(smem + kAttnLongSmemV, &tmVt, &bars[0], (kt0 + j) * 128, head * kHeadDim);
      tma_load_2d(smem + kAttnLongSmemV + 8192, &tmVt, &bars[0], (kt0 + j) * 128 + 64, head * kHeadDim);
      mbar_wait(&bars[0], par, 13);
      tc_fence_after_sync();
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
      const uint32_t qa = smem_u32(smem + kAttnLongSmemQ), ka = smem_u32(smem + kAttnLongSmemK);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tmem_base, umma_smem_desc(qa + k * 32, kDescKMajorSW128),
                     umma_smem_desc(ka + k * 32, kDescKMajorSW128), idesc_s, k != 0 ? 1u : 0u);
      umma_commit(&bars[1]);
    }
    mbar_wait_warp(&bars[1], par, 14);
    tc_fence_after_sync();

    // ---- running max over this key tile ----
    float m_j = __int_as_float(0xff800000);
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(taddr + c4 * 32, raw);
      tmem_ld_wait();
      const uint32_t aw = row_valid ? s_kb[j * 4 + c4] : 0u;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float sc = __uint_as_float(raw[i]) * p.scale_log2;
        if (has_rel) sc += s_rel[rel0 + j * 128 + c4 * 32 + i];
        if (!(aw & (1u << i))) sc = __int_as_float(0xff800000);
        m_j = fmaxf(m_j, sc);
      }
    }
    const float m_new = fmaxf(m_run, m_j);
    const bool dead = !(m_new > __int_as_float(0xff800000));  // no allowed key seen so far
    const float mm = dead ? 0.f : m_new;
    const float alpha = (m_run > __int_as_float(0xff800000)) ? ex2_approx(m_run - mm) : 0.f;
    m_run = m_new;
    sum *= alpha;
    const float neg_mm = -mm;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(taddr + c4 * 32, raw);
      tmem_ld_wait();
      const uint32_t aw = (row_valid && !dead) ? s_kb[j * 4 + c4] : 0u;
      uint32_t packed[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float pv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float sc = fmaf(__uint_as_float(raw[i + u]), p.scale_log2, neg_mm);
          if (has_rel) sc += s_rel[rel0 + j * 128 + c4 * 32 + i + u];
          pv[u] = (aw & (1u << (i + u))) ? ex2_approx(sc) : 0.f;
        }
        sum += pv[0] + pv[1];
        packed[i >> 1] = pack_bf16x2(pv[0], pv[1]);
      }
      uint8_t* blk = sP + (c4 >> 1) * 16384 + r * 128;  // K-major SWIZZLE_128B, see attn_kernel
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int chunk = ((c4 & 1) * 4 + q4) ^ (r & 7);
        *reinterpret_cast<uint4*>(blk + chunk * 16) =
            make_uint4(packed[4 * q4], packed[4 * q4 + 1], packed[4 * q4 + 2], packed[4 * q4 + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after_sync();
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64);
      const uint32_t pa = smem_u32(sP), va = smem_u32(smem + kAttnLongSmemV);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16_ss(tmem_base + 128, umma_smem_desc(pa + (k >> 2) * 16384 + (k & 3) * 32, kDescKMajorSW128),
                     umma_smem_desc(va + (k >> 2) * 8192 + (k & 3) * 32, kDescKMajorSW128), idesc_o, k != 0 ? 1u : 0u);
      umma_commit(&bars[2]);
    }
    mbar_wait_warp(&bars[2], par, 15);
    tc_fence_after_sync();
#pragma unroll
    for (int c2 = 0; c2 < 2; ++c2) {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(taddr + 128 + c2 * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[c2 * 32 + i] = fmaf(acc[c2 * 32 + i], alpha, __uint_as_float(raw[i]));
    }
    tc_fence_before_sync();  // the next iteration's MMAs overwrite S (after this thread's reads above)
  }

  const float inv = sum > 0.f ? 1.0f / sum : 0.f;
  if (row_valid) {
    uint4* dst = reinterpret_cast<uint4*>(p.ctx + static_cast<int64_t>(row0 + r) * p.I + head * kHeadDim);
#pragma unroll
    for (int q8 = 0; q8 < 8; ++q8) {
      uint32_t w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) w[u] = pack_bf16x2(acc[8 * q8 + 2 * u] * inv, acc[8 * q8 + 2 * u + 1] * inv);
      dst[q8] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kAttnLongTmemCols);
  }
}

// ===================================================================================================
// pooling / head / normalise (fp32)
// ===================================================================================================
// pooled[b, :] = hidden[b, 0, :]  or  sum_l hidden[b,l,:] m[b,l] / clamp(sum_l m[b,l], 1e-9)
__global__ void pool_kernel(const float* hidden, const int64_t* mask, int L, int H, int mean, float* pooled) {
  const int b = blockIdx.x;
  const float* src = hidden + static_cast<int64_t>(b) * L * H;
  if (!mean) {
    for (int c = threadIdx.x; c < H; c += blockDim.x) pooled[static_cast<int64_t>(b) * H + c] = src[c];
    return;
  }
  __shared__ float sm[kMaxLongL];
  for (int l = threadIdx.x; l < L; l += blockDim.x) sm[l] = mask[static_cast<int64_t>(b) * L + l] != 0 ? 1.f : 0.f;
  __syncthreads();
  float cnt = 0.f;
  for (int l = 0; l < L; ++l) cnt += sm[l];
  const float denom = fmaxf(cnt, 1e-9f);
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc += src[static_cast<int64_t>(l) * H + c] * sm[l];
    pooled[static_cast<int64_t>(b) * H + c] = acc / denom;
  }
}

// out[b, o] = sum_i in[b, i] * W[o, i]  (bias-free LinearHead).  One warp per (o, group of 8 rows).
__global__ void __launch_bounds__(256) head_kernel(const float* in, const float* W, int B, int Hin, int Hout, float* out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int groups = (B + 7) / 8;
  if (warp >= Hout * groups) return;
  const int o = warp % Hout, g = warp / Hout;
  const float* w = W + static_cast<int64_t>(o) * Hin;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = lane; i < Hin; i += 32) {
    const float wv = __ldg(w + i);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int b = g * 8 + u;
      if (b < B) acc[u] = fmaf(in[static_cast<int64_t>(b) * Hin + i], wv, acc[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], s);
    const int b = g * 8 + u;
    if (lane == 0 && b < B) out[static_cast<int64_t>(b) * Hout + o] = acc[u];
  }
}

// F.normalize(x, dim=1) = x / max(||x||_2, 1e-12) (optional) and store as fp32 / bf16 with a row pitch
template <typename OutT>
__global__ void finish_reps_kernel(const float* in, int B, int D, int normalize, OutT* out, int64_t pitch) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* src = in + static_cast<int64_t>(b) * D;
  float scale = 1.f;
  if (normalize) {
    float q = 0.f;
    for (int i = lane; i < D; i += 32) q = fmaf(src[i], src[i], q);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) q += __shfl_xor_sync(0xffffffffu, q, s);
    scale = 1.0f / fmaxf(sqrtf(q), 1e-12f);
  }
  for (int i = lane; i < D; i += 32) out[static_cast<int64_t>(b) * pitch + i] = static_cast<OutT>(src[i] * scale);
}

__global__ void f32_to_bf16_kernel(const float* src, __nv_bfloat16* dst, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

// T5 relative-position bucket (bidirectional), same fp32 arithmetic as modeling_t5.py:188-234
static int t5_bucket(int rel, int num_buckets, int max_distance) {
  const int nb = num_buckets / 2;
  const int out = rel > 0 ? nb : 0;
  const int n = rel < 0 ? -rel : rel;
  const int max_exact = nb / 2;
  if (n < max_exact) return out + n;
  const float v = logf(static_cast<float>(n) / static_cast<float>(max_exact)) /
                  static_cast<float>(log(static_cast<double>(max_distance) / static_cast<double>(max_exact))) *
                  static_cast<float>(nb - max_exact);
  int large = max_exact + static_cast<int>(v);
  if (large > nb - 1) large = nb - 1;
  return out + large;
}

}  // namespace om

using namespace om;

// ===================================================================================================
// host side
// ===================================================================================================
struct LayerW {
  __nv_bfloat16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;  // [3I,H] [H,I] [F,H] [H,F]
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;          // BERT only
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;  // BERT: post-attn / post-ffn LN
                                                                                 // T5  : pre-attn / pre-ffn RMS (g only)
  // normalisation folded into the consuming GEMM (om_encoder_finalize): wqkv / w1 hold W diag(gamma) in bf16 (rows
  // centred for LayerNorm), b*_fold = W beta + b (BERT); fp32 staging freed after folding
  float *wqkv_f32 = nullptr, *w1_f32 = nullptr;
  float *bqkv_fold = nullptr, *b1_fold = nullptr;
};

struct om_encoder {
  om_encoder_desc d;
  int I = 0;  // heads * 64
  std::vector<LayerW> layers;
  float *word = nullptr, *pos = nullptr, *type = nullptr, *emb_g = nullptr, *emb_b = nullptr;  // BERT embeddings
  float* final_g = nullptr;                                                                    // T5 final RMSNorm
  float* rel_w = nullptr;         // T5 [buckets, heads] (host copy kept in rel_host)
  std::vector<float> rel_host;
  float* relbias_log2 = nullptr;  // [heads, 255]
  float* relbias_long_log2 = nullptr;  // [heads, 1023] (sequences longer than one tile)
  float* head_w = nullptr;        // [head_out, H]
  std::vector<std::string> missing;
  std::vector<std::pair<std::string, bool>> required;  // name -> set?
  bool finalized = false;
  // workspace
  int Tmax = 0, Tld = 0;
  float *h = nullptr, *kmask = nullptr, *pooled = nullptr, *headed = nullptr;
  __nv_bfloat16 *xb = nullptr, *qk = nullptr, *vt = nullptr, *ctx = nullptr, *inter = nullptr;
  float* stats[2] = {nullptr, nullptr};  // [Tmax, kStatParts, 2] row statistics of the residual stream (ping-pong)
  std::vector<void*> allocs;
};

namespace {

template <typename T>
int dev_alloc(om_encoder* e, T** p, size_t count) {
  void* q = nullptr;
  OM_CUDA(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  e->allocs.push_back(q);
  *p = static_cast<T*>(q);
  return 0;
}

void require(om_encoder* e, const std::string& name) { e->required.emplace_back(name, false); }

int mark(om_encoder* e, const std::string& name) {
  for (auto& r : e->required)
    if (r.first == name) {
      r.second = true;
      return 0;
    }
  return 1;
}

// copies a fp32 [rows, cols] block (host or device) into dst (+ optional bf16 conversion)
int upload(const void* data, om_memkind kind, size_t count, float* dst_f32, __nv_bfloat16* dst_bf16) {
  float* staged = dst_f32;
  float* tmp = nullptr;
  if (!staged) {
    OM_CUDA(cudaMalloc(&tmp, count * sizeof(float)));
    staged = tmp;
  }
  cudaError_t err =
      cudaMemcpy(staged, data, count * sizeof(float), kind == OM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice);
  if (err == cudaSuccess && dst_bf16) {
    f32_to_bf16_kernel<<<static_cast<int>(std::min<size_t>((count + 255) / 256, 4096)), 256>>>(staged, dst_bf16,
                                                                                              (int64_t)count);
    err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
  }
  if (tmp) cudaFree(tmp);
  if (err != cudaSuccess) return fail(OM_ECUDA, "weight upload failed: %s", cudaGetErrorString(err));
  return 0;
}

bool shape_is(const int64_t* shape, int ndim, int64_t a, int64_t b = -1) {
  if (b < 0) return ndim == 1 && shape[0] == a;
  return ndim == 2 && shape[0] == a && shape[1] == b;
}

int bad_shape(const char* name) { return fail(OM_EINVAL, "om_encoder_set_weight: unexpected shape for '%s'", name); }

}  // namespace

extern "C" {

int om_encoder_create(const om_encoder_desc* desc, om_encoder** out) {
  if (!desc || !out) return fail(OM_EINVAL, "om_encoder_create: null argument");
  OM_TRY(device_sm_count());
  const om_encoder_desc& d = *desc;
  if (d.arch != OM_ARCH_BERT && d.arch != OM_ARCH_T5ENC) return fail(OM_EINVAL, "unknown arch %d", d.arch);
  if (d.hidden <= 0 || d.hidden % 128 != 0 || d.hidden > 1024)
    return fail(OM_EINVAL, "hidden=%d unsupported (multiple of 128, <= 1024)", d.hidden);
  if (d.heads <= 0 || d.heads * kHeadDim > 2048 || (d.heads * kHeadDim) % 128 != 0)
    return fail(OM_EINVAL, "heads=%d unsupported (head width is 64; heads*64 must be a multiple of 128)", d.heads);
  if (d.arch == OM_ARCH_BERT && d.heads * kHeadDim != d.hidden)
    return fail(OM_EINVAL, "BERT requires hidden == heads * 64 (got hidden=%d heads=%d)", d.hidden, d.heads);
  if (d.ffn <= 0 || d.ffn % 64 != 0) return fail(OM_EINVAL, "ffn=%d must be a positive multiple of 64", d.ffn);
  if (d.layers <= 0 || d.vocab <= 0) return fail(OM_EINVAL, "layers/vocab must be positive");
  if (d.has_head && (d.head_out <= 0)) return fail(OM_EINVAL, "head_out must be positive when has_head=1");
  if (d.max_batch_tokens <= 0) return fail(OM_EINVAL, "max_batch_tokens must be positive");
  om_encoder* e = new (std::nothrow) om_encoder();
  if (!e) return fail(OM_ENOMEM, "out of host memory");
  e->d = d;
  e->I = d.heads * kHeadDim;
  const int H = d.hidden, I = e->I, F = d.ffn;
  e->layers.resize(d.layers);
  int rc = 0;
  auto A = [&](auto** p, size_t n) {
    if (rc == 0) rc = dev_alloc(e, p, n);
  };
  if (d.arch == OM_ARCH_BERT) {
    A(&e->word, (size_t)d.vocab * H);
    A(&e->pos, (size_t)d.max_pos * H);
    A(&e->type, (size_t)std::max(d.type_vocab, 1) * H);
    A(&e->emb_g, H);
    A(&e->emb_b, H);
    require(e, "embeddings.word_embeddings.weight");
    require(e, "embeddings.position_embeddings.weight");
    require(e, "embeddings.token_type_embeddings.weight");
    require(e, "embeddings.LayerNorm.weight");
    require(e, "embeddings.LayerNorm.bias");
  } else {
    A(&e->word, (size_t)d.vocab * H);
    A(&e->final_g, H);
    A(&e->rel_w, (size_t)d.rel_buckets * d.heads);
    A(&e->relbias_log2, (size_t)d.heads * (2 * kMaxL - 1));
    A(&e->relbias_long_log2, (size_t)d.heads * (2 * kMaxLongL - 1));
    require(e, "shared.weight");
    require(e, "encoder.final_layer_norm.weight");
    require(e, "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight");
  }
  for (int i = 0; i < d.layers; ++i) {
    LayerW& w = e->layers[i];
    A(&w.wqkv, (size_t)3 * I * H);
    A(&w.wo, (size_t)H * I);
    A(&w.w1, (size_t)F * H);
    A(&w.w2, (size_t)H * F);
    A(&w.ln1_g, H);
    A(&w.ln2_g, H);
    if (rc == 0 && (cudaMalloc(&w.wqkv_f32, (size_t)3 * I * H * 4) != cudaSuccess ||
                    cudaMalloc(&w.w1_f32, (size_t)F * H * 4) != cudaSuccess)) {
      cudaGetLastError();
      rc = fail(OM_ENOMEM, "om_encoder_create: out of device memory (weight staging)");
    }
    char buf[160];
    if (d.arch == OM_ARCH_BERT) {
      A(&w.bqkv_fold, (size_t)3 * I);
      A(&w.b1_fold, F);
      A(&w.bqkv, (size_t)3 * I);
      A(&w.bo, H);
      A(&w.b1, F);
      A(&w.b2, H);
      A(&w.ln1_b, H);
      A(&w.ln2_b, H);
      static const char* names[] = {"attention.self.query", "attention.self.key", "attention.self.value",
                                    "attention.output.dense", "attention.output.LayerNorm", "intermediate.dense",
                                    "output.dense", "output.LayerNorm"};
      for (const char* n : names)
        for (const char* suffix : {"weight", "bias"}) {
          snprintf(buf, sizeof buf, "encoder.layer.%d.%s.%s", i, n, suffix);
          require(e, buf);
        }
    } else {
      static const char* names[] = {"layer.0.SelfAttention.q", "layer.0.SelfAttention.k", "layer.0.SelfAttention.v",
                                    "layer.0.SelfAttention.o", "layer.0.layer_norm", "layer.1.DenseReluDense.wi",
                                    "layer.1.DenseReluDense.wo", "layer.1.layer_norm"};
      for (const char* n : names) {
        snprintf(buf, sizeof buf, "encoder.block.%d.%s.weight", i, n);
        require(e, buf);
      }
    }
  }
  if (d.has_head) {
    A(&e->head_w, (size_t)d.head_out * H);
    require(e, "head.linear.weight");
  }
  // workspace
  e->Tmax = d.max_batch_tokens;
  e->Tld = static_cast<int>(round_up(2 * static_cast<int64_t>(e->Tmax) + 128, 8));  // V^T pitch: <= 128 columns per tile
  const size_t T = e->Tmax;
  A(&e->h, T * H);
  A(&e->kmask, T);
  A(&e->xb, T * H);
  A(&e->qk, T * 2 * I);
  A(&e->vt, (size_t)I * e->Tld);
  A(&e->ctx, T * I);
  A(&e->inter, T * F);
  A(&e->stats[0], T * 2 * kStatParts);
  A(&e->stats[1], T * 2 * kStatParts);
  A(&e->pooled, T * H);  // at most Tmax sequences (L >= 1)
  A(&e->headed, (size_t)e->Tmax * std::max(d.head_out, 1));
  if (rc != 0) {
    om_encoder_destroy(e);
    return rc;
  }
  // statistics slots a GEMM configuration never writes must read as zero (see RowNorm)
  if (cudaMemset(e->stats[0], 0, T * 2 * kStatParts * 4) != cudaSuccess ||
      cudaMemset(e->stats[1], 0, T * 2 * kStatParts * 4) != cudaSuccess ||
      cudaMemset(e->vt, 0, (size_t)I * e->Tld * 2) != cudaSuccess) {
    om_encoder_destroy(e);
    return fail(OM_ECUDA, "workspace memset failed");
  }
  *out = e;
  return 0;
}

void om_encoder_destroy(om_encoder* e) {
  if (!e) return;
  for (void* p : e->allocs) cudaFree(p);
  for (LayerW& w : e->layers) {
    cudaFree(w.wqkv_f32);
    cudaFree(w.w1_f32);
  }
  delete e;
}

int om_encoder_rep_dim(const om_encoder* e) { return e ? (e->d.has_head ? e->d.head_out : e->d.hidden) : 0; }

int om_encoder_set_weight(om_encoder* e, const char* name_c, const void* data, om_memkind kind, const int64_t* shape,
                          int ndim) {
  if (!e || !name_c || !data || !shape) return fail(OM_EINVAL, "om_encoder_set_weight: null argument");
  std::string name(name_c);
  if (name.rfind("bert.", 0) == 0) name = name.substr(5);  // BertFor* checkpoints prefix the backbone
  const om_encoder_desc& d = e->d;
  const int H = d.hidden, I = e->I, F = d.ffn;
  e->finalized = false;
  if (name == "encoder.embed_tokens.weight") name = "shared.weight";
  if (name == "head.linear.weight" || name == "linear.weight") {
    if (!d.has_head) return 1;
    if (!shape_is(shape, ndim, d.head_out, H)) return bad_shape(name_c);
    OM_TRY(upload(data, kind, (size_t)d.head_out * H, e->head_w, nullptr));
    mark(e, "head.linear.weight");
    return 0;
  }
  if (d.arch == OM_ARCH_BERT) {
    if (name == "embeddings.word_embeddings.weight") {
      if (!shape_is(shape, ndim, d.vocab, H)) return bad_shape(name_c);
      OM_TRY(upload(data, kind, (size_t)d.vocab * H, e->word, nullptr));
    } else if (name == "embeddings.position_embeddings.weight") {
      if (!shape_is(shape, ndim, d.max_pos, H)) return bad_shape(name_c);
      OM_TRY(upload(data, kind, (size_t)d.max_pos * H, e->pos, nullptr));
    } else if (name == "embeddings.token_type_embeddings.weight") {
      if (!shape_is(shape, ndim, d.type_vocab, H)) return bad_shape(name_c);
      OM_TRY(upload(data, kind, (size_t)d.type_vocab * H, e->type, nullptr));
    } else if (name == "embeddings.LayerNorm.weight" || name == "embeddings.LayerNorm.bias") {
      if (!shape_is(shape, ndim, H)) return bad_shape(name_c);
      OM_TRY(upload(data, kind, H, name.back() == 't' ? e->emb_g : e->emb_b, nullptr));
    } else if (name.rfind("encoder.layer.", 0) == 0) {
      int li = -1, consumed = 0;
      if (sscanf(name.c_str(), "encoder.layer.%d.%n", &li, &consumed) != 1 || li < 0 || li >= d.layers) return 1;
      const std::string rest = name.substr(consumed);
      LayerW& w = e->layers[li];
      const bool is_w = rest.size() > 7 && rest.compare(rest.size() - 7, 7, ".weight") == 0;
      const std::string mod = rest.substr(0, rest.rfind('.'));
      int slot = mod == "attention.self.query" ? 0 : mod == "attention.self.key" ? 1 : mod == "attention.self.value" ? 2 : -1;
      if (slot >= 0) {
        if (is_w) {
          if (!shape_is(shape, ndim, I, H)) return bad_shape(name_c);
          if (!w.wqkv_f32) return fail(OM_ESTATE, "om_encoder_set_weight after om_encoder_finalize");
          OM_TRY(upload(data, kind, (size_t)I * H, w.wqkv_f32 + (size_t)slot * I * H, nullptr));
        } else {
          if (!shape_is(shape, ndim, I)) return bad_shape(name_c);
          OM_TRY(upload(data, kind, I, w.bqkv + (size_t)slot * I, nullptr));
        }
      } else if (mod == "attention.output.dense") {
        if (is_w) {
          if (!shape_is(shape, ndim, H, I)) return bad_shape(name_c);
          OM_TRY(upload(data, kind, (size_t)H * I, nullptr, w.wo));
        } else {
          if (!shape_is(shape, ndim, H)) return bad_shape(name_c);
          OM_TRY(upload(data, kind, H, w.bo, nullptr));
        }
      } else if (mod == "attention.output.LayerNorm" || mod == "output.LayerNorm") {
        if (!shape_is(shape, ndim, H)) return bad_shape(name_c);
        float* dst = mod[0] == 'a' ? (is_w ? w.ln1_g : w.ln1_b) : (is_w ? w.ln2_g : w.ln2_b);
        OM_TRY(upload(data, kind, H, dst, nullptr));
      } else if (mod == "intermediate.dense") {
        if (is_w) {
          if (!shape_is(shape, ndim, F, H)) return bad_shape(name_c);
          if (!w.w1_f32) return fail(OM_ESTATE, "om_encoder_set_weight after om_encoder_finalize");
          OM_TRY(upload(data, kind, (size_t)F * H, w.w1_f32, nullptr));
        } else {
          if (!shape_is(shape, ndim, F)) return bad_shape(name_c);
          OM_TRY(upload(data, kind, F, w.b1, nullptr));
        }
      } else if (mod == "output.dense") {
        if (is_w) {
          if (!shape_is(shape, ndim, H, F)) return bad_shape(name_c);
          OM_TRY(upload(data, kind, (size_t)H * F, nullptr, w.w2));
        } else {
          if (!shape_is(shape, ndim, H)) return bad_shape(name_c);
          OM_TRY(upload(data, kind, H, w.b2, nullptr));
        }
      } else {
        return 1;
      }
    } else {
      return 1;  // e.g. pooler.*: computed by HF, never used by OpenMatch
    }
  } else {
    if (name == "shared.weight") {
      if (!shape_is(shape, ndim, d.vocab, H)) return bad_shape(name_c);
      OM_TRY(upload(data, kind, (size_t)d.vocab * H, e->word, nullptr));
    } else if (name == "encoder.final_layer_norm.weight") {
      if (!shape_is(shape, ndim, H)) return bad_shape(name_c);
      OM_TRY(upload(data, kind, H, e->final_g, nullptr));
    } else if (name.rfind("encoder.block.", 0) == 0) {
      int li = -1, consumed = 0;
      if (sscanf(name.c_str(), "encoder.block.%d.%n", &li, &consumed) != 1 || li < 0 || li >= d.layers) return 1;
      const std::string rest = name.substr(consumed);
      LayerW& w = e->layers[li];
      if (rest == "layer.0.SelfAttention.relative_attention_bias.weight") {
        if (li != 0) return 1;
        if (!shape_is(shape, ndim, d.rel_buckets, d.heads)) return bad_shape(name_c);
        OM_TRY(upload(data, kind, (size_t)d.rel_buckets * d.heads, e->rel_w, nullptr));
        e->rel_host.resize((size_t)d.rel_buckets * d.heads);
        OM_CUDA(cudaMemcpy(e->rel_host.data(), e->rel_w, e->rel_host.size() * 4, cudaMemcpyDeviceToHost));
      } else if (rest == "layer.0.SelfAttention.q.weight" || rest == "layer.0.SelfAttention.k.weight" ||
                 rest == "layer.0.SelfAttention.v.weight") {
        const int slot = rest[22] == 'q' ? 0 : rest[22] == 'k' ? 1 : 2;
        if (!shape_is(shape, ndim, I, H)) return bad_shape(name_c);
        if (!w.wqkv_f32) return fail(OM_ESTATE, "om_encoder_set_weight after om_encoder_finalize");
        OM_TRY(upload(data, kind, (size_t)I * H, w.wqkv_f32 + (size_t)slot * I * H, nullptr));
      } else if (rest == "layer.0.SelfAttention.o.weight") {
        if (!shape_is(shape, ndim, H, I)) return bad_shape(name_c);
        OM_TRY(upload(data, kind, (size_t)H * I, nullptr, w.wo));
      } else if (rest == "layer.0.layer_norm.weight" || rest == "layer.1.layer_norm.weight") {
        if (!shape_is(shape, ndim, H)) return bad_shape(name_c);
        OM_TRY(upload(data, kind, H, rest[6] == '0' ? w.ln1_g : w.ln2_g, nullptr));
      } else if (rest == "layer.1.DenseReluDense.wi.weight") {
        if (!shape_is(shape, ndim, F, H)) return bad_shape(name_c);
        if (!w.w1_f32) return fail(OM_ESTATE, "om_encoder_set_weight after om_encoder_finalize");
        OM_TRY(upload(data, kind, (size_t)F * H, w.w1_f32, nullptr));
      } else if (rest == "layer.1.DenseReluDense.wo.weight") {
        if (!shape_is(shape, ndim, H, F)) return bad_shape(name_c);
        OM_TRY(upload(data, kind, (size_t)H * F, nullptr, w.w2));
      } else if (rest == "layer.1.DenseReluDense.wi_0.weight" || rest == "layer.1.DenseReluDense.wi_1.weight") {
        return fail(OM_EINVAL, "gated-GELU T5 feed-forward (t5 v1.1) is not supported by this build");
      } else {
        return 1;
      }
    } else {
      return 1;
    }
  }
  return mark(e, name);
}

int om_encoder_finalize(om_encoder* e) {
  if (!e) return fail(OM_EINVAL, "om_encoder_finalize: null encoder");
  std::string miss;
  int nmiss = 0;
  for (auto& r : e->required)
    if (!r.second) {
      if (nmiss < 4) miss += (nmiss ? ", " : "") + r.first;
      ++nmiss;
    }
  if (nmiss) return fail(OM_ESTATE, "om_encoder_finalize: %d parameter(s) missing: %s%s", nmiss, miss.c_str(), nmiss > 4 ? ", ..." : "");
  if (e->d.arch == OM_ARCH_T5ENC) {
    const int nh = e->d.heads;
    for (int pass = 0; pass < 2; ++pass) {  // one table per attention kernel
      const int maxl = pass == 0 ? kMaxL : kMaxLongL, W = 2 * maxl - 1;
      std::vector<float> table((size_t)nh * W);
      for (int rel = -(maxl - 1); rel <= maxl - 1; ++rel) {
        const int b = t5_bucket(rel, e->d.rel_buckets, e->d.rel_max_distance);
        for (int h = 0; h < nh; ++h) table[(size_t)h * W + rel + maxl - 1] = e->rel_host[(size_t)b * nh + h] * kLog2e;
      }
      OM_CUDA(cudaMemcpy(pass == 0 ? e->relbias_log2 : e->relbias_long_log2, table.data(), table.size() * 4,
                         cudaMemcpyHostToDevice));
    }
  }
  // fold every normalisation into the GEMM that consumes it (file header): BERT layer l's QKV takes the LayerNorm that
  // produced its input (embeddings.LayerNorm for l = 0, else layer l-1's output.LayerNorm) and W1 takes
  // attention.output.LayerNorm; T5 block l's QKV / wi take its own pre-norm RMS weights (no beta, no mean term).
  {
    const bool bert = e->d.arch == OM_ARCH_BERT;
    const int H = e->d.hidden, I = e->I, F = e->d.ffn;
    for (int li = 0; li < e->d.layers; ++li) {
      LayerW& w = e->layers[li];
      if (!w.wqkv_f32 || !w.w1_f32) return fail(OM_ESTATE, "om_encoder_finalize called twice");
      const float* g_in = bert ? (li == 0 ? e->emb_g : e->layers[li - 1].ln2_g) : w.ln1_g;
      const float* b_in = bert ? (li == 0 ? e->emb_b : e->layers[li - 1].ln2_b) : nullptr;
      fold_norm_kernel<<<(3 * I + 7) / 8, 256>>>(w.wqkv_f32, g_in, b_in, bert ? w.bqkv : nullptr, 3 * I, H, bert ? 1 : 0,
                                                 w.wqkv, w.bqkv_fold);
      fold_norm_kernel<<<(F + 7) / 8, 256>>>(w.w1_f32, bert ? w.ln1_g : w.ln2_g, bert ? w.ln1_b : nullptr,
                                             bert ? w.b1 : nullptr, F, H, bert ? 1 : 0, w.w1, w.b1_fold);
    }
    OM_CUDA(cudaGetLastError());
    OM_CUDA(cudaDeviceSynchronize());
    for (LayerW& w : e->layers) {
      cudaFree(w.wqkv_f32);
      cudaFree(w.w1_f32);
      w.wqkv_f32 = w.w1_f32 = nullptr;
    }
  }
  static bool attr = false;
  if (!attr) {
    OM_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    OM_CUDA(cudaFuncSetAttribute(attn_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnLongSmemBytes));
    attr = true;
  }
  e->finalized = true;
  return 0;
}

int om_encode(om_encoder* e, const int64_t* input_ids, const int64_t* attention_mask, const int64_t* token_type_ids,
              int B, int L, void* out_reps, om_dtype out_dtype, int64_t out_row_stride, float* out_hidden, void* stream) {
  if (!e || !input_ids || !attention_mask || !out_reps) return fail(OM_EINVAL, "om_encode: null argument");
  if (!e->finalized) return fail(OM_ESTATE, "om_encode: call om_encoder_finalize first");
  if (B <= 0 || L <= 0) return fail(OM_EINVAL, "om_encode: B and L must be positive");
  const bool long_seq = L > kMaxL;
  if (long_seq && (L > kMaxLongL || L % 128 != 0))
    return fail(OM_EINVAL, "om_encode: L=%d unsupported (at most %d tokens, or 256 / 384 / 512: pad to a multiple of 128)", L,
                kMaxL);
  const om_encoder_desc& d = e->d;
  if (d.arch == OM_ARCH_BERT && L > d.max_pos) return fail(OM_EINVAL, "om_encode: L=%d exceeds max_position_embeddings", L);
  const int64_t T64 = static_cast<int64_t>(B) * L;
  if (T64 > e->Tmax) return fail(OM_EINVAL, "om_encode: B*L=%lld exceeds max_batch_tokens=%d", (long long)T64, e->Tmax);
  if (out_dtype != OM_F32 && out_dtype != OM_BF16) return fail(OM_EINVAL, "om_encode: out dtype must be f32 or bf16");
  const int rep_dim = om_encoder_rep_dim(e);
  if (out_row_stride < rep_dim) return fail(OM_EINVAL, "om_encode: out_row_stride < rep_dim");
  const int sms = device_sm_count();
  if (sms < 0) return sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T = static_cast<int>(T64), H = d.hidden, I = e->I, F = d.ffn;
  const bool bert = d.arch == OM_ARCH_BERT;
  const int rows4 = (T + 3) / 4;

  NvtxRange nvtx("om.encode");
  keymask_kernel<<<(T + 255) / 256, 256, 0, st>>>(attention_mask, e->kmask, T);
  // the residual stream s lives in e->h (fp32, un-normalised) with a bf16 copy in e->xb and row statistics in
  // e->stats[0] (input of a layer: from the embedding or the previous FFN2) / e->stats[1] (after the attention block)
  if (bert)
    bert_embed_kernel<<<rows4, 128, 0, st>>>(input_ids, token_type_ids, e->word, e->type, e->pos, T, L, H, d.vocab,
                                             std::max(d.type_vocab, 1), e->h, e->xb, e->stats[0]);
  else
    t5_embed_kernel<<<rows4, 128, 0, st>>>(input_ids, e->word, T, H, d.vocab, e->h, e->xb, e->stats[0]);
  OM_CUDA(cudaGetLastError());

  // attention geometry
  const int spt = L > 64 ? 1 : kMaxL / L;  // sequences per 128-row tile (long sequences: L / 128 tiles per sequence)
  AttnParams ap;
  ap.T = T;
  ap.L = L;
  ap.spt = spt;
  ap.I = I;
  ap.Tvalid_rows = long_seq ? 128 : spt * L;
  ap.scale_log2 = (bert ? 0.125f : 1.0f) * kLog2e;
  ap.kmask = e->kmask;
  ap.relbias_log2 = bert ? nullptr : (long_seq ? e->relbias_long_log2 : e->relbias_log2);
  ap.ctx = e->ctx;
  const int n_tiles = long_seq ? T / 128 : (B + spt - 1) / spt;
  CUtensorMap tmQK, tmVt;
  if (make_tmap_bf16_2d(&tmQK, e->qk, (uint64_t)2 * I, (uint64_t)T, (uint64_t)2 * I * 2, 64, 128) != 0 ||
      make_tmap_bf16_2d(&tmVt, e->vt, (uint64_t)n_tiles * 128, (uint64_t)I, (uint64_t)e->Tld * 2, 64, 64) != 0)
    return fail(OM_ECUDA, "om_encode: tensor map creation failed");

  // TMA-store tensor maps of the bf16 GEMM outputs (box = 64 columns x 32 rows = one epilogue warp's chunk pair) and the
  // residual stream's maps (fp32 load + store, bf16 store; box = 32 columns x 32 rows = one chunk)
  CUtensorMap tmQKout, tmInter, tmS, tmXb;
  if (make_tmap_bf16_2d(&tmQKout, e->qk, (uint64_t)2 * I, (uint64_t)T, (uint64_t)2 * I * 2, 64, 32) != 0 ||
      make_tmap_bf16_2d(&tmInter, e->inter, (uint64_t)F, (uint64_t)T, (uint64_t)F * 2, 64, 32) != 0 ||
      make_tmap_2d(&tmS, e->h, 4, (uint64_t)H, (uint64_t)T, (uint64_t)H * 4, 32, 32, 128) != 0 ||
      make_tmap_2d(&tmXb, e->xb, 2, (uint64_t)H, (uint64_t)T, (uint64_t)H * 2, 32, 32, 64) != 0)
    return fail(OM_ECUDA, "om_encode: output tensor map creation failed");
  const float inv_h = 1.0f / static_cast<float>(H);
  const int rms = bert ? 0 : 1;
  const RowNorm normA{e->stats[0], inv_h, d.ln_eps, rms}, normB{e->stats[1], inv_h, d.ln_eps, rms};
  const RowNorm ident{nullptr, inv_h, d.ln_eps, rms};
  // wide GEMMs (QKV, FFN1: N >= 2 H, plain bf16 epilogues): CTA pairs (cta_group::2, 256 x 256 tiles: half the operand
  // bytes per SM and FLOP; 1 700 vs 1 595 TFLOP/s on the FFN1 shape, profiles/r02_2sm_product_core.log), single-CTA
  // tiles when the device cannot host a pair
  static const bool pair_gemm = getenv("OM_ENCODER_SINGLE_CTA") == nullptr;  // measurement switch, read once
  auto wide_gemm = [&](const __nv_bfloat16* A, int K, const __nv_bfloat16* W, int M, int N, const auto& epi) -> cudaError_t {
    cudaError_t err = cudaErrorNotSupported;
    if (pair_gemm) err = launch_gemm2<5, false, 8>(A, K, W, K, M, N, K, epi, sms, st);
    if (err == cudaErrorNotSupported) err = launch_gemm<256, 4, false, 8>(A, K, W, K, M, N, K, epi, sms, st);
    return err;
  };
  // residual GEMMs (N = H): 192-wide tiles divide 768 into 4 (1024 tiles = 6.9 waves of 3/4-size tiles instead of 5.2
  // waves of full tiles); 8 epilogue warps = 2 column groups per tile -> (H / BN) * 2 <= kStatParts statistics slots
  // CTA pairs: 256-wide tiles with 8 epilogue warps (two column groups) fit next to a 5-stage ring of 32 KB stages
  // (bert-base 6.59 vs 6.80 ms, bert-large 20.4 vs 21.7 ms per batch of 256: profiles/r02_encoder_pair_resid_probe.log)
  auto resid_gemm = [&](const __nv_bfloat16* A, int K, const __nv_bfloat16* W, EpiResidNorm epi) -> cudaError_t {
    if (pair_gemm && ((H + 255) / 256) * 2 <= kStatParts) {
      EpiResidNorm e2 = epi;
      e2.parts = 2;
      e2.bn = 0;
      const cudaError_t err = launch_gemm2<5, false, 8>(A, K, W, K, T, H, K, e2, sms, st);
      if (err != cudaErrorNotSupported) return err;
    }
    if (H % 192 == 0) return launch_gemm<192, 4, false, 8>(A, K, W, K, T, H, K, epi, sms, st);
    epi.parts = 1;  // 256-wide tiles: a 4-stage ring leaves room for 4 epilogue warps (one column group)
    return launch_gemm<256, 4, false, 4>(A, K, W, K, T, H, K, epi, sms, st);
  };
  if ((H % 192 == 0 ? H / 192 : (H + 255) / 256) * 2 > kStatParts)
    return fail(OM_EINVAL, "om_encode: hidden=%d needs more statistics slots than kStatParts", H);
  for (int li = 0; li < d.layers; ++li) {
    const LayerW& w = e->layers[li];
    NvtxRange nvtx_layer("om.encode.layer");
    // LayerNorm that produced this layer's input (BERT; applied on the fly wherever the input is consumed)
    const float* g_in = bert ? (li == 0 ? e->emb_g : e->layers[li - 1].ln2_g) : nullptr;
    const float* b_in = bert ? (li == 0 ? e->emb_b : e->layers[li - 1].ln2_b) : nullptr;
    {
      EpiQKV epi{tmQKout, e->qk, e->vt, e->Tld, bert ? w.bqkv_fold : nullptr, T, 2 * I, ap.Tvalid_rows, normA};
      // (192-wide tiles with 12 epilogue warps measured 3 % slower here and on FFN1: r02 tile A/B in profiles/README.md)
      cudaError_t err = wide_gemm(e->xb, H, w.wqkv, T, 3 * I, epi);
      if (err != cudaSuccess) return fail(OM_ECUDA, "QKV GEMM launch failed: %s", cudaGetErrorString(err));
    }
    if (long_seq)
      attn_long_kernel<<<dim3(n_tiles, d.heads), 128, kAttnLongSmemBytes, st>>>(tmQK, tmVt, ap);
    else
      attn_kernel<<<std::min(n_tiles * d.heads, sms * 4), 128, kAttnSmemBytes, st>>>(tmQK, tmVt, ap, n_tiles, d.heads);
    OM_CUDA(cudaGetLastError());
    {
      // s <- ctx Wo^T + bo + LN_in(s) (BERT) / + s (T5); statistics of the new s -> stats[1]
      EpiResidNorm epi{tmS, tmXb, bert ? w.bo : nullptr, bert ? normA : ident, g_in, b_in, e->stats[1], 2, T, H,
                       H % 192 == 0 ? 192 : 256};
      cudaError_t err = resid_gemm(e->ctx, I, w.wo, epi);
      if (err != cudaSuccess) return fail(OM_ECUDA, "O-proj GEMM launch failed: %s", cudaGetErrorString(err));
    }
    {
      cudaError_t err;
      if (bert) {
        EpiBiasActBf16<ACT_GELU> epi{tmInter, e->inter, F, w.b1_fold, T, F, normB};
        err = wide_gemm(e->xb, H, w.w1, T, F, epi);
      } else {
        EpiBiasActBf16<ACT_RELU> epi{tmInter, e->inter, F, nullptr, T, F, normB};
        err = wide_gemm(e->xb, H, w.w1, T, F, epi);
      }
      if (err != cudaSuccess) return fail(OM_ECUDA, "FFN1 GEMM launch failed: %s", cudaGetErrorString(err));
    }
    {
      // s <- inter W2^T + b2 + LN_attn(s) (BERT) / + s (T5); statistics -> stats[0] (the next layer's input)
      EpiResidNorm epi{tmS, tmXb, bert ? w.b2 : nullptr, bert ? normB : ident, w.ln1_g, w.ln1_b, e->stats[0], 2, T, H,
                       H % 192 == 0 ? 192 : 256};
      cudaError_t err = resid_gemm(e->inter, F, w.w2, epi);
      if (err != cudaSuccess) return fail(OM_ECUDA, "FFN2 GEMM launch failed: %s", cudaGetErrorString(err));
    }
  }
  // the one normalisation that runs as a kernel: last_hidden_state = LN_out(s) (BERT: last layer's output.LayerNorm,
  // T5: final_layer_norm), in place in e->h, for pooling and the optional out_hidden copy
  if (bert)
    norm_kernel<false, false><<<rows4, 128, 0, st>>>(e->h, nullptr, e->layers[d.layers - 1].ln2_g,
                                                     e->layers[d.layers - 1].ln2_b, d.ln_eps, T, H, e->h, nullptr);
  else
    norm_kernel<true, false><<<rows4, 128, 0, st>>>(e->h, nullptr, e->final_g, nullptr, d.ln_eps, T, H, e->h, nullptr);
  OM_CUDA(cudaGetLastError());
  if (out_hidden)
    OM_CUDA(cudaMemcpyAsync(out_hidden, e->h, static_cast<size_t>(T) * H * 4, cudaMemcpyDeviceToDevice, st));

  NvtxRange nvtx_pool("om.encode.pool_head_normalize");
  pool_kernel<<<B, 256, 0, st>>>(e->h, attention_mask, L, H, d.pooling == OM_POOL_MEAN ? 1 : 0, e->pooled);
  const float* reps = e->pooled;
  if (d.has_head) {
    const int warps = d.head_out * ((B + 7) / 8);
    head_kernel<<<(warps * 32 + 255) / 256, 256, 0, st>>>(e->pooled, e->head_w, B, H, d.head_out, e->headed);
    reps = e->headed;
  }
  if (out_dtype == OM_F32)
    finish_reps_kernel<float><<<(B + 3) / 4, 128, 0, st>>>(reps, B, rep_dim, d.normalize, static_cast<float*>(out_reps),
                                                          out_row_stride);
  else
    finish_reps_kernel<__nv_bfloat16><<<(B + 3) / 4, 128, 0, st>>>(reps, B, rep_dim, d.normalize,
                                                                  static_cast<__nv_bfloat16*>(out_reps), out_row_stride);
  OM_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
