/* openmatch_b200 — C ABI of the B200-native dense-retrieval hot path (libopenmatch_b200.so).
 *
 * The reference (thunlp/OpenMatch, 100 % Python) has no FFI layer; the seams where its hot path crosses
 * into third-party compute are three Python call sites, and each entry point below replaces one of them
 * (paths relative to the reference tree):
 *
 *   encoder  : DRModel.encode -> model(**items)          src/openmatch/modeling/dense_retrieval_model.py:133-155
 *              + mean_pooling                             src/openmatch/utils.py:233-235
 *              + LinearHead.forward                       src/openmatch/modeling/linear.py:22-23
 *   index    : faiss.IndexFlatIP(dim) / .add / .search / .reset / .ntotal
 *                                                         src/openmatch/retriever/dense_retriever.py:38-41,105,133-137,180
 *              IndexShards merge behind index_cpu_to_gpu_multiple(shard=True)   ...:43-58
 *   loss     : torch.matmul + cross_entropy (+ autograd)  src/openmatch/loss.py:7-15,
 *                                                         src/openmatch/modeling/dense_retrieval_model.py:113-125
 *
 * Conventions: plain C symbols; every function returns 0 on success and a negative OM_E* code on
 * failure, om_last_error() then holds a thread-local message.  The caller owns all tensor memory and
 * passes raw pointers (host or device as stated) plus a CUDA stream handle (cudaStream_t cast to void*,
 * NULL = legacy default stream); the library owns only its opaque handles.  One process drives one GPU;
 * a handle is not thread-safe, distinct handles are.  There is no CPU fallback: every compute entry
 * point fails with OM_ENODEVICE when no sm_100 device is present.
 */
#ifndef OPENMATCH_B200_H_
#define OPENMATCH_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OM_ABI_VERSION 1

enum { OM_OK = 0, OM_EINVAL = -1, OM_ECUDA = -2, OM_ENOMEM = -3, OM_ENODEVICE = -4, OM_ESTATE = -5, OM_EFAULT = -6 };

typedef enum { OM_F32 = 0, OM_BF16 = 1, OM_F16 = 2 } om_dtype;
typedef enum { OM_HOST = 0, OM_DEVICE = 1 } om_memkind;
typedef enum { OM_ARCH_BERT = 0, OM_ARCH_T5ENC = 1 } om_arch;
typedef enum { OM_POOL_FIRST = 0, OM_POOL_MEAN = 1 } om_pooling;
typedef enum { OM_REDUCE_MEAN = 0, OM_REDUCE_SUM = 1 } om_reduction;

typedef struct om_encoder om_encoder;
typedef struct om_index om_index;
typedef struct om_comm om_comm;

/* ---- library ------------------------------------------------------------------------------------- */
int om_abi_version(void);
const char* om_last_error(void);
/* number of SMs of the current device, or a negative error (OM_ENODEVICE without a GPU) */
int om_device_sm_count(void);

/* ---- encoder: replaces HF BertModel / T5EncoderModel forward + pooling + head + normalise ---------- */
typedef struct om_encoder_desc {
  int32_t arch;             /* om_arch */
  int32_t layers;           /* num_hidden_layers / num_layers */
  int32_t hidden;           /* hidden_size / d_model (multiple of 64) */
  int32_t heads;            /* attention heads; head width is fixed at 64 (bert-base/large, t5-base) */
  int32_t ffn;              /* intermediate_size / d_ff (multiple of 64) */
  int32_t vocab;            /* vocab_size */
  int32_t max_pos;          /* max_position_embeddings (BERT); ignored for T5 */
  int32_t type_vocab;       /* type_vocab_size (BERT); ignored for T5 */
  float ln_eps;             /* layer_norm_eps (1e-12 BERT) / layer_norm_epsilon (1e-6 T5) */
  int32_t pooling;          /* om_pooling: DRModel.pooling 'first' | 'mean' */
  int32_t has_head;         /* 1: bias-free LinearHead follows pooling */
  int32_t head_out;         /* LinearHead output_dim (multiple of 8) */
  int32_t normalize;        /* 1: F.normalize(reps, dim=1) */
  int32_t rel_buckets;      /* T5 relative_attention_num_buckets (32) */
  int32_t rel_max_distance; /* T5 relative_attention_max_distance (128) */
  int32_t max_batch_tokens; /* workspace sizing: max B*L per om_encode call (e.g. 256*128) */
} om_encoder_desc;

int om_encoder_create(const om_encoder_desc* desc, om_encoder** out);
/* Parameter by its HuggingFace state_dict name (e.g. "encoder.layer.3.attention.self.query.weight",
 * "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight", or "head.linear.weight" for
 * the LinearHead).  Data is fp32, row-major, host or device; the library keeps its own packed copy, the
 * caller's tensor may be freed afterwards.  Unknown names (e.g. "pooler.*", which OpenMatch never uses)
 * are ignored and reported through the return value 1. */
int om_encoder_set_weight(om_encoder* enc, const char* name, const void* data, om_memkind kind,
                          const int64_t* shape, int ndim);
/* Verifies that every required parameter was supplied and builds derived tables. */
int om_encoder_finalize(om_encoder* enc);
/* input_ids / attention_mask / token_type_ids (nullable => zeros; ignored for T5): int64 [B, L] device,
 * row-major, exactly what DRInferenceCollator / QPCollator hand to the model; L <= 128.
 * out_reps: device [B, rep_dim] fp32 or bf16 with row pitch out_row_stride (elements) — may point into an
 * index shard obtained from om_index_reserve().  out_hidden: nullable device fp32 [B, L, hidden]
 * (last_hidden_state).  Asynchronous on `stream`. */
int om_encode(om_encoder* enc, const int64_t* input_ids, const int64_t* attention_mask,
              const int64_t* token_type_ids, int B, int L, void* out_reps, om_dtype out_dtype,
              int64_t out_row_stride, float* out_hidden, void* stream);
int om_encoder_rep_dim(const om_encoder* enc);
void om_encoder_destroy(om_encoder* enc);

/* ---- index: replaces faiss.IndexFlatIP (exact inner-product top-k) -------------------------------- */
int om_index_create(int d, om_index** out); /* faiss.IndexFlatIP(d); lives on the current device */
/* index.add(x): x [n, d] row-major, fp32 (host or device).  Rows get ids ntotal .. ntotal+n-1. */
int om_index_add(om_index* idx, const void* x, om_memkind kind, om_dtype dtype, int64_t n, void* stream);
/* Zero-copy ingest: reserve room for n more rows and get the device address of the fp32 row block
 * (row pitch = d floats) so the encoder can write embeddings in place; om_index_commit(n) publishes
 * them (builds the fp16 scan copy and updates the error-norm maxima the exactness certificate uses).  */
int om_index_reserve(om_index* idx, int64_t n, float** dev_rows);
int om_index_commit(om_index* idx, int64_t n, void* stream);
int64_t om_index_ntotal(const om_index* idx);
int om_index_dim(const om_index* idx);
int om_index_reset(om_index* idx);
/* D, I = index.search(q, k): q [nq, d] fp32 (host or device); D fp32 [nq, k], I int64 [nq, k] written to
 * host or device memory (out_kind).  Rows are ordered by (score descending, id ascending); missing
 * slots (k > ntotal) hold id -1 and score -FLT_MAX, as faiss does.  Reported ids are id_offset + local
 * row, so a rank of a row-sharded index passes the global id of its first row.  k <= 4096.
 * Exactness: candidates (k + slack per query) are selected on fp16 tensor-core scores and re-scored in fp32;
 * an a-posteriori certificate (measured quantisation-error norms of corpus and query, see csrc/search.cu
 * certify_kernel) then PROVES per query that no row outside the candidate list can reach the k-th fp32 score.
 * Queries that fail it are re-run with the widest candidate list (4096) and, if still unproven (e.g. thousands
 * of near-duplicate rows), answered by an exact fp32 scan with the same summation order as the re-score.  The
 * result is therefore always the exact top-k by fp32 inner product with ties by ascending id.
 * Synchronous with respect to `stream` on return. */
int om_index_search(om_index* idx, const void* q, om_memkind q_kind, int nq, int k, float* D, int64_t* I,
                    om_memkind out_kind, int64_t id_offset, void* stream);
/* Row-sharded search with the exchange inside the library — replaces faiss.index_cpu_to_gpu_multiple(shard=True) +
 * IndexShards (src/openmatch/retriever/dense_retriever.py:43-58).  One process per GPU; every rank holds a contiguous
 * row shard and calls om_index_search_sharded with the same queries and its own id_offset; every rank receives the
 * same global (D, I).  Each shard keeps a candidate list sized for its share of the answer (m + 6 sqrt(m) + 32 with m = (k + slack) / world
 * rows), re-scores it in fp32 and ships it whole together with the list's stage-score floor and the shard's error-norm
 * maxima in ONE packed NCCL all-gather per query chunk (issued on `stream` between the kernels, single host
 * synchronisation at the end of a level); every rank merges the lists and runs the exactness certificate of
 * om_index_search with tau = the largest floor.  Skewed shards fail the certificate and are answered by the wider levels.
 *   om_comm_unique_id : rank 0 obtains 128 opaque bytes (ncclGetUniqueId) and ships them to the other ranks
 *                       (e.g. torch.distributed.broadcast_object_list)
 *   om_comm_init      : collective over the `world` ranks (ncclCommInitRank) on the current device
 * NCCL is bound at run time (dlopen of libnccl.so.2; inside PyTorch that is the copy torch already loaded). */
int om_comm_unique_id(char* out128);
int om_comm_init(const char* unique_id128, int rank, int world, om_comm** out);
void om_comm_destroy(om_comm* comm);
int om_index_search_sharded(om_index* idx, om_comm* comm, const void* q, om_memkind q_kind, int nq, int k, float* D,
                            int64_t* I, om_memkind out_kind, int64_t id_offset, void* stream);

/* Three-phase variant for row-sharded indexes (one shard per process, at most 16384 queries per round trip).
 * Replaces the per-shard full top-k that faiss IndexShards computes before merging (dense_retriever.py:43-58):
 * the shards agree on a per-query score floor first, so each one re-scores and ships ~k / n_shards rows.
 *   begin : fp16 tensor-core scan of the local shard.  local_range (device fp32 [2, nq]) receives per query the local
 *           (k + slack)-th best candidate-stage score (row 0; -inf when the shard has fewer rows) and the local best
 *           (row 1).  The caller MAX-reduces it over the shards.
 *   count : local_hist (device int32 [nq, om_search_floor_bins()]) receives the histogram of the local
 *           candidates over equal-width bins of [global_range[0][q], global_range[1][q]].  The caller SUM-reduces.
 *   finish: re-scores in fp32 only the local candidates in or above the bin holding the global (k + slack)-th
 *           score (every member of the global top-k is among them) and writes them sorted to device D fp32
 *           [nq, k] / I int64 [nq, k], padded with -FLT_MAX / -1.  global_hist == NULL: floor = global_range[0];
 *           global_range == NULL: no pruning.  kept_max (device int32, nullable) receives the longest valid prefix
 *           over the queries, so the caller can exchange [nq, kept_max] instead of [nq, k].
 * These are the uncertified building blocks (the caller reduces between the phases, e.g. over gloo in the CPU
 * protocol test); om_index_search_sharded runs the same phases plus the certificate and its escalation. */
int om_index_search_begin(om_index* idx, const void* q, om_memkind q_kind, int nq, int k, float* local_range,
                          void* stream);
int om_index_search_count(om_index* idx, const float* global_range, int* local_hist, void* stream);
int om_index_search_finish(om_index* idx, const float* global_range, const int* global_hist, float* D, int64_t* I,
                           int64_t id_offset, int* kept_max, void* stream);
int om_search_floor_bins(void);
/* Tunables: "rescore_slack" (extra candidate-stage rows kept per query; default max(128, k/5)),
 * "force_safe_rounds" (1 = always use the overflow-proof fixed-size round schedule; testing),
 * "round_growth" (2..8: each scan round covers (g-1) x the rows already seen; default 0 = auto: 2, or 8 for <= 256 queries),
 * "certify" (default 1; 0 = skip the exactness certificate and its escalation: top-k of the fp16 candidate stage),
 * "exact_only" (1 = answer every query with the exact fp32 CUDA-core scan; testing),
 * "debug_stage_scores" (1 = D holds candidate-stage scores instead of fp32 re-scores; error-model measurement),
 * "pair_scan" (default 1: the scan GEMM runs on CTA pairs, tcgen05 cta_group::2; 0 = single-CTA tiles),
 * "profile" (1 = bracket every kernel launch of a search with CUDA events on the launching stream). */
int om_index_set_param(om_index* idx, const char* name, int64_t value);
/* Statistics of the last search: "rounds", "overflow_retries", "candidates" (per query capacity),
 * "launches" (kernels launched), "uncertified" (queries the first level could not prove exact),
 * "uncertified_wide" (still unproven with 4096 candidates), "exact_queries" (answered by the exact fp32 scan),
 * and with "profile" on: "scan_ns", "select_ns",
 * "finalize_ns", "other_ns" (device time summed over the launches of each kind; other = exchange + merge + certify). */
int64_t om_index_get_stat(const om_index* idx, const char* name);
void om_index_destroy(om_index* idx);

/* Exchange step of the row-sharded search: merge `nparts` per-shard results laid out as
 * D_parts [nparts, nq, k], I_parts [nparts, nq, k] (device) into the global top-k by (score desc, id asc);
 * ids < 0 are padding.  Matches merge semantics of faiss IndexShards / utils.py:215-229. */
int om_topk_merge(const float* D_parts, const int64_t* I_parts, int nparts, int nq, int k, float* D, int64_t* I,
                  void* stream);
/* Same with input lists of width k_in (e.g. the kept_max prefix of om_index_search_finish) and k_out results;
 * more than 8192 entries per query (nparts * k_in) are merged hierarchically. */
int om_topk_merge_n(const float* D_parts, const int64_t* I_parts, int nparts, int nq, int k_in, int k_out, float* D,
                    int64_t* I, void* stream);

/* ---- loss: replaces matmul + cross_entropy + autograd backward ------------------------------------ */
/* Q [nq, d], P [np, d] device, fp32 or bf16 (both the same dtype), row-major.
 * target: nullable int64 [nq] device (NULL => i * (np / nq), loss.py:11-13).
 * loss_out: device fp32 scalar = loss_scale * reduce_i(logsumexp_j s_ij - s_i,target_i).
 * dQ [nq, d], dP [np, d]: nullable device fp32 gradients of loss_out.  scores_out: nullable device fp32
 * [nq, np] logits (DROutput.scores).  Asynchronous on `stream`. */
int om_contrastive_loss_fwd_bwd(const void* Q, const void* P, om_dtype dtype, int nq, int np, int d,
                                const int64_t* target, int reduction, float loss_scale, float* loss_out,
                                float* dQ, float* dP, float* scores_out, void* stream);
/* Diagnostics: device time (ns, %globaltimer) the most recent loss call spent in its four phases
 * {PREP, LOGITS, SOFTMAX, GRADS}.  Synchronous. */
int om_debug_loss_phase_ns(uint64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* OPENMATCH_B200_H_ */
