/* mrs_b200_model.h — C ABI of the B200-native Llama-family decode layer stack.
 *
 * This is the host-side caller of the hot path (the role of `Llama::forward` /
 * `Block::forward` in the reference: mistralrs-core/src/models/llama.rs:243-260,475-…),
 * restricted to what the benchmark needs: it owns no memory, takes raw device pointers, and
 * enqueues the per-token kernel chain on the given stream (CUDA-graph capturable: no
 * allocation, no host sync).  Per decoder layer it issues
 *     [RMSNorm + Q8_1 + fused QKV GEMV] -> RoPE -> KV-cache write -> paged decode attention
 *     -> [Q8_1 + o_proj GEMV + residual] -> [RMSNorm + Q8_1 + gate/up GEMV + SiLU*mul]
 *     -> [Q8_1 + down GEMV + residual]
 * with the bracketed groups being single `mrs_mmvq_fused` launches.
 */
#ifndef MRS_B200_MODEL_H
#define MRS_B200_MODEL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { const void *data; int32_t ggml_type; int32_t rows; int32_t cols; } mrs_qweight;

typedef struct {
  mrs_qweight wq, wk, wv, wo, w_gate, w_up, w_down; /* local (possibly TP-sharded) shapes */
  const void *attn_norm, *ffn_norm;                  /* [hidden] in the activation dtype */
  void *k_cache, *v_cache;                           /* HND [num_blocks, kv_heads, block_size, head_dim] */
} mrs_llama_layer;

/* Tensor-parallel context of the peer-memory all-reduce (one process per GPU; the buffers are
 * symmetric allocations whose peer mappings the host obtained by rendezvous, e.g.
 * torch.distributed._symmetric_memory): every rank owns [flags | slot 0 | slot 1] at the same offsets. */
typedef struct {
  int32_t world, rank;
  void *peer_base[8];            /* base of rank r's buffer as mapped in THIS process (own included) */
  int64_t flags_offset;          /* >= world uint32 flag words, zero-initialised */
  int64_t slot_offset[2];        /* two partial buffers of >= batch * hidden activation elements each */
  void *seq_counter;             /* local device uint32[2], zero-initialised: [0] all-reduces so far, [1] set to 1 when a
                                  * low-latency all-reduce gave up waiting for a peer (~0.25 s): results are invalid */
  /* low-latency protocol (used when ll_offset != 0): every rank PUSHES its partial to every peer as 8-byte words
   * {two activation elements, sequence number} — one NVLink store hop, no flag round trip, no fence; the receiver
   * polls the words themselves.  Region: [2 slots][world source ranks][ll_src_stride bytes], zero-initialised,
   * ll_src_stride >= 4 bytes per element of the largest all-reduce. */
  int64_t ll_offset, ll_slot_stride, ll_src_stride;
} mrs_tp_ctx;
/* out = T(T(sum over ranks of slot partials, rank order, f32) + residual); count % 8 == 0; one CTA, in-graph,
 * PDL-chained.  Stands in for SumAllReduce::sum_all_reduce + the residual add (REF distributed/mod.rs:436-453).
 * Two protocols, same arithmetic and bit-identical results on every rank: flags + pull (ll_offset == 0) and the
 * low-latency push above. */
int32_t mrs_tp_allreduce_residual(const mrs_tp_ctx *ctx, int32_t slot, const void *residual, void *out, int32_t count,
                                  int32_t dtype, int32_t pdl, void *stream);

typedef struct {
  int32_t hidden, n_layers, n_heads, n_kv_heads, head_dim, vocab; /* heads are LOCAL counts under TP */
  int32_t block_size, act_dtype;   /* act_dtype: 0 f16, 1 bf16 */
  float rms_eps, sm_scale;
  int32_t rope_neox, pdl;
  const mrs_llama_layer *layers;   /* host array [n_layers] */
  mrs_qweight tok_embd, lm_head;
  const void *final_norm, *rope_cos, *rope_sin; /* rope tables [max_pos, head_dim/2] act dtype */
  /* per-step device metadata (see mrs_decode_advance) */
  int32_t batch, padded_tiles, max_blocks_per_seq;
  int32_t skip_mask;         /* measurement only: bit0 skip rope/cache/attention, bit1 skip the GEMVs */
  int32_t fused_attention;   /* 1: mrs_paged_decode_fused; 0: rotary + reshape_and_cache + flashinfer_decode */
  int32_t reserved0;
  int32_t *token_ids;        /* [batch] in: token to process; out_token may alias it */
  int32_t *positions;        /* [batch] */
  int64_t *slot_mapping;     /* [batch] */
  int32_t *kv_indptr, *kv_indices, *kv_last_page_len;
  int32_t *request_indices, *kv_tile_indices, *o_indptr, *kv_chunk_size;
  uint8_t *block_valid_mask;
  /* scratch (activation dtype unless noted) */
  void *x, *x2, *q, *k, *v, *attn_out, *act, *logits; /* x, x2: [batch, hidden] residual stream ping-pong */
  void *tmp_v; float *tmp_s;
  int32_t *out_token;        /* [batch] argmax of the logits */
  int32_t *attn_counters;    /* zeroed int32 [batch * n_kv_heads * 2] (fused attention merge) */
  void *argmax_scratch;      /* zeroed, >= 16 * batch + 16 bytes */
  /* tensor parallel: called after the row-parallel projections when non-NULL */
  void (*all_reduce)(void *buf, int64_t count, int32_t dtype, void *stream, void *user);
  void *all_reduce_user;
  const mrs_tp_ctx *tp;      /* non-NULL with world > 1: peer-memory sum (takes precedence over all_reduce) */
} mrs_llama_step;

/* Enqueue one decode step (all layers + lm_head + argmax) on `stream`. Returns cudaError. */
int32_t mrs_llama_decode_step(const mrs_llama_step *s, void *stream);

/* On-device restatement of the scheduler-side index producers for a running decode batch
 * (REF inputs_processor.rs:896-923, flashinfer/metadata.rs:88-216): context_lens[b] += 1,
 * positions, slot_mapping, the paged-KV CSR and the split-KV tile plan for the new lengths,
 * all from the dense block table — so a whole generation replays as one CUDA graph.
 * batch <= 256.  A sequence at min(max_blocks_per_seq*block_size, max_pos) (max_pos <= 0: no RoPE
 * bound) stops growing: slot -1 (no KV write) and bit 0 of *error_flag (nullable) is set. */
int32_t mrs_decode_advance(const int32_t *block_tables, int32_t max_blocks_per_seq, int32_t *context_lens,
                           int32_t batch, int32_t block_size, int32_t split_pages, int32_t padded_tiles,
                           int32_t *positions, int64_t *slot_mapping, int32_t *kv_indptr, int32_t *kv_indices,
                           int32_t *kv_last_page_len, int32_t *request_indices, int32_t *kv_tile_indices,
                           int32_t *o_indptr, int32_t *kv_chunk_size, uint8_t *block_valid_mask,
                           int32_t max_pos, int32_t *error_flag, void *stream);

/* rows of a quantised table -> activation dtype (embedding gather). ids on device. */
int32_t mrs_embedding_gather(int32_t ggml_type, const void *table, int32_t cols, const int32_t *ids, int32_t n,
                             void *out, int32_t act_dtype, void *stream);
/* out[b] = argmax_v logits[b, v] (first maximum), logits in act dtype */
int32_t mrs_argmax(const void *logits, int32_t rows, int32_t cols, int32_t act_dtype, int32_t *out, void *scratch,
                   int32_t pdl, void *stream);

/* ---- GPTQ / AWQ int4 decode stack (BASELINE config 4: Mistral-7B GPTQ g128, batch 32) --------------
 * Linears are int4 tiles produced by gptq_marlin_repack / awq_marlin_repack (mrs_b200_quant.h) with
 * UNPERMUTED scales [K/group, N] in the activation dtype; q||k||v and gate||up are concatenated along
 * N at load time.  cache_layout 1: HND cache + fused RoPE/KV-write/attention; 0: vLLM layout
 * (K [NB,KVH,D/8,BS,8], V [NB,KVH,D,BS]) through rotary + reshape_and_cache + paged_attention_v1. */
typedef struct { const void *tiles; const void *scales; const void *qzeros; int32_t k, n; } mrs_w4_weight;
typedef struct {
  mrs_w4_weight wqkv, wo, w_gate_up, w_down;
  const void *attn_norm, *ffn_norm;
  void *k_cache, *v_cache;
} mrs_gptq_layer;
typedef struct {
  int32_t hidden, n_layers, n_heads, n_kv_heads, head_dim, vocab, block_size, act_dtype, group_size;
  float rms_eps, sm_scale;
  int32_t rope_neox, cache_layout, batch, padded_tiles, max_blocks_per_seq, skip_mask;
  const mrs_gptq_layer *layers;            /* host array [n_layers] */
  const void *tok_embd, *lm_head;          /* dense [vocab, hidden] in the activation dtype */
  const void *final_norm, *rope_cos, *rope_sin;
  int32_t *token_ids, *positions; int64_t *slot_mapping;
  int32_t *kv_indptr, *kv_indices, *kv_last_page_len, *request_indices, *kv_tile_indices, *o_indptr, *kv_chunk_size;
  uint8_t *block_valid_mask;
  int32_t *block_tables, *context_lens;    /* dense table + lengths (vLLM-layout attention) */
  void *x, *x2, *h, *qkv, *attn_out, *o, *gate_up, *act, *logits, *tmp_v; float *tmp_s;
  int32_t *out_token, *attn_counters; void *argmax_scratch;
} mrs_gptq_step;
/* skip_mask: bit0 skip rope/cache/attention, bit1 skip the linears (measurement only), bit2 plain stream order instead of
 * the programmatic-dependent-launch chain (HND layout: every launch of the layer loop triggers its dependents at start
 * and waits for the upstream grid before touching its inputs/outputs, so each W4A16 GEMM streams weights while the
 * small kernel before it still runs). */
int32_t mrs_gptq_decode_step(const mrs_gptq_step *s, void *stream);
/* The chain's links (not reference ABI): mrs_w4a16_gemm / mrs_dense_linear / add_rms_norm / fused_split_glu with a
 * `pdl` flag.  pdl != 0 requires that the launch before it on `stream` is also a link (or a plain kernel). */
int32_t mrs_w4a16_gemm_pdl(const void *x, const void *w_tiles, const void *scales, const int32_t *qzeros, void *y,
                           int32_t M, int32_t K, int32_t N, int32_t group, int32_t dtype, int32_t scale_perm,
                           int32_t pdl, void *stream);
int32_t mrs_dense_linear_pdl(const void *x, const void *w, void *y, int32_t M, int32_t K, int32_t N, int32_t dtype,
                             int32_t pdl, void *stream);
void mrs_add_rms_norm_pdl(const void *x, const void *residual, const void *weight, void *residual_dst, void *norm_dst,
                          int32_t nrows, int32_t ncols, float eps, int32_t dtype, int32_t pdl, void *stream);
void mrs_split_glu_pdl(const void *input, void *output, uint32_t rows, uint32_t split_size, int32_t activation,
                       int32_t dtype, int32_t pdl, void *stream);

#ifdef __cplusplus
}
#endif
#endif
