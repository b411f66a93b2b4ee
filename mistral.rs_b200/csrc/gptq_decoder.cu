// gptq_decoder.cu — decode layer stack for GPTQ / AWQ int4 checkpoints (Mistral / Llama family)
// at serving batch sizes (BASELINE config 4: Mistral-7B GPTQ g128, batch 32, paged KV block 16).
//
// REF structure mirrored: mistralrs-core/src/models/mistral.rs (Attention / MLP / DecoderLayer
// forward over `QuantMethod` linears), with the linears on the reference's Marlin symbols
// (gptq/marlin_ffi.rs) — here the swap-AB tcgen05 kernel of w4a16.cu — and the lm_head as the dense
// 16-bit linear the checkpoint keeps.  Per layer (8 launches):
//   [RMSNorm] -> fused QKV GEMM -> RoPE + KV write + paged decode attention (one launch, HND; or the
//   rotary -> reshape_and_cache -> paged_attention_v1 chain for the vLLM layout) -> o_proj GEMM ->
//   add + RMSNorm -> gate||up GEMM -> SiLU*mul -> down GEMM -> add + RMSNorm (next layer's norm)
#include "common.cuh"
#include "mrs_b200_model.h"

#include <stdio.h>

extern "C" int32_t mrs_w4a16_gemm(const void *x, const void *w_tiles, const void *scales, const int32_t *qzeros, void *y,
                                  int32_t M, int32_t K, int32_t N, int32_t group, int32_t dtype, int32_t scale_perm,
                                  void *stream);
extern "C" int32_t mrs_dense_linear(const void *x, const void *w, void *y, int32_t M, int32_t K, int32_t N, int32_t dtype,
                                    void *stream);
extern "C" void mrs_rms_norm_f16(const void *x, const void *weight, void *dst, const int nrows, const int ncols, const float eps, int64_t stream);
extern "C" void mrs_rms_norm_bf16(const void *x, const void *weight, void *dst, const int nrows, const int ncols, const float eps, int64_t stream);
extern "C" void add_rms_norm_f16(const void *x, const void *residual, const void *weight, void *residual_dst, void *norm_dst, const int nrows, const int ncols, const float eps, int64_t stream);
extern "C" void add_rms_norm_bf16(const void *x, const void *residual, const void *weight, void *residual_dst, void *norm_dst, const int nrows, const int ncols, const float eps, int64_t stream);
extern "C" void fused_split_glu_f16(const void *input, void *output, uint32_t rows, uint32_t split_size, int activation, cudaStream_t stream);
extern "C" void fused_split_glu_bf16(const void *input, void *output, uint32_t rows, uint32_t split_size, int activation, cudaStream_t stream);
extern "C" int32_t mrs_paged_decode_fused_strided(void *q, void *k_new, void *v_new, void *key_cache, void *value_cache,
                                                  const void *rope_cos, const void *rope_sin, const int32_t *positions,
                                                  const int64_t *slot_mapping, const int32_t *kv_indptr,
                                                  const int32_t *kv_indices, const int32_t *kv_last_page_len,
                                                  const int32_t *request_indices, const int32_t *kv_tile_indices,
                                                  const int32_t *o_indptr, const int32_t *kv_chunk_size_ptr,
                                                  const uint8_t *block_valid_mask, void *o, void *tmp_v, float *tmp_s,
                                                  int32_t *counters, int32_t batch_size, int32_t padded_batch_size,
                                                  int32_t num_qo_heads, int32_t num_kv_heads, int32_t head_size,
                                                  int32_t page_size, float sm_scale, uint32_t dtype, int32_t pdl,
                                                  int64_t q_stride_n, int64_t kv_new_stride, void *stream);
extern "C" void rotary_embedding_positions(void *query, void *key, void *cos_cache, void *sin_cache, void *positions,
                                           int32_t is_neox, int32_t head_size, int64_t num_tokens, int32_t rot_dim,
                                           int32_t seq_len, int32_t num_heads, int32_t num_kv_heads, int64_t query_stride,
                                           int64_t key_stride, uint32_t dtype, int64_t stream);
extern "C" void reshape_and_cache(void *key, void *value, void *key_cache, void *value_cache, int64_t *slot_mapping,
                                  int32_t num_tokens, int32_t num_heads, int32_t head_size, int32_t block_size, int32_t x,
                                  int32_t key_stride, int32_t value_stride, cudaStream_t stream, uint32_t dtype,
                                  uint32_t cache_dtype, float *k_scale, float *v_scale);
extern "C" void paged_attention_v1_f16(void *out, void *query, void *key_cache, void *value_cache, void *alibi_slopes,
                                       int32_t num_kv_heads, float scale, float softcapping, uint32_t *block_tables,
                                       uint32_t *context_lens, int32_t block_size, int32_t max_context_len, int32_t num_seqs,
                                       int32_t num_heads, int32_t head_size, int32_t max_num_blocks_per_seq, int32_t q_stride,
                                       int32_t kv_block_stride, int32_t kv_head_stride, cudaStream_t stream,
                                       uint32_t cache_dtype, float *k_scale, float *v_scale, const float *sinks);
extern "C" void paged_attention_v1_bf16(void *out, void *query, void *key_cache, void *value_cache, void *alibi_slopes,
                                        int32_t num_kv_heads, float scale, float softcapping, uint32_t *block_tables,
                                        uint32_t *context_lens, int32_t block_size, int32_t max_context_len, int32_t num_seqs,
                                        int32_t num_heads, int32_t head_size, int32_t max_num_blocks_per_seq, int32_t q_stride,
                                        int32_t kv_block_stride, int32_t kv_head_stride, cudaStream_t stream,
                                        uint32_t cache_dtype, float *k_scale, float *v_scale, const float *sinks);
extern "C" int32_t mrs_argmax(const void *logits, int32_t rows, int32_t cols, int32_t act_dtype, int32_t *out, void *scratch,
                              int32_t pdl, void *stream);

namespace mrs {
// dense embedding rows: out[b, :] = table[ids[b], :]   (16-bit elements, 16-byte vectors)
__global__ void dense_embedding_kernel(const uint4 *__restrict__ table, int cols8, const int32_t *__restrict__ ids,
                                       uint4 *__restrict__ out) {
  const int64_t row = ids[blockIdx.x];
  for (int i = threadIdx.x; i < cols8; i += blockDim.x) out[(int64_t)blockIdx.x * cols8 + i] = table[row * cols8 + i];
}
}  // namespace mrs

#define MRS_TRY(expr)                                                \
  do {                                                               \
    const int _e = (int)(expr);                                      \
    if (_e != 0) {                                                   \
      fprintf(stderr, "mrs_b200: %s -> cudaError %d\n", #expr, _e);  \
      return _e;                                                     \
    }                                                                \
  } while (0)

extern "C" int32_t mrs_w4a16_gemm_pdl(const void *x, const void *w_tiles, const void *scales, const int32_t *qzeros, void *y,
                                      int32_t M, int32_t K, int32_t N, int32_t group, int32_t dtype, int32_t scale_perm,
                                      int32_t pdl, void *stream);
extern "C" int32_t mrs_dense_linear_pdl(const void *x, const void *w, void *y, int32_t M, int32_t K, int32_t N, int32_t dtype,
                                        int32_t pdl, void *stream);
extern "C" void mrs_add_rms_norm_pdl(const void *x, const void *residual, const void *weight, void *residual_dst, void *norm_dst,
                                     int32_t nrows, int32_t ncols, float eps, int32_t dtype, int32_t pdl, void *stream);
extern "C" void mrs_split_glu_pdl(const void *input, void *output, uint32_t rows, uint32_t split_size, int32_t activation,
                                  int32_t dtype, int32_t pdl, void *stream);

extern "C" int32_t mrs_gptq_decode_step(const mrs_gptq_step *s, void *stream) {
  const int dt = s->act_dtype, B = s->batch, H = s->hidden;
  const int nq = s->n_heads * s->head_dim, nkv = s->n_kv_heads * s->head_dim, nqkv = nq + 2 * nkv;
  cudaStream_t st = (cudaStream_t)stream;
  if (B < 1 || B > 256 || (dt != MRS_F16 && dt != MRS_BF16) || H % 8) return (int32_t)cudaErrorInvalidValue;
  const bool f16 = dt == MRS_F16;
  auto rms = [&](const void *x, const void *w, void *dst) {
    if (f16) mrs_rms_norm_f16(x, w, dst, B, H, s->rms_eps, (int64_t)stream); else mrs_rms_norm_bf16(x, w, dst, B, H, s->rms_eps, (int64_t)stream);
  };
  // Every launch of the layer loop is a link of ONE programmatic-dependent-launch chain (skip_mask bit 2 turns it
  // off): each kernel triggers its dependents when it starts and waits for the upstream grid before touching its
  // inputs / outputs, so the W4A16 GEMMs stream their weights while the small kernels before them still run.
  // The HND attention launch is PDL-capable; the vLLM-layout chain (reference-ABI kernels, no PDL forms) is not,
  // so that layout keeps plain stream order.
  const int pdl = ((s->skip_mask & 4) || s->cache_layout != 1) ? 0 : 1;
  auto add_rms = [&](const void *x, const void *res, const void *w, void *res_dst, void *norm_dst) {
    mrs_add_rms_norm_pdl(x, res, w, res_dst, norm_dst, B, H, s->rms_eps, dt, pdl, stream);
  };
  auto linear = [&](const mrs_w4_weight &w, const void *x, void *y) -> int {
    return mrs_w4a16_gemm_pdl(x, w.tiles, w.scales, (const int32_t *)w.qzeros, y, B, w.k, w.n, s->group_size, dt, 0, pdl, stream);
  };
  const bool do_attn = !(s->skip_mask & 1), do_lin = !(s->skip_mask & 2);

  mrs::dense_embedding_kernel<<<B, 256, 0, st>>>((const uint4 *)s->tok_embd, H / 8, s->token_ids, (uint4 *)s->x);
  void *x = s->x, *x2 = s->x2;   // residual stream ping-pong
  rms(x, s->layers[0].attn_norm, s->h);
  for (int l = 0; l < s->n_layers; l++) {
    const mrs_gptq_layer &L = s->layers[l];
    if (do_lin) MRS_TRY(linear(L.wqkv, s->h, s->qkv));
    void *q = s->qkv, *k = (char *)s->qkv + (size_t)nq * 2, *v = (char *)s->qkv + (size_t)(nq + nkv) * 2;
    if (do_attn && s->cache_layout == 1) {
      MRS_TRY(mrs_paged_decode_fused_strided(q, k, v, L.k_cache, L.v_cache, s->rope_cos, s->rope_sin, s->positions,
                                             s->slot_mapping, s->kv_indptr, s->kv_indices, s->kv_last_page_len,
                                             s->request_indices, s->kv_tile_indices, s->o_indptr, s->kv_chunk_size,
                                             s->block_valid_mask, s->attn_out, s->padded_tiles > B ? s->tmp_v : nullptr,
                                             s->padded_tiles > B ? s->tmp_s : nullptr, s->attn_counters, B, s->padded_tiles,
                                             s->n_heads, s->n_kv_heads, s->head_dim, s->block_size, s->sm_scale, (uint32_t)dt,
                                             (s->rope_neox ? 0 : 2) | pdl, nqkv, nqkv, stream));
    } else if (do_attn) {
      // vLLM cache layout (REF MISTRALRS_FLASHINFER_DECODE=0): rotary -> reshape_and_cache -> paged_attention_v1
      rotary_embedding_positions(q, k, (void *)s->rope_cos, (void *)s->rope_sin, s->positions, s->rope_neox, s->head_dim, B,
                                 s->head_dim / 2, 0, s->n_heads, s->n_kv_heads, nqkv, nqkv, (uint32_t)dt, (int64_t)stream);
      reshape_and_cache(k, v, L.k_cache, L.v_cache, s->slot_mapping, B, s->n_kv_heads, s->head_dim, s->block_size, 8, nqkv,
                        nqkv, st, (uint32_t)dt, (uint32_t)dt, nullptr, nullptr);
      const int kv_block_stride = s->n_kv_heads * s->head_dim * s->block_size, kv_head_stride = s->head_dim * s->block_size;
      if (f16)
        paged_attention_v1_f16(s->attn_out, q, L.k_cache, L.v_cache, nullptr, s->n_kv_heads, s->sm_scale, 1.0f,
                               (uint32_t *)s->block_tables, (uint32_t *)s->context_lens, s->block_size,
                               s->max_blocks_per_seq * s->block_size, B, s->n_heads, s->head_dim, s->max_blocks_per_seq, nqkv,
                               kv_block_stride, kv_head_stride, st, (uint32_t)dt, nullptr, nullptr, nullptr);
      else
        paged_attention_v1_bf16(s->attn_out, q, L.k_cache, L.v_cache, nullptr, s->n_kv_heads, s->sm_scale, 1.0f,
                                (uint32_t *)s->block_tables, (uint32_t *)s->context_lens, s->block_size,
                                s->max_blocks_per_seq * s->block_size, B, s->n_heads, s->head_dim, s->max_blocks_per_seq, nqkv,
                                kv_block_stride, kv_head_stride, st, (uint32_t)dt, nullptr, nullptr, nullptr);
    }
    if (do_lin) MRS_TRY(linear(L.wo, s->attn_out, s->o));
    add_rms(s->o, x, L.ffn_norm, x2, s->h);                                   // x2 = o + x ; h = norm(x2)
    if (do_lin) MRS_TRY(linear(L.w_gate_up, s->h, s->gate_up));
    mrs_split_glu_pdl(s->gate_up, s->act, B, L.w_down.k, 0, dt, pdl, stream);
    if (do_lin) MRS_TRY(linear(L.w_down, s->act, s->o));
    const void *next_norm = (l + 1 < s->n_layers) ? s->layers[l + 1].attn_norm : s->final_norm;
    add_rms(s->o, x2, next_norm, x, s->h);                                     // x = down + x2 ; h = next norm(x)
  }
  if (do_lin) MRS_TRY(mrs_dense_linear_pdl(s->h, s->lm_head, s->logits, B, H, s->vocab, dt, pdl, stream));
  MRS_TRY(mrs_argmax(s->logits, B, s->vocab, dt, s->out_token, s->argmax_scratch, 0, stream));
  return (int32_t)cudaGetLastError();
}
