/* mrs_b200_ops.h — C ABI of the small fused ops between the GEMMs (libmrs_b200.so). */
#ifndef MRS_B200_OPS_H
#define MRS_B200_OPS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct CUstream_st *mrs_ops_stream_t;

/* REF mistralrs-quant/src/rotary/ffi.rs:4-43, kernels/rotary/rotary.cu:110-196.
 * `rot_dim` is the number of rotated PAIRS (cos/sin row length), as in the reference. */
void rotary_embedding(void *query, void *key, void *cos_cache, void *sin_cache, int32_t is_neox, int32_t head_size,
                      int64_t num_tokens, int32_t rot_dim, int32_t num_heads, int32_t num_kv_heads,
                      int64_t query_stride, int64_t key_stride, uint32_t dtype, int64_t stream);
void rotary_embedding_positions(void *query, void *key, void *cos_cache, void *sin_cache, void *positions,
                                int32_t is_neox, int32_t head_size, int64_t num_tokens, int32_t rot_dim,
                                int32_t seq_len, int32_t num_heads, int32_t num_kv_heads, int64_t query_stride,
                                int64_t key_stride, uint32_t dtype, int64_t stream);

/* REF mistralrs-quant/src/utils/ffi.rs:274-330, kernels/ops/ops.cu:848-1060 */
#define MRS_GLU_DECL(t)                                                                                        \
  void fused_glu_##t(const void *a, const void *b, void *output, uint32_t rows, uint32_t cols,                 \
                     uint32_t a_row_stride, uint32_t b_row_stride, int activation, mrs_ops_stream_t stream);   \
  void fused_split_glu_##t(const void *input, void *output, uint32_t rows, uint32_t split_size, int activation, \
                           mrs_ops_stream_t stream);
MRS_GLU_DECL(f16) MRS_GLU_DECL(bf16) MRS_GLU_DECL(f32)

/* REF mistralrs-core/src/cuda/ffi.rs (add_rms_norm_*), sort.cu:403-461,701-727; mrs_rms_norm_* is
 * the plain RMSNorm the reference gets from candle_nn::ops::rms_norm (core/src/layers.rs:403-413). */
#define MRS_RMS_DECL(t)                                                                                         \
  void add_rms_norm_##t(const void *x, const void *residual, const void *weight, void *residual_dst,            \
                        void *norm_dst, const int nrows, const int ncols, const float eps, int64_t stream);     \
  void mrs_rms_norm_##t(const void *x, const void *weight, void *dst, const int nrows, const int ncols,         \
                        const float eps, int64_t stream);
MRS_RMS_DECL(f16) MRS_RMS_DECL(bf16) MRS_RMS_DECL(f32)

/* REF mistralrs-core/src/cuda/ffi.rs:183-227, sort.cu:619-672: per-head RMSNorm of a strided
 * [batch, heads, seq, head_dim] view (element strides) into a contiguous tensor of that shape. */
#define MRS_RMS4D_DECL(t)                                                                                       \
  void rms_norm_strided_4d_##t(const void *x, const void *weight, void *dst, int64_t stride_b, int64_t stride_h, \
                               int64_t stride_s, int64_t stride_d, int32_t batch, int32_t heads, int32_t seq_len, \
                               int32_t head_dim, float eps, int64_t stream);
MRS_RMS4D_DECL(f16) MRS_RMS4D_DECL(bf16) MRS_RMS4D_DECL(f32)
/* ---- sampling tail — REF mistralrs-core/src/cuda/ffi.rs:581-665, sort.cu:1503-2262 (csrc/sampler.cu).
 * top-k of raw f32 logits in (value desc, index asc) order + softmax statistics at inv_temperature:
 * packed_out = [values k | indices as f32 k | denom | global_max]; scratch arrays as the Rust callers
 * size them (block_values / block_indices: nblocks*k per row, block_maxes / block_sums: nblocks per row);
 * chunk_size <= 2048, nblocks * k <= 47 * 1024.  top-1: packed [value, token id] and / or token ids;
 * a NaN logit poisons the row (NaN / UINT32_MAX). */
void topk_large_f32(const float *input, float *block_values, uint32_t *block_indices, float *block_maxes, float *block_sums,
                    float *values_out, uint32_t *indices_out, float *softmax_info_out, int ncols, int k, int chunk_size,
                    int nblocks, float inv_temperature, int64_t stream);
void topk_large_f32_packed(const float *input, float *block_values, uint32_t *block_indices, float *block_maxes,
                           float *block_sums, float *packed_out, int ncols, int k, int chunk_size, int nblocks,
                           float inv_temperature, int64_t stream);
void topk_large_f32_packed_batched(const float *input, const float *inv_temperatures, float *block_values,
                                   uint32_t *block_indices, float *block_maxes, float *block_sums, float *packed_out,
                                   int nrows, int ncols, int k, int chunk_size, int nblocks, int64_t stream);
void top1_large_f32_packed(const float *input, float *block_values, uint32_t *block_indices, float *packed_out,
                           uint32_t *token_ids_out, int ncols, int chunk_size, int nblocks, int64_t stream);
void top1_large_f32_packed_batched(const float *input, float *block_values, uint32_t *block_indices, float *packed_out,
                                   uint32_t *token_ids_out, int nrows, int ncols, int chunk_size, int nblocks, int64_t stream);
#ifdef __cplusplus
}
#endif
#endif
