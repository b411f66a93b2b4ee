// cache_ops.cu — paged KV-cache data movement behind the reference's C symbols
// (REF: mistralrs-paged-attn/src/cuda/ffi.rs:96-116,159-176,212-268 and the kernels in
//  reshape_and_cache_kernel.cu:32-87, flashinfer_decode.cu:20-100, gather_kv_cache_kernel.cu,
//  copy_blocks_kernel.cu).
//
// Two cache layouts, both the reference's:
//   vLLM      K [NB, KVH, D/x, BS, x]   V [NB, KVH, D, BS]
//   HND       K,V [NB, KVH, BS, D]      (FlashInfer; the default for Llama/Mistral shapes)
// dtype codes 0=f16 1=bf16 2=f32 3=fp8_e4m3 (cache only).  16-bit types are moved as raw
// uint16 (bit-exact); FP8 follows the reference's scaled conversion.
#include "common.cuh"

#include <cuda_fp8.h>
#include <stdio.h>
#include <stdlib.h>

namespace mrs {

template <typename T> struct Conv;
template <> struct Conv<uint16_t> {  // raw 16-bit (f16 or bf16: pure copy)
  __device__ static __forceinline__ uint16_t id(uint16_t v) { return v; }
};

// value -> float for the fp8 paths
template <int DT> __device__ __forceinline__ float widen(const void *p, int64_t i) {
  if constexpr (DT == MRS_F16) return __half2float(((const __half *)p)[i]);
  else if constexpr (DT == MRS_BF16) return __bfloat162float(((const __nv_bfloat16 *)p)[i]);
  else return ((const float *)p)[i];
}
template <int DT> __device__ __forceinline__ void narrow(void *p, int64_t i, float v) {
  if constexpr (DT == MRS_F16) ((__half *)p)[i] = __float2half_rn(v);
  else if constexpr (DT == MRS_BF16) ((__nv_bfloat16 *)p)[i] = __float2bfloat16_rn(v);
  else ((float *)p)[i] = v;
}

// ---------------------------------------------------------------- reshape_and_cache
// LAYOUT 0 vLLM, 1 HND.  ESZ: element bytes of a same-dtype copy (2 or 4); FP8: cache is e4m3.
template <int LAYOUT, typename E>
__global__ void reshape_and_cache_copy(const E *__restrict__ key, const E *__restrict__ value, E *__restrict__ kc,
                                       E *__restrict__ vc, const int64_t *__restrict__ slot_mapping, int key_stride,
                                       int value_stride, int num_heads, int head_size, int block_size, int x) {
  const int64_t t = blockIdx.x;
  const int64_t slot = slot_mapping[t];
  if (slot < 0) return;  // padding token (REF reshape_and_cache_kernel.cu:45-48)
  const int64_t blk = slot / block_size, off = slot % block_size;
  const int n = num_heads * head_size;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int h = i / head_size, d = i - h * head_size;
    int64_t ki, vi;
    if constexpr (LAYOUT == 0) {
      ki = blk * num_heads * head_size * block_size + (int64_t)h * head_size * block_size +
           (int64_t)(d / x) * block_size * x + off * x + (d % x);
      vi = blk * num_heads * head_size * block_size + (int64_t)h * head_size * block_size + (int64_t)d * block_size + off;
    } else {
      ki = vi = ((blk * num_heads + h) * block_size + off) * head_size + d;
    }
    kc[ki] = key[t * key_stride + i];
    vc[vi] = value[t * value_stride + i];
  }
}

template <int LAYOUT, int DT>
__global__ void reshape_and_cache_fp8(const void *__restrict__ key, const void *__restrict__ value,
                                      __nv_fp8_e4m3 *__restrict__ kc, __nv_fp8_e4m3 *__restrict__ vc,
                                      const int64_t *__restrict__ slot_mapping, int key_stride, int value_stride,
                                      int num_heads, int head_size, int block_size, int x, float k_scale, float v_scale,
                                      const float *k_scale_p, const float *v_scale_p) {
  const int64_t t = blockIdx.x;
  const int64_t slot = slot_mapping[t];
  if (slot < 0) return;
  if (k_scale_p != nullptr) { k_scale = *k_scale_p; v_scale = *v_scale_p; }
  const int64_t blk = slot / block_size, off = slot % block_size;
  const int n = num_heads * head_size;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int h = i / head_size, d = i - h * head_size;
    int64_t ki, vi;
    if constexpr (LAYOUT == 0) {
      ki = blk * num_heads * head_size * block_size + (int64_t)h * head_size * block_size +
           (int64_t)(d / x) * block_size * x + off * x + (d % x);
      vi = blk * num_heads * head_size * block_size + (int64_t)h * head_size * block_size + (int64_t)d * block_size + off;
    } else {
      ki = vi = ((blk * num_heads + h) * block_size + off) * head_size + d;
    }
    kc[ki] = __nv_fp8_e4m3(widen<DT>(key, t * key_stride + i) / k_scale);
    vc[vi] = __nv_fp8_e4m3(widen<DT>(value, t * value_stride + i) / v_scale);
  }
}

template <int LAYOUT>
static void launch_reshape_and_cache(void *key, void *value, void *kc, void *vc, const int64_t *slot_mapping,
                                     int num_tokens, int num_heads, int head_size, int block_size, int x,
                                     int key_stride, int value_stride, uint32_t dtype, uint32_t cache_dtype,
                                     float k_scale, float v_scale, const float *ksp, const float *vsp,
                                     cudaStream_t st) {
  if (num_tokens <= 0) return;
  dim3 grid(num_tokens);
  int threads = num_heads * head_size;
  threads = threads < 512 ? ((threads + 31) / 32) * 32 : 512;
  if (cache_dtype == 3) {
    auto K = (__nv_fp8_e4m3 *)kc;
    auto V = (__nv_fp8_e4m3 *)vc;
    if (dtype == 0) reshape_and_cache_fp8<LAYOUT, MRS_F16><<<grid, threads, 0, st>>>(key, value, K, V, slot_mapping, key_stride, value_stride, num_heads, head_size, block_size, x, k_scale, v_scale, ksp, vsp);
    else if (dtype == 1) reshape_and_cache_fp8<LAYOUT, MRS_BF16><<<grid, threads, 0, st>>>(key, value, K, V, slot_mapping, key_stride, value_stride, num_heads, head_size, block_size, x, k_scale, v_scale, ksp, vsp);
    else if (dtype == 2) reshape_and_cache_fp8<LAYOUT, MRS_F32><<<grid, threads, 0, st>>>(key, value, K, V, slot_mapping, key_stride, value_stride, num_heads, head_size, block_size, x, k_scale, v_scale, ksp, vsp);
  } else if (dtype == cache_dtype && (dtype == 0 || dtype == 1)) {
    reshape_and_cache_copy<LAYOUT, uint16_t><<<grid, threads, 0, st>>>((const uint16_t *)key, (const uint16_t *)value, (uint16_t *)kc, (uint16_t *)vc, slot_mapping, key_stride, value_stride, num_heads, head_size, block_size, x);
  } else if (dtype == 2 && cache_dtype == 2) {
    reshape_and_cache_copy<LAYOUT, uint32_t><<<grid, threads, 0, st>>>((const uint32_t *)key, (const uint32_t *)value, (uint32_t *)kc, (uint32_t *)vc, slot_mapping, key_stride, value_stride, num_heads, head_size, block_size, x);
  } else {
    fprintf(stderr, "mrs_b200: reshape_and_cache received unsupported dtype pair %u/%u\n", dtype, cache_dtype);
  }
}

// ---------------------------------------------------------------- gather (paged -> dense)
// One block per output token; cu_seq_lens [num_seqs+1] locates the sequence
// (REF flashinfer_decode.cu:56-100, gather_kv_cache_kernel.cu).
template <int LAYOUT, typename E>
__global__ void gather_kv_copy(const E *__restrict__ kc, const E *__restrict__ vc, E *__restrict__ k_out,
                               E *__restrict__ v_out, const int32_t *__restrict__ block_table,
                               const int32_t *__restrict__ cu_seq_lens, int num_tokens, int num_seqs, int block_size,
                               int block_table_stride, int num_kv_heads, int head_size, int x) {
  const int t = blockIdx.x;
  if (t >= num_tokens) return;
  int lo = 0, hi = num_seqs;  // largest seq with cu[seq] <= t
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cu_seq_lens[mid] <= t) lo = mid; else hi = mid;
  }
  const int seq = lo;
  const int so = t - cu_seq_lens[seq];
  const int64_t blk = block_table[(int64_t)seq * block_table_stride + so / block_size];
  const int off = so % block_size;
  const int n = num_kv_heads * head_size;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int h = i / head_size, d = i - h * head_size;
    int64_t ki, vi;
    if constexpr (LAYOUT == 0) {
      ki = blk * num_kv_heads * head_size * block_size + (int64_t)h * head_size * block_size +
           (int64_t)(d / x) * block_size * x + (int64_t)off * x + (d % x);
      vi = blk * num_kv_heads * head_size * block_size + (int64_t)h * head_size * block_size + (int64_t)d * block_size + off;
    } else {
      ki = vi = ((blk * num_kv_heads + h) * block_size + off) * head_size + d;
    }
    k_out[(int64_t)t * n + i] = kc[ki];
    v_out[(int64_t)t * n + i] = vc[vi];
  }
}

template <int LAYOUT, int DT>
__global__ void gather_kv_fp8(const __nv_fp8_e4m3 *__restrict__ kc, const __nv_fp8_e4m3 *__restrict__ vc,
                              void *__restrict__ k_out, void *__restrict__ v_out,
                              const int32_t *__restrict__ block_table, const int32_t *__restrict__ cu_seq_lens,
                              int num_tokens, int num_seqs, int block_size, int block_table_stride, int num_kv_heads,
                              int head_size, int x, float k_scale, float v_scale, const float *k_scale_p,
                              const float *v_scale_p) {
  const int t = blockIdx.x;
  if (t >= num_tokens) return;
  if (k_scale_p != nullptr) { k_scale = *k_scale_p; v_scale = *v_scale_p; }
  int lo = 0, hi = num_seqs;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cu_seq_lens[mid] <= t) lo = mid; else hi = mid;
  }
  const int seq = lo;
  const int so = t - cu_seq_lens[seq];
  const int64_t blk = block_table[(int64_t)seq * block_table_stride + so / block_size];
  const int off = so % block_size;
  const int n = num_kv_heads * head_size;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int h = i / head_size, d = i - h * head_size;
    int64_t ki, vi;
    if constexpr (LAYOUT == 0) {
      ki = blk * num_kv_heads * head_size * block_size + (int64_t)h * head_size * block_size +
           (int64_t)(d / x) * block_size * x + (int64_t)off * x + (d % x);
      vi = blk * num_kv_heads * head_size * block_size + (int64_t)h * head_size * block_size + (int64_t)d * block_size + off;
    } else {
      ki = vi = ((blk * num_kv_heads + h) * block_size + off) * head_size + d;
    }
    narrow<DT>(k_out, (int64_t)t * n + i, (float)kc[ki] * k_scale);
    narrow<DT>(v_out, (int64_t)t * n + i, (float)vc[vi] * v_scale);
  }
}

template <int LAYOUT>
static void launch_gather(void *kc, void *vc, void *k_out, void *v_out, const int32_t *block_table,
                          const int32_t *cu_seq_lens, int num_tokens, int num_seqs, int block_size,
                          int block_table_stride, int num_kv_heads, int head_size, int x, uint32_t out_dtype,
                          uint32_t cache_dtype, float k_scale, float v_scale, const float *ksp, const float *vsp,
                          cudaStream_t st) {
  if (num_tokens <= 0) return;
  dim3 grid(num_tokens);
  int threads = num_kv_heads * head_size;
  threads = threads < 512 ? ((threads + 31) / 32) * 32 : 512;
  if (cache_dtype == 3) {
    auto K = (const __nv_fp8_e4m3 *)kc;
    auto V = (const __nv_fp8_e4m3 *)vc;
    if (out_dtype == 0) gather_kv_fp8<LAYOUT, MRS_F16><<<grid, threads, 0, st>>>(K, V, k_out, v_out, block_table, cu_seq_lens, num_tokens, num_seqs, block_size, block_table_stride, num_kv_heads, head_size, x, k_scale, v_scale, ksp, vsp);
    else if (out_dtype == 1) gather_kv_fp8<LAYOUT, MRS_BF16><<<grid, threads, 0, st>>>(K, V, k_out, v_out, block_table, cu_seq_lens, num_tokens, num_seqs, block_size, block_table_stride, num_kv_heads, head_size, x, k_scale, v_scale, ksp, vsp);
    else if (out_dtype == 2) gather_kv_fp8<LAYOUT, MRS_F32><<<grid, threads, 0, st>>>(K, V, k_out, v_out, block_table, cu_seq_lens, num_tokens, num_seqs, block_size, block_table_stride, num_kv_heads, head_size, x, k_scale, v_scale, ksp, vsp);
  } else if (out_dtype == cache_dtype && (out_dtype == 0 || out_dtype == 1)) {
    gather_kv_copy<LAYOUT, uint16_t><<<grid, threads, 0, st>>>((const uint16_t *)kc, (const uint16_t *)vc, (uint16_t *)k_out, (uint16_t *)v_out, block_table, cu_seq_lens, num_tokens, num_seqs, block_size, block_table_stride, num_kv_heads, head_size, x);
  } else if (out_dtype == 2 && cache_dtype == 2) {
    gather_kv_copy<LAYOUT, uint32_t><<<grid, threads, 0, st>>>((const uint32_t *)kc, (const uint32_t *)vc, (uint32_t *)k_out, (uint32_t *)v_out, block_table, cu_seq_lens, num_tokens, num_seqs, block_size, block_table_stride, num_kv_heads, head_size, x);
  } else {
    fprintf(stderr, "mrs_b200: gather_kv_cache received unsupported dtype pair %u/%u\n", out_dtype, cache_dtype);
  }
}

// ---------------------------------------------------------------- copy_blocks (copy-on-write)
// key/value_cache_ptrs: device arrays of per-layer cache base addresses; block_mapping
// [num_pairs, 2] int64 (src, dst).  Grid (layers, pairs) — REF copy_blocks_kernel.cu:1-37.
template <typename E>
__global__ void copy_blocks_kernel(int64_t *key_cache_ptrs, int64_t *value_cache_ptrs,
                                   const int64_t *__restrict__ block_mapping, int numel_key, int numel_value) {
  E *kc = (E *)key_cache_ptrs[blockIdx.x];
  E *vc = (E *)value_cache_ptrs[blockIdx.x];
  const int64_t src = block_mapping[2 * blockIdx.y], dst = block_mapping[2 * blockIdx.y + 1];
  for (int i = threadIdx.x; i < numel_key; i += blockDim.x) kc[dst * numel_key + i] = kc[src * numel_key + i];
  for (int i = threadIdx.x; i < numel_value; i += blockDim.x) vc[dst * numel_value + i] = vc[src * numel_value + i];
}

}  // namespace mrs

using namespace mrs;

extern "C" void reshape_and_cache(void *key, void *value, void *key_cache, void *value_cache, int64_t *slot_mapping,
                                  int32_t num_tokens, int32_t num_heads, int32_t head_size, int32_t block_size,
                                  int32_t x, int32_t key_stride, int32_t value_stride, cudaStream_t stream,
                                  uint32_t dtype, uint32_t cache_dtype, float *k_scale, float *v_scale) {
  // vLLM-layout ABI: FP8 scales are device pointers, dereferenced inside the kernel
  launch_reshape_and_cache<0>(key, value, key_cache, value_cache, slot_mapping, num_tokens, num_heads, head_size,
                              block_size, x, key_stride, value_stride, dtype, cache_dtype, 1.f, 1.f, k_scale, v_scale,
                              stream);
}

extern "C" void reshape_and_cache_flashinfer(void *key, void *value, void *key_cache, void *value_cache,
                                             int64_t *slot_mapping, int32_t num_tokens, int32_t num_heads,
                                             int32_t head_size, int32_t block_size, int32_t key_stride,
                                             int32_t value_stride, float k_scale, float v_scale, uint32_t dtype,
                                             uint32_t cache_dtype, cudaStream_t stream) {
  launch_reshape_and_cache<1>(key, value, key_cache, value_cache, slot_mapping, num_tokens, num_heads, head_size,
                              block_size, 1, key_stride, value_stride, dtype, cache_dtype, k_scale, v_scale, nullptr, nullptr, stream);
}

extern "C" void gather_kv_cache_flashinfer(void *key_cache, void *value_cache, void *k_out, void *v_out,
                                           const int32_t *block_table, const int32_t *cu_seq_lens, int32_t num_tokens,
                                           int32_t num_seqs, int32_t block_size, int32_t block_table_stride,
                                           int32_t num_kv_heads, int32_t head_size, uint32_t out_dtype,
                                           uint32_t cache_dtype, float k_scale, float v_scale, cudaStream_t stream) {
  launch_gather<1>(key_cache, value_cache, k_out, v_out, block_table, cu_seq_lens, num_tokens, num_seqs, block_size,
                   block_table_stride, num_kv_heads, head_size, 1, out_dtype, cache_dtype, k_scale, v_scale, nullptr, nullptr, stream);
}

extern "C" void gather_kv_cache(void *key_cache, void *value_cache, void *k_out, void *v_out, const float *k_scale,
                                const float *v_scale, const int32_t *block_table, const int32_t *cu_seq_lens,
                                int32_t num_tokens, int32_t num_seqs, int32_t block_size, int32_t block_table_stride,
                                int32_t num_kv_heads, int32_t head_size, int32_t x, cudaStream_t stream,
                                uint32_t out_dtype, uint32_t cache_dtype) {
  launch_gather<0>(key_cache, value_cache, k_out, v_out, block_table, cu_seq_lens, num_tokens, num_seqs, block_size,
                   block_table_stride, num_kv_heads, head_size, x, out_dtype, cache_dtype, 1.f, 1.f, k_scale, v_scale,
                   stream);
}

#define MRS_COPY_BLOCKS(tag, E)                                                                              \
  extern "C" void copy_blocks_##tag(int64_t *key_cache_ptrs, int64_t *value_cache_ptrs,                      \
                                    const int64_t *block_mapping, int32_t num_layers, int32_t num_pairs,     \
                                    int32_t numel_per_block_key, int32_t numel_per_block_value,              \
                                    int64_t stream) {                                                        \
    if (num_layers <= 0 || num_pairs <= 0) return;                                                           \
    dim3 grid(num_layers, num_pairs);                                                                        \
    copy_blocks_kernel<E><<<grid, 512, 0, (cudaStream_t)stream>>>(key_cache_ptrs, value_cache_ptrs,          \
                                                                  block_mapping, numel_per_block_key,        \
                                                                  numel_per_block_value);                    \
  }
MRS_COPY_BLOCKS(f32, uint32_t)
MRS_COPY_BLOCKS(f16, uint16_t)
MRS_COPY_BLOCKS(bf16, uint16_t)
MRS_COPY_BLOCKS(u8, uint8_t)
