// sampler.cu — the sampling tail of a decode step on the device: top-k (+ softmax statistics for the
// host's top-p / temperature sampling) and greedy top-1 over f32 logits, behind the reference's symbols
//   topk_large_f32, topk_large_f32_packed, topk_large_f32_packed_batched,
//   top1_large_f32_packed, top1_large_f32_packed_batched
// REF mistralrs-core/src/cuda/ffi.rs:581-665, kernels mistralrs-core/src/cuda/sort.cu:1503-2262, callers
// mistralrs-core/src/ops.rs:690-830,1915 (chunk_size 2048, k <= 128).  SURVEY §8(f) rank 3.
//
// Output contract restated from the reference kernels (bit-exact for values and indices):
//   top-k  : the k largest RAW logits in (value descending, token index ascending) order — NaN and
//            -inf never selected (missing entries: value -inf, index 0) — as values[k] / indices[k], or
//            packed [values k | indices-as-float k | denom | global_max] with
//            global_max = max_j x_j * inv_T,  denom = sum_j exp(x_j * inv_T - global_max)
//            (per-chunk partial sums rescaled; f32 summation order differs -> tolerance 1e-6 rel).
//   top-1  : packed [max value, token id as float] and / or token_ids_out; any NaN in the row gives
//            NaN / UINT32_MAX.
// Same two-launch shape and the caller's scratch arrays (block_values / block_indices / block_maxes /
// block_sums, sized by the Rust side as in the reference), but the per-chunk selection keeps each
// thread's 8 candidates in registers as packed 64-bit keys (order-preserving value bits << 32 |
// ~index) instead of re-reading the chunk from global memory k times.
#include "common.cuh"

#include <math_constants.h>
#include <stdio.h>

namespace mrs {

constexpr int SP_THREADS = 256, SP_PER_THREAD = 8;   // chunk_size <= 2048

__device__ __forceinline__ unsigned long long sp_key(float v, uint32_t idx) {
  if (v != v || v == -INFINITY) return 0ull;           // never selected (REF: `candidate > local_max`, NaN skipped)
  uint32_t b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float sp_key_value(unsigned long long k) {
  const uint32_t b = (uint32_t)(k >> 32);
  return __uint_as_float((b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b);
}
__device__ __forceinline__ uint32_t sp_key_index(unsigned long long k) { return 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull); }

__device__ __forceinline__ unsigned long long sp_block_max(unsigned long long v, unsigned long long *sm) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
    v = o > v ? o : v;
  }
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long best = 0ull;
#pragma unroll
  for (int w = 0; w < SP_THREADS / 32; w++) best = sm[w] > best ? sm[w] : best;
  return best;
}
__device__ __forceinline__ float sp_block_sum(float v, float *sm) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < SP_THREADS / 32; w++) t += sm[w];
  __syncthreads();
  return t;
}

// stage 1: per chunk top-k (sorted) + chunk max / sum of exp
__global__ void __launch_bounds__(SP_THREADS) topk_stage1_kernel(const float *__restrict__ input, float *__restrict__ block_values,
                                                               uint32_t *__restrict__ block_indices, float *__restrict__ block_maxes,
                                                               float *__restrict__ block_sums, int ncols, int k, int chunk_size,
                                                               const float *__restrict__ inv_temperatures, float scalar_inv_t) {
  __shared__ unsigned long long sm_key[2][SP_THREADS / 32];
  __shared__ float sm_f[SP_THREADS / 32];
  const size_t row = blockIdx.y;
  const int chunk = blockIdx.x, nblocks = gridDim.x, tid = threadIdx.x;
  input += row * (size_t)ncols;
  block_values += row * (size_t)nblocks * k; block_indices += row * (size_t)nblocks * k;
  block_maxes += row * (size_t)nblocks; block_sums += row * (size_t)nblocks;
  const float inv_t = inv_temperatures ? inv_temperatures[row] : scalar_inv_t;
  const int start = chunk * chunk_size, end = min(start + chunk_size, ncols);
  unsigned long long key[SP_PER_THREAD];
  float xv[SP_PER_THREAD];
#pragma unroll
  for (int i = 0; i < SP_PER_THREAD; i++) {
    const int idx = start + tid + i * SP_THREADS;
    const bool in = idx < end;
    xv[i] = in ? input[idx] : -INFINITY;
    key[i] = in ? sp_key(xv[i], (uint32_t)idx) : 0ull;
  }
  float top_value = -INFINITY;
  for (int ki = 0; ki < k; ki++) {
    unsigned long long mine = 0ull;
#pragma unroll
    for (int i = 0; i < SP_PER_THREAD; i++) mine = key[i] > mine ? key[i] : mine;
    const unsigned long long best = sp_block_max(mine, sm_key[ki & 1]);
    if (best != 0ull && mine == best) {
#pragma unroll
      for (int i = 0; i < SP_PER_THREAD; i++) if (key[i] == best) key[i] = 0ull;
    }
    if (tid == 0) {
      block_values[chunk * k + ki] = best ? sp_key_value(best) : -INFINITY;
      block_indices[chunk * k + ki] = best ? sp_key_index(best) : 0u;
    }
    if (ki == 0) top_value = best ? sp_key_value(best) : -INFINITY;
  }
  const float block_max = (end > start) ? top_value * inv_t : -INFINITY;
  float local = 0.f;
#pragma unroll
  for (int i = 0; i < SP_PER_THREAD; i++) {
    if (start + tid + i * SP_THREADS < end) {
      if (xv[i] != xv[i]) local = CUDART_NAN_F;
      else if (block_max != -INFINITY) local += expf(xv[i] * inv_t - block_max);
    }
  }
  __syncthreads();
  const float s = sp_block_sum(local, sm_f);
  if (tid == 0) { block_maxes[chunk] = block_max; block_sums[chunk] = s; }
}

// stage 2: global top-k among the chunk candidates + softmax statistics
__global__ void __launch_bounds__(SP_THREADS) topk_stage2_kernel(const float *__restrict__ block_values, const uint32_t *__restrict__ block_indices,
                                                               const float *__restrict__ block_maxes, const float *__restrict__ block_sums,
                                                               float *__restrict__ values_out, uint32_t *__restrict__ indices_out,
                                                               float *__restrict__ info_out, float *__restrict__ packed_out, int nblocks,
                                                               int k) {
  extern __shared__ unsigned char sp_used[];
  __shared__ unsigned long long sm_key[2][SP_THREADS / 32];
  __shared__ float sm_f[SP_THREADS / 32];
  const size_t row = blockIdx.x;
  const int tid = threadIdx.x, n = nblocks * k;
  block_values += row * (size_t)n; block_indices += row * (size_t)n;
  block_maxes += row * (size_t)nblocks; block_sums += row * (size_t)nblocks;
  if (packed_out) packed_out += row * (size_t)(2 * k + 2);
  for (int i = tid; i < n; i += SP_THREADS) sp_used[i] = 0;
  float gm = -INFINITY;
  for (int b = tid; b < nblocks; b += SP_THREADS) gm = fmaxf(gm, block_maxes[b]);
  gm = warp_max(gm);
  if ((tid & 31) == 0) sm_f[tid >> 5] = gm;
  __syncthreads();
  gm = -INFINITY;
#pragma unroll
  for (int w = 0; w < SP_THREADS / 32; w++) gm = fmaxf(gm, sm_f[w]);
  __syncthreads();
  float ld = 0.f;
  if (gm != -INFINITY)
    for (int b = tid; b < nblocks; b += SP_THREADS) ld += block_sums[b] * expf(block_maxes[b] - gm);
  const float denom = sp_block_sum(ld, sm_f);
  if (tid == 0) {
    if (packed_out) { packed_out[2 * k] = denom; packed_out[2 * k + 1] = gm; }
    if (info_out) { info_out[0] = denom; info_out[1] = gm; }
  }
  for (int ki = 0; ki < k; ki++) {
    unsigned long long mine = 0ull;
    int mpos = -1;
    for (int pos = tid; pos < n; pos += SP_THREADS) {
      if (sp_used[pos]) continue;
      const unsigned long long kk = sp_key(block_values[pos], block_indices[pos]);
      if (kk > mine) { mine = kk; mpos = pos; }
    }
    const unsigned long long best = sp_block_max(mine, sm_key[ki & 1]);
    if (best != 0ull && mine == best) sp_used[mpos] = 1;   // keys are unique per token index
    if (tid == 0) {
      const float v = best ? sp_key_value(best) : -INFINITY;
      const uint32_t ix = best ? sp_key_index(best) : 0u;
      if (packed_out) { packed_out[ki] = v; packed_out[k + ki] = (float)ix; }
      if (values_out) { values_out[ki] = v; indices_out[ki] = ix; }
    }
    __syncthreads();   // sp_used update visible to the next round's scan
  }
}

// greedy: chunk maxima, then the row maximum (first index on ties); NaN anywhere poisons the row
__global__ void __launch_bounds__(SP_THREADS) top1_stage1_kernel(const float *__restrict__ input, float *__restrict__ block_values,
                                                               uint32_t *__restrict__ block_indices, int ncols, int chunk_size) {
  __shared__ unsigned long long sm_key[SP_THREADS / 32];
  const size_t row = blockIdx.y;
  const int chunk = blockIdx.x, nblocks = gridDim.x, tid = threadIdx.x;
  input += row * (size_t)ncols; block_values += row * (size_t)nblocks; block_indices += row * (size_t)nblocks;
  const int start = chunk * chunk_size, end = min(start + chunk_size, ncols);
  unsigned long long mine = 0ull;
  int has_nan = 0;
  for (int idx = start + tid; idx < end; idx += SP_THREADS) {
    const float v = input[idx];
    if (v != v) has_nan = 1;
    const unsigned long long kk = sp_key(v, (uint32_t)idx);
    mine = kk > mine ? kk : mine;
  }
  const int any_nan = __syncthreads_or(has_nan);
  const unsigned long long best = sp_block_max(mine, sm_key);
  if (tid == 0) {
    block_values[chunk] = any_nan ? CUDART_NAN_F : (best ? sp_key_value(best) : -INFINITY);
    block_indices[chunk] = (any_nan || !best) ? 0u : sp_key_index(best);
  }
}
__global__ void __launch_bounds__(SP_THREADS) top1_stage2_kernel(const float *__restrict__ block_values, const uint32_t *__restrict__ block_indices,
                                                               float *__restrict__ packed_out, uint32_t *__restrict__ token_ids_out, int nblocks) {
  __shared__ unsigned long long sm_key[SP_THREADS / 32];
  const size_t row = blockIdx.x;
  const int tid = threadIdx.x;
  block_values += row * (size_t)nblocks; block_indices += row * (size_t)nblocks;
  unsigned long long mine = 0ull;
  int has_nan = 0;
  for (int pos = tid; pos < nblocks; pos += SP_THREADS) {
    const float v = block_values[pos];
    if (v != v) has_nan = 1;
    const unsigned long long kk = sp_key(v, block_indices[pos]);
    mine = kk > mine ? kk : mine;
  }
  const int any_nan = __syncthreads_or(has_nan);
  const unsigned long long best = sp_block_max(mine, sm_key);
  if (tid == 0) {
    const uint32_t tok = any_nan ? 0xFFFFFFFFu : (best ? sp_key_index(best) : 0u);
    if (packed_out) {
      packed_out[row * 2] = any_nan ? CUDART_NAN_F : (best ? sp_key_value(best) : -INFINITY);
      packed_out[row * 2 + 1] = any_nan ? CUDART_NAN_F : (float)tok;
    }
    if (token_ids_out) token_ids_out[row] = tok;
  }
}

static bool sp_check(int chunk_size, int k, int nblocks, const char *what) {
  if (chunk_size < 1 || chunk_size > SP_THREADS * SP_PER_THREAD || k < 1 || nblocks < 1 || (size_t)nblocks * k > 47 * 1024) {
    fprintf(stderr, "mrs_b200: %s: unsupported shape (chunk_size %d <= %d, k %d, nblocks %d)\n", what, chunk_size,
            SP_THREADS * SP_PER_THREAD, k, nblocks);
    return false;
  }
  return true;
}

}  // namespace mrs

using namespace mrs;

extern "C" void topk_large_f32(const float *input, float *block_values, uint32_t *block_indices, float *block_maxes,
                               float *block_sums, float *values_out, uint32_t *indices_out, float *softmax_info_out,
                               int ncols, int k, int chunk_size, int nblocks, float inv_temperature, int64_t stream) {
  if (!sp_check(chunk_size, k, nblocks, "topk_large_f32")) return;
  cudaStream_t st = (cudaStream_t)stream;
  topk_stage1_kernel<<<dim3(nblocks, 1), SP_THREADS, 0, st>>>(input, block_values, block_indices, block_maxes, block_sums, ncols, k,
                                                             chunk_size, nullptr, inv_temperature);
  topk_stage2_kernel<<<1, SP_THREADS, (size_t)nblocks * k, st>>>(block_values, block_indices, block_maxes, block_sums, values_out,
                                                                indices_out, softmax_info_out, nullptr, nblocks, k);
}
extern "C" void topk_large_f32_packed(const float *input, float *block_values, uint32_t *block_indices, float *block_maxes,
                                      float *block_sums, float *packed_out, int ncols, int k, int chunk_size, int nblocks,
                                      float inv_temperature, int64_t stream) {
  if (!sp_check(chunk_size, k, nblocks, "topk_large_f32_packed")) return;
  cudaStream_t st = (cudaStream_t)stream;
  topk_stage1_kernel<<<dim3(nblocks, 1), SP_THREADS, 0, st>>>(input, block_values, block_indices, block_maxes, block_sums, ncols, k,
                                                             chunk_size, nullptr, inv_temperature);
  topk_stage2_kernel<<<1, SP_THREADS, (size_t)nblocks * k, st>>>(block_values, block_indices, block_maxes, block_sums, nullptr, nullptr,
                                                                nullptr, packed_out, nblocks, k);
}
extern "C" void topk_large_f32_packed_batched(const float *input, const float *inv_temperatures, float *block_values,
                                              uint32_t *block_indices, float *block_maxes, float *block_sums, float *packed_out,
                                              int nrows, int ncols, int k, int chunk_size, int nblocks, int64_t stream) {
  if (nrows < 1 || !sp_check(chunk_size, k, nblocks, "topk_large_f32_packed_batched")) return;
  cudaStream_t st = (cudaStream_t)stream;
  topk_stage1_kernel<<<dim3(nblocks, nrows), SP_THREADS, 0, st>>>(input, block_values, block_indices, block_maxes, block_sums, ncols, k,
                                                                 chunk_size, inv_temperatures, 1.0f);
  topk_stage2_kernel<<<nrows, SP_THREADS, (size_t)nblocks * k, st>>>(block_values, block_indices, block_maxes, block_sums, nullptr,
                                                                    nullptr, nullptr, packed_out, nblocks, k);
}
extern "C" void top1_large_f32_packed(const float *input, float *block_values, uint32_t *block_indices, float *packed_out,
                                      uint32_t *token_ids_out, int ncols, int chunk_size, int nblocks, int64_t stream) {
  if (nblocks < 1 || chunk_size < 1) return;
  cudaStream_t st = (cudaStream_t)stream;
  top1_stage1_kernel<<<dim3(nblocks, 1), SP_THREADS, 0, st>>>(input, block_values, block_indices, ncols, chunk_size);
  top1_stage2_kernel<<<1, SP_THREADS, 0, st>>>(block_values, block_indices, packed_out, token_ids_out, nblocks);
}
extern "C" void top1_large_f32_packed_batched(const float *input, float *block_values, uint32_t *block_indices, float *packed_out,
                                              uint32_t *token_ids_out, int nrows, int ncols, int chunk_size, int nblocks,
                                              int64_t stream) {
  if (nblocks < 1 || chunk_size < 1 || nrows < 1) return;
  cudaStream_t st = (cudaStream_t)stream;
  top1_stage1_kernel<<<dim3(nblocks, nrows), SP_THREADS, 0, st>>>(input, block_values, block_indices, ncols, chunk_size);
  top1_stage2_kernel<<<nrows, SP_THREADS, 0, st>>>(block_values, block_indices, packed_out, token_ids_out, nblocks);
}
