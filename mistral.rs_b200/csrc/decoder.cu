// decoder.cu — the caller of the hot path: a Llama-family decode layer stack over the kernels in
// this library (see include/mrs_b200_model.h), plus the few glue kernels a token step needs
// (quantised embedding gather, argmax, on-device KV index advance).
//
// REF structure being mirrored: mistralrs-core/src/models/llama.rs:68-135 (attention),
// :243-260 (block), :475-… (model forward); embedding gather over ggml blocks:
// mistralrs-quant/src/gguf/mod.rs:815-845; KV indices: see mrs_decode_advance in the header.
#include "common.cuh"
#include "dequant.cuh"
#include "mrs_b200_model.h"

#include <stdio.h>

extern "C" int mrs_mmvq_fused(int ggml_type, int mode, int dt, const void *w0, const void *w1, const void *w2,
                              const void *x, const void *norm_w, float eps, const void *residual, void *dst0,
                              void *dst1, void *dst2, int K, int n0, int n1, int n2, int b_size, int activation,
                              int pdl, void *stream);
extern "C" int mrs_mmvq_fused_qkv_mixed(int type_qk, int type_v, int dt, const void *wq, const void *wk, const void *wv,
                                        const void *x, const void *norm_w, float eps, void *q, void *k, void *v,
                                        int K, int nq, int nk, int nv, int b_size, int pdl, void *stream);
extern "C" void rotary_embedding_positions(void *query, void *key, void *cos_cache, void *sin_cache, void *positions,
                                           int32_t is_neox, int32_t head_size, int64_t num_tokens, int32_t rot_dim,
                                           int32_t seq_len, int32_t num_heads, int32_t num_kv_heads,
                                           int64_t query_stride, int64_t key_stride, uint32_t dtype, int64_t stream);
extern "C" void reshape_and_cache_flashinfer(void *key, void *value, void *key_cache, void *value_cache,
                                             int64_t *slot_mapping, int32_t num_tokens, int32_t num_heads,
                                             int32_t head_size, int32_t block_size, int32_t key_stride,
                                             int32_t value_stride, float k_scale, float v_scale, uint32_t dtype,
                                             uint32_t cache_dtype, cudaStream_t stream);
extern "C" int32_t mrs_paged_decode_fused(void *q, void *k_new, void *v_new, void *key_cache, void *value_cache,
                                          const void *rope_cos, const void *rope_sin, const int32_t *positions,
                                          const int64_t *slot_mapping, const int32_t *kv_indptr,
                                          const int32_t *kv_indices, const int32_t *kv_last_page_len,
                                          const int32_t *request_indices, const int32_t *kv_tile_indices,
                                          const int32_t *o_indptr, const int32_t *kv_chunk_size_ptr,
                                          const uint8_t *block_valid_mask, void *o, void *tmp_v, float *tmp_s,
                                          int32_t *counters, int32_t batch_size, int32_t padded_batch_size,
                                          int32_t num_qo_heads, int32_t num_kv_heads, int32_t head_size,
                                          int32_t page_size, float sm_scale, uint32_t dtype, int32_t pdl,
                                          void *stream);
extern "C" int32_t flashinfer_decode(void *q, void *key_cache, void *value_cache, const int32_t *kv_indptr,
                                     const int32_t *kv_indices, const int32_t *kv_last_page_len,
                                     const int32_t *request_indices, const int32_t *kv_tile_indices,
                                     const int32_t *o_indptr, const int32_t *kv_chunk_size_ptr,
                                     const bool *block_valid_mask, void *o, void *tmp_v, void *tmp_s,
                                     int32_t batch_size, int32_t padded_batch_size, int32_t num_qo_heads,
                                     int32_t num_kv_heads, int32_t head_size, int32_t page_size, int32_t q_stride_n,
                                     int32_t q_stride_h, float sm_scale, int32_t window_left, float logits_soft_cap,
                                     float k_scale, float v_scale, uint32_t dtype, uint32_t cache_dtype,
                                     cudaStream_t stream);

namespace mrs {

__global__ void embedding_gather_kernel(int type, const uint8_t *__restrict__ table, int cols,
                                        const int32_t *__restrict__ ids, void *__restrict__ out, int act_dtype) {
  const int row = ids[blockIdx.x];
  const int be = blk_elems(type), bb = blk_bytes(type);
  const uint8_t *rp = table + (size_t)row * (cols / be) * bb;
  for (int i = threadIdx.x; i < cols; i += blockDim.x)
    store_act(out, (int64_t)blockIdx.x * cols + i, dequant_elem(type, rp + (size_t)(i / be) * bb, i % be), act_dtype);
}

// first-maximum argmax: grid (chunks, rows); each CTA reduces a chunk and folds it into a packed
// 64-bit key (order-preserving float bits << 32 | ~index) with atomicMax; the last CTA of a row
// publishes the index and re-zeroes the scratch (scratch: u64 key[rows] then u32 count[rows]).
__device__ __forceinline__ unsigned long long pack_key(float v, int idx) {
  unsigned int b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (unsigned int)(0xFFFFFFFFu - (unsigned int)idx);
}
__global__ void argmax_kernel(const void *__restrict__ logits, int cols, int act_dtype, int32_t *__restrict__ out,
                              unsigned long long *__restrict__ keys, unsigned int *__restrict__ counts, int pdl) {
  __shared__ unsigned long long sk[32];
  if (pdl) pdl_wait();
  const int row = blockIdx.y;
  const int64_t base = (int64_t)row * cols;
  unsigned long long best = 0ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cols; i += gridDim.x * blockDim.x) {
    const unsigned long long k = pack_key(load_act(logits, base + i, act_dtype), i);
    best = k > best ? k : best;
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, m);
    best = o > best ? o : best;
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sk[w] = best;
  __syncthreads();
  if (w == 0) {
    best = (l < (blockDim.x >> 5)) ? sk[l] : 0ull;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, m);
      best = o > best ? o : best;
    }
    if (l == 0) {
      atomicMax(&keys[row], best);
      __threadfence();
      const unsigned int done = atomicAdd(&counts[row], 1u);
      if (done == gridDim.x - 1) {
        __threadfence();
        const unsigned long long k = atomicExch(&keys[row], 0ull);
        out[row] = (int32_t)(0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFull));
        counts[row] = 0u;
      }
    }
  }
}

// dst = T(dst + res) — the residual add after a row-parallel all-reduce
__global__ void add_residual_kernel(void *__restrict__ dst, const void *__restrict__ res, int64_t n, int dt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) store_act(dst, i, load_act(dst, i, dt) + load_act(res, i, dt), dt);
}

// ---- tensor-parallel sum of the row-parallel partials (REF SumAllReduce::sum_all_reduce,
// mistralrs-quant/src/distributed/mod.rs:436-453: ncclAllReduce(sum) in the activation dtype, then the
// residual add of the decoder layer) as ONE small kernel over NVLink peer memory, in the CUDA graph and
// on the PDL chain — a latency-bound 8-16 KB message does not need a collective library:
//   every rank's row-parallel GEMV wrote its partial [count] (activation dtype) into ITS slot buffer;
//   thread 0 publishes "rank r reached all-reduce #seq" into every peer's flag word r (system-scope
//   release), waits until its own flag words show every peer at #seq (acquire), then all threads pull
//   the peers' partials with 16-byte loads (ld.volatile: the same addresses are reused every second
//   all-reduce), sum them in f32 IN RANK ORDER (every rank computes bit-identical sums), round to the
//   dtype like the collective's output, add the residual, round again.
// Two slot buffers alternate: a rank can only overwrite slot s two all-reduces later, after a barrier
// every peer reaches only once it has finished reading s.
struct ArCtxDev {
  int world, rank;
  const uint8_t *peer_base[8];     // peers' symmetric buffers (own included), mapped into this process
  uint32_t *flags_local;           // [world] flag words in the own buffer
  unsigned long long flags_off, slot_off[2];
  uint32_t *seq;                   // [0] device counter of all-reduces done (graph replays continue it); [1] time-out flag
  unsigned long long ll_off, ll_slot_stride, ll_src_stride;   // low-latency region (ll_off == 0: flags + pull)
};
__global__ void __launch_bounds__(1024) tp_allreduce_residual_kernel(const ArCtxDev c, int slot, const void *__restrict__ residual,
                                                                     void *__restrict__ out, int count, int dt, int pdl) {
  __shared__ uint32_t s_seq;
  if (pdl && threadIdx.x == 0) pdl_launch_dependents();
  if (pdl) pdl_wait();                       // the own partial is complete and flushed
  if (threadIdx.x == 0) {
    const uint32_t seq = *c.seq + 1u;
    *c.seq = seq;
    __threadfence_system();
    for (int r = 0; r < c.world; r++) {
      uint32_t *flag = (uint32_t *)(c.peer_base[r] + c.flags_off) + c.rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
    }
    for (int r = 0; r < c.world; r++) {
      uint32_t v;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(c.flags_local + r) : "memory");
      } while ((int32_t)(v - seq) < 0);
    }
    s_seq = seq;
  }
  __syncthreads();
  for (int i = threadIdx.x * 8; i < count; i += blockDim.x * 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < c.world; r++) {
      const uint4 raw = __ldcv((const uint4 *)(c.peer_base[r] + c.slot_off[slot]) + (i >> 3));
      float v[8];
      unpack_act8(raw, dt, v);
#pragma unroll
      for (int k = 0; k < 8; k++) acc[k] += v[k];
    }
    float res[8];
    load_act8(residual, i, dt, res);
#pragma unroll
    for (int k = 0; k < 8; k++) store_act(out, (int64_t)i + k, round_act(acc[k], dt) + res[k], dt);
  }
}

// Low-latency form (NCCL's "LL" idea on our own buffers): the partial travels as 8-byte words {two activation elements,
// sequence number}.  An 8-byte store is delivered whole, so the receiver needs no flag and no fence — it polls the word
// until the sequence number is the current one.  One NVLink store hop instead of flag round trip + pull round trip.
// Thread t owns element pairs t, t + blockDim, ...: reads its own partial (local), pushes it to every peer, then collects
// the peers' words from its own region, sums in rank order in f32, rounds, adds the residual.  Two slots alternate, the
// sequence number grows monotonically: a word of the all-reduce two steps back can never be mistaken for the current one,
// and a rank can only overwrite slot s after every peer has sent it the all-reduce in between, i.e. finished reading s.
__global__ void __launch_bounds__(1024) tp_allreduce_ll_kernel(const ArCtxDev c, int slot, const void *__restrict__ partial,
                                                               const void *__restrict__ residual, void *__restrict__ out, int count,
                                                               int dt, int pdl) {
  if (pdl && threadIdx.x == 0) pdl_launch_dependents();
  if (pdl) pdl_wait();                       // the own partial is complete and flushed
  const uint32_t seq = *(volatile const uint32_t *)c.seq + 1u;
  const bool dead = *(volatile const uint32_t *)(c.seq + 1) != 0u;   // an earlier all-reduce gave up on a peer: never wait again
  __syncthreads();
  if (threadIdx.x == 0) *(volatile uint32_t *)c.seq = seq;
  const unsigned long long area = c.ll_off + (unsigned long long)slot * c.ll_slot_stride;
  const int npairs = count >> 1;
  for (int i = threadIdx.x; i < npairs; i += blockDim.x) {
    const uint32_t mine = ((const uint32_t *)partial)[i];
    for (int r = 0; r < c.world; r++) {
      if (r == c.rank) continue;
      uint2 *dst = (uint2 *)(c.peer_base[r] + area + (unsigned long long)c.rank * c.ll_src_stride) + i;
      asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(mine), "r"(seq) : "memory");
    }
  }
  for (int i = threadIdx.x; i < npairs; i += blockDim.x) {
    const uint32_t mine = ((const uint32_t *)partial)[i];
    float a0 = 0.f, a1 = 0.f;
    for (int r = 0; r < c.world; r++) {
      uint32_t w = mine;
      if (r != c.rank) {
        const uint2 *src = (const uint2 *)(c.peer_base[c.rank] + area + (unsigned long long)r * c.ll_src_stride) + i;
        // bounded wait: a peer that never shows up (ranks out of step) must not hang the GPU — after ~0.25 s the
        // kernel gives up on the word, raises the flag next to the sequence counter and carries on with stale data
        uint32_t fl, spins = 0;
        const long long t0 = clock64();
        for (;;) {
          asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(w), "=r"(fl) : "l"(src) : "memory");
          if (fl == seq || dead) break;
          if ((++spins & 1023u) == 0u &&
              (clock64() - t0 > 500000000ll || *(volatile const uint32_t *)(c.seq + 1) != 0u)) { c.seq[1] = 1u; break; }
        }
      }
      float v0, v1;
      if (dt == MRS_BF16) { const float2 f = __bfloat1622float2(*(const __nv_bfloat162 *)&w); v0 = f.x; v1 = f.y; }
      else { const float2 f = __half22float2(*(const __half2 *)&w); v0 = f.x; v1 = f.y; }
      a0 += v0; a1 += v1;
    }
    const float r0 = load_act(residual, 2 * (int64_t)i, dt), r1 = load_act(residual, 2 * (int64_t)i + 1, dt);
    store_act(out, 2 * (int64_t)i, round_act(a0, dt) + r0, dt);
    store_act(out, 2 * (int64_t)i + 1, round_act(a1, dt) + r1, dt);
  }
}

// one CTA: integers only, must match the host producers.  Thread b owns sequence b (lengths, slot,
// chunk count), thread 0 turns the per-sequence counts into the two prefix sums, then all threads
// fill the tile list and the page indices.  A sequence that has used up its block table or the RoPE
// table (pos >= min(max_blocks*bs, max_pos)) is frozen: its context does not grow, its KV write is
// skipped (slot -1, the reference's _PAD_SLOT_ID) and *error_flag gets bit 0 — generation past the
// allocated context must never turn into an out-of-bounds cache write.
constexpr int ADV_MAX_BATCH = 256;
__global__ void decode_advance_kernel(const int32_t *__restrict__ block_tables, int max_blocks,
                                      int32_t *__restrict__ context_lens, int batch, int bs, int split_pages,
                                      int padded_tiles, int max_pos, int32_t *positions, int64_t *slot_mapping,
                                      int32_t *kv_indptr, int32_t *kv_indices, int32_t *kv_last_page_len,
                                      int32_t *request_indices, int32_t *kv_tile_indices, int32_t *o_indptr,
                                      int32_t *kv_chunk_size, uint8_t *block_valid_mask, int32_t *error_flag) {
  __shared__ int s_nb[ADV_MAX_BATCH], s_chunks[ADV_MAX_BATCH], s_indptr[ADV_MAX_BATCH + 1], s_oind[ADV_MAX_BATCH + 1];
  const int cap = min(max_blocks * bs, max_pos > 0 ? max_pos : max_blocks * bs);
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    int pos = context_lens[b];             // position of the token being processed now
    const bool full = pos >= cap;
    if (full) { pos = cap - 1; if (error_flag != nullptr) atomicOr(error_flag, 1); }
    const int ctx = pos + 1;               // context length including it
    context_lens[b] = ctx;
    positions[b] = pos;
    slot_mapping[b] = full ? (int64_t)-1 : (int64_t)block_tables[(int64_t)b * max_blocks + pos / bs] * bs + pos % bs;
    const int nb = (ctx + bs - 1) / bs;
    s_nb[b] = nb;
    kv_last_page_len[b] = ctx - (nb - 1) * bs;
    s_chunks[b] = (split_pages > 0) ? ((nb < 1 ? 1 : nb) + split_pages - 1) / split_pages : 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nnz = 0, tiles = 0;
    s_indptr[0] = 0; s_oind[0] = 0;
    for (int b = 0; b < batch; b++) {
      nnz += s_nb[b];
      tiles = min(tiles + s_chunks[b], padded_tiles);
      s_indptr[b + 1] = nnz; s_oind[b + 1] = tiles;
    }
    kv_chunk_size[0] = (split_pages > 0 ? split_pages : 1) * bs;
  }
  __syncthreads();
  const int tiles = s_oind[batch];
  for (int b = threadIdx.x; b <= batch; b += blockDim.x) { kv_indptr[b] = s_indptr[b]; o_indptr[b] = s_oind[b]; }
  for (int b = 0; b < batch; b++) {
    const int t0 = s_oind[b], nt = s_oind[b + 1] - t0;
    for (int t = threadIdx.x; t < nt; t += blockDim.x) { request_indices[t0 + t] = b; kv_tile_indices[t0 + t] = t; }
    // page indices: parallel over (b, i)
    const int p0 = s_indptr[b], nb = s_indptr[b + 1] - p0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) kv_indices[p0 + i] = block_tables[(int64_t)b * max_blocks + i];
  }
  for (int t = threadIdx.x; t < padded_tiles; t += blockDim.x) {
    block_valid_mask[t] = t < tiles ? 1 : 0;
    if (t >= tiles) { request_indices[t] = 0; kv_tile_indices[t] = 0; }
  }
}

}  // namespace mrs

using namespace mrs;

extern "C" int32_t mrs_embedding_gather(int32_t ggml_type, const void *table, int32_t cols, const int32_t *ids,
                                        int32_t n, void *out, int32_t act_dtype, void *stream) {
  if (n <= 0) return 0;
  if (blk_bytes(ggml_type) == 0 || cols % blk_elems(ggml_type)) return (int32_t)cudaErrorInvalidValue;
  embedding_gather_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(ggml_type, (const uint8_t *)table, cols, ids, out, act_dtype);
  return (int32_t)cudaGetLastError();
}

extern "C" int32_t mrs_argmax(const void *logits, int32_t rows, int32_t cols, int32_t act_dtype, int32_t *out,
                              void *scratch, int32_t pdl, void *stream) {
  if (rows <= 0) return 0;
  if (scratch == nullptr) return (int32_t)cudaErrorInvalidValue;
  unsigned long long *keys = (unsigned long long *)scratch;
  unsigned int *counts = (unsigned int *)(keys + rows);
  int chunks = (cols + 4095) / 4096;
  if (chunks > 64) chunks = 64;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(chunks, rows); cfg.blockDim = dim3(256); cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return (int32_t)cudaLaunchKernelEx(&cfg, argmax_kernel, logits, (int)cols, (int)act_dtype, out, keys, counts, (int)pdl);
}

extern "C" int32_t mrs_decode_advance(const int32_t *block_tables, int32_t max_blocks_per_seq, int32_t *context_lens,
                                      int32_t batch, int32_t block_size, int32_t split_pages, int32_t padded_tiles,
                                      int32_t *positions, int64_t *slot_mapping, int32_t *kv_indptr,
                                      int32_t *kv_indices, int32_t *kv_last_page_len, int32_t *request_indices,
                                      int32_t *kv_tile_indices, int32_t *o_indptr, int32_t *kv_chunk_size,
                                      uint8_t *block_valid_mask, int32_t max_pos, int32_t *error_flag, void *stream) {
  if (batch < 1 || batch > ADV_MAX_BATCH || max_blocks_per_seq < 1 || block_size < 1) return (int32_t)cudaErrorInvalidValue;
  decode_advance_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(block_tables, max_blocks_per_seq, context_lens, batch,
                                                             block_size, split_pages, padded_tiles, max_pos, positions,
                                                             slot_mapping, kv_indptr, kv_indices, kv_last_page_len,
                                                             request_indices, kv_tile_indices, o_indptr, kv_chunk_size,
                                                             block_valid_mask, error_flag);
  return (int32_t)cudaGetLastError();
}

extern "C" int32_t mrs_tp_allreduce_residual(const mrs_tp_ctx *ctx, int32_t slot, const void *residual, void *out,
                                             int32_t count, int32_t dtype, int32_t pdl, void *stream) {
  if (ctx == nullptr || ctx->world < 1 || ctx->world > 8 || (slot & ~1) || count % 8 || (dtype != MRS_F16 && dtype != MRS_BF16))
    return (int32_t)cudaErrorInvalidValue;
  ArCtxDev c = {};
  c.world = ctx->world; c.rank = ctx->rank;
  for (int r = 0; r < ctx->world; r++) c.peer_base[r] = (const uint8_t *)ctx->peer_base[r];
  c.flags_off = (unsigned long long)ctx->flags_offset;
  c.slot_off[0] = (unsigned long long)ctx->slot_offset[0]; c.slot_off[1] = (unsigned long long)ctx->slot_offset[1];
  c.flags_local = (uint32_t *)((uint8_t *)ctx->peer_base[ctx->rank] + ctx->flags_offset);
  c.seq = (uint32_t *)ctx->seq_counter;
  c.ll_off = (unsigned long long)ctx->ll_offset; c.ll_slot_stride = (unsigned long long)ctx->ll_slot_stride;
  c.ll_src_stride = (unsigned long long)ctx->ll_src_stride;
  if (ctx->ll_offset != 0) {
    if ((int64_t)count * 4 > ctx->ll_src_stride) return (int32_t)cudaErrorInvalidValue;
    int th = (count / 2 + 31) / 32 * 32;
    if (th > 1024) th = 1024;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(th); cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    // the partial of this rank: its slot in the symmetric buffer (what the row-parallel GEMV wrote)
    const void *partial = (const uint8_t *)ctx->peer_base[ctx->rank] + ctx->slot_offset[slot];
    return (int32_t)cudaLaunchKernelEx(&cfg, tp_allreduce_ll_kernel, c, (int)slot, partial, residual, out, (int)count, (int)dtype, (int)pdl);
  }
  int threads = (count / 8 + 31) / 32 * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 32) threads = 32;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1); cfg.blockDim = dim3(threads); cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return (int32_t)cudaLaunchKernelEx(&cfg, tp_allreduce_residual_kernel, c, (int)slot, residual, out, (int)count, (int)dtype, (int)pdl);
}

#define MRS_TRY(expr)                                  \
  do {                                                 \
    const int _e = (int)(expr);                        \
    if (_e != 0) {                                     \
      fprintf(stderr, "mrs_b200: %s -> cudaError %d\n", #expr, _e); \
      return _e;                                       \
    }                                                  \
  } while (0)

extern "C" int32_t mrs_llama_decode_step(const mrs_llama_step *s, void *stream) {
  const int dt = s->act_dtype, B = s->batch, H = s->hidden, pdl = s->pdl;
  const int nq = s->n_heads * s->head_dim, nkv = s->n_kv_heads * s->head_dim;
  cudaStream_t st = (cudaStream_t)stream;
  if (B < 1 || B > 8) return (int32_t)cudaErrorInvalidValue;
  const bool do_attn = !(s->skip_mask & 1), do_gemv = !(s->skip_mask & 2);

  MRS_TRY(mrs_embedding_gather(s->tok_embd.ggml_type, s->tok_embd.data, H, s->token_ids, B, s->x, dt, stream));
  void *hidden = s->x, *hidden2 = s->x2;
  for (int l = 0; l < s->n_layers; l++) {
    const mrs_llama_layer &L = s->layers[l];
    // --- attention block: x = x + o_proj(attn(rope(qkv(norm(x)))))
    // the fused launches decode every matrix of a group with ONE ggml type: per-tensor type
    // overrides (GGUF --tensor-type, per-layer UQFF topologies) take separate launches
    if (!do_gemv) {
    } else if (L.wq.ggml_type != L.wk.ggml_type) {
      MRS_TRY(mrs_mmvq_fused(L.wq.ggml_type, 0, dt, L.wq.data, nullptr, nullptr, hidden, L.attn_norm, s->rms_eps,
                             nullptr, s->q, nullptr, nullptr, H, nq, 0, 0, B, 0, pdl, stream));
      MRS_TRY(mrs_mmvq_fused(L.wk.ggml_type, 0, dt, L.wk.data, nullptr, nullptr, hidden, L.attn_norm, s->rms_eps,
                             nullptr, s->k, nullptr, nullptr, H, nkv, 0, 0, B, 0, pdl, stream));
      MRS_TRY(mrs_mmvq_fused(L.wv.ggml_type, 0, dt, L.wv.data, nullptr, nullptr, hidden, L.attn_norm, s->rms_eps,
                             nullptr, s->v, nullptr, nullptr, H, nkv, 0, 0, B, 0, pdl, stream));
    } else if (L.wk.ggml_type == L.wv.ggml_type) {
      MRS_TRY(mrs_mmvq_fused(L.wq.ggml_type, 2, dt, L.wq.data, L.wk.data, L.wv.data, hidden, L.attn_norm, s->rms_eps,
                             nullptr, s->q, s->k, s->v, H, nq, nkv, nkv, B, 0, pdl, stream));
    } else {
      // Q4_K_M keeps attn_v in Q6_K on some layers: q∥k fused, v on its own.  (The one-grid form,
      // mrs_mmvq_fused_qkv_mixed, measured 0.7 % slower here: under PDL the small v launch already
      // hides behind the q∥k tail, and q∥k loses the CTAs it hands to v.)
      MRS_TRY(mrs_mmvq_fused(L.wq.ggml_type, 2, dt, L.wq.data, L.wk.data, nullptr, hidden, L.attn_norm, s->rms_eps,
                             nullptr, s->q, s->k, nullptr, H, nq, nkv, 0, B, 0, pdl, stream));
      MRS_TRY(mrs_mmvq_fused(L.wv.ggml_type, 0, dt, L.wv.data, nullptr, nullptr, hidden, L.attn_norm, s->rms_eps,
                             nullptr, s->v, nullptr, nullptr, H, nkv, 0, 0, B, 0, pdl, stream));
    }
    if (do_attn && s->fused_attention) {
      MRS_TRY(mrs_paged_decode_fused(s->q, s->k, s->v, L.k_cache, L.v_cache, s->rope_cos, s->rope_sin, s->positions,
                                     s->slot_mapping, s->kv_indptr, s->kv_indices, s->kv_last_page_len,
                                     s->request_indices, s->kv_tile_indices, s->o_indptr, s->kv_chunk_size,
                                     s->block_valid_mask, s->attn_out, s->padded_tiles > B ? s->tmp_v : nullptr,
                                     s->padded_tiles > B ? s->tmp_s : nullptr, s->attn_counters, B, s->padded_tiles,
                                     s->n_heads, s->n_kv_heads, s->head_dim, s->block_size, s->sm_scale, (uint32_t)dt,
                                     pdl | (s->rope_neox ? 0 : 2), stream));
    } else if (do_attn) {
    rotary_embedding_positions(s->q, s->k, (void *)s->rope_cos, (void *)s->rope_sin, s->positions, s->rope_neox,
                               s->head_dim, B, s->head_dim / 2, 0, s->n_heads, s->n_kv_heads, nq, nkv, (uint32_t)dt,
                               (int64_t)stream);
    reshape_and_cache_flashinfer(s->k, s->v, L.k_cache, L.v_cache, s->slot_mapping, B, s->n_kv_heads, s->head_dim,
                                 s->block_size, nkv, nkv, 1.f, 1.f, (uint32_t)dt, (uint32_t)dt, st);
    MRS_TRY(flashinfer_decode(s->q, L.k_cache, L.v_cache, s->kv_indptr, s->kv_indices, s->kv_last_page_len,
                              s->request_indices, s->kv_tile_indices, s->o_indptr, s->kv_chunk_size,
                              (const bool *)s->block_valid_mask, s->attn_out, s->padded_tiles > B ? s->tmp_v : nullptr,
                              s->padded_tiles > B ? s->tmp_s : nullptr, B, s->padded_tiles, s->n_heads, s->n_kv_heads,
                              s->head_dim, s->block_size, nq, s->head_dim, s->sm_scale, -1, 0.f, 1.f, 1.f,
                              (uint32_t)dt, (uint32_t)dt, st));
    }
    if (!do_gemv) continue;
    if (s->tp != nullptr && s->tp->world > 1) {
      // row-parallel partial -> this rank's slot buffer -> peer-memory sum + residual (one kernel, on the PDL chain)
      void *part = (uint8_t *)s->tp->peer_base[s->tp->rank] + s->tp->slot_offset[0];
      MRS_TRY(mrs_mmvq_fused(L.wo.ggml_type, 0, dt, L.wo.data, nullptr, nullptr, s->attn_out, nullptr, 0.f, nullptr,
                             part, nullptr, nullptr, nq, H, 0, 0, B, 0, pdl, stream));
      MRS_TRY(mrs_tp_allreduce_residual(s->tp, 0, hidden, hidden2, B * H, dt, pdl, stream));
    } else if (s->all_reduce == nullptr) {
      MRS_TRY(mrs_mmvq_fused(L.wo.ggml_type, 0, dt, L.wo.data, nullptr, nullptr, s->attn_out, nullptr, 0.f, hidden,
                             hidden2, nullptr, nullptr, nq, H, 0, 0, B, 0, pdl, stream));
    } else {  // row-parallel: partial sums -> all-reduce -> residual add (REF distributed/layers.rs:965-975)
      MRS_TRY(mrs_mmvq_fused(L.wo.ggml_type, 0, dt, L.wo.data, nullptr, nullptr, s->attn_out, nullptr, 0.f, nullptr,
                             hidden2, nullptr, nullptr, nq, H, 0, 0, B, 0, pdl, stream));
      s->all_reduce(hidden2, (int64_t)B * H, dt, stream, s->all_reduce_user);
      add_residual_kernel<<<(unsigned)(((int64_t)B * H + 255) / 256), 256, 0, st>>>(hidden2, hidden, (int64_t)B * H, dt);
    }
    // --- MLP block: x = x + down(silu(gate(norm(x))) * up(norm(x)))
    if (L.w_gate.ggml_type != L.w_up.ggml_type || L.w_gate.rows != L.w_up.rows) return (int32_t)cudaErrorInvalidValue;
    MRS_TRY(mrs_mmvq_fused(L.w_gate.ggml_type, 1, dt, L.w_gate.data, L.w_up.data, nullptr, hidden2, L.ffn_norm,
                           s->rms_eps, nullptr, s->act, nullptr, nullptr, H, L.w_gate.rows, L.w_gate.rows, 0, B, 0,
                           pdl, stream));
    if (s->tp != nullptr && s->tp->world > 1) {
      void *part = (uint8_t *)s->tp->peer_base[s->tp->rank] + s->tp->slot_offset[1];
      MRS_TRY(mrs_mmvq_fused(L.w_down.ggml_type, 0, dt, L.w_down.data, nullptr, nullptr, s->act, nullptr, 0.f, nullptr,
                             part, nullptr, nullptr, L.w_down.cols, H, 0, 0, B, 0, pdl, stream));
      MRS_TRY(mrs_tp_allreduce_residual(s->tp, 1, hidden2, hidden, B * H, dt, pdl, stream));
    } else if (s->all_reduce == nullptr) {
      MRS_TRY(mrs_mmvq_fused(L.w_down.ggml_type, 0, dt, L.w_down.data, nullptr, nullptr, s->act, nullptr, 0.f, hidden2,
                             hidden, nullptr, nullptr, L.w_down.cols, H, 0, 0, B, 0, pdl, stream));
    } else {
      MRS_TRY(mrs_mmvq_fused(L.w_down.ggml_type, 0, dt, L.w_down.data, nullptr, nullptr, s->act, nullptr, 0.f, nullptr,
                             hidden, nullptr, nullptr, L.w_down.cols, H, 0, 0, B, 0, pdl, stream));
      s->all_reduce(hidden, (int64_t)B * H, dt, stream, s->all_reduce_user);
      add_residual_kernel<<<(unsigned)(((int64_t)B * H + 255) / 256), 256, 0, st>>>(hidden, hidden2, (int64_t)B * H, dt);
    }
  }
  if (do_gemv)
  MRS_TRY(mrs_mmvq_fused(s->lm_head.ggml_type, 0, dt, s->lm_head.data, nullptr, nullptr, hidden, s->final_norm,
                         s->rms_eps, nullptr, s->logits, nullptr, nullptr, H, s->vocab, 0, 0, B, 0, pdl, stream));
  MRS_TRY(mrs_argmax(s->logits, B, s->vocab, dt, s->out_token, s->argmax_scratch, pdl, stream));
  return 0;
}
