// mmq_tc.cu — prefill GEMM over ggml quant blocks on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Replaces the reference's batch>8 path `fast_mmq::plain` (REF mistralrs-quant/src/gguf/
// fast_mmq.rs:762-826 -> launch_mmq_quantize_q8_1_* + launch_mmq_gguf_<q>, int8 `mma.sync`
// m16n8k32 with Q8_1 activations; kernels/mmq_gguf/mmq_gguf.cuh) and the opt-in Marlin repack
// path (gguf/packed_affine.rs:604): Y[M,N] = X[M,K] . W[N,K]^T with W in raw ggml blocks.
//
// Design (per CTA, one 256-token x 256-row output tile, 320 threads, 1 CTA/SM):
//   warp 0      TMA producer: activation tiles X[128 x 64] (x2 token sub-tiles) through a
//               2-D tensor map (SWIZZLE_128B) into the stage ring, mbarrier complete_tx.
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//               (M=128, N=256, K=16) from shared-memory descriptors; accumulators live in
//               TMEM (2 x 256 f32 columns = all 512 columns); tcgen05.commit frees the stage.
//   warps 2..9  dequantisers: each thread owns one weight row of the tile, reads its ggml
//               blocks straight from global/L2 (each weight byte is touched once per 256-token
//               tile), expands 64 weights per K-step to f16 and writes them K-major into the
//               128-byte-swizzled B stage; fence.proxy.async + mbarrier hand-off to the MMA
//               warp.  After the K loop the same warps are the epilogue: tcgen05.ld -> bf16/f16
//               -> global.
// Numerics: every weight is dequantised with the reference's f32 formula and rounded ONCE to the activations' 16-bit
// format (f16 with f16 activations, bf16 with bf16 ones: the MMA takes one format for both operands); activations are
// used as they are — no activation quantisation — f32 accumulation in TMEM.  Closer to the exact product than the
// reference's int8-activation MMQ.
// (Round 2: mrs_mmq_gguf sends Q8_0 / Q4_K / Q6_K launches with K % 256 == 0 to csrc/mmq_ts.cu; this kernel keeps the
// other seven types, odd shapes and the GPTQ/AWQ checkpoint-layout GEMM, and is required to agree with it bit for bit.)
#include "affine.cuh"
#include "dequant.cuh"
#include "tc_common.cuh"

#include <stdio.h>

namespace mrs {

constexpr int TC_BM = 128;        // UMMA M (tokens per sub-tile)
constexpr int TC_MT = 2;          // token sub-tiles per CTA
constexpr int TC_BN = 256;        // UMMA N (weight rows per CTA)
constexpr int TC_BK = 64;         // K per stage (128 bytes of 16-bit -> one swizzle atom row)
constexpr int TC_STAGES = 3;
constexpr int TC_DQ_WARPS = 16;                       // dequantiser warps: two threads per weight row
constexpr int TC_THREADS = 64 + TC_DQ_WARPS * 32;
constexpr int A_STAGE_BYTES = TC_MT * TC_BM * TC_BK * 2;   // 32 KB
constexpr int B_STAGE_BYTES = TC_BN * TC_BK * 2;           // 32 KB
constexpr int TC_SMEM = 1024 + TC_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256;

// ---- exact block decoders: 64 consecutive weights of one row -> f16 ------------------------
// (formulas identical to oracle/mrs_oracle.c unpack_block; REF layouts mmvq_gguf.cu:134-225)
__device__ __forceinline__ float h2f_at(const uint8_t *p) {
  return __half2float(__ushort_as_half(*(const unsigned short *)p));
}
__device__ __forceinline__ void k4_scale_min(int j, const uint8_t *q, int &sc, int &m) {
  if (j < 4) { sc = q[j] & 63; m = q[j + 4] & 63; }
  else { sc = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}

// out: 64 floats for elements [k0, k0+64) of the row starting at `row` (k0 multiple of 64)
template <int TYPE>
__device__ __forceinline__ void dequant64(const uint8_t *row, int k0, float *out) {
  if constexpr (TYPE == MRS_Q8_0) {
    const uint8_t *b = row + (size_t)(k0 / 32) * 34;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float d = h2f_at(b + 34 * h);
      const uint16_t *q16 = (const uint16_t *)(b + 34 * h + 2);
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const uint16_t v = q16[i];
        out[32 * h + 2 * i] = d * (float)(int8_t)(v & 0xFF);
        out[32 * h + 2 * i + 1] = d * (float)(int8_t)(v >> 8);
      }
    }
  } else if constexpr (TYPE == MRS_Q4_K) {
    const uint8_t *b = row + (size_t)(k0 / 256) * 144;
    const int j = (k0 % 256) / 64;
    const float d = h2f_at(b), dmin = h2f_at(b + 2);
    int sc0, m0, sc1, m1;
    k4_scale_min(2 * j, b + 4, sc0, m0);
    k4_scale_min(2 * j + 1, b + 4, sc1, m1);
    const float d0 = d * (float)sc0, d1 = d * (float)sc1, o0 = dmin * (float)m0, o1 = dmin * (float)m1;
    const uint4 *qs = (const uint4 *)(b + 16 + 32 * j);
    const uint4 qa = qs[0], qb = qs[1];
    const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t byte = (w[i] >> (8 * k)) & 0xFF;
        out[4 * i + k] = d0 * (float)(byte & 0xF) - o0;
        out[32 + 4 * i + k] = d1 * (float)(byte >> 4) - o1;
      }
  } else if constexpr (TYPE == MRS_Q6_K) {
    const uint8_t *b = row + (size_t)(k0 / 256) * 210;
    const int n = (k0 % 256) / 128, hi = ((k0 % 128) / 64);  // hi: elements 64..127 of the half use the high nibbles
    const uint16_t *ql = (const uint16_t *)(b + 64 * n);
    const uint16_t *qh = (const uint16_t *)(b + 128 + 32 * n);
    const int8_t *sc = (const int8_t *)(b + 192 + 8 * n + 4 * hi);
    const float d = h2f_at(b + 208);
#pragma unroll
    for (int i = 0; i < 16; i++) {  // l = 2i, 2i+1
      const uint32_t la = ql[i], lb = ql[16 + i], h = qh[i];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const int l = 2 * i + k;
        const uint32_t qa8 = (la >> (8 * k)) & 0xFF, qb8 = (lb >> (8 * k)) & 0xFF, h8 = (h >> (8 * k)) & 0xFF;
        const int q1 = (int)((hi ? (qa8 >> 4) : (qa8 & 0xF)) | (((h8 >> (4 * hi)) & 3) << 4)) - 32;
        const int q2 = (int)((hi ? (qb8 >> 4) : (qb8 & 0xF)) | (((h8 >> (4 * hi + 2)) & 3) << 4)) - 32;
        out[l] = d * (float)sc[l / 16] * (float)q1;
        out[32 + l] = d * (float)sc[2 + l / 16] * (float)q2;
      }
    }
  } else {
    // other ggml types: correct but unoptimised per-element path
    const int be = (TYPE >= MRS_Q2_K) ? 256 : 32;
    constexpr int bb = (TYPE == MRS_Q4_0) ? 18 : (TYPE == MRS_Q4_1) ? 20 : (TYPE == MRS_Q5_0) ? 22 : (TYPE == MRS_Q5_1) ? 24
                     : (TYPE == MRS_Q2_K) ? 84 : (TYPE == MRS_Q3_K) ? 110 : (TYPE == MRS_Q5_K) ? 176 : 0;
    for (int i = 0; i < 64; i++) {
      const int e = k0 + i;
      out[i] = dequant_elem(TYPE, row + (size_t)(e / be) * bb, e % be);
    }
  }
}


// ---- register-staged decoding of 32 consecutive weights (two threads share a row's K-step) ----
// load_raw32 only issues loads (so the next K-step's bytes are in flight while the current one
// is expanded); expand32 turns them into 32 floats.
__device__ __forceinline__ void ld_unaligned_words9(const uint8_t *a, uint32_t *w, uint32_t &phase) {
  const uintptr_t u = (uintptr_t)a;
  const uint32_t *a0 = (const uint32_t *)(u & ~(uintptr_t)3);
  phase = (uint32_t)(u & 3) * 8;
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = a0[i];
  // a 34-byte Q8_0 block that starts on a word boundary ends in the middle of word 8: fetch only
  // its two valid bytes, so the LAST block of a tensor is never read past
  w[8] = (phase != 0) ? a0[8] : (uint32_t)*(const uint16_t *)(a0 + 8);
}
// byte stream starting at the unaligned address: word i of the stream
__device__ __forceinline__ uint32_t stream_word(const uint32_t *w, int i, uint32_t phase) {
  return __funnelshift_r(w[i], (i + 1 < 9) ? w[i + 1] : 0u, phase);
}

template <int TYPE> struct Raw32 { uint32_t w[1]; };
template <> struct Raw32<MRS_Q4_K> { uint4 hdr, q0, q1; };
template <> struct Raw32<MRS_Q8_0> { uint32_t w[9]; uint32_t phase; };
template <> struct Raw32<MRS_Q6_K> { uint32_t ql[9], qh[9]; uint32_t pl, ph; uint32_t sc; uint32_t d; };

template <int TYPE>
__device__ __forceinline__ void load_raw32(const uint8_t *row, int k, Raw32<TYPE> &r) {
  if constexpr (TYPE == MRS_Q4_K) {
    const uint8_t *b = row + (size_t)(k / 256) * 144;
    const int j = ((k % 256) / 32) >> 1;
    r.hdr = *(const uint4 *)b;
    r.q0 = *(const uint4 *)(b + 16 + 32 * j);
    r.q1 = *(const uint4 *)(b + 32 + 32 * j);
  } else if constexpr (TYPE == MRS_Q8_0) {
    ld_unaligned_words9(row + (size_t)(k / 32) * 34, r.w, r.phase);
  } else if constexpr (TYPE == MRS_Q6_K) {
    const uint8_t *b = row + (size_t)(k / 256) * 210;
    const int n = (k % 256) / 128, kk = (k % 128) / 32;
    ld_unaligned_words9(b + 64 * n + 32 * (kk & 1), r.ql, r.pl);
    ld_unaligned_words9(b + 128 + 32 * n, r.qh, r.ph);
    r.sc = *(const uint16_t *)(b + 192 + 8 * n + 2 * kk);
    r.d = *(const uint16_t *)(b + 208);
  }
}

template <int TYPE>
__device__ __forceinline__ void expand32(const Raw32<TYPE> &r, const uint8_t *row, int k, float *out) {
  if constexpr (TYPE == MRS_Q4_K) {
    const int sub = (k % 256) / 32, hi = sub & 1;
    const float d = __half2float(__ushort_as_half((unsigned short)(r.hdr.x & 0xFFFF)));
    const float dmin = __half2float(__ushort_as_half((unsigned short)(r.hdr.x >> 16)));
    // 6-bit scale / min of sub-block `sub` from the 12 scale bytes held in hdr.y/z/w (registers only)
    auto qb = [&](int i) -> int {
      const uint32_t wsel = (i < 4) ? r.hdr.y : ((i < 8) ? r.hdr.z : r.hdr.w);
      return (int)((wsel >> (8 * (i & 3))) & 0xFF);
    };
    int sc, m;
    if (sub < 4) { sc = qb(sub) & 63; m = qb(sub + 4) & 63; }
    else { sc = (qb(sub + 4) & 0xF) | ((qb(sub - 4) >> 6) << 4); m = (qb(sub + 4) >> 4) | ((qb(sub) >> 6) << 4); }
    const float ds = d * (float)sc, om = dmin * (float)m;
    const uint32_t w[8] = {r.q0.x, r.q0.y, r.q0.z, r.q0.w, r.q1.x, r.q1.y, r.q1.z, r.q1.w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t v = hi ? (w[i] >> 4) : w[i];
#pragma unroll
      for (int b = 0; b < 4; b++) out[4 * i + b] = ds * (float)((v >> (8 * b)) & 0xF) - om;
    }
  } else if constexpr (TYPE == MRS_Q8_0) {
    const uint32_t s0 = stream_word(r.w, 0, r.phase);
    const float d = __half2float(__ushort_as_half((unsigned short)(s0 & 0xFFFF)));
#pragma unroll
    for (int i = 0; i < 8; i++) {
      // qs word i = stream bytes 2+4i .. 5+4i = high half of stream word i, low half of word i+1
      const uint32_t lo = stream_word(r.w, i, r.phase), hi2 = stream_word(r.w, i + 1, r.phase);
      const uint32_t q = __funnelshift_r(lo, hi2, 16);
#pragma unroll
      for (int b = 0; b < 4; b++) out[4 * i + b] = d * (float)(int8_t)((q >> (8 * b)) & 0xFF);
    }
  } else if constexpr (TYPE == MRS_Q6_K) {
    const int kk = (k % 128) / 32, hs = kk >> 1;   // hs: 1 -> high nibbles, qh bits shift by 4
    const float d = __half2float(__ushort_as_half((unsigned short)r.d));
    const float d0 = d * (float)(int8_t)(r.sc & 0xFF), d1 = d * (float)(int8_t)(r.sc >> 8);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t l = stream_word(r.ql, i, r.pl), h = stream_word(r.qh, i, r.ph);
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const uint32_t lb = (l >> (8 * b)) & 0xFF, hb = (h >> (8 * b)) & 0xFF;
        const int q = (int)((hs ? (lb >> 4) : (lb & 0xF)) | (((hb >> (2 * kk)) & 3) << 4)) - 32;
        out[4 * i + b] = ((4 * i + b) < 16 ? d0 : d1) * (float)q;
      }
    }
  } else {
    // other ggml types: exact per-element decode (correct, not tuned)
    const int be = (TYPE >= MRS_Q2_K) ? 256 : 32;
    constexpr int bb = (TYPE == MRS_Q4_0) ? 18 : (TYPE == MRS_Q4_1) ? 20 : (TYPE == MRS_Q5_0) ? 22 : (TYPE == MRS_Q5_1) ? 24
                     : (TYPE == MRS_Q2_K) ? 84 : (TYPE == MRS_Q3_K) ? 110 : (TYPE == MRS_Q5_K) ? 176 : 1;
    for (int i = 0; i < 32; i++) {
      const int e = k + i;
      out[i] = dequant_elem(TYPE, row + (size_t)(e / be) * bb, e % be);
    }
  }
}

struct TcParams;
// "checkpoint" types: weights are not ggml blocks addressed by row_bytes but separate arrays read through dequant32_ckpt
template <int TYPE> struct IsInt4Ckpt { static constexpr bool value = (TYPE >= 100 && TYPE <= 103); };

struct TcParams {
  const uint8_t *w;
  void *y;
  int M, N, K, row_bytes, out_dtype;
  int b_fmt;  // operand format of the dequantised weights: 0 f16, 1 bf16
  // GPTQ / AWQ int4 checkpoints (TYPE_GPTQ4 / TYPE_AWQ4): raw HF tensors, no Marlin repack
  const int32_t *qweight;   // GPTQ [K/8, N] (nibbles along K); AWQ [K, N/8] (nibbles along N, order 0,2,4,6,1,3,5,7)
  const __half *scales;     // [K/group, N]
  const int32_t *qzeros;    // AWQ [K/group, N/8]; GPTQ: ignored (symmetric, w = (q-8)*s — REF marlin kU4B8)
  const int32_t *g_idx;     // GPTQ act-order group of each k, or nullptr (k / group)
  int group;
  // packed-affine ggml weights (TYPE_AFF4 / TYPE_AFF8; layout in affine.cuh): w = scale * q - offset, group 16 or 32
  const uint8_t *aff_payload;
  const uint16_t *aff_scales, *aff_offsets;
  int aff_bf16;             // 16-bit format of scales / offsets
};
constexpr int TYPE_GPTQ4 = 100, TYPE_AWQ4 = 101, TYPE_AFF4 = 102, TYPE_AFF8 = 103;

// 64 weights k0..k0+63 of output channel n from a GPTQ / AWQ int4 checkpoint, as the Marlin
// path of the reference computes them: w = f16((q - 8) * s) (GPTQ, REF marlin_matmul_f16.cu /
// marlin_kernel.cuh dequant kU4B8) or f16((q - z) * s) (AWQ, kU4 + zero points).
template <int TYPE>
__device__ __forceinline__ void dequant64_ckpt(const TcParams &p, int n, int k0, float *out) {
  if constexpr (TYPE == TYPE_GPTQ4) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t w = (uint32_t)p.qweight[(size_t)(k0 / 8 + i) * p.N + n];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int k = k0 + 8 * i + j;
        const int g = p.g_idx ? p.g_idx[k] : k / p.group;
        const float sc = __half2float(p.scales[(size_t)g * p.N + n]);
        out[8 * i + j] = (float)((int)((w >> (4 * j)) & 0xF) - 8) * sc;
      }
    }
  } else {
    const int sh = 4 * ((n & 7) == 0 ? 0 : (n & 7) == 1 ? 4 : (n & 7) == 2 ? 1 : (n & 7) == 3 ? 5 : (n & 7) == 4 ? 2 : (n & 7) == 5 ? 6 : (n & 7) == 6 ? 3 : 7);
#pragma unroll 8
    for (int i = 0; i < 64; i++) {
      const int k = k0 + i, g = k / p.group;
      const int q = (int)(((uint32_t)p.qweight[(size_t)k * (p.N / 8) + n / 8] >> sh) & 0xF);
      const int z = (int)(((uint32_t)p.qzeros[(size_t)g * (p.N / 8) + n / 8] >> sh) & 0xF);
      out[i] = (float)(q - z) * __half2float(p.scales[(size_t)g * p.N + n]);
    }
  }
}

template <int TYPE>
__device__ __forceinline__ void dequant32_ckpt(const TcParams &p, int n, int k0, float *out) {
  if constexpr (TYPE == TYPE_AFF4 || TYPE == TYPE_AFF8) {
    affine::dequant32(p.aff_payload, p.aff_scales, p.aff_offsets, TYPE == TYPE_AFF4 ? 4 : 8, p.group, p.aff_bf16 != 0, p.K, n, k0, out);
  } else if constexpr (TYPE == TYPE_GPTQ4) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t w = (uint32_t)p.qweight[(size_t)(k0 / 8 + i) * p.N + n];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int k = k0 + 8 * i + j;
        const int g = p.g_idx ? p.g_idx[k] : k / p.group;
        out[8 * i + j] = (float)((int)((w >> (4 * j)) & 0xF) - 8) * __half2float(p.scales[(size_t)g * p.N + n]);
      }
    }
  } else {
    const int sh = 4 * ((n & 7) == 0 ? 0 : (n & 7) == 1 ? 4 : (n & 7) == 2 ? 1 : (n & 7) == 3 ? 5 : (n & 7) == 4 ? 2 : (n & 7) == 5 ? 6 : (n & 7) == 6 ? 3 : 7);
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
      const int k = k0 + i, g = k / p.group;
      const int q = (int)(((uint32_t)p.qweight[(size_t)k * (p.N / 8) + n / 8] >> sh) & 0xF);
      const int z = (int)(((uint32_t)p.qzeros[(size_t)g * (p.N / 8) + n / 8] >> sh) & 0xF);
      out[i] = (float)(q - z) * __half2float(p.scales[(size_t)g * p.N + n]);
    }
  }
}

template <int TYPE>
__global__ void __launch_bounds__(TC_THREADS, 1)
mmq_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned carve-up (SWIZZLE_128B atoms)
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *a_st = smem;                                   // [STAGES][2][128 rows][128 B]
  uint8_t *b_st = smem + TC_STAGES * A_STAGE_BYTES;       // [STAGES][256 rows][128 B]
  uint64_t *bars = (uint64_t *)(smem + TC_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t *a_full = bars, *b_full = bars + TC_STAGES, *empty = bars + 2 * TC_STAGES, *acc_full = bars + 3 * TC_STAGES;
  uint32_t *tmem_slot = (uint32_t *)(bars + 3 * TC_STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * (TC_MT * TC_BM), n0 = blockIdx.x * TC_BN;
  const int nk = p.K / TC_BK;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; s++) { mbar_init(&a_full[s], 1); mbar_init(&b_full[s], TC_DQ_WARPS); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (activations) =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int kb = 0; kb < nk; kb++) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&a_full[stage], A_STAGE_BYTES);
#pragma unroll
        for (int t = 0; t < TC_MT; t++)
          tma_load_2d(a_st + (size_t)stage * A_STAGE_BYTES + (size_t)t * (TC_BM * 128), &tmap_x, kb * TC_BK, m0 + t * TC_BM, &a_full[stage]);
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: c=f32 (1<<4), a=f16/bf16 (bit 7), b=f16 (0), K-major both,
    // N>>3 at bit 17, M>>4 at bit 24
    const uint32_t a_fmt = (p.out_dtype == MRS_BF16) ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (a_fmt << 7) | ((uint32_t)p.b_fmt << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    int stage = 0, phase = 0;
    for (int kb = 0; kb < nk; kb++) {
      mbar_wait(&a_full[stage], phase);
      mbar_wait(&b_full[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint8_t *bs = b_st + (size_t)stage * B_STAGE_BYTES;
#pragma unroll
        for (int t = 0; t < TC_MT; t++) {
          const uint8_t *as = a_st + (size_t)stage * A_STAGE_BYTES + (size_t)t * (TC_BM * 128);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; k++) {
            // advancing 16 elements (32 bytes) along K inside the swizzle atom = +2 in the
            // 16-byte-granular start address
            const uint64_t ad = umma_desc_sw128(as) + (uint64_t)(2 * k);
            const uint64_t bd = umma_desc_sw128(bs) + (uint64_t)(2 * k);
            umma_f16(tmem_base + (uint32_t)(t * TC_BN), ad, bd, idesc, (kb | k) ? 1u : 0u);
          }
        }
        umma_commit(&empty[stage]);             // frees the smem stage when these MMAs retire
        if (kb == nk - 1) umma_commit(acc_full);  // accumulators complete
      }
      __syncwarp();
      if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================== dequantisers (16 warps, two threads per weight row) =====================
    const int dt_ = tid - 64;               // 0..511
    const int r = dt_ >> 1, hf = dt_ & 1;   // row within the tile, which 32-weight half of the K-step
    const int row = n0 + r;
    const bool live = row < p.N;
    const uint8_t *wrow = p.w + (size_t)(live ? row : 0) * p.row_bytes;
    int stage = 0, phase = 0;
    Raw32<TYPE> raw;
    if constexpr (!IsInt4Ckpt<TYPE>::value) { if (live) load_raw32<TYPE>(wrow, 32 * hf, raw); }
    for (int kb = 0; kb < nk; kb++) {
      const int k = kb * TC_BK + 32 * hf;
      float v[32];
      if (live) {
        if constexpr (IsInt4Ckpt<TYPE>::value) {
          dequant32_ckpt<TYPE>(p, row, k, v);
        } else {
          const Raw32<TYPE> cur = raw;
          if (kb + 1 < nk) load_raw32<TYPE>(wrow, k + TC_BK, raw);   // next K-step in flight
          expand32<TYPE>(cur, wrow, k, v);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = 0.f;
      }
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t *dst = b_st + (size_t)stage * B_STAGE_BYTES + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128;
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {       // 16-byte chunk c of the row lands at chunk c ^ (r % 8)
        const int c = 4 * hf + cc;
        uint4 pk;
        if (p.b_fmt == 0) {
          __half2 *h = (__half2 *)&pk;
#pragma unroll
          for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(v[8 * cc + 2 * i], v[8 * cc + 2 * i + 1]);
        } else {
          __nv_bfloat162 *h = (__nv_bfloat162 *)&pk;
#pragma unroll
          for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[8 * cc + 2 * i], v[8 * cc + 2 * i + 1]);
        }
        *(uint4 *)(dst + ((c ^ (r & 7)) << 4)) = pk;
      }
      fence_proxy_async();                   // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[stage]);
      if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
    }

    // ===================== epilogue: TMEM -> registers -> global =====================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int cg = (warp - 2) >> 2;          // four warps per quarter: 64 columns each
#pragma unroll 1
    for (int t = 0; t < TC_MT; t++) {
      const int tok = m0 + t * TC_BM + q * 32 + lane;
#pragma unroll 1
      for (int cb = 0; cb < 2; cb++) {
        const int col0 = cg * 64 + cb * 32;
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * TC_BN + col0), acc);
        if (tok < p.M) {
          if (p.out_dtype == MRS_BF16) {
            __nv_bfloat16 *yo = (__nv_bfloat16 *)p.y + (size_t)tok * p.N + n0 + col0;
            if (n0 + col0 + 32 <= p.N && (p.N % 8) == 0) {
#pragma unroll
              for (int i = 0; i < 4; i++) {
                uint4 pk;
                __nv_bfloat162 *h = (__nv_bfloat162 *)&pk;
#pragma unroll
                for (int k = 0; k < 4; k++) h[k] = __floats2bfloat162_rn(__uint_as_float(acc[8 * i + 2 * k]), __uint_as_float(acc[8 * i + 2 * k + 1]));
                *(uint4 *)(yo + 8 * i) = pk;
              }
            } else {
              for (int i = 0; i < 32; i++) if (n0 + col0 + i < p.N) yo[i] = __float2bfloat16_rn(__uint_as_float(acc[i]));
            }
          } else {
            __half *yo = (__half *)p.y + (size_t)tok * p.N + n0 + col0;
            for (int i = 0; i < 32; i++) if (n0 + col0 + i < p.N) yo[i] = __float2half_rn(__uint_as_float(acc[i]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- host -----------------------------------------------------------------------------------
template <int TYPE>
static cudaError_t launch_tc(const TcParams &p, const CUtensorMap &tmap, cudaStream_t st) {
  auto kern = mmq_tc_kernel<TYPE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);  // per device: set on every launch
  dim3 grid((p.N + TC_BN - 1) / TC_BN, (p.M + TC_MT * TC_BM - 1) / (TC_MT * TC_BM));
  kern<<<grid, TC_THREADS, TC_SMEM, st>>>(tmap, p);
  return cudaGetLastError();
}

}  // namespace mrs

using namespace mrs;

static int tc_row_bytes(int type, int K) {
  switch (type) {
  case MRS_Q4_0: return K / 32 * 18; case MRS_Q4_1: return K / 32 * 20; case MRS_Q5_0: return K / 32 * 22;
  case MRS_Q5_1: return K / 32 * 24; case MRS_Q8_0: return K / 32 * 34; case MRS_Q2_K: return K / 256 * 84;
  case MRS_Q3_K: return K / 256 * 110; case MRS_Q4_K: return K / 256 * 144; case MRS_Q5_K: return K / 256 * 176;
  case MRS_Q6_K: return K / 256 * 210; default: return 0;
  }
}

// Y[M, N] (dtype) = X[M, K] (dtype, row-major contiguous) . W[N, K]^T (ggml blocks).
// dtype 0 = f16, 1 = bf16.  K must be a multiple of 64 (and of the type's block size).
static int g_tc_b_fmt = -1;  // -1: same 16-bit format as the activations; 0 force f16; 1 force bf16
extern "C" void mrs_mmq_set_weight_format(int32_t fmt) { g_tc_b_fmt = fmt; }

extern "C" int32_t mrs_mmq_gguf_ts(int32_t ggml_type, const void *w, const void *x, void *y, int32_t M, int32_t N, int32_t K,
                                   int32_t dtype, void *stream);   // mmq_ts.cu

extern "C" int32_t mrs_mmq_gguf(int32_t ggml_type, const void *w, const void *x, void *y, int32_t M, int32_t N,
                                int32_t K, int32_t dtype, void *stream) {
  if (M <= 0 || N <= 0) return 0;
  if (g_tc_b_fmt < 0) {   // the second-generation kernel (swap-AB, A operand in tensor memory) when the launch fits it
    const int32_t e = mrs_mmq_gguf_ts(ggml_type, w, x, y, M, N, K, dtype, stream);
    if (e != (int32_t)cudaErrorNotSupported) return e;
  }
  const int rb = tc_row_bytes(ggml_type, K);
  if (rb == 0 || K % 64 != 0 || (dtype != 0 && dtype != 1)) return (int32_t)cudaErrorInvalidValue;
  if (ggml_type >= MRS_Q2_K && K % 256 != 0) return (int32_t)cudaErrorInvalidValue;
  PFN_encodeTiled enc = tc_get_encode();
  if (enc == nullptr) return (int32_t)cudaErrorNotSupported;
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tmap, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                         const_cast<void *>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "mrs_b200: cuTensorMapEncodeTiled failed (%d)\n", (int)r); return (int32_t)cudaErrorInvalidValue; }
  if (((uintptr_t)w & 15) || ((uintptr_t)x & 15) || ((uintptr_t)y & 15)) return (int32_t)cudaErrorMisalignedAddress;
  TcParams p = {(const uint8_t *)w, y, M, N, K, rb, dtype, g_tc_b_fmt < 0 ? dtype : g_tc_b_fmt};
  cudaStream_t st = (cudaStream_t)stream;
  switch (ggml_type) {
  case MRS_Q4_0: return (int32_t)launch_tc<MRS_Q4_0>(p, tmap, st);
  case MRS_Q4_1: return (int32_t)launch_tc<MRS_Q4_1>(p, tmap, st);
  case MRS_Q5_0: return (int32_t)launch_tc<MRS_Q5_0>(p, tmap, st);
  case MRS_Q5_1: return (int32_t)launch_tc<MRS_Q5_1>(p, tmap, st);
  case MRS_Q8_0: return (int32_t)launch_tc<MRS_Q8_0>(p, tmap, st);
  case MRS_Q2_K: return (int32_t)launch_tc<MRS_Q2_K>(p, tmap, st);
  case MRS_Q3_K: return (int32_t)launch_tc<MRS_Q3_K>(p, tmap, st);
  case MRS_Q4_K: return (int32_t)launch_tc<MRS_Q4_K>(p, tmap, st);
  case MRS_Q5_K: return (int32_t)launch_tc<MRS_Q5_K>(p, tmap, st);
  case MRS_Q6_K: return (int32_t)launch_tc<MRS_Q6_K>(p, tmap, st);
  default: return (int32_t)cudaErrorInvalidValue;
  }
}

// GPTQ / AWQ int4 linear on the tcgen05 path, straight from the checkpoint tensors (no Marlin
// repack): Y[M,N] f16 = X[M,K] f16 . W, W[k,n] = (q-8)*s (GPTQ, symmetric) or (q-z)*s (AWQ).
// Mirrors GptqLayer::forward_raw -> marlin_matmul (REF mistralrs-quant/src/gptq/gptq_cuda.rs:357-398,
// marlin_backend.rs); activations are f16 (`quantized_act_type`), bits == 4 only.
extern "C" int32_t mrs_gptq_gemm(const void *x, const int32_t *qweight, const void *scales, const int32_t *qzeros,
                                 const int32_t *g_idx, void *y, int32_t M, int32_t K, int32_t N, int32_t group_size,
                                 int32_t is_awq, void *stream) {
  if (M <= 0 || N <= 0) return 0;
  if (K % 64 != 0 || N % 8 != 0 || group_size <= 0 || K % group_size != 0) return (int32_t)cudaErrorInvalidValue;
  if (is_awq && qzeros == nullptr) return (int32_t)cudaErrorInvalidValue;
  PFN_encodeTiled enc = tc_get_encode();
  if (enc == nullptr) return (int32_t)cudaErrorNotSupported;
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(x), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return (int32_t)cudaErrorInvalidValue;
  TcParams p = {};
  p.y = y; p.M = M; p.N = N; p.K = K; p.out_dtype = MRS_F16; p.b_fmt = 0;
  p.qweight = qweight; p.scales = (const __half *)scales; p.qzeros = qzeros; p.g_idx = g_idx; p.group = group_size;
  return (int32_t)(is_awq ? launch_tc<TYPE_AWQ4>(p, tmap, (cudaStream_t)stream) : launch_tc<TYPE_GPTQ4>(p, tmap, (cudaStream_t)stream));
}

// Packed-affine ggml linear (a5): Y[m, n] = X[m, k] . W^T with W held as unsigned 4- / 8-bit payload + per-group 16-bit
// scale and offset, as mrs_gguf_affine_repack_* (affine.cu) wrote them.  Behind the reference's symbols
// `marlin_affine_{u4,u8}_{f16,bf16}` (REF mistralrs-quant/src/gguf/packed_affine.rs:1458-1509 declarations, :697-752 call
// site: n is the PADDED width, the output is [m, padded_n]; `workspace` is the reference kernel's lock array — unused here).
// Same tcgen05 kernel as the checkpoint-layout int4 GEMM: each weight is fma(q, scale, -offset) in f32, rounded once to the
// activations' 16-bit format, f32 accumulation in tensor memory.  0 ok, -1 bad shape, else a cudaError.
static int32_t affine_gemm(const void *x, const void *payload, const void *scales, const void *offsets, void *y, int m, int k, int n, int group,
                           int bits, int dtype, cudaStream_t st) {
  if (m <= 0 || n <= 0) return 0;
  if (k <= 0 || k % 64 != 0 || (group != 16 && group != 32) || n % 8 != 0) return -1;
  if (x == nullptr || payload == nullptr || scales == nullptr || offsets == nullptr || y == nullptr) return -1;
  if (((uintptr_t)x & 15) || ((uintptr_t)payload & 15) || ((uintptr_t)y & 15) || ((uintptr_t)scales & 1) || ((uintptr_t)offsets & 1))
    return (int32_t)cudaErrorMisalignedAddress;
  PFN_encodeTiled enc = tc_get_encode();
  if (enc == nullptr) return (int32_t)cudaErrorNotSupported;
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)m};
  const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&tmap, dtype == MRS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(x), dims, strides,
          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return (int32_t)cudaErrorInvalidValue;
  TcParams p = {};
  p.y = y; p.M = m; p.N = n; p.K = k; p.out_dtype = dtype; p.b_fmt = dtype; p.group = group;
  p.aff_payload = (const uint8_t *)payload; p.aff_scales = (const uint16_t *)scales; p.aff_offsets = (const uint16_t *)offsets;
  p.aff_bf16 = dtype == MRS_BF16;
  return (int32_t)(bits == 4 ? launch_tc<TYPE_AFF4>(p, tmap, st) : launch_tc<TYPE_AFF8>(p, tmap, st));
}
#define MRS_AFFINE_ENTRY(NAME, BITS, DT)                                                                                                      \
  extern "C" int32_t NAME(const void *input, const void *weight, void *scales, void *offsets, void *output, int32_t m, int32_t k, int32_t n,   \
                          int32_t group_size, void *workspace, int64_t stream) {                                                              \
    (void)workspace;                                                                                                                          \
    return affine_gemm(input, weight, scales, offsets, output, m, k, n, group_size, BITS, DT, (cudaStream_t)stream);                          \
  }
MRS_AFFINE_ENTRY(marlin_affine_u4_f16, 4, MRS_F16)
MRS_AFFINE_ENTRY(marlin_affine_u4_bf16, 4, MRS_BF16)
MRS_AFFINE_ENTRY(marlin_affine_u8_f16, 8, MRS_F16)
MRS_AFFINE_ENTRY(marlin_affine_u8_bf16, 8, MRS_BF16)
