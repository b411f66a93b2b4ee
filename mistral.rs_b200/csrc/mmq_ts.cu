// mmq_ts.cu — prefill GEMM over ggml quant blocks, second generation: swap-AB, A operand in tensor memory.
//
// Same contract as mmq_tc.cu (Y[M,N] = X[M,K] . W[N,K]^T, W in raw ggml blocks, X / Y 16-bit, f32 accumulate;
// replaces REF mistralrs-quant/src/gguf/fast_mmq.rs:762-826 `fast_mmq::plain`), for the types the k-quant
// "M" files and UQFF q8 are made of: Q8_0, Q4_K, Q6_K.  mrs_mmq_gguf routes here when it can (K % 256 == 0,
// 16-byte-multiple rows); every other case stays on mmq_tc.cu.
//
// What changed against mmq_tc.cu, and why (profiles/r02_experiments.md has the measurements):
//   * the WEIGHT rows are the UMMA M dimension (128 rows = 128 TMEM lanes), the tokens the N dimension (256):
//     the dequantised tile is then the A operand, which tcgen05.mma may read from TENSOR MEMORY — the
//     dequantisers write their 16-bit pairs with tcgen05.st (lane = weight row, two consecutive-k values per
//     32-bit column) and nothing of the dequantised tile ever touches shared memory: no swizzled STS, no
//     generic->async proxy fence (MEMBAR.ALL.CTA) per stage, no 32 KB B stage competing with the activations;
//   * the raw ggml bytes come in through the TMA (one 2-D box of 128 rows x one 256-k superblock span per
//     stage) instead of 4-byte global loads from 512 threads: the dequantisers only ever read shared memory,
//     16 bytes at a time, and the loads of stage s+1 are in flight while s is expanded;
//   * MMA issue, TMA issue and barrier traffic are warp-convergent with one elected lane and warp-uniform
//     operands (no per-instruction ELECT/R2UR waterfalls).
// One MMA group (128 k) is 8 x tcgen05.mma M=128 N=256 K=16 = 1024 tensor-pipe clocks; the 16 dequantiser warps
// need ~450 (Q8_0) to ~900 (Q6_K) issue clocks for the same 128 x 128 weights: the kernel is MMA-paced.
//
// Numerics are those of mmq_tc.cu: every weight is dequantised with the reference's f32 formula and rounded
// once to the activation format; activations are used as they are; f32 accumulation in TMEM.
#include "dequant.cuh"
#include "tc_common.cuh"

#include <stdio.h>

namespace mrs {

constexpr int TS_BM = 128;                 // weight rows per CTA (UMMA M, TMEM lanes)
constexpr int TS_AK = 128;                 // k per A stage (64 TMEM columns)
constexpr int TS_RAWK = 256;               // k per raw stage (one superblock span)
constexpr int TS_XK = 64;                  // k per X stage (one SWIZZLE_128B atom row)
constexpr int TS_AS = 4;                   // A stages in tensor memory
constexpr int TS_XS = 4;                   // X stages
constexpr int TS_DQ_WARPS = 16;
constexpr int TS_THREADS = 32 * (3 + TS_DQ_WARPS);   // warp 0 raw TMA | warp 1 MMA | warp 2 X TMA | warps 3..18 dequantisers

template <int TYPE> struct TsFmt;
template <> struct TsFmt<MRS_Q8_0> { static constexpr int SPAN = 272, BOX = 272; };   // 8 blocks of 34 B
template <> struct TsFmt<MRS_Q4_K> { static constexpr int SPAN = 144, BOX = 144; };   // one superblock
template <> struct TsFmt<MRS_Q6_K> { static constexpr int SPAN = 210, BOX = 240; };   // one superblock.  The TMA wants a 16-byte
// aligned box start, a 210-byte superblock starts at 210 s: the box starts at the aligned address below it and the
// dequantisers read from byte phase (210 s) % 16 on (even, the same for every row: rows are 16-byte multiples).
// Pitch 240 keeps 16-byte reads of 8 consecutive rows on distinct banks.
template <int TYPE, int NT> struct TsPlan {
  static constexpr int RAW_BYTES = TS_BM * TsFmt<TYPE>::BOX;
  static constexpr int X_BYTES = NT * 128;
  static constexpr int RS = (TYPE == MRS_Q4_K) ? 3 : 2;
  static constexpr int RING_BYTES = TS_XS * X_BYTES + RS * RAW_BYTES;
  static constexpr int SMEM = 1024 + RING_BYTES + 512;
  static_assert(SMEM <= 227 * 1024, "shared memory plan");
};

struct TsParams {
  void *y;
  int M, N, K, dtype;     // dtype: 0 f16, 1 bf16 (activations, outputs and the dequantised weights)
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *(const uint32_t *)&h;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *(const uint32_t *)&h;
}
// byte i of `t` (0..255) minus `off` as f32: 0x4B000000 | byte = 2^23 + byte, minus (2^23 + off) — exact
template <int OFF = 0>
__device__ __forceinline__ float byte_f32(uint32_t t, int i) {
  return __uint_as_float(prmt(t, 0x4B000000u, 0x7440u | (uint32_t)i)) - (8388608.0f + (float)OFF);
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t *r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
          "r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ---- 32 consecutive weights (unit u of the row's 256-k span) -> 16 packed pairs in the activation format ----
// row: shared-space address of the row's span in the raw stage.  Formulas as dequant.cuh / oracle unpack_block.
// 32 bytes at shared address `a` (2-byte aligned) as 8 words: three aligned 16-byte reads + one funnel shift per word;
// the phase is warp-uniform, so the word-offset switch does not diverge
__device__ __forceinline__ void lds_run32(uint32_t a, uint32_t *W) {
  const uint32_t a0 = a & ~15u, ph = a & 15u;
  const uint4 x0 = lds128(a0), x1 = lds128(a0 + 16u), x2 = lds128(a0 + 32u);
  const uint32_t w[12] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w};
  const uint32_t sh = (ph & 3u) * 8u;
  switch (ph >> 2) {
  case 0:
#pragma unroll
    for (int i = 0; i < 8; i++) W[i] = __funnelshift_r(w[i], w[i + 1], sh);
    break;
  case 1:
#pragma unroll
    for (int i = 0; i < 8; i++) W[i] = __funnelshift_r(w[i + 1], w[i + 2], sh);
    break;
  case 2:
#pragma unroll
    for (int i = 0; i < 8; i++) W[i] = __funnelshift_r(w[i + 2], w[i + 3], sh);
    break;
  default:
#pragma unroll
    for (int i = 0; i < 8; i++) W[i] = __funnelshift_r(w[i + 3], w[i + 4], sh);
    break;
  }
}

template <int TYPE, bool BF>
__device__ __forceinline__ void ts_dequant32(uint32_t row, int u, uint32_t *o) {
  if constexpr (TYPE == MRS_Q8_0) {
    // block u: 34 bytes at 34 u (2-byte aligned): f16 d, then 32 int8.  The thread reads the 9 aligned words that
    // cover it; the phase (0 or 2 bytes) depends on u's parity only — warp-uniform.
    const uint32_t a0 = (row + 34u * (uint32_t)u) & ~3u;
    uint32_t w[9];
#pragma unroll
    for (int i = 0; i < 9; i++) w[i] = lds32(a0 + 4u * i);
    uint32_t d16, q[8];
    if (u & 1) {
      d16 = w[0] >> 16;
#pragma unroll
      for (int i = 0; i < 8; i++) q[i] = w[i + 1];
    } else {
      d16 = w[0] & 0xFFFFu;
#pragma unroll
      for (int i = 0; i < 8; i++) q[i] = __funnelshift_r(w[i], w[i + 1], 16);
    }
    if constexpr (!BF) {
      // f16: (q ^ 0x80) | 0x6400 = 1024 + 128 + q exactly; minus 1152 -> q; times d: one rounding (the product of an
      // 8-bit integer and an 11-bit significand rounds exactly like f32 d * q rounded to f16)
      const uint32_t d2 = d16 * 0x00010001u, c1152 = 0x64806480u;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const uint32_t t = q[i] ^ 0x80808080u;
        const uint32_t lo = prmt(t, 0x64646464u, 0x4140u), hi = prmt(t, 0x64646464u, 0x4342u);
        __half2 a = __hsub2(*(const __half2 *)&lo, *(const __half2 *)&c1152);
        __half2 b = __hsub2(*(const __half2 *)&hi, *(const __half2 *)&c1152);
        a = __hmul2(a, *(const __half2 *)&d2);
        b = __hmul2(b, *(const __half2 *)&d2);
        o[2 * i] = *(const uint32_t *)&a;
        o[2 * i + 1] = *(const uint32_t *)&b;
      }
    } else {
      const float d = __half2float(__ushort_as_half((unsigned short)d16));
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const uint32_t t = q[i] ^ 0x80808080u;
        o[2 * i] = pack_bf16x2(d * byte_f32<128>(t, 0), d * byte_f32<128>(t, 1));
        o[2 * i + 1] = pack_bf16x2(d * byte_f32<128>(t, 2), d * byte_f32<128>(t, 3));
      }
    }
  } else if constexpr (TYPE == MRS_Q4_K) {
    // superblock: [d f16][dmin f16][scales 12 B][qs 128 B]; sub-block u: low (u even) or high nibbles of qs[32 (u >> 1) ..]
    const uint4 hdr = lds128(row);
    const uint4 q0 = lds128(row + 16u + 32u * (uint32_t)(u >> 1)), q1 = lds128(row + 32u + 32u * (uint32_t)(u >> 1));
    const float d = __half2float(__ushort_as_half((unsigned short)(hdr.x & 0xFFFFu)));
    const float dmin = __half2float(__ushort_as_half((unsigned short)(hdr.x >> 16)));
    auto qb = [&](int i) -> int {
      const uint32_t wsel = (i < 4) ? hdr.y : ((i < 8) ? hdr.z : hdr.w);
      return (int)((wsel >> (8 * (i & 3))) & 0xFFu);
    };
    int sc, m;
    if (u < 4) { sc = qb(u) & 63; m = qb(u + 4) & 63; }
    else { sc = (qb(u + 4) & 0xF) | ((qb(u - 4) >> 6) << 4); m = (qb(u + 4) >> 4) | ((qb(u) >> 6) << 4); }
    const float ds = d * (float)sc, nom = -(dmin * (float)m);
    const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    const int sh = (u & 1) ? 4 : 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t t = (w[i] >> sh) & 0x0F0F0F0Fu;
      const float v0 = fmaf(ds, byte_f32(t, 0), nom), v1 = fmaf(ds, byte_f32(t, 1), nom);
      const float v2 = fmaf(ds, byte_f32(t, 2), nom), v3 = fmaf(ds, byte_f32(t, 3), nom);
      o[2 * i] = BF ? pack_bf16x2(v0, v1) : pack_f16x2(v0, v1);
      o[2 * i + 1] = BF ? pack_bf16x2(v2, v3) : pack_f16x2(v2, v3);
    }
  } else {
    // Q6_K superblock: [ql 128][qh 64][scales 16 x i8][d f16]; unit u: half n = u >> 2, quarter kk = u & 3:
    // ql bytes 64 n + 32 (kk & 1) .. +32 (low nibbles for kk < 2, high otherwise), qh bytes 128 + 32 n .. +32
    // (bits 2 kk, 2 kk + 1), scales 192 + 8 n + 2 kk (two, 16 weights each), d at 208
    const int n = u >> 2, kk = u & 3;
    // (`row` carries the stage's byte phase: 2-byte aligned only)
    uint32_t lw[8], hw[8];
    lds_run32(row + 64u * (uint32_t)n + 32u * (uint32_t)(kk & 1), lw);
    lds_run32(row + 128u + 32u * (uint32_t)n, hw);
    const uint32_t sc2 = lds_u16s(row + 192u + 8u * (uint32_t)n + 2u * (uint32_t)kk);
    const float d = __half2float(__ushort_as_half((unsigned short)lds_u16s(row + 208u)));
    const float d0 = d * (float)(int8_t)(sc2 & 0xFFu), d1 = d * (float)(int8_t)(sc2 >> 8);
    const int lsh = (kk >> 1) ? 4 : 0, hsh = 2 * kk;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      // byte-parallel: 6-bit q = low/high nibble of ql | two bits of qh << 4
      const uint32_t t = ((lw[i] >> lsh) & 0x0F0F0F0Fu) | (((hw[i] >> hsh) & 0x03030303u) << 4);
      const float ds = (i < 4) ? d0 : d1;
      const float v0 = ds * byte_f32<32>(t, 0), v1 = ds * byte_f32<32>(t, 1);
      const float v2 = ds * byte_f32<32>(t, 2), v3 = ds * byte_f32<32>(t, 3);
      o[2 * i] = BF ? pack_bf16x2(v0, v1) : pack_f16x2(v0, v1);
      o[2 * i + 1] = BF ? pack_bf16x2(v2, v3) : pack_f16x2(v2, v3);
    }
  }
}

template <int TYPE, int NT, bool BF>
__global__ void __launch_bounds__(TS_THREADS, 1)
mmq_ts_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, const TsParams p) {
  using P = TsPlan<TYPE, NT>;
  using F = TsFmt<TYPE>;
  constexpr int RS = P::RS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *x_ring = smem, *r_ring = smem + TS_XS * P::X_BYTES;
  uint64_t *bars = (uint64_t *)(smem + P::RING_BYTES);
  uint64_t *raw_full = bars, *raw_empty = bars + 4, *x_full = bars + 8, *x_empty = bars + 12, *a_full = bars + 16,
           *a_empty = bars + 20, *acc_full = bars + 24;
  uint32_t *tmem_slot = (uint32_t *)(bars + 25);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * TS_BM, m0 = blockIdx.y * NT;
  const int na = p.K / TS_AK;            // A stages (MMA groups) of this tile
  const int nr = p.K / TS_RAWK;          // raw stages

  if (warp == 0 && lane < 25) {
    uint32_t cnt = 1u;
    if (lane >= 4 && lane < 8) cnt = TS_DQ_WARPS;        // raw_empty
    if (lane >= 16 && lane < 20) cnt = TS_DQ_WARPS;      // a_full
    mbar_init(&bars[lane], cnt);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);   // uniform register
  const uint32_t tmem_a = tmem_base + (uint32_t)NT;                        // A ring after the accumulator

  if (warp == 0) {
    // ===================== raw producer: one 2-D box (128 rows x one superblock span) per stage =====================
    int stage = 0, phase = 0;
    for (int s = 0; s < nr; s++) {
      mbar_wait(&raw_empty[stage], phase ^ 1);
      mbar_arrive_expect_tx_warp(&raw_full[stage], P::RAW_BYTES);
      tma_load_2d_warp(r_ring + (size_t)stage * P::RAW_BYTES, &tmap_w, ((s * F::SPAN) & ~15) / 2, n0, &raw_full[stage]);
      if (++stage == RS) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 2) {
    // ===================== X producer: [NT tokens x 64 k] tiles, SWIZZLE_128B =====================
    int stage = 0, phase = 0;
    const int nx = p.K / TS_XK;
    for (int s = 0; s < nx; s++) {
      mbar_wait(&x_empty[stage], phase ^ 1);
      mbar_arrive_expect_tx_warp(&x_full[stage], P::X_BYTES);
      tma_load_2d_warp(x_ring + (size_t)stage * P::X_BYTES, &tmap_x, s * TS_XK, m0, &x_full[stage]);
      if (++stage == TS_XS) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: D[128 rows, NT tokens] += A[tmem] . X[smem]^T =====================
    const uint32_t fmt = BF ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TS_BM >> 4) << 24);
    int xs = 0, xph = 0, as = 0, aph = 0;
    for (int i = 0; i < na; i++) {
      mbar_wait(&a_full[as], aph);
      const uint32_t ta = tmem_a + (uint32_t)as * 64u;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        mbar_wait(&x_full[xs], xph);
        tc_fence_after();
        const uint64_t xd = umma_desc_sw128(x_ring + (size_t)xs * P::X_BYTES);
#pragma unroll
        for (int j = 0; j < 4; j++)
          umma_f16_ts_warp(tmem_base, ta + (uint32_t)(32 * h + 8 * j), xd + (uint64_t)(2 * j), idesc, (i | h | j) ? 1u : 0u);
        umma_commit_warp(&x_empty[xs]);
        if (++xs == TS_XS) { xs = 0; xph ^= 1; }
      }
      umma_commit_warp(&a_empty[as]);
      if (i == na - 1) umma_commit_warp(acc_full);
      if (++as == TS_AS) { as = 0; aph ^= 1; }
    }
  } else {
    // ===================== dequantisers: warp -> (TMEM lane quarter, 32-k part of the A stage) =====================
    const int q4 = warp & 3, part = (warp - 3) >> 2;
    const int r = q4 * 32 + lane;                      // weight row in the tile == TMEM lane
    const uint32_t row0 = smem_u32(r_ring) + (uint32_t)r * F::BOX;
    const uint32_t ta_base = tmem_a + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(part * 16);
    int rs = 0, rph = 0, as = 0, aph = 0;
    for (int i = 0; i < na; i++) {
      const int half = i & 1;                          // which 128 k of the raw stage
      if (half == 0) mbar_wait(&raw_full[rs], rph);
      uint32_t o[16];
      const uint32_t phase = (uint32_t)(((i >> 1) * F::SPAN) & 15);   // byte phase of this span inside its aligned box (0 unless Q6_K)
      ts_dequant32<TYPE, BF>(row0 + (uint32_t)rs * P::RAW_BYTES + phase, half * 4 + part, o);
      mbar_wait(&a_empty[as], aph ^ 1);                // the MMAs that read this A stage have retired
      tc_fence_after();
      tmem_st_32x16(ta_base + (uint32_t)as * 64u, o);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&a_full[as]);
        if (half == 1) mbar_arrive(&raw_empty[rs]);    // both halves of the span are in registers (and past them)
      }
      if (half == 1 && ++rs == RS) { rs = 0; rph ^= 1; }
      if (++as == TS_AS) { as = 0; aph ^= 1; }
    }

    // ===================== epilogue: TMEM -> registers -> y[token][row] =====================
    // warp (q4, part): rows 32 q4 + lane, token columns part * NT/4 .. +NT/4; for one token the 32 lanes write 32
    // consecutive 16-bit outputs (64 B)
    mbar_wait(acc_full, 0);
    tc_fence_after();
    constexpr int CW = NT / 4;
    const int nrow = n0 + r;
#pragma unroll 1
    for (int c0 = 0; c0 < CW; c0 += 32) {
      uint32_t acc[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(part * CW + c0), acc);
      if (nrow < p.N) {
#pragma unroll
        for (int c = 0; c < 32; c++) {
          const int tok = m0 + part * CW + c0 + c;
          if (tok < p.M) {
            const float v = __uint_as_float(acc[c]);
            if constexpr (BF) ((__nv_bfloat16 *)p.y)[(size_t)tok * p.N + nrow] = __float2bfloat16_rn(v);
            else ((__half *)p.y)[(size_t)tok * p.N + nrow] = __float2half_rn(v);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int TYPE, int NT>
static cudaError_t launch_ts(const void *w, const void *x, void *y, int M, int N, int K, int row_bytes, int dtype, cudaStream_t st) {
  using P = TsPlan<TYPE, NT>;
  using F = TsFmt<TYPE>;
  PFN_encodeTiled enc = tc_get_encode();
  if (enc == nullptr) return cudaErrorNotSupported;
  CUtensorMap tw, tx;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)(row_bytes / 2), (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)(F::BOX / 2), (cuuint32_t)TS_BM};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&tw, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void *>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  if (!tc_make_map_2d(&tx, x, (uint64_t)M, (uint64_t)K, TS_XK, (uint32_t)NT, dtype)) return cudaErrorInvalidValue;
  TsParams p = {y, M, N, K, dtype};
  dim3 grid((N + TS_BM - 1) / TS_BM, (M + NT - 1) / NT);
  if (dtype == 1) {
    auto kern = mmq_ts_kernel<TYPE, NT, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM);
    kern<<<grid, TS_THREADS, P::SMEM, st>>>(tw, tx, p);
  } else {
    auto kern = mmq_ts_kernel<TYPE, NT, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM);
    kern<<<grid, TS_THREADS, P::SMEM, st>>>(tw, tx, p);
  }
  return cudaGetLastError();
}

}  // namespace mrs

using namespace mrs;

// 0: take this kernel whenever the shape allows (default); 1: never (mmq_tc.cu for everything) — A/B and tests
static int g_mmq_path = 0;
extern "C" void mrs_mmq_set_path(int32_t path) { g_mmq_path = path; }

// returns cudaErrorNotSupported when the launch does not fit this kernel (the caller falls back to mmq_tc.cu)
extern "C" int32_t mrs_mmq_gguf_ts(int32_t ggml_type, const void *w, const void *x, void *y, int32_t M, int32_t N, int32_t K,
                                   int32_t dtype, void *stream) {
  if (g_mmq_path == 1) return (int32_t)cudaErrorNotSupported;
  if (ggml_type != MRS_Q8_0 && ggml_type != MRS_Q4_K && ggml_type != MRS_Q6_K) return (int32_t)cudaErrorNotSupported;
  if (M <= 0 || N <= 0 || K % TS_RAWK != 0 || (dtype != 0 && dtype != 1)) return (int32_t)cudaErrorNotSupported;
  const int row_bytes = (ggml_type == MRS_Q8_0) ? K / 32 * 34 : (ggml_type == MRS_Q4_K) ? K / 256 * 144 : K / 256 * 210;
  if (row_bytes % 16 != 0 || ((uintptr_t)w & 15) || ((uintptr_t)x & 15) || ((uintptr_t)y & 1)) return (int32_t)cudaErrorNotSupported;
  cudaStream_t st = (cudaStream_t)stream;
#define MRS_TS(T)                                                                                              \
  return (int32_t)(M <= 128 ? launch_ts<T, 128>(w, x, y, M, N, K, row_bytes, dtype, st)                        \
                            : launch_ts<T, 256>(w, x, y, M, N, K, row_bytes, dtype, st))
  if (ggml_type == MRS_Q8_0) MRS_TS(MRS_Q8_0);
  if (ggml_type == MRS_Q4_K) MRS_TS(MRS_Q4_K);
  MRS_TS(MRS_Q6_K);
#undef MRS_TS
}
