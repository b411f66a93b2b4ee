// prefill_attn_tc.cu — causal (var-len) prompt attention on the 5th-gen tensor cores (tcgen05 + TMEM), head size 128.
//
// Same contract and arithmetic class as prefill_attn.cu (REF mistralrs-core/src/paged_attention/layers/
// paged_attention.rs:1413-1475 -> flash_attn_varlen): S = QK^T and O = PV on 16-bit MMAs with
// f32 accumulation, online softmax in f32 (base-2 exponent, scale folded), P rounded to the activation dtype before
// the second GEMM.  mrs_prefill_attention routes here for head_dim 128 without window / softcap; prefill_attn.cu
// (mma.sync) keeps every other case.
//
// One CTA = 128 query rows of one (sequence, head), 128-token K/V tiles, 192 threads:
//   warp 0      TMA: Q once (two SWIZZLE_128B boxes of 64 d), then K and V tiles through two 2-stage rings;
//   warp 1      MMA issuer (one elected lane, warp-convergent):
//                 S_j  = Q K_j^T      SS form, both operands K-major in shared memory, M128 N128 K16 x 8
//                 O   += P_j V_j      TS form: P (16-bit pairs) is the A operand in TENSOR MEMORY, V is the B operand
//                                     straight from its row-major [token][d] tile = MN-major SWIZZLE_128B
//                                     (no transpose pass): M128 N128 K16 x 8
//   warps 2..5  softmax + correction + epilogue, thread = query row = TMEM lane: tcgen05.ld the S row (two passes:
//               max, then exp2 / sum / pack), tcgen05.st P IN PLACE over the first 64 columns of S (a row is private
//               to its thread, and chunk c of P lands on columns the thread has already read), rescale O in TMEM only
//               when some row's maximum moved (warp vote), finally O / l -> global.
// The per-tile chain S -> softmax -> PV is latency-bound (~5000 clocks for 1024 of tensor work), so the kernel is
// laid out for TWO CTAs per SM instead of a deeper pipeline inside one: single K / V stages (96 KB of shared memory),
// 256 TMEM columns — S / P [0,128), O [128,256) — and the other CTA's MMAs fill this one's softmax.
#include "tc_common.cuh"

#include <math.h>
#include <stdio.h>

namespace mrs {

constexpr int FT_BM = 128, FT_BN = 128, FT_D = 128;
constexpr int FT_THREADS = 32 * 6;
constexpr int FT_TILE_BYTES = FT_BN * FT_D * 2;                       // 32 KB: two boxes [128 rows][64 d]
constexpr int FT_SMEM = 1024 + 3 * FT_TILE_BYTES + 256;              // Q + K + V + barriers
constexpr uint32_t FT_TCOLS = 256;

struct FtParams {
  void *o;
  const int32_t *cu_seqlens;   // [B + 1] or nullptr (single sequence of length T)
  int T, H, KVH;
  int64_t o_stride;            // elements between consecutive tokens of the output
  float scale_log2;            // softmax_scale * log2(e)
  int causal, bf16;
  uint32_t v_lbo, v_sbo;       // MN-major descriptor strides of the V operand, bytes
};

// shared-memory operand descriptor, SWIZZLE_128B, explicit leading / stride byte offsets
__device__ __forceinline__ uint64_t umma_desc_sw128_ex(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint64_t addr = (uint64_t)(saddr >> 4) & 0x3FFF;
  return addr | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_f16_ss_warp(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t *r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
          "r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__global__ void __launch_bounds__(FT_THREADS, 2)
prefill_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_v, const FtParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *q_s = smem, *k_s = smem + FT_TILE_BYTES, *v_s = smem + 2 * FT_TILE_BYTES;
  uint64_t *bars = (uint64_t *)(smem + 3 * FT_TILE_BYTES);
  uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 2, *v_full = bars + 3, *v_empty = bars + 4, *s_full = bars + 5,
           *p_full = bars + 6, *pv_done = bars + 7;
  uint32_t *tmem_slot = (uint32_t *)(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int kvh = h / (p.H / p.KVH);
  int seq0 = 0, len = p.T;
  if (p.cu_seqlens != nullptr) { seq0 = p.cu_seqlens[b]; len = p.cu_seqlens[b + 1] - seq0; }
  const int ntile_q = (len + FT_BM - 1) / FT_BM;
  const int qt = ntile_q - 1 - (int)blockIdx.x;      // heavy (late) query tiles first
  if (qt < 0) return;
  const int q0 = qt * FT_BM;
  const int q_hi = min(len, q0 + FT_BM) - 1;
  const int kv_end = p.causal ? (q_hi + 1) : len;
  const int nt = (kv_end + FT_BN - 1) / FT_BN;

  if (warp == 0 && lane < 8) {
    mbar_init(&bars[lane], lane == 6 ? 4u : 1u);      // p_full: the four softmax warps
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, FT_TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);
  const uint32_t t_s = tmem_base, t_p = tmem_base, t_o = tmem_base + 128u;   // P overwrites the head of S

  if (warp == 0) {
    // ===================== TMA producer =====================
    mbar_arrive_expect_tx_warp(q_full, FT_TILE_BYTES);
    tma_load_2d_warp(q_s, &tmap_q, h * FT_D, seq0 + q0, q_full);
    tma_load_2d_warp(q_s + FT_TILE_BYTES / 2, &tmap_q, h * FT_D + 64, seq0 + q0, q_full);
    for (int j = 0; j < nt; j++) {
      const int ph = j & 1;
      const int row = seq0 + j * FT_BN;
      mbar_wait(k_empty, ph ^ 1);
      mbar_arrive_expect_tx_warp(k_full, FT_TILE_BYTES);
      tma_load_2d_warp(k_s, &tmap_k, kvh * FT_D, row, k_full);
      tma_load_2d_warp(k_s + FT_TILE_BYTES / 2, &tmap_k, kvh * FT_D + 64, row, k_full);
      mbar_wait(v_empty, ph ^ 1);
      mbar_arrive_expect_tx_warp(v_full, FT_TILE_BYTES);
      tma_load_2d_warp(v_s, &tmap_v, kvh * FT_D, row, v_full);
      tma_load_2d_warp(v_s + FT_TILE_BYTES / 2, &tmap_v, kvh * FT_D + 64, row, v_full);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t fmt = p.bf16 ? 1u : 0u;
    // c = f32 (bit 4), a / b format (bits 7, 10), b_major = MN (bit 16) for the PV product, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc_s = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(FT_BN >> 3) << 17) | ((uint32_t)(FT_BM >> 4) << 24);
    const uint32_t idesc_o = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) | ((uint32_t)(FT_D >> 3) << 17) | ((uint32_t)(FT_BM >> 4) << 24);
    mbar_wait(q_full, 0);
    for (int j = 0; j < nt; j++) {
      const int ph = j & 1;
      mbar_wait(k_full, ph);
      if (j > 0) mbar_wait(pv_done, ph ^ 1);           // PV_{j-1} has read P out of the S columns S_j is about to overwrite
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const uint64_t ad = umma_desc_sw128(q_s + (c >> 2) * (FT_TILE_BYTES / 2)) + (uint64_t)(2 * (c & 3));
        const uint64_t bd = umma_desc_sw128(k_s + (c >> 2) * (FT_TILE_BYTES / 2)) + (uint64_t)(2 * (c & 3));
        umma_f16_ss_warp(t_s, ad, bd, idesc_s, c ? 1u : 0u);
      }
      umma_commit_warp(s_full);
      umma_commit_warp(k_empty);
      mbar_wait(p_full, ph);
      mbar_wait(v_full, ph);
      tc_fence_after();
      const uint32_t vs = smem_u32(v_s);
#pragma unroll
      for (int c = 0; c < 8; c++)     // 16 tokens per MMA: two 8-row groups of 1024 B
        umma_f16_ts_warp(t_o, t_p + (uint32_t)(8 * c), umma_desc_sw128_ex(vs + (uint32_t)c * 2048u, p.v_lbo, p.v_sbo), idesc_o,
                         (j | c) ? 1u : 0u);
      umma_commit_warp(pv_done);
      umma_commit_warp(v_empty);
    }
  } else {
    // ===================== softmax / correction / epilogue: thread = query row =====================
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const int qg = q0 + r;                                 // row index inside the sequence
    // m_ref is the exponent's reference point, not necessarily the running maximum: it only moves when a row's
    // maximum outgrows it by more than 2^8 (then O and l are rescaled).  The quotient O / l does not depend on the
    // reference, P just lives in [0, 256] instead of [0, 1] — and the O round trip through TMEM, with its wait for the
    // previous PV, almost never happens (with a plain running maximum some row of a warp moves in most tiles).
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nt; j++) {
      const int ph = j & 1;
      const uint32_t ts = t_s + lane_off;
      const int kv0 = j * FT_BN;
      const int lim = min(p.causal ? qg : len - 1, len - 1) - kv0;   // columns > lim are masked
      mbar_wait(s_full, ph);
      tc_fence_after();
      // the whole S row in registers: four loads in flight, one wait
      uint32_t v[128];
      tmem_ld_32x32_nowait(ts, v);
      tmem_ld_32x32_nowait(ts + 32u, v + 32);
      tmem_ld_32x32_nowait(ts + 64u, v + 64);
      tmem_ld_32x32_nowait(ts + 96u, v + 96);
      tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; i++) mx = fmaxf(mx, (i <= lim) ? __uint_as_float(v[i]) : -INFINITY);
      mx *= p.scale_log2;                                   // (scale > 0: max commutes with the scaling)
      const bool move = (m_ref == -INFINITY) || (mx > m_ref + 8.0f);
      const float m_use = move ? fmaxf(m_ref, mx) : m_ref;
      const float corr = (m_use == m_ref) ? 1.f : ex2_approx(m_ref - m_use);   // (m_ref = -inf: 0, nothing accumulated yet)
      const float msub = (m_use == -INFINITY) ? 0.f : m_use;
      if (j > 0 && __any_sync(0xffffffffu, corr != 1.f)) {
        mbar_wait(pv_done, ph ^ 1);                         // (already complete: S_j was issued after it)
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < FT_D; c0 += 32) {
          uint32_t o[32];
          tmem_ld_32x32(t_o + lane_off + (uint32_t)c0, o);
#pragma unroll
          for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
          tmem_st_32x32(t_o + lane_off + (uint32_t)c0, o);
        }
      }
      // p = 2^(s * scale - m_use), row sum in f32, P rounded to the activation format, written over the head of S
      float rs = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < FT_BN; c0 += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = (c0 + i <= lim) ? ex2_approx(fmaf(__uint_as_float(v[c0 + i]), p.scale_log2, -msub)) : 0.f;
          const float p1 = (c0 + i + 1 <= lim) ? ex2_approx(fmaf(__uint_as_float(v[c0 + i + 1]), p.scale_log2, -msub)) : 0.f;
          rs += p0 + p1;
          if (p.bf16) { const __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1); pk[i >> 1] = *(const uint32_t *)&hh; }
          else { const __half2 hh = __floats2half2_rn(p0, p1); pk[i >> 1] = *(const uint32_t *)&hh; }
        }
        tmem_st_x16(t_p + lane_off + (uint32_t)(c0 >> 1), pk);
      }
      l_run = l_run * corr + rs;
      m_ref = m_use;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / l -> global, one 256-byte row per thread
    if (nt > 0) {
      mbar_wait(pv_done, (nt - 1) & 1);
      tc_fence_after();
    }
    const float inv = (l_run > 0.f) ? 1.f / l_run : 0.f;
    uint8_t *orow = (uint8_t *)p.o + ((int64_t)(seq0 + qg) * p.o_stride + (int64_t)h * FT_D) * 2;
#pragma unroll 1
    for (int c0 = 0; c0 < FT_D; c0 += 32) {
      uint32_t v[32];
      if (nt > 0) tmem_ld_32x32(t_o + lane_off + (uint32_t)c0, v);
      if (qg < len) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk;
          uint32_t *w = (uint32_t *)&pk;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const float a = (nt > 0) ? __uint_as_float(v[i + 2 * k]) * inv : 0.f, bq = (nt > 0) ? __uint_as_float(v[i + 2 * k + 1]) * inv : 0.f;
            if (p.bf16) { const __nv_bfloat162 hh = __floats2bfloat162_rn(a, bq); w[k] = *(const uint32_t *)&hh; }
            else { const __half2 hh = __floats2half2_rn(a, bq); w[k] = *(const uint32_t *)&hh; }
          }
          *(uint4 *)(orow + (size_t)(c0 + i) * 2) = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, FT_TCOLS);
}

}  // namespace mrs

using namespace mrs;

static uint32_t g_ft_lbo = 16384u, g_ft_sbo = 1024u;
static int g_ft_enable = 1;
// dev knobs: MN-major descriptor strides of V (bytes); enable = 0 keeps mrs_prefill_attention on prefill_attn.cu
extern "C" void mrs_prefill_attn_tc_debug(int32_t enable, uint32_t lbo, uint32_t sbo) {
  g_ft_enable = enable;
  if (lbo) g_ft_lbo = lbo;
  if (sbo) g_ft_sbo = sbo;
}

// returns cudaErrorNotSupported when the call does not fit this kernel (the caller falls back to prefill_attn.cu)
extern "C" int32_t mrs_prefill_attention_tc(const void *q, const void *k, const void *v, void *out, const int32_t *cu_seqlens,
                                            int32_t batch, int32_t total_tokens, int32_t max_seqlen, int32_t num_heads,
                                            int32_t num_kv_heads, int32_t head_dim, int64_t q_stride, int64_t kv_stride,
                                            int64_t o_stride, float softmax_scale, int32_t causal, int32_t window_left,
                                            float softcap, uint32_t dtype, void *stream) {
  if (!g_ft_enable || head_dim != 128 || window_left >= 0 || softcap > 0.f || (dtype != 0 && dtype != 1)) return (int32_t)cudaErrorNotSupported;
  if (total_tokens <= 0) return 0;
  if (num_kv_heads <= 0 || num_heads % num_kv_heads || (q_stride | kv_stride | o_stride) % 8) return (int32_t)cudaErrorNotSupported;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) return (int32_t)cudaErrorNotSupported;
  if (q_stride < (int64_t)num_heads * 128 || kv_stride < (int64_t)num_kv_heads * 128) return (int32_t)cudaErrorNotSupported;
  PFN_encodeTiled enc = tc_get_encode();
  if (enc == nullptr) return (int32_t)cudaErrorNotSupported;
  auto make = [&](CUtensorMap *m, const void *base, int64_t cols, int64_t stride) -> bool {
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)total_tokens};
    const cuuint64_t strides[1] = {(cuuint64_t)stride * 2};
    const cuuint32_t box[2] = {64u, 128u};
    const cuuint32_t estr[2] = {1, 1};
    return enc(m, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims,
               strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  CUtensorMap tq, tk, tv;
  if (!make(&tq, q, (int64_t)num_heads * 128, q_stride) || !make(&tk, k, (int64_t)num_kv_heads * 128, kv_stride) ||
      !make(&tv, v, (int64_t)num_kv_heads * 128, kv_stride))
    return (int32_t)cudaErrorNotSupported;
  FtParams p = {};
  p.o = out; p.cu_seqlens = cu_seqlens; p.T = total_tokens; p.H = num_heads; p.KVH = num_kv_heads; p.o_stride = o_stride;
  p.scale_log2 = softmax_scale * 1.4426950408889634f; p.causal = causal; p.bf16 = (dtype == 1);
  p.v_lbo = g_ft_lbo; p.v_sbo = g_ft_sbo;
  const int nb = cu_seqlens ? batch : 1, ml = cu_seqlens ? max_seqlen : total_tokens;
  dim3 grid((ml + FT_BM - 1) / FT_BM, num_heads, nb);
  cudaFuncSetAttribute(prefill_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
  prefill_attn_tc_kernel<<<grid, FT_THREADS, FT_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, p);
  return (int32_t)cudaGetLastError();
}
