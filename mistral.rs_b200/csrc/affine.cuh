// affine.cuh — ggml blocks as "unsigned payload + per-group scale and offset":  w = scale * q - offset.
//
// The reference's opt-in packed path for batched GGUF linears (REF mistralrs-quant/src/gguf/packed_affine.rs:44-69
// `AffineFormatSpec`: payload width and group size per source type; :94-135 the plan: payload padded_n*k*bits/8 bytes,
// k/group*padded_n 16-bit scales and as many offsets) re-tiles every block format into one of two shapes — 4-bit or
// 8-bit unsigned quants with a 16-bit scale and offset per 16 or 32 weights — so that a single GEMM serves all of them.
// This header is the per-format part of that: one 32-weight segment of one weight row -> 32 quants + the scale/offset
// of its one or two groups.  It is plain integer/float code shared by the device repack kernel (affine.cu), by the
// GEMM's dequantiser (mmq_tc.cu) and — compiled for the host by tests/shims — by the CPU tests, which is how the
// per-format arithmetic is checked against the oracle's dequantiser without a GPU.
//
// Payload layout (ours; opaque to the caller, who only allocates the three buffers): row-major per output channel,
//   u4: byte b of a row holds q[2b] | q[2b+1] << 4;   u8: byte k holds q[k];
//   scales / offsets: [padded_n][k / group] 16-bit (f16 or bf16, the dtype of the entry point).
// Block layouts: REF mistralrs-quant/kernels/mmvq_gguf/mmvq_gguf.cu:134-225; the same decoders as dequant.cuh /
// oracle/mrs_oracle.c, restated per 32-weight segment.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#define MRS_AFF_HD __host__ __device__ __forceinline__
#else
#define MRS_AFF_HD inline
#endif

namespace mrs {
namespace affine {

// ggml / UQFF type codes (REF gguf/mod.rs:54-72)
enum : int { F_Q4_0 = 2, F_Q4_1 = 3, F_Q5_0 = 6, F_Q5_1 = 7, F_Q8_0 = 8, F_Q8_1 = 9, F_Q2_K = 10, F_Q3_K = 11, F_Q4_K = 12, F_Q5_K = 13,
              F_Q6_K = 14, F_Q8_K = 15 };

struct Spec {
  int block_elems, block_bytes, bits, group;
};

MRS_AFF_HD bool spec_for(int format, Spec &s) {
  switch (format) {
  case F_Q4_0: s = {32, 18, 4, 32}; return true;
  case F_Q4_1: s = {32, 20, 4, 32}; return true;
  case F_Q5_0: s = {32, 22, 8, 32}; return true;
  case F_Q5_1: s = {32, 24, 8, 32}; return true;
  case F_Q8_0: s = {32, 34, 8, 32}; return true;
  case F_Q8_1: s = {32, 36, 8, 32}; return true;
  case F_Q2_K: s = {256, 84, 4, 16}; return true;
  case F_Q3_K: s = {256, 110, 4, 16}; return true;
  case F_Q4_K: s = {256, 144, 4, 32}; return true;
  case F_Q5_K: s = {256, 176, 8, 32}; return true;
  case F_Q6_K: s = {256, 210, 8, 16}; return true;
  case F_Q8_K: s = {256, 292, 8, 32}; return true;
  default: return false;
  }
}

// ---- 16-bit float <-> f32 on both sides of the compiler ----
MRS_AFF_HD float f16_bits_to_f32(uint16_t h) {
#if defined(__CUDACC__)
  return __half2float(__ushort_as_half(h));   // cuda_fp16.h serves both passes of nvcc
#else
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else {   // subnormal: normalise
      int e = -1;
      do { e++; man <<= 1; } while ((man & 0x400u) == 0);
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
    }
  } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
  else bits = sign | ((exp + 112u) << 23) | (man << 13);
  float f;
  memcpy(&f, &bits, 4);
  return f;
#endif
}
MRS_AFF_HD uint16_t f32_to_f16_bits(float f) {
#if defined(__CUDACC__)
  return __half_as_ushort(__float2half_rn(f));
#else
  const _Float16 h = (_Float16)f;   // round to nearest even
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
#endif
}
MRS_AFF_HD uint16_t f32_to_bf16_bits(float f) {   // round to nearest even; the inputs here are finite
  uint32_t b;
  memcpy(&b, &f, 4);
  return (uint16_t)((b + 0x7FFFu + ((b >> 16) & 1u)) >> 16);
}
MRS_AFF_HD float bf16_bits_to_f32(uint16_t h) {
  const uint32_t b = (uint32_t)h << 16;
  float f;
  memcpy(&f, &b, 4);
  return f;
}
MRS_AFF_HD float ld_f16(const uint8_t *p) { return f16_bits_to_f32((uint16_t)(p[0] | (p[1] << 8))); }

// 6-bit scale / min of sub-block j from the 12 packed bytes of Q4_K / Q5_K
MRS_AFF_HD void k4_scale_min(int j, const uint8_t *s, int &sc, int &m) {
  if (j < 4) { sc = s[j] & 63; m = s[j + 4] & 63; }
  else { sc = (s[j + 4] & 0xF) | ((s[j - 4] >> 6) << 4); m = (s[j + 4] >> 4) | ((s[j] >> 6) << 4); }
}

// 32 weights k0 .. k0+31 (k0 a multiple of 32) of one row: quants q[32] in the payload's unsigned range, and the f32
// scale / offset of the segment's two 16-weight halves (equal for the 32-group formats):  w = sc * q - of.
MRS_AFF_HD void decompose32(int format, const uint8_t *row, int k0, uint8_t *q, float *sc, float *of) {
  switch (format) {
  case F_Q4_0: case F_Q4_1: {
    const uint8_t *b = row + (size_t)(k0 / 32) * (format == F_Q4_0 ? 18 : 20);
    const float d = ld_f16(b);
    const uint8_t *qs = b + (format == F_Q4_0 ? 2 : 4);
    for (int i = 0; i < 16; i++) { q[i] = qs[i] & 0xF; q[i + 16] = qs[i] >> 4; }
    sc[0] = sc[1] = d;
    of[0] = of[1] = (format == F_Q4_0) ? 8.0f * d : -ld_f16(b + 2);
    break;
  }
  case F_Q5_0: case F_Q5_1: {
    const uint8_t *b = row + (size_t)(k0 / 32) * (format == F_Q5_0 ? 22 : 24);
    const float d = ld_f16(b);
    const uint8_t *hp = b + (format == F_Q5_0 ? 2 : 4);
    const uint32_t qh = (uint32_t)hp[0] | ((uint32_t)hp[1] << 8) | ((uint32_t)hp[2] << 16) | ((uint32_t)hp[3] << 24);
    const uint8_t *qs = hp + 4;
    for (int i = 0; i < 16; i++) {
      q[i] = (uint8_t)((qs[i] & 0xF) | (((qh >> i) & 1u) << 4));
      q[i + 16] = (uint8_t)((qs[i] >> 4) | (((qh >> (i + 16)) & 1u) << 4));
    }
    sc[0] = sc[1] = d;
    of[0] = of[1] = (format == F_Q5_0) ? 16.0f * d : -ld_f16(b + 2);
    break;
  }
  case F_Q8_0: case F_Q8_1: {
    const uint8_t *b = row + (size_t)(k0 / 32) * (format == F_Q8_0 ? 34 : 36);
    const float d = ld_f16(b);
    const uint8_t *qs = b + (format == F_Q8_0 ? 2 : 4);
    for (int i = 0; i < 32; i++) q[i] = (uint8_t)(qs[i] ^ 0x80);   // int8 + 128
    sc[0] = sc[1] = d;
    of[0] = of[1] = 128.0f * d;
    break;
  }
  case F_Q8_K: {
    const uint8_t *b = row + (size_t)(k0 / 256) * 292;
    float d;
    memcpy(&d, b, 4);
    const uint8_t *qs = b + 4 + (k0 % 256);
    for (int i = 0; i < 32; i++) q[i] = (uint8_t)(qs[i] ^ 0x80);
    sc[0] = sc[1] = d;
    of[0] = of[1] = 128.0f * d;
    break;
  }
  case F_Q2_K: {
    const uint8_t *b = row + (size_t)(k0 / 256) * 84;
    const int e = k0 % 256, n = e / 128, j = (e % 128) / 32;
    const float d = ld_f16(b + 80), dmin = ld_f16(b + 82);
    const uint8_t *qs = b + 16 + 32 * n;
    for (int l = 0; l < 32; l++) q[l] = (qs[l] >> (2 * j)) & 3;
    for (int h = 0; h < 2; h++) {
      const uint8_t s = b[8 * n + 2 * j + h];
      sc[h] = d * (float)(s & 0xF);
      of[h] = dmin * (float)(s >> 4);
    }
    break;
  }
  case F_Q3_K: {
    const uint8_t *b = row + (size_t)(k0 / 256) * 110;
    const int e = k0 % 256, n = e / 128, j = (e % 128) / 32;
    const float d = ld_f16(b + 108);
    const uint8_t *hm = b, *qs = b + 32 + 32 * n, *s = b + 96;
    for (int l = 0; l < 32; l++) q[l] = (uint8_t)(((qs[l] >> (2 * j)) & 3) | (((hm[l] >> (4 * n + j)) & 1) << 2));
    for (int h = 0; h < 2; h++) {
      const int is = 8 * n + 2 * j + h;
      const int lo = is < 8 ? (s[is] & 0xF) : (s[is - 8] >> 4);
      const int hi = (s[8 + (is & 3)] >> (2 * (is >> 2))) & 3;
      sc[h] = d * (float)((lo | (hi << 4)) - 32);
      of[h] = 4.0f * sc[h];
    }
    break;
  }
  case F_Q4_K: case F_Q5_K: {
    const bool five = format == F_Q5_K;
    const uint8_t *b = row + (size_t)(k0 / 256) * (five ? 176 : 144);
    const int e = k0 % 256, c = e / 64, half = (e % 64) / 32;
    const float d = ld_f16(b), dmin = ld_f16(b + 2);
    const uint8_t *qs = b + (five ? 48 : 16) + 32 * c, *qh = b + 16;
    for (int l = 0; l < 32; l++) {
      uint8_t v = half ? (qs[l] >> 4) : (qs[l] & 0xF);
      if (five) v |= (uint8_t)(((qh[l] >> (2 * c + half)) & 1) << 4);
      q[l] = v;
    }
    int s, m;
    k4_scale_min(2 * c + half, b + 4, s, m);
    sc[0] = sc[1] = d * (float)s;
    of[0] = of[1] = dmin * (float)m;
    break;
  }
  case F_Q6_K: {
    const uint8_t *b = row + (size_t)(k0 / 256) * 210;
    const int e = k0 % 256, n = e / 128, j = (e % 128) / 32;
    const float d = ld_f16(b + 208);
    const uint8_t *ql = b + 64 * n + 32 * (j & 1), *qh = b + 128 + 32 * n;
    const int8_t *s = (const int8_t *)(b + 192) + 8 * n + 2 * j;
    for (int l = 0; l < 32; l++) {
      const uint8_t nib = (j >= 2) ? (ql[l] >> 4) : (ql[l] & 0xF);
      q[l] = (uint8_t)(nib | (((qh[l] >> (2 * j)) & 3) << 4));
    }
    for (int h = 0; h < 2; h++) {
      sc[h] = d * (float)s[h];
      of[h] = 32.0f * sc[h];
    }
    break;
  }
  default:
    for (int i = 0; i < 32; i++) q[i] = 0;
    sc[0] = sc[1] = of[0] = of[1] = 0.f;
  }
}

// One (row, 32-weight segment) of the repack: what a single thread of the device kernel does.
// row == nullptr writes a padding row (all zeros).  bf16 selects the 16-bit format of scales / offsets.
MRS_AFF_HD void repack_segment(int format, const Spec &sp, const uint8_t *row, int seg, uint8_t *payload_row, uint16_t *scales_row,
                               uint16_t *offsets_row, bool bf16) {
  uint8_t q[32];
  float sc[2], of[2];
  if (row) decompose32(format, row, 32 * seg, q, sc, of);
  else {
    for (int i = 0; i < 32; i++) q[i] = 0;
    sc[0] = sc[1] = of[0] = of[1] = 0.f;
  }
  if (sp.bits == 4) {
    uint8_t *dst = payload_row + 16 * seg;
    for (int i = 0; i < 16; i++) dst[i] = (uint8_t)(q[2 * i] | (q[2 * i + 1] << 4));
  } else {
    uint8_t *dst = payload_row + 32 * seg;
    for (int i = 0; i < 32; i++) dst[i] = q[i];
  }
  const int per = 32 / sp.group;   // 1 or 2 groups in the segment
  for (int h = 0; h < per; h++) {
    scales_row[per * seg + h] = bf16 ? f32_to_bf16_bits(sc[h]) : f32_to_f16_bits(sc[h]);
    offsets_row[per * seg + h] = bf16 ? f32_to_bf16_bits(of[h]) : f32_to_f16_bits(of[h]);
  }
}

// The GEMM side: 32 weights k0.. of packed row `n` as f32, from the three packed arrays.  One fused multiply-add per
// weight: w = fma(q, scale, -offset).
MRS_AFF_HD void dequant32(const uint8_t *payload, const uint16_t *scales, const uint16_t *offsets, int bits, int group, bool bf16, int K,
                          int n, int k0, float *out) {
  const int per = 32 / group, gpr = K / group;
  float sc[2], of[2];
  for (int h = 0; h < per; h++) {
    const uint16_t s = scales[(size_t)n * gpr + k0 / group + h], o = offsets[(size_t)n * gpr + k0 / group + h];
    sc[h] = bf16 ? bf16_bits_to_f32(s) : f16_bits_to_f32(s);
    of[h] = bf16 ? bf16_bits_to_f32(o) : f16_bits_to_f32(o);
  }
  if (per == 1) { sc[1] = sc[0]; of[1] = of[0]; }
  if (bits == 4) {
    const uint8_t *src = payload + (size_t)n * (K / 2) + k0 / 2;
#if defined(__CUDA_ARCH__)
    const uint4 raw = *(const uint4 *)src;
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#else
    uint32_t w[4];
    memcpy(w, src, 16);
#endif
    for (int i = 0; i < 32; i++) {
      const float q = (float)((w[i >> 3] >> (4 * (i & 7))) & 0xFu);
#if defined(__CUDA_ARCH__)
      out[i] = __fmaf_rn(q, sc[i >> 4], -of[i >> 4]);
#else
      out[i] = __builtin_fmaf(q, sc[i >> 4], -of[i >> 4]);
#endif
    }
  } else {
    const uint8_t *src = payload + (size_t)n * K + k0;
#if defined(__CUDA_ARCH__)
    const uint4 r0 = *(const uint4 *)src, r1 = *(const uint4 *)(src + 16);
    const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#else
    uint32_t w[8];
    memcpy(w, src, 32);
#endif
    for (int i = 0; i < 32; i++) {
      const float q = (float)((w[i >> 2] >> (8 * (i & 3))) & 0xFFu);
#if defined(__CUDA_ARCH__)
      out[i] = __fmaf_rn(q, sc[i >> 4], -of[i >> 4]);
#else
      out[i] = __builtin_fmaf(q, sc[i >> 4], -of[i >> 4]);
#endif
    }
  }
}

}  // namespace affine
}  // namespace mrs
