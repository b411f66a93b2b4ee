// dequant.cuh — exact per-element decoders of the ten ggml block types (device side).
#pragma once
#include "common.cuh"

namespace mrs {

// ------------------------------------------------------------------ exact block decoders
// (same formulas as oracle/mrs_oracle.c unpack_block; layouts REF mmvq_gguf.cu:134-225)
__device__ __forceinline__ float h2f(const uint8_t *p) {
  return __half2float(__ushort_as_half((unsigned short)(p[0] | (p[1] << 8))));
}
__device__ __forceinline__ void scale_min_k4(int j, const uint8_t *q, int &sc, int &m) {
  if (j < 4) { sc = q[j] & 63; m = q[j + 4] & 63; }
  else { sc = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}

__device__ __forceinline__ float dequant_elem(int type, const uint8_t *b, int e) {
  switch (type) {
  case MRS_Q4_0: { const int q = (e < 16) ? (b[2 + e] & 0xF) : (b[2 + e - 16] >> 4); return h2f(b) * (float)(q - 8); }
  case MRS_Q4_1: { const int q = (e < 16) ? (b[4 + e] & 0xF) : (b[4 + e - 16] >> 4); return h2f(b) * (float)q + h2f(b + 2); }
  case MRS_Q5_0: case MRS_Q5_1: {
    const int o = (type == MRS_Q5_0) ? 2 : 4;
    const uint32_t qh = b[o] | (b[o + 1] << 8) | (b[o + 2] << 16) | ((uint32_t)b[o + 3] << 24);
    const int j = e & 15;
    const int lo = (e < 16) ? (b[o + 4 + j] & 0xF) : (b[o + 4 + j] >> 4);
    const int hi = (qh >> e) & 1;
    const int q = lo | (hi << 4);
    return (type == MRS_Q5_0) ? h2f(b) * (float)(q - 16) : h2f(b) * (float)q + h2f(b + 2);
  }
  case MRS_Q8_0: return h2f(b) * (float)(int8_t)b[2 + e];
  case MRS_Q2_K: {
    const int n = e / 128, j = (e % 128) / 32, l = e % 32, g = e / 16;
    const int q = (b[16 + 32 * n + l] >> (2 * j)) & 3;
    return h2f(b + 80) * (float)(b[g] & 0xF) * (float)q - h2f(b + 82) * (float)(b[g] >> 4);
  }
  case MRS_Q3_K: {
    const int n = e / 128, j = (e % 128) / 32, l = e % 32, is = e / 16;
    const int lo = (b[32 + 32 * n + l] >> (2 * j)) & 3;
    const int hb = (b[l] >> (4 * n + j)) & 1;
    const uint8_t *s = b + 96;
    const int sc = (((s[is & 7] >> (4 * (is >> 3))) & 0xF) | (((s[8 + (is & 3)] >> (2 * (is >> 2))) & 3) << 4)) - 32;
    return h2f(b + 108) * (float)sc * (float)(lo - (hb ? 0 : 4));
  }
  case MRS_Q4_K: case MRS_Q5_K: {
    const int j = e / 64, hi = (e % 64) / 32, l = e % 32;
    const uint8_t *qs = b + ((type == MRS_Q4_K) ? 16 : 48);
    int v = hi ? (qs[32 * j + l] >> 4) : (qs[32 * j + l] & 0xF);
    if (type == MRS_Q5_K) v |= ((b[16 + l] >> (2 * j + hi)) & 1) << 4;
    int sc, m;
    scale_min_k4(2 * j + hi, b + 4, sc, m);
    return h2f(b) * (float)sc * (float)v - h2f(b + 2) * (float)m;
  }
  case MRS_Q6_K: {
    const int n = e / 128, k = (e % 128) / 32, l = e % 32;
    const uint8_t *ql = b + 64 * n, *qh = b + 128 + 32 * n;
    int lo = (k & 1) ? ql[l + 32] : ql[l];
    lo = (k & 2) ? (lo >> 4) : (lo & 0xF);
    const int q = (lo | (((qh[l] >> (2 * k)) & 3) << 4)) - 32;
    return h2f(b + 208) * (float)(int8_t)b[192 + e / 16] * (float)q;
  }
  default: return 0.f;
  }
}

__host__ __device__ __forceinline__ int blk_elems(int t) { return (t >= MRS_Q2_K) ? 256 : 32; }
__host__ __device__ __forceinline__ int blk_bytes(int t) {
  switch (t) {
  case MRS_Q4_0: return 18; case MRS_Q4_1: return 20; case MRS_Q5_0: return 22; case MRS_Q5_1: return 24;
  case MRS_Q8_0: return 34; case MRS_Q2_K: return 84; case MRS_Q3_K: return 110; case MRS_Q4_K: return 144;
  case MRS_Q5_K: return 176; case MRS_Q6_K: return 210; default: return 0;
  }
}


}  // namespace mrs
