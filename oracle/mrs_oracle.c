/*
 * mrs_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT).  See mrs_oracle.h.
 *
 * Every function cites the reference file:line whose behaviour it restates.  `REF:` paths are
 * relative to /root/reference (EricLBuehler/mistral.rs @ 9e92262).
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -pthread (oracle/Makefile).  -ffp-contract=off
 * matters: the quantiser's f32 butterfly sums must not be fused.
 */
#define _GNU_SOURCE
#include "mrs_oracle.h"

#include <math.h>
#include <sched.h>
#include <unistd.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define QK_K 256

/* ------------------------------------------------------------------ scalar conversions */

float mrs_f16_to_f32(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000) << 16;
  uint32_t exp = (h >> 10) & 0x1F;
  uint32_t man = h & 0x3FF;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else { /* subnormal */
      int e = -1;
      do { man <<= 1; e++; } while ((man & 0x400) == 0);
      man &= 0x3FF;
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7F800000u | (man << 13);
  } else {
    bits = sign | ((exp + 112) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

uint16_t mrs_f32_to_f16(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000;
  uint32_t ax = x & 0x7FFFFFFFu;
  if (ax >= 0x7F800000u) { /* inf / nan */
    return (uint16_t)(sign | 0x7C00 | ((ax > 0x7F800000u) ? 0x200 : 0));
  }
  if (ax >= 0x477FF000u) { /* rounds to >= 65520 -> inf */
    return (uint16_t)(sign | 0x7C00);
  }
  if (ax < 0x33000001u) { /* < 2^-25 (rounds to zero); exactly 2^-25 ties to even=0 */
    return (uint16_t)sign;
  }
  int32_t e = (int32_t)(ax >> 23) - 127;
  uint32_t m = (ax & 0x7FFFFF) | 0x800000;
  if (e < -14) { /* subnormal half */
    int shift = -14 - e + 13; /* bits to drop from the 24-bit mantissa */
    uint32_t q = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1))) q++;
    return (uint16_t)(sign | q);
  }
  uint32_t q = m >> 13;
  uint32_t rem = m & 0x1FFF;
  uint32_t he = (uint32_t)(e + 15);
  uint32_t out = (he << 10) | (q & 0x3FF);
  if (rem > 0x1000 || (rem == 0x1000 && (q & 1))) out++;
  return (uint16_t)(sign | out);
}

float mrs_bf16_to_f32(uint16_t h) {
  uint32_t bits = (uint32_t)h << 16;
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

uint16_t mrs_f32_to_bf16(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  if ((x & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((x >> 16) | 0x40); /* nan */
  uint32_t lsb = (x >> 16) & 1;
  x += 0x7FFFu + lsb;
  return (uint16_t)(x >> 16);
}

float mrs_round_dtype(float f, int dtype) {
  if (dtype == MRS_F16) return mrs_f16_to_f32(mrs_f32_to_f16(f));
  if (dtype == MRS_BF16) return mrs_bf16_to_f32(mrs_f32_to_bf16(f));
  return f;
}

/* ------------------------------------------------------------------ block layouts
 * REF: mistralrs-quant/kernels/mmvq_gguf/mmvq_gguf.cu:134-225 */
#pragma pack(push, 1)
typedef struct { uint16_t d; uint8_t qs[16]; } blk_q4_0;                     /* 18 */
typedef struct { uint16_t d, m; uint8_t qs[16]; } blk_q4_1;                  /* 20 */
typedef struct { uint16_t d; uint8_t qh[4]; uint8_t qs[16]; } blk_q5_0;      /* 22 */
typedef struct { uint16_t d, m; uint8_t qh[4]; uint8_t qs[16]; } blk_q5_1;   /* 24 */
typedef struct { uint16_t d; int8_t qs[32]; } blk_q8_0;                      /* 34 */
typedef struct { uint16_t d, s; int8_t qs[32]; } blk_q8_1;                   /* 36 */
typedef struct { uint8_t scales[16]; uint8_t qs[64]; uint16_t d, dmin; } blk_q2_k;           /* 84 */
typedef struct { uint8_t hmask[32]; uint8_t qs[64]; uint8_t scales[12]; uint16_t d; } blk_q3_k; /* 110 */
typedef struct { uint16_t d, dmin; uint8_t scales[12]; uint8_t qs[128]; } blk_q4_k;          /* 144 */
typedef struct { uint16_t d, dmin; uint8_t scales[12]; uint8_t qh[32]; uint8_t qs[128]; } blk_q5_k; /* 176 */
typedef struct { uint8_t ql[128]; uint8_t qh[64]; int8_t scales[16]; uint16_t d; } blk_q6_k; /* 210 */
#pragma pack(pop)

int mrs_block_elems(int t) {
  switch (t) {
  case MRS_Q4_0: case MRS_Q4_1: case MRS_Q5_0: case MRS_Q5_1: case MRS_Q8_0: case MRS_Q8_1: return 32;
  case MRS_Q2_K: case MRS_Q3_K: case MRS_Q4_K: case MRS_Q5_K: case MRS_Q6_K: return 256;
  default: return 0;
  }
}
int mrs_block_bytes(int t) {
  switch (t) {
  case MRS_Q4_0: return 18; case MRS_Q4_1: return 20; case MRS_Q5_0: return 22;
  case MRS_Q5_1: return 24; case MRS_Q8_0: return 34; case MRS_Q8_1: return 36;
  case MRS_Q2_K: return 84; case MRS_Q3_K: return 110; case MRS_Q4_K: return 144;
  case MRS_Q5_K: return 176; case MRS_Q6_K: return 210;
  default: return 0;
  }
}

/* 6-bit scale/min unpack of Q4_K/Q5_K — REF: mmvq_gguf.cu:598-609 (the aux[] construction) */
static void scale_min_k4(int j, const uint8_t *q, int *sc, int *m) {
  if (j < 4) {
    *sc = q[j] & 63;
    *m = q[j + 4] & 63;
  } else {
    *sc = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4);
    *m = (q[j + 4] >> 4) | ((q[j - 0] >> 6) << 4);
  }
}

/* Q3_K 6-bit signed scales — REF: mmvq_gguf.cu:366-377 (sc_low | sc_high) - 32 */
static int q3k_scale(const uint8_t *scales, int is) {
  int low = (scales[is % 8] >> (4 * (is / 8))) & 0xF;
  int high = (scales[8 + is % 4] >> (2 * (is / 4))) & 3;
  return (low | (high << 4)) - 32;
}

/* Integer quants of one block in natural element order + the per-group affine description
 * used by both the exact decoder and the Q8_1 arithmetic. */
typedef struct {
  int n;          /* elements (32 / 256) */
  int group;      /* elements per scale group (32, 16) */
  int q[256];     /* integer quant (after the type's fixed offset, e.g. q-8, q-32) */
  double scale[16]; /* per group multiplicative scale  (d * sc) */
  double min[16];   /* per group additive term (-dmin*m or +m) */
} unpacked_t;

static int unpack_block(int type, const uint8_t *p, unpacked_t *u) {
  memset(u->min, 0, sizeof u->min);
  switch (type) {
  case MRS_Q4_0: {
    const blk_q4_0 *b = (const blk_q4_0 *)p;
    u->n = 32; u->group = 32;
    for (int j = 0; j < 16; j++) { u->q[j] = (b->qs[j] & 0xF) - 8; u->q[j + 16] = (b->qs[j] >> 4) - 8; }
    u->scale[0] = mrs_f16_to_f32(b->d);
    return 0;
  }
  case MRS_Q4_1: {
    const blk_q4_1 *b = (const blk_q4_1 *)p;
    u->n = 32; u->group = 32;
    for (int j = 0; j < 16; j++) { u->q[j] = b->qs[j] & 0xF; u->q[j + 16] = b->qs[j] >> 4; }
    u->scale[0] = mrs_f16_to_f32(b->d); u->min[0] = mrs_f16_to_f32(b->m);
    return 0;
  }
  case MRS_Q5_0: case MRS_Q5_1: {
    const uint8_t *qh8, *qs;
    if (type == MRS_Q5_0) {
      const blk_q5_0 *b = (const blk_q5_0 *)p; qh8 = b->qh; qs = b->qs;
      u->scale[0] = mrs_f16_to_f32(b->d);
    } else {
      const blk_q5_1 *b = (const blk_q5_1 *)p; qh8 = b->qh; qs = b->qs;
      u->scale[0] = mrs_f16_to_f32(b->d); u->min[0] = mrs_f16_to_f32(b->m);
    }
    uint32_t qh; memcpy(&qh, qh8, 4);
    u->n = 32; u->group = 32;
    for (int j = 0; j < 16; j++) {
      int h0 = ((qh >> j) << 4) & 0x10;
      int h1 = (qh >> (j + 12)) & 0x10;
      int off = (type == MRS_Q5_0) ? 16 : 0;
      u->q[j] = ((qs[j] & 0xF) | h0) - off;
      u->q[j + 16] = ((qs[j] >> 4) | h1) - off;
    }
    return 0;
  }
  case MRS_Q8_0: {
    const blk_q8_0 *b = (const blk_q8_0 *)p;
    u->n = 32; u->group = 32;
    for (int j = 0; j < 32; j++) u->q[j] = b->qs[j];
    u->scale[0] = mrs_f16_to_f32(b->d);
    return 0;
  }
  case MRS_Q2_K: {
    const blk_q2_k *b = (const blk_q2_k *)p;
    u->n = 256; u->group = 16;
    double d = mrs_f16_to_f32(b->d), dmin = mrs_f16_to_f32(b->dmin);
    for (int g = 0; g < 16; g++) {
      u->scale[g] = d * (b->scales[g] & 0xF);
      u->min[g] = -dmin * (b->scales[g] >> 4);
    }
    /* element e: chunk n=e/128, shift j=(e%128)/32, l=e%32 -> qs[32n+l] >> 2j */
    for (int e = 0; e < 256; e++) {
      int n = e / 128, j = (e % 128) / 32, l = e % 32;
      u->q[e] = (b->qs[32 * n + l] >> (2 * j)) & 3;
    }
    return 0;
  }
  case MRS_Q3_K: {
    const blk_q3_k *b = (const blk_q3_k *)p;
    u->n = 256; u->group = 16;
    double d = mrs_f16_to_f32(b->d);
    for (int g = 0; g < 16; g++) u->scale[g] = d * q3k_scale(b->scales, g);
    for (int e = 0; e < 256; e++) {
      int n = e / 128, j = (e % 128) / 32, l = e % 32;
      int lo = (b->qs[32 * n + l] >> (2 * j)) & 3;
      int hbit = (b->hmask[l] >> (4 * n + j)) & 1;
      u->q[e] = lo - (hbit ? 0 : 4);
    }
    return 0;
  }
  case MRS_Q4_K: case MRS_Q5_K: {
    const uint8_t *scales, *qs, *qh = NULL;
    double d, dmin;
    if (type == MRS_Q4_K) {
      const blk_q4_k *b = (const blk_q4_k *)p;
      scales = b->scales; qs = b->qs; d = mrs_f16_to_f32(b->d); dmin = mrs_f16_to_f32(b->dmin);
    } else {
      const blk_q5_k *b = (const blk_q5_k *)p;
      scales = b->scales; qs = b->qs; qh = b->qh; d = mrs_f16_to_f32(b->d); dmin = mrs_f16_to_f32(b->dmin);
    }
    u->n = 256; u->group = 32;
    for (int g = 0; g < 8; g++) {
      int sc, m; scale_min_k4(g, scales, &sc, &m);
      u->scale[g] = d * sc; u->min[g] = -dmin * m;
    }
    for (int e = 0; e < 256; e++) {
      int j = e / 64, hi = (e % 64) / 32, l = e % 32;
      int v = hi ? (qs[32 * j + l] >> 4) : (qs[32 * j + l] & 0xF);
      if (qh) v |= ((qh[l] >> (2 * j + hi)) & 1) << 4;
      u->q[e] = v;
    }
    return 0;
  }
  case MRS_Q6_K: {
    const blk_q6_k *b = (const blk_q6_k *)p;
    u->n = 256; u->group = 16;
    double d = mrs_f16_to_f32(b->d);
    for (int g = 0; g < 16; g++) u->scale[g] = d * b->scales[g];
    for (int e = 0; e < 256; e++) {
      int n = e / 128, k = (e % 128) / 32, l = e % 32;
      const uint8_t *ql = b->ql + 64 * n, *qh = b->qh + 32 * n;
      int lo = (k & 1) ? ql[l + 32] : ql[l];
      lo = (k & 2) ? (lo >> 4) : (lo & 0xF);
      int hi = (qh[l] >> (2 * k)) & 3;
      u->q[e] = (lo | (hi << 4)) - 32;
    }
    return 0;
  }
  default:
    return -1;
  }
}

int mrs_dequantize(int type, const void *blocks, float *out, int64_t n) {
  int be = mrs_block_elems(type), bb = mrs_block_bytes(type);
  if (!be || type == MRS_Q8_1 || n % be) return -1;
  const uint8_t *p = (const uint8_t *)blocks;
  unpacked_t u;
  for (int64_t i = 0; i < n / be; i++, p += bb) {
    if (unpack_block(type, p, &u)) return -1;
    for (int e = 0; e < u.n; e++) {
      int g = e / u.group;
      out[i * be + e] = (float)(u.scale[g] * u.q[e] + u.min[g]);
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ Q8_1 quantiser
 * REF: mmvq_gguf.cu:1220-1251.  One warp == one 32-element block: amax by butterfly max,
 * sum by butterfly add (xor masks 16,8,4,2,1), d = amax/127, q = (int8)roundf(x/d),
 * ds = (half d, half sum).  NOTE the reference is compiled with --use_fast_math
 * (mistralrs-quant/build.rs:38): its two divisions are approximate on the GPU; this oracle
 * uses IEEE division, so q may differ by +-1 on exact rounding ties (tests bound that). */
void mrs_quantize_q8_1(const float *x, void *y, int kx, int kx_padded, int rows) {
  blk_q8_1 *out = (blk_q8_1 *)y;
  for (int r = 0; r < rows; r++) {
    for (int ib = 0; ib < kx_padded / 32; ib++) {
      float v[32], s[32], t[32];
      float amax = 0.0f;
      for (int l = 0; l < 32; l++) {
        int ix = ib * 32 + l;
        v[l] = (ix < kx) ? x[(int64_t)r * kx + ix] : 0.0f;
        s[l] = v[l];
        amax = fmaxf(amax, fabsf(v[l]));
      }
      for (int mask = 16; mask > 0; mask >>= 1) {
        for (int l = 0; l < 32; l++) t[l] = s[l] + s[l ^ mask];
        memcpy(s, t, sizeof s);
      }
      const float d = amax / 127.0f;
      blk_q8_1 *b = &out[(int64_t)r * (kx_padded / 32) + ib];
      for (int l = 0; l < 32; l++) b->qs[l] = (amax == 0.0f) ? 0 : (int8_t)roundf(v[l] / d);
      b->d = mrs_f32_to_f16(d);
      b->s = mrs_f32_to_f16(s[0]);
    }
  }
}

/* ------------------------------------------------------------------ Q8_1 GEMV arithmetic
 * REF: mmvq_gguf.cu:244-450 (vec_dot_*_impl) and :458-684 (per-type wrappers).  Written in
 * "sum over scale groups" form; the derivation that each reference thread-partial adds up to
 * exactly these integer dots is in DESIGN.md §Oracle.  Integer dots are exact; the float
 * factors (half d/dmin/d8, 6-bit scales) are multiplied exactly in double, so this is the
 * infinitely-precise value of the reference's arithmetic — the reference and our kernels
 * differ from it only by f32 accumulation order. */
static double block_dot_q8_1(int type, const uint8_t *wb, const blk_q8_1 *y) {
  unpacked_t u;
  unpack_block(type, wb, &u);
  double acc = 0.0;
  switch (type) {
  case MRS_Q4_0: case MRS_Q5_0: {
    /* d*(sumi*d8 - off*s8): q here carries the -8/-16 offset already, undo it to follow
     * the reference form (unsigned q, offset applied through s8). REF :244-258, :284-306 */
    int off = (type == MRS_Q4_0) ? 8 : 16;
    long sumi = 0;
    for (int e = 0; e < 32; e++) sumi += (long)(u.q[e] + off) * y->qs[e];
    double d8 = mrs_f16_to_f32(y->d), s8 = mrs_f16_to_f32(y->s);
    acc = u.scale[0] * ((double)sumi * d8 - (double)off * s8);
    break;
  }
  case MRS_Q4_1: case MRS_Q5_1: { /* sumi*d*d8 + m*s8 — REF :260-277, :308-334 */
    long sumi = 0;
    for (int e = 0; e < 32; e++) sumi += (long)u.q[e] * y->qs[e];
    double d8 = mrs_f16_to_f32(y->d), s8 = mrs_f16_to_f32(y->s);
    acc = (double)sumi * ((double)(float)u.scale[0] * d8) + (double)(float)u.min[0] * s8;
    break;
  }
  case MRS_Q8_0: { /* sumi*d8_0*d8_1 — REF :336-346 */
    long sumi = 0;
    for (int e = 0; e < 32; e++) sumi += (long)u.q[e] * y->qs[e];
    acc = (double)sumi * u.scale[0] * (double)mrs_f16_to_f32(y->d);
    break;
  }
  case MRS_Q2_K: case MRS_Q4_K: case MRS_Q5_K: {
    /* dm.x*sum_g d8*(idot*sc) - dm.y*sum_g d8*(isum*m) — REF :348-366, :386-432 */
    int ng = 256 / u.group;
    for (int g = 0; g < ng; g++) {
      const blk_q8_1 *yb = &y[(g * u.group) / 32];
      int base = (g * u.group) % 32;
      long idot = 0, isum = 0;
      for (int e = 0; e < u.group; e++) {
        idot += (long)u.q[g * u.group + e] * yb->qs[base + e];
        isum += yb->qs[base + e];
      }
      double d8 = mrs_f16_to_f32(yb->d);
      acc += d8 * ((double)idot * u.scale[g] + (double)isum * u.min[g]);
    }
    break;
  }
  case MRS_Q3_K: case MRS_Q6_K: { /* d*sum_g d8*(idot*sc) — REF :368-384, :434-450 */
    for (int g = 0; g < 16; g++) {
      const blk_q8_1 *yb = &y[g / 2];
      int base = (g % 2) * 16;
      long idot = 0;
      for (int e = 0; e < 16; e++) idot += (long)u.q[g * 16 + e] * yb->qs[base + e];
      acc += (double)mrs_f16_to_f32(yb->d) * (double)idot * u.scale[g];
    }
    break;
  }
  default: break;
  }
  return acc;
}

int mrs_mmvq_q8_1(int type, const void *w, const void *yq, double *out, int ncols, int nrows,
                  int stride_col_y_blocks, int batch) {
  int be = mrs_block_elems(type), bb = mrs_block_bytes(type);
  if (!be || type == MRS_Q8_1 || ncols % be) return -1;
  int nb = ncols / be;
  const uint8_t *wp = (const uint8_t *)w;
  const blk_q8_1 *y = (const blk_q8_1 *)yq;
  for (int b = 0; b < batch; b++)
    for (int r = 0; r < nrows; r++) {
      double acc = 0.0;
      for (int k = 0; k < nb; k++)
        acc += block_dot_q8_1(type, wp + ((int64_t)r * nb + k) * bb,
                              y + (int64_t)b * stride_col_y_blocks + (int64_t)k * (be / 32));
      out[(int64_t)b * nrows + r] = acc;
    }
  return 0;
}

int mrs_matmul_exact(int type, const void *w, const float *x, double *out, int ncols, int nrows,
                     int batch) {
  int be = mrs_block_elems(type), bb = mrs_block_bytes(type);
  if (!be || type == MRS_Q8_1 || ncols % be) return -1;
  int nb = ncols / be;
  float *row = (float *)malloc(sizeof(float) * (size_t)ncols);
  for (int r = 0; r < nrows; r++) {
    mrs_dequantize(type, (const uint8_t *)w + (int64_t)r * nb * bb, row, ncols);
    for (int b = 0; b < batch; b++) {
      double acc = 0.0;
      for (int k = 0; k < ncols; k++) acc += (double)row[k] * (double)x[(int64_t)b * ncols + k];
      out[(int64_t)b * nrows + r] = acc;
    }
  }
  free(row);
  return 0;
}

/* ------------------------------------------------------------------ reference CPU path
 * REF call site: mistralrs-quant/src/gguf/mod.rs:465-478 -> candle QMatMul::forward (candle
 * v0.11.0 @35d7ae7c, candle-core/src/quantized/k_quants.rs — NOT under /root/reference).
 * Algorithm restated from the published llama.cpp/candle k-quants: activations quantised per
 * 256 to Q8_K (iscale=-128/max, bsums per 16) for K-quants, per 32 to Q8_0 (d=amax/127) for
 * the 32-wide types; integer block dots; f32 accumulation.  Parity vs candle: UNPINNED. */
typedef struct { float d; int8_t qs[256]; int16_t bsums[16]; } blk_q8_k;

static inline int nearest_int(float f) { return (int)lrintf(f); }

static void quantize_row_q8_k(const float *x, blk_q8_k *y, int k) {
  for (int i = 0; i < k / 256; i++, x += 256) {
    float max = 0, amax = 0;
    for (int j = 0; j < 256; j++) { float ax = fabsf(x[j]); if (ax > amax) { amax = ax; max = x[j]; } }
    if (amax == 0) { memset(&y[i], 0, sizeof(blk_q8_k)); continue; }
    const float iscale = -128.f / max;
    for (int j = 0; j < 256; j++) { int v = nearest_int(iscale * x[j]); y[i].qs[j] = (int8_t)(v > 127 ? 127 : v); }
    for (int j = 0; j < 16; j++) { int s = 0; for (int l = 0; l < 16; l++) s += y[i].qs[16 * j + l]; y[i].bsums[j] = (int16_t)s; }
    y[i].d = 1.f / iscale;
  }
}

static void quantize_row_q8_0(const float *x, blk_q8_0 *y, float *yd, int k) {
  for (int i = 0; i < k / 32; i++, x += 32) {
    float amax = 0;
    for (int j = 0; j < 32; j++) amax = fmaxf(amax, fabsf(x[j]));
    const float d = amax / 127.f, id = d ? 1.f / d : 0.f;
    y[i].d = mrs_f32_to_f16(d);
    yd[i] = mrs_f16_to_f32(y[i].d);
    for (int j = 0; j < 32; j++) y[i].qs[j] = (int8_t)roundf(x[j] * id);
  }
}

#if defined(__x86_64__) && defined(__GNUC__)
#define MRS_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define MRS_CLONES
#endif

MRS_CLONES static float dot_q4k_q8k(const blk_q4_k *w, const blk_q8_k *y, int nb) {
  float sumf = 0;
  for (int i = 0; i < nb; i++) {
    int sc[8], mn[8];
    for (int g = 0; g < 8; g++) scale_min_k4(g, w[i].scales, &sc[g], &mn[g]);
    int summ = 0;
    for (int g = 0; g < 8; g++) summ += (y[i].bsums[2 * g] + y[i].bsums[2 * g + 1]) * mn[g];
    int32_t sumi = 0;
    for (int j = 0; j < 4; j++) {
      const uint8_t *q = w[i].qs + 32 * j;
      const int8_t *a = y[i].qs + 64 * j;
      int32_t s0 = 0, s1 = 0;
      for (int l = 0; l < 32; l++) { s0 += (q[l] & 0xF) * a[l]; s1 += (q[l] >> 4) * a[l + 32]; }
      sumi += s0 * sc[2 * j] + s1 * sc[2 * j + 1];
    }
    const float d = mrs_f16_to_f32(w[i].d) * y[i].d, dmin = mrs_f16_to_f32(w[i].dmin) * y[i].d;
    sumf += d * (float)sumi - dmin * (float)summ;
  }
  return sumf;
}

MRS_CLONES static float dot_q6k_q8k(const blk_q6_k *w, const blk_q8_k *y, int nb) {
  float sumf = 0;
  for (int i = 0; i < nb; i++) {
    int32_t sumi = 0;
    for (int n = 0; n < 2; n++) {
      const uint8_t *ql = w[i].ql + 64 * n, *qh = w[i].qh + 32 * n;
      const int8_t *a = y[i].qs + 128 * n, *sc = w[i].scales + 8 * n;
      /* 16-element runs with scalar accumulators (same integers as the l/16-indexed form of
       * REF k_quants vec_dot_q6_K, written so the compiler can vectorise it) */
      for (int is = 0; is < 2; is++) {
        int32_t s1 = 0, s2 = 0, s3 = 0, s4 = 0;
        for (int l = 16 * is; l < 16 * is + 16; l++) {
          const int q1 = ((ql[l] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
          const int q2 = ((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
          const int q3 = ((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
          const int q4 = ((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
          s1 += q1 * a[l]; s2 += q2 * a[l + 32]; s3 += q3 * a[l + 64]; s4 += q4 * a[l + 96];
        }
        sumi += s1 * sc[is] + s2 * sc[is + 2] + s3 * sc[is + 4] + s4 * sc[is + 6];
      }
    }
    sumf += mrs_f16_to_f32(w[i].d) * y[i].d * (float)sumi;
  }
  return sumf;
}

MRS_CLONES static float dot_q8_0_q8_0(const blk_q8_0 *w, const blk_q8_0 *y, const float *yd, int nb) {
  float sumf = 0;
  for (int i = 0; i < nb; i++) {
    int32_t s = 0;
    for (int l = 0; l < 32; l++) s += w[i].qs[l] * y[i].qs[l];
    sumf += (float)s * mrs_f16_to_f32(w[i].d) * yd[i];
  }
  return sumf;
}

/* generic (slower) fallbacks through the unpacker for the remaining types */
static float dot_generic_q8k(int type, const uint8_t *w, const blk_q8_k *y, int nb) {
  float sumf = 0; unpacked_t u; int bb = mrs_block_bytes(type);
  for (int i = 0; i < nb; i++) {
    unpack_block(type, w + (size_t)i * bb, &u);
    double acc = 0;
    for (int g = 0; g < 256 / u.group; g++) {
      long idot = 0, isum = 0;
      for (int e = 0; e < u.group; e++) { idot += (long)u.q[g * u.group + e] * y[i].qs[g * u.group + e]; isum += y[i].qs[g * u.group + e]; }
      acc += (double)idot * u.scale[g] + (double)isum * u.min[g];
    }
    sumf += (float)(acc * y[i].d);
  }
  return sumf;
}
static float dot_generic_q8_0(int type, const uint8_t *w, const blk_q8_0 *y, const float *yd, int nb) {
  float sumf = 0; unpacked_t u; int bb = mrs_block_bytes(type);
  for (int i = 0; i < nb; i++) {
    unpack_block(type, w + (size_t)i * bb, &u);
    long idot = 0, isum = 0;
    for (int e = 0; e < 32; e++) { idot += (long)u.q[e] * y[i].qs[e]; isum += y[i].qs[e]; }
    sumf += (float)(((double)idot * u.scale[0] + (double)isum * u.min[0]) * yd[i]);
  }
  return sumf;
}

typedef struct {
  int type, ncols, r0, r1, nrows, batch;
  const uint8_t *w; const void *yq; const float *yd; float *out;
} cpu_job_t;

static void cpu_job_run(cpu_job_t *j) {
  int be = mrs_block_elems(j->type), bb = mrs_block_bytes(j->type), nb = j->ncols / be;
  for (int r = j->r0; r < j->r1; r++) {
    const uint8_t *wr = j->w + (size_t)r * nb * bb;
    for (int b = 0; b < j->batch; b++) {
      float v;
      if (be == 256) {
        const blk_q8_k *y = (const blk_q8_k *)j->yq + (size_t)b * nb;
        if (j->type == MRS_Q4_K) v = dot_q4k_q8k((const blk_q4_k *)wr, y, nb);
        else if (j->type == MRS_Q6_K) v = dot_q6k_q8k((const blk_q6_k *)wr, y, nb);
        else v = dot_generic_q8k(j->type, wr, y, nb);
      } else {
        const blk_q8_0 *y = (const blk_q8_0 *)j->yq + (size_t)b * nb;
        const float *yd = j->yd + (size_t)b * nb;
        if (j->type == MRS_Q8_0) v = dot_q8_0_q8_0((const blk_q8_0 *)wr, y, yd, nb);
        else v = dot_generic_q8_0(j->type, wr, y, yd, nb);
      }
      j->out[(size_t)b * j->nrows + r] = v;
    }
  }
}

/* persistent pool (the reference's CPU path runs on rayon's persistent pool; spawning threads
 * per GEMV would dominate a 1-row-per-thread decode step) */
#define MRS_MAX_THREADS 256
#if defined(__x86_64__) && defined(__GNUC__)
#define MRS_CPU_RELAX() __builtin_ia32_pause()
#else
#define MRS_CPU_RELAX() do { } while (0)
#endif
static struct {
  pthread_t th[MRS_MAX_THREADS];
  cpu_job_t jobs[MRS_MAX_THREADS];
  pthread_mutex_t mu;
  pthread_cond_t cv_start, cv_done;
  int nthreads, generation, pending, ids[MRS_MAX_THREADS];
} g_pool = {.mu = PTHREAD_MUTEX_INITIALIZER, .cv_start = PTHREAD_COND_INITIALIZER, .cv_done = PTHREAD_COND_INITIALIZER};

static void *pool_worker(void *arg) {
  const int id = *(int *)arg;
  int seen = 0;
  { /* Workers are pinned one per allowed CPU.  First widen the mask (the embedding process —
     * numpy/OpenBLAS — may have pinned the creating thread to one core; every id the mask can hold,
     * the kernel intersects with the cpuset we are allowed), read back what we really got, then take
     * the (id+1)-th CPU of it.  Unpinned sleepers were woken "affine" to the dispatching thread's CPU
     * and ran one after another on it (measured: 8 workers, one core, zero speed-up). */
    cpu_set_t all, got;
    CPU_ZERO(&all);
    for (int c = 0; c < CPU_SETSIZE; c++) CPU_SET(c, &all);
    pthread_setaffinity_np(pthread_self(), sizeof all, &all);
    if (pthread_getaffinity_np(pthread_self(), sizeof got, &got) == 0) {
      const int n = CPU_COUNT(&got);
      if (n > 1) {
        int want = (id + 1) % n, k = 0;
        for (int c = 0; c < CPU_SETSIZE; c++) {
          if (!CPU_ISSET(c, &got)) continue;
          if (k++ == want) {
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(c, &one);
            pthread_setaffinity_np(pthread_self(), sizeof one, &one);
            break;
          }
        }
      }
    }
  }
  for (;;) {
    /* spin briefly before sleeping: a decode step is a chain of sub-millisecond GEMVs */
    for (int spin = 0; spin < 20000 && __atomic_load_n(&g_pool.generation, __ATOMIC_ACQUIRE) == seen; spin++)
      MRS_CPU_RELAX();
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.generation == seen) pthread_cond_wait(&g_pool.cv_start, &g_pool.mu);
    seen = g_pool.generation;
    cpu_job_t job = g_pool.jobs[id];
    pthread_mutex_unlock(&g_pool.mu);
    if (job.r1 > job.r0) cpu_job_run(&job);
    pthread_mutex_lock(&g_pool.mu);
    if (--g_pool.pending == 0) pthread_cond_signal(&g_pool.cv_done);
    pthread_mutex_unlock(&g_pool.mu);
  }
  return NULL;
}

static void pool_ensure(int n) {
  while (g_pool.nthreads < n) {
    g_pool.ids[g_pool.nthreads] = g_pool.nthreads;
    pthread_create(&g_pool.th[g_pool.nthreads], NULL, pool_worker, &g_pool.ids[g_pool.nthreads]);
    g_pool.nthreads++;
  }
}

int mrs_qmatmul_cpu(int type, const void *w, const float *x, float *out, int ncols, int nrows,
                    int batch, int threads) {
  int be = mrs_block_elems(type);
  if (!be || type == MRS_Q8_1 || ncols % be) return -1;
  int nb = ncols / be;
  void *yq; float *yd = NULL;
  if (be == 256) {
    yq = malloc(sizeof(blk_q8_k) * (size_t)nb * batch);
    for (int b = 0; b < batch; b++) quantize_row_q8_k(x + (size_t)b * ncols, (blk_q8_k *)yq + (size_t)b * nb, ncols);
  } else {
    yq = malloc(sizeof(blk_q8_0) * (size_t)nb * batch);
    yd = (float *)malloc(sizeof(float) * (size_t)nb * batch);
    for (int b = 0; b < batch; b++) quantize_row_q8_0(x + (size_t)b * ncols, (blk_q8_0 *)yq + (size_t)b * nb, yd + (size_t)b * nb, ncols);
  }
  if (threads < 1) threads = 1;
  if (threads > nrows) threads = nrows;
  if (threads > MRS_MAX_THREADS) threads = MRS_MAX_THREADS;
  if (threads == 1) {
    cpu_job_t job = {type, ncols, 0, nrows, nrows, batch, (const uint8_t *)w, yq, yd, out};
    cpu_job_run(&job);
  } else {
    /* the calling thread takes the first allowed CPU for the duration of the call (workers sit on
     * the others and spin between jobs; an unpinned caller gets scheduled behind one of them) */
    cpu_set_t saved, one;
    int restore = 0;
    if (sched_getaffinity(0, sizeof saved, &saved) == 0 && CPU_COUNT(&saved) > 1) {
      for (int c = 0; c < CPU_SETSIZE; c++)
        if (CPU_ISSET(c, &saved)) { CPU_ZERO(&one); CPU_SET(c, &one); restore = sched_setaffinity(0, sizeof one, &one) == 0; break; }
    }
    pthread_mutex_lock(&g_pool.mu);
    pool_ensure(threads - 1);
    const int nw = g_pool.nthreads;  /* all pool threads wake; extra ones get empty jobs */
    for (int t = 0; t < nw; t++) {
      cpu_job_t job = {type, ncols, 0, 0, nrows, batch, (const uint8_t *)w, yq, yd, out};
      if (t < threads - 1) { job.r0 = (int)((int64_t)nrows * (t + 1) / threads); job.r1 = (int)((int64_t)nrows * (t + 2) / threads); }
      g_pool.jobs[t] = job;
    }
    g_pool.pending = nw;
    g_pool.generation++;
    pthread_cond_broadcast(&g_pool.cv_start);
    pthread_mutex_unlock(&g_pool.mu);
    cpu_job_t mine = {type, ncols, 0, (int)((int64_t)nrows / threads), nrows, batch, (const uint8_t *)w, yq, yd, out};
    cpu_job_run(&mine);
    for (int spin = 0; spin < 200000 && __atomic_load_n(&g_pool.pending, __ATOMIC_ACQUIRE) > 0; spin++)
      MRS_CPU_RELAX();
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.pending > 0) pthread_cond_wait(&g_pool.cv_done, &g_pool.mu);
    pthread_mutex_unlock(&g_pool.mu);
    if (restore) sched_setaffinity(0, sizeof saved, &saved);
  }
  free(yq); free(yd);
  return 0;
}

/* ------------------------------------------------------------------ GLU
 * REF: mmvq_gguf.cu:52-88; ops.cu:806-847.  (The reference's expf / division are fast-math
 * approximations on the GPU; tolerance in tests covers that.) */
float mrs_glu_act(float x, int act) {
  switch (act) {
  case 1: { const float x3 = x * x * x; return 0.5f * x * (1.0f + tanhf(0.7978845608f * (x + 0.044715f * x3))); }
  case 2: return fmaxf(x, 0.0f);
  case 3: return x * 0.5f * erfcf(-x * 0.70710678f);
  case 4: return 1.0f / (1.0f + expf(-x));
  case 0: default: return x / (1.0f + expf(-x));
  }
}

void mrs_fused_glu(const float *a, const float *b, float *out, int64_t n, int act, int dtype) {
  for (int64_t i = 0; i < n; i++) {
    float activated = mrs_round_dtype(mrs_glu_act(a[i], act), dtype);
    out[i] = mrs_round_dtype(activated * b[i], dtype);
  }
}

/* ------------------------------------------------------------------ RMSNorm */
void mrs_rms_norm(const float *x, const float *w, float *out, int rows, int cols, float eps, int dtype) {
  for (int r = 0; r < rows; r++) {
    double ss = 0;
    for (int c = 0; c < cols; c++) ss += (double)x[(size_t)r * cols + c] * x[(size_t)r * cols + c];
    float inv = 1.0f / sqrtf((float)(ss / cols) + eps);
    for (int c = 0; c < cols; c++) out[(size_t)r * cols + c] = mrs_round_dtype(x[(size_t)r * cols + c] * inv * w[c], dtype);
  }
}

void mrs_add_rms_norm(const float *x, const float *res, const float *w, float *sum_out, float *norm_out,
                      int rows, int cols, float eps, int dtype) {
  for (int r = 0; r < rows; r++) {
    double ss = 0;
    for (int c = 0; c < cols; c++) {
      size_t i = (size_t)r * cols + c;
      float v = mrs_round_dtype(x[i] + res[i], dtype);
      sum_out[i] = v;
      ss += (double)v * v;
    }
    float inv = 1.0f / sqrtf((float)(ss / cols) + eps);
    for (int c = 0; c < cols; c++) {
      size_t i = (size_t)r * cols + c;
      norm_out[i] = mrs_round_dtype(sum_out[i] * inv * w[c], dtype);
    }
  }
}

/* ------------------------------------------------------------------ RoPE
 * REF: rotary.cu:10-34: arr[x] = x*cos - y*sin; arr[y] = y*cos + x*sin with the scalar type's
 * operators.  As compiled for the GPU (verified against the reference kernel's outputs,
 * tests/golden/ref_golden.npz) the second product is rounded to dtype and the first is fused:
 *   out_x = fma(x, cos, -round(y*sin)),  out_y = fma(y, cos, round(x*sin))   (one rounding each). */
static float fma_round(float a, float b, float c, int dtype) {
  if (dtype == MRS_F32) return fmaf(a, b, c);
  /* 16-bit inputs: a*b and the sum with c are exact in double.  Round once to dtype by going
   * through float with round-to-odd (24 bits >= dtype bits + 2, so no double rounding). */
  const double e = (double)a * (double)b + (double)c;
  float f = (float)e;
  if ((double)f != e) {
    float t = (fabs((double)f) > fabs(e)) ? nextafterf(f, 0.0f) : f; /* truncate toward zero */
    uint32_t bits;
    memcpy(&bits, &t, 4);
    bits |= 1u;                                                      /* sticky: make it odd */
    memcpy(&t, &bits, 4);
    f = t;
  }
  return mrs_round_dtype(f, dtype);
}

static void rope_pair(float *arr, int xi, int yi, float c, float s, int dtype) {
  const float x = arr[xi], y = arr[yi];
  const float ys = mrs_round_dtype(y * s, dtype), xs = mrs_round_dtype(x * s, dtype);
  arr[xi] = fma_round(x, c, -ys, dtype);
  arr[yi] = fma_round(y, c, xs, dtype);
}

void mrs_rotary(float *q, float *k, const float *cosb, const float *sinb, const uint32_t *positions,
                int is_neox, int head_size, int64_t tokens, int rot_half, int num_heads,
                int num_kv_heads, int64_t q_stride, int64_t k_stride, int dtype) {
  for (int64_t t = 0; t < tokens; t++) {
    int64_t pos = positions ? positions[t] : t;
    const float *c = cosb + pos * rot_half, *s = sinb + pos * rot_half;
    for (int pass = 0; pass < 2; pass++) {
      float *base = pass ? k + t * k_stride : q + t * q_stride;
      int nh = pass ? num_kv_heads : num_heads;
      for (int h = 0; h < nh; h++)
        for (int o = 0; o < rot_half; o++) {
          if (is_neox) rope_pair(base + (int64_t)h * head_size, o, rot_half + o, c[o], s[o], dtype);
          else rope_pair(base + (int64_t)h * head_size, 2 * o, 2 * o + 1, c[o], s[o], dtype);
        }
    }
  }
}

/* REF: mistralrs-core/src/layers.rs:1071-1160 (+ calculate_default_inv_freq) */
void mrs_llama3_rope_table(float *cosb, float *sinb, int max_pos, int head_dim, float theta,
                           int use_scaling, float factor, float low_freq_factor,
                           float high_freq_factor, int original_max_pos) {
  int half = head_dim / 2;
  float *inv = (float *)malloc(sizeof(float) * (size_t)half);
  for (int i = 0; i < half; i++) {
    float freq = 1.0f / powf(theta, (float)(2 * i) / (float)head_dim);
    if (use_scaling) {
      float low_wl = (float)original_max_pos / low_freq_factor;
      float high_wl = (float)original_max_pos / high_freq_factor;
      float wavelen = 2.0f * (float)M_PI / freq;
      if (wavelen < high_wl) {
      } else if (wavelen > low_wl) {
        freq = freq / factor;
      } else {
        float smooth = ((float)original_max_pos / wavelen - low_freq_factor) / (high_freq_factor - low_freq_factor);
        freq = (1.0f - smooth) * freq / factor + smooth * freq;
      }
    }
    inv[i] = freq;
  }
  for (int p = 0; p < max_pos; p++)
    for (int i = 0; i < half; i++) {
      float f = (float)p * inv[i];
      cosb[(size_t)p * half + i] = cosf(f);
      sinb[(size_t)p * half + i] = sinf(f);
    }
  free(inv);
}

/* ------------------------------------------------------------------ KV cache scatter */
void mrs_reshape_and_cache(const uint16_t *key, const uint16_t *value, uint16_t *kc, uint16_t *vc,
                           const int64_t *slot_mapping, int num_tokens, int num_heads, int head_size,
                           int block_size, int x, int key_stride, int value_stride, int layout) {
  for (int t = 0; t < num_tokens; t++) {
    int64_t slot = slot_mapping[t];
    if (slot < 0) continue;
    int64_t blk = slot / block_size, off = slot % block_size;
    for (int i = 0; i < num_heads * head_size; i++) {
      int h = i / head_size, d = i % head_size;
      int64_t ki, vi;
      if (layout == 0) {
        ki = blk * num_heads * (head_size / x) * block_size * x + (int64_t)h * (head_size / x) * block_size * x +
             (int64_t)(d / x) * block_size * x + off * x + (d % x);
        vi = blk * num_heads * head_size * block_size + (int64_t)h * head_size * block_size + (int64_t)d * block_size + off;
      } else {
        ki = vi = ((blk * num_heads + h) * block_size + off) * head_size + d;
      }
      kc[ki] = key[(int64_t)t * key_stride + i];
      vc[vi] = value[(int64_t)t * value_stride + i];
    }
  }
}

/* ------------------------------------------------------------------ paged decode attention */
static float ld16(const uint16_t *p, int64_t i, int dtype) {
  return dtype == MRS_F16 ? mrs_f16_to_f32(p[i]) : mrs_bf16_to_f32(p[i]);
}

void mrs_paged_attention(const float *q, const uint16_t *kc, const uint16_t *vc, const int32_t *block_tables,
                         const int32_t *context_lens, float *out, int num_seqs, int num_heads,
                         int num_kv_heads, int head_size, int block_size, int max_blocks, int q_stride,
                         float scale, float softcap, int layout, int x, int dtype) {
  int group = num_heads / num_kv_heads;
  for (int s = 0; s < num_seqs; s++) {
    int ctx = context_lens[s];
    double *logits = (double *)malloc(sizeof(double) * (size_t)(ctx > 0 ? ctx : 1));
    for (int h = 0; h < num_heads; h++) {
      int kvh = h / group;
      const float *qv = q + (int64_t)s * q_stride + (int64_t)h * head_size;
      double mx = -INFINITY;
      for (int t = 0; t < ctx; t++) {
        int64_t blk = block_tables[(int64_t)s * max_blocks + t / block_size];
        int off = t % block_size;
        double dot = 0;
        for (int d = 0; d < head_size; d++) {
          int64_t ki = layout == 0
              ? blk * num_kv_heads * head_size * block_size + (int64_t)kvh * head_size * block_size +
                    (int64_t)(d / x) * block_size * x + (int64_t)off * x + (d % x)
              : ((blk * num_kv_heads + kvh) * block_size + off) * head_size + d;
          dot += (double)qv[d] * ld16(kc, ki, dtype);
        }
        double l = dot * scale;
        if (softcap > 0.0f && softcap != 1.0f) l = tanh(l / softcap) * softcap;
        logits[t] = l;
        if (l > mx) mx = l;
      }
      double den = 0;
      for (int t = 0; t < ctx; t++) { logits[t] = exp(logits[t] - mx); den += logits[t]; }
      for (int d = 0; d < head_size; d++) {
        double acc = 0;
        for (int t = 0; t < ctx; t++) {
          int64_t blk = block_tables[(int64_t)s * max_blocks + t / block_size];
          int off = t % block_size;
          int64_t vi = layout == 0
              ? blk * num_kv_heads * head_size * block_size + (int64_t)kvh * head_size * block_size +
                    (int64_t)d * block_size + off
              : ((blk * num_kv_heads + kvh) * block_size + off) * head_size + d;
          acc += logits[t] * ld16(vc, vi, dtype);
        }
        out[((int64_t)s * num_heads + h) * head_size + d] = ctx > 0 ? (float)(acc / den) : 0.0f;
      }
    }
    free(logits);
  }
}
