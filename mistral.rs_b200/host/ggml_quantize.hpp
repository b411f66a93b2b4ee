// f32 -> ggml blocks for the 32-wide block types (Q4_0, Q4_1, Q5_0, Q5_1, Q8_0): what the reference's in-situ
// quantisation (ISQ) runs on the host when a GGUF layer is re-quantised to another type.
//   REF mistralrs-quant/src/gguf/mod.rs:633-708 (`GgufMatMul::apply_isq`: same type -> keep the blocks; otherwise
//       dequantise and `QTensor::quantize` to the requested type), :601-604.
// The arithmetic is candle-core `quantized::k_quants` `from_float` (third-party, pinned at candle 0.11.0 — not under
// /root/reference), which follows ggml's `quantize_row_*_ref`; the offline pin used here is gguf-py's bit-exact
// restatement of the same routines (tests/test_isq_host.py).  K-quant quantisation (Q2_K..Q6_K: iterative scale search)
// is NOT provided: no offline reference to pin it against (SURVEY §8c lists it as unpinned).
// Compile with -ffp-contract=off: `x * id + 8.5f` must round twice, as the reference does.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace mrs {

inline uint16_t q_f32_to_f16_bits(float f) {
  const _Float16 h = (_Float16)f;   // round to nearest even
  uint16_t b;
  std::memcpy(&b, &h, 2);
  return b;
}
inline void q_put16(uint8_t *p, uint16_t v) { p[0] = (uint8_t)(v & 0xFF); p[1] = (uint8_t)(v >> 8); }

// bytes per 32-weight block, or 0 for a type without a quantiser here
inline int quantize_block_bytes(int ggml_type) {
  switch (ggml_type) {
  case 2: return 18; case 3: return 20; case 6: return 22; case 7: return 24; case 8: return 34;
  default: return 0;
  }
}

inline void quantize_block32(int type, const float *x, uint8_t *y) {
  if (type == 8) {   // Q8_0: d = amax / 127, q = round-half-away(x / d)
    float amax = 0.f;
    for (int i = 0; i < 32; i++) amax = std::fmax(amax, std::fabs(x[i]));
    const float d = amax / 127.0f, id = d != 0.f ? 1.0f / d : 0.f;
    q_put16(y, q_f32_to_f16_bits(d));
    for (int i = 0; i < 32; i++) y[2 + i] = (uint8_t)(int8_t)std::round(x[i] * id);
    return;
  }
  const bool sym = (type == 2 || type == 6);      // Q4_0 / Q5_0: signed-max scaling; Q4_1 / Q5_1: min + range
  const bool five = (type == 6 || type == 7);
  const int levels = five ? 31 : 15;
  float d, lo = 0.f;
  if (sym) {
    float amax = 0.f, mx = 0.f;                    // the value of largest magnitude, first occurrence, with its sign
    for (int i = 0; i < 32; i++) if (amax < std::fabs(x[i])) { amax = std::fabs(x[i]); mx = x[i]; }
    d = mx / (five ? -16.0f : -8.0f);
  } else {
    float mn = x[0], mx = x[0];
    for (int i = 1; i < 32; i++) { mn = std::fmin(mn, x[i]); mx = std::fmax(mx, x[i]); }
    d = (mx - mn) / (float)levels;
    lo = mn;
  }
  const float id = d != 0.f ? 1.0f / d : 0.f;
  const float bias = sym ? (five ? 16.5f : 8.5f) : 0.5f;
  uint8_t q[32];
  for (int i = 0; i < 32; i++) {
    const float v = sym ? x[i] * id : (x[i] - lo) * id;
    int t = (int)std::trunc(v + bias);
    q[i] = (uint8_t)(t < 0 ? 0 : t > levels ? levels : t);
  }
  uint8_t *p = y;
  q_put16(p, q_f32_to_f16_bits(d)); p += 2;
  if (!sym) { q_put16(p, q_f32_to_f16_bits(lo)); p += 2; }
  if (five) {
    uint32_t qh = 0;
    for (int i = 0; i < 32; i++) qh |= (uint32_t)(q[i] >> 4) << i;
    p[0] = (uint8_t)qh; p[1] = (uint8_t)(qh >> 8); p[2] = (uint8_t)(qh >> 16); p[3] = (uint8_t)(qh >> 24);
    p += 4;
  }
  for (int i = 0; i < 16; i++) p[i] = (uint8_t)((q[i] & 0xF) | (q[i + 16] << 4));
}

// n must be a multiple of 32; returns bytes written, or -1
inline int64_t quantize_row(int type, const float *x, int64_t n, uint8_t *out) {
  const int bb = quantize_block_bytes(type);
  if (bb == 0 || n < 0 || n % 32) return -1;
  for (int64_t b = 0; b < n / 32; b++) quantize_block32(type, x + 32 * b, out + (size_t)b * bb);
  return n / 32 * bb;
}

}  // namespace mrs
