// Host tail of the on-device sampler: turns the packed rows the top-k / top-1 kernels write (csrc/sampler.cu, behind
// the reference's sort.cu symbols) into a token and its log-probability.  Host-only, no allocation per call beyond a
// k-sized scratch vector.
//   REF mistralrs-core/src/sampler.rs:1172-1273 (sample_topk_on_device: probabilities from the packed values and the
//       softmax normaliser, nucleus cut on the mass renormalised over the kept set :532-534, min-p against the first
//       probability, weighted draw), :666-742 (the batched row form with a per-row k), :1284-1297 (top-1 row check).
// The random number stays with the caller (the reference draws from its Isaac64 stream through rand's WeightedIndex):
// this takes one uniform variate u in [0,1) per row and applies the same rule — the first entry whose running weight
// exceeds u * total.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace mrs {

enum SampleStatus : int {
  SAMPLE_OK = 0,
  SAMPLE_BAD_LENGTH = -1,       // row is not 2*packed_k + 2 floats, or k out of range
  SAMPLE_BAD_NORMALIZER = -2,   // denominator <= 0 / not finite, or max not finite
  SAMPLE_ALL_ZERO = -3,         // nothing survives the filters
  SAMPLE_BAD_WEIGHT = -4,       // a negative or non-finite probability (NaN logits upstream)
  SAMPLE_BAD_TOP1 = -5,         // top-1 row is not (finite max, non-negative integer index)
};

// packed = [values k | indices-as-float k | denom | global max]; row_k <= packed_k entries are used (1 = greedy)
inline int sample_topk_packed_row(const float *packed, int64_t packed_len, int64_t packed_k, int64_t row_k, float inv_temperature,
                                  float top_p, float min_p, double u, uint32_t *token, float *logprob, int32_t *selected_out = nullptr) {
  if (packed_k <= 0 || packed_len != 2 * packed_k + 2 || row_k <= 0) return SAMPLE_BAD_LENGTH;
  if (row_k > packed_k) row_k = packed_k;
  const float denom = packed[2 * packed_k], gmax = packed[2 * packed_k + 1];
  if (!(denom > 0.0f) || !std::isfinite(denom) || !std::isfinite(gmax)) return SAMPLE_BAD_NORMALIZER;
  std::vector<float> report((size_t)row_k);
  for (int64_t i = 0; i < row_k; i++) report[(size_t)i] = std::exp(packed[i] * inv_temperature - gmax) / denom;
  int64_t sel = 0;
  if (row_k > 1) {
    std::vector<float> w(report);
    if (top_p > 0.0f && top_p < 1.0f) {
      float kept = 0.0f;
      for (float p : w) kept += p;
      const float cutoff = top_p * kept;
      float run = 0.0f;
      for (float &p : w) {
        if (run >= cutoff) p = 0.0f; else run += p;
      }
    }
    if (min_p > 0.0f && min_p < 1.0f) {
      const float thr = w[0] * min_p;
      for (float &p : w) if (thr >= p) p = 0.0f;
    }
    float total = 0.0f;
    for (float p : w) {
      if (!(p >= 0.0f) || !std::isfinite(p)) return SAMPLE_BAD_WEIGHT;
      total += p;
    }
    if (!(total > 0.0f)) return SAMPLE_ALL_ZERO;
    const float chosen = (float)(u * (double)total);
    float run = 0.0f;
    sel = row_k - 1;
    for (int64_t i = 0; i + 1 < row_k; i++) {   // running weights of all but the last entry, as the reference's sampler keeps them
      run += w[(size_t)i];
      if (run > chosen) { sel = i; break; }
    }
    while (sel > 0 && w[(size_t)sel] == 0.0f) sel--;   // u*total rounding up to the total must not land on a filtered entry
  }
  *token = (uint32_t)packed[packed_k + sel];
  *logprob = std::log(report[(size_t)sel]);
  if (selected_out) *selected_out = (int32_t)sel;
  return SAMPLE_OK;
}

inline int sample_top1_row(const float *packed, uint32_t *token) {
  const float mx = packed[0], idx = packed[1];
  if (!std::isfinite(mx) || !std::isfinite(idx) || idx < 0.0f || idx != std::floor(idx)) return SAMPLE_BAD_TOP1;
  *token = (uint32_t)idx;
  return SAMPLE_OK;
}

}  // namespace mrs
