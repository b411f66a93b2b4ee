// TEST SHIM (not part of the product): compiles csrc/affine.cuh for the host so the CPU suite can check the
// per-format arithmetic the device repack kernel and the GEMM dequantiser run, element for element, against the oracle.
// The loops below stand in for the CUDA grids: one iteration = one device thread.
#include "../../mistral.rs_b200/csrc/affine.cuh"

using namespace mrs::affine;

extern "C" {

int aff_spec(int format, int *out4) {
  Spec s;
  if (!spec_for(format, s)) return -1;
  out4[0] = s.block_elems; out4[1] = s.block_bytes; out4[2] = s.bits; out4[3] = s.group;
  return 0;
}

int aff_repack_host(int format, const uint8_t *src, uint8_t *payload, uint16_t *scales, uint16_t *offsets, int k, int n, int padded_n,
                    int bf16) {
  Spec s;
  if (!spec_for(format, s) || k % s.block_elems != 0 || k % 64 != 0) return -1;
  const size_t row_bytes = (size_t)(k / s.block_elems) * s.block_bytes, prow = (size_t)k * s.bits / 8, gpr = (size_t)k / s.group;
  for (int r = 0; r < padded_n; r++)
    for (int seg = 0; seg < k / 32; seg++)
      repack_segment(format, s, r < n ? src + r * row_bytes : nullptr, seg, payload + r * prow, scales + r * gpr, offsets + r * gpr, bf16 != 0);
  return 0;
}

void aff_dequant_host(const uint8_t *payload, const uint16_t *scales, const uint16_t *offsets, int bits, int group, int bf16, int k, int rows,
                      float *out) {
  for (int r = 0; r < rows; r++)
    for (int k0 = 0; k0 < k; k0 += 32) dequant32(payload, scales, offsets, bits, group, bf16 != 0, k, r, k0, out + (size_t)r * k + k0);
}
}
