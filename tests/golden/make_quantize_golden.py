"""Writes tests/golden/quantize_golden.npz: f32 inputs and the ggml blocks gguf-py (its bit-exact restatement of ggml's
`quantize_row_*_ref`) produces for them, for the five 32-wide block types.  Run here (needs the `gguf` package):
    python tests/golden/make_quantize_golden.py
The CPU suite checks host/ggml_quantize.hpp against this file, so the pin holds where gguf-py is absent."""
import os

import gguf
import numpy as np
from gguf import quants

rng = np.random.default_rng(20260923)
x = (rng.standard_normal((12, 256)) * np.repeat(rng.choice([1e-3, 1.0, 40.0], size=(12, 8)), 32, axis=1)).astype(np.float32)
x[0, :32] = 0.0                   # all-zero block
x[1, 40] = -x[1, 35]              # equal magnitudes, opposite signs
x[2, 64:96] = -2.5                # constant block
x[3, 96:128] = np.float32(65504.0) * np.linspace(-1, 1, 32, dtype=np.float32)   # scales near the top of f16
out = {"x": x}
for name in ("Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q8_0"):
    out[name.lower()] = quants.quantize(x, getattr(gguf.GGMLQuantizationType, name)).reshape(-1)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "quantize_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
