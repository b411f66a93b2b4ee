"""Generates tests/golden/gguf_dequant.npz: random ggml blocks of all ten types and their
dequantisation by gguf-py (pip `gguf` 0.19.0, `gguf.quants.dequantize`) — the independent
published restatement the oracle's block decoders are pinned against.  Run in the build
container:  python tests/golden/make_gguf_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gguf  # noqa: E402
from gguf import GGMLQuantizationType as T  # noqa: E402
from gguf import quants  # noqa: E402

import oracle  # noqa: E402

TYPES = {"q4_0": T.Q4_0, "q4_1": T.Q4_1, "q5_0": T.Q5_0, "q5_1": T.Q5_1, "q8_0": T.Q8_0, "q2_k": T.Q2_K,
         "q3_k": T.Q3_K, "q4_k": T.Q4_K, "q5_k": T.Q5_K, "q6_k": T.Q6_K}
out = {"gguf_version": np.array(gguf.__version__ if hasattr(gguf, "__version__") else "0.19.0")}
rng = np.random.default_rng(20260922)
for name, t in TYPES.items():
    blocks = oracle.random_blocks(name, 12, rng, scale_exp=(-12, 2))
    out[f"{name}_blocks"] = blocks
    out[f"{name}_deq"] = quants.dequantize(blocks, t).reshape(-1).astype(np.float32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gguf_dequant.npz"), **out)
print("wrote gguf_dequant.npz")
