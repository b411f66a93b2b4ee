"""Shared helpers for the parity tests (CUDA path vs oracle/ on the same seeded inputs)."""
import numpy as np
import torch

import oracle

ALL_TYPES = ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0", "q2_k", "q3_k", "q4_k", "q5_k", "q6_k"]
TORCH_DT = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}


def make_weight(dtype, nrows, ncols, seed):
    rng = np.random.default_rng(seed)
    nb = nrows * ncols // oracle.BLOCK_ELEMS[dtype]
    return oracle.random_blocks(dtype, nb, rng)


def make_acts(batch, k, seed, dt):
    """N(0,1) activations rounded through the activation dtype (what the kernel will see)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((batch, k)).astype(np.float32)
    return oracle.round_dtype(x, dt)


def to_dev(a, dev, dt=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t.to(TORCH_DT[dt]) if dt else t


def ulp_report(got, want_f64, dt):
    """Compare kernel output (already in dtype `dt`, widened to f32) with the oracle's
    infinitely-precise value.  Returns (max relative-to-scale error, fraction within 1 ulp)."""
    want = want_f64.astype(np.float64)
    scale = np.abs(want).max() + 1e-30
    err = np.abs(got.astype(np.float64) - want)
    eps = {"bf16": 2.0 ** -8, "f16": 2.0 ** -11, "f32": 2.0 ** -24}[dt]
    tol = eps * np.maximum(np.abs(want), scale * 1e-3) * 1.01 + scale * 2e-6
    return (err / scale).max(), float((err <= tol).mean())
