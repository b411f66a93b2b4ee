"""Host quantisers behind `QTensor.quantize` / `GgufMatMul.apply_isq` (C++ host/ggml_quantize.hpp) against gguf-py's
bit-exact restatement of ggml's `quantize_row_*_ref` (the routines candle's `from_float` follows).  Bytes: bit-exact."""
import numpy as np
import pytest
import torch

import oracle
from mistralrs_b200 import quant

try:
    import gguf
    from gguf import quants
except ImportError:                 # the committed fixture below still pins the quantisers
    gguf = quants = None

TYPES = {"q4_0": "Q4_0", "q4_1": "Q4_1", "q5_0": "Q5_0", "q5_1": "Q5_1", "q8_0": "Q8_0"}


@pytest.mark.parametrize("dtype", list(TYPES))
def test_quantize_matches_committed_golden(dtype):   # tests/golden/quantize_golden.npz (make_quantize_golden.py)
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quantize_golden.npz"))
    w = quant.QTensor.quantize(torch.from_numpy(g["x"]), dtype)
    assert np.array_equal(w.data.numpy(), g[dtype])


@pytest.mark.skipif(gguf is None, reason="gguf-py not installed")
@pytest.mark.parametrize("dtype", list(TYPES))
def test_quantize_matches_gguf_py_bit_for_bit(dtype):
    rng = np.random.default_rng(sum(map(ord, dtype)))
    qt = getattr(gguf.GGMLQuantizationType, TYPES[dtype])
    for trial in range(40):
        x = (rng.standard_normal((6, 320)) * rng.choice([1e-4, 1.0, 250.0])).astype(np.float32)
        if trial % 4 == 0:
            x[0, :32] = 0.0                       # an all-zero block: d = 0, id = 0
        if trial % 5 == 0:
            x[1, 35] = -x[1, 33]                  # equal magnitudes: the first occurrence sets the sign of d
        if trial % 6 == 0:
            x[2, 64:96] = 3.25                    # a constant block: zero range for the min/max types
        w = quant.QTensor.quantize(torch.from_numpy(x), dtype)
        assert w.dtype == dtype and tuple(w.shape) == (6, 320) and w.data.dtype == torch.uint8
        assert np.array_equal(w.data.numpy(), quants.quantize(x, qt).reshape(-1))
        # and the oracle's decoder reads the blocks back to within half a step of the block's scale
        back = oracle.dequantize(dtype, w.data.numpy()).reshape(6, 320)
        step = np.abs(x).reshape(6, 10, 32).max(axis=2, keepdims=True) / (7 if dtype[1] == "4" else 15 if dtype[1] == "5" else 127)
        ok = np.abs(back.reshape(6, 10, 32) - x.reshape(6, 10, 32)) <= step * 1.02 + 1e-12
        assert (ok | (step < 1e-3)).all()       # (a scale in f16's subnormal range is itself coarsely rounded: not checked)


def test_quantize_rejects_what_it_cannot_do():
    x = torch.zeros(4, 256)
    for dtype in ("q4_k", "q6_k", "q2_k"):
        with pytest.raises(NotImplementedError, match="K-quant"):
            quant.QTensor.quantize(x, dtype)
    with pytest.raises(ValueError):
        quant.QTensor.quantize(torch.zeros(4, 40), "q8_0")
    with pytest.raises(ValueError):
        quant.QTensor.quantize(torch.zeros(256), "q8_0")


def test_apply_isq_same_type_keeps_the_blocks():   # gguf/mod.rs:641-653: no requantisation when the type already matches
    rng = np.random.default_rng(2)
    blocks = torch.from_numpy(oracle.random_blocks("q4_k", 8 * 2, rng).reshape(-1))
    bias = torch.arange(8, dtype=torch.float32)
    layer = quant.GgufMatMul(quant.QTensor(blocks, "q4_k", (8, 512)), bias)
    for dt in (None, "q4_k"):
        out = layer.apply_isq(dt, "cpu")
        assert out.w.dtype == "q4_k" and torch.equal(out.w.data, blocks) and torch.equal(out.b, bias)
    with pytest.raises(RuntimeError, match="CUDA"):          # a different type needs the device decoders: no host fallback
        layer.apply_isq("q8_0", "cpu")
