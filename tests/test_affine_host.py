"""Packed-affine view of ggml blocks (a5), CPU side.
(1) the numpy restatement `scale * q - offset` reproduces the oracle's dequantiser for the ten ggml types it decodes;
(2) csrc/affine.cuh — the code the device repack kernel and the GEMM dequantiser run per thread — compiled for the host
    by tests/shims, gives bit-identical payload / scales / offsets and weights to the numpy restatement, all twelve
    source formats, f16 and bf16 metadata, with N padding;
(3) the host-side plan (format table, shape rule, padding) against the reference's own unit tests
    (packed_affine.rs `format_specs_cover_all_quantized_gguf_types`, `marlin_shape_filter_matches_available_tiles`).
Integers and 16-bit patterns: bit-exact."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from oracle import affine_np as A

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "shims", "affine_host.cpp")
HDR = os.path.join(HERE, "..", "mistral.rs_b200", "csrc", "affine.cuh")
OUT = os.path.join(HERE, "shims", "_build", "libaffine_host.so")


@pytest.fixture(scope="module")
def shim():
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", SRC, "-o", OUT])
    return ctypes.CDLL(OUT)


def _blocks(dtype, nblocks, rng):
    if dtype in oracle.F16_FIELDS:
        return oracle.random_blocks(dtype, nblocks, rng)
    bb = A.SPECS[dtype][2]
    raw = rng.integers(0, 256, size=(nblocks, bb), dtype=np.uint8)
    d = np.exp2(rng.uniform(-9, -7, size=nblocks))
    if dtype == "q8_1":
        raw[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(nblocks, 2)
        raw[:, 2:4] = np.zeros(nblocks, np.float16).view(np.uint8).reshape(nblocks, 2)
    else:   # q8_k: f32 scale
        raw[:, 0:4] = d.astype(np.float32).view(np.uint8).reshape(nblocks, 4)
    return raw


@pytest.mark.parametrize("dtype", ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0", "q2_k", "q3_k", "q4_k", "q5_k", "q6_k"])
def test_decomposition_reproduces_the_dequantiser(dtype):
    rng = np.random.default_rng(A.SPECS[dtype][0])
    _, elems, bb, bits, group, _ = A.SPECS[dtype]
    blocks = _blocks(dtype, 96, rng)
    q, sc, of = A.decompose(dtype, blocks)
    assert q.dtype == np.uint8 and int(q.max()) < (1 << bits) and sc.shape == (96, elems // group)
    w = np.repeat(sc, group, axis=1).astype(np.float64) * q - np.repeat(of, group, axis=1)
    ref = oracle.dequantize(dtype, blocks).reshape(96, elems).astype(np.float64)
    assert np.allclose(w, ref, rtol=3e-7, atol=1e-9)


def test_q8_1_and_q8_k_decomposition():
    rng = np.random.default_rng(3)
    b = _blocks("q8_1", 40, rng)
    q, sc, of = A.decompose("q8_1", b)
    d = b[:, 0:2].copy().view(np.float16).astype(np.float64)
    assert np.array_equal(sc * q - of, d * b[:, 4:].view(np.int8))
    b = _blocks("q8_k", 10, rng)
    q, sc, of = A.decompose("q8_k", b)
    d = b[:, 0:4].copy().view(np.float32).astype(np.float64)
    w = np.repeat(sc, 32, axis=1).astype(np.float64) * q - np.repeat(of, 32, axis=1)
    assert np.allclose(w, d * b[:, 4:260].view(np.int8), rtol=1e-6)


@pytest.mark.parametrize("bf16", [0, 1])
@pytest.mark.parametrize("dtype", list(A.SPECS))
def test_device_code_on_the_host_vs_oracle(shim, dtype, bf16):
    code, elems, bb, bits, group, _ = A.SPECS[dtype]
    spec = (ctypes.c_int * 4)()
    assert shim.aff_spec(code, spec) == 0 and list(spec) == [elems, bb, bits, group]
    rng = np.random.default_rng(100 + code)
    n, k = 37, 512
    padded_n = A.padded_n_for_shape(n, k)
    assert padded_n == 64
    blocks = _blocks(dtype, n * k // elems, rng)
    pay = np.full((padded_n, k * bits // 8), 0xAA, np.uint8)
    sc = np.full((padded_n, k // group), 0xAAAA, np.uint16)
    of = np.full((padded_n, k // group), 0xAAAA, np.uint16)
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert shim.aff_repack_host(code, P(blocks), P(pay), P(sc), P(of), k, n, padded_n, bf16) == 0
    epay, esc, eof = A.repack(dtype, blocks, n, k, padded_n, bool(bf16))
    assert np.array_equal(pay, epay) and np.array_equal(sc, esc) and np.array_equal(of, eof)
    assert not pay[n:].any() and not sc[n:].any() and not of[n:].any()          # padding rows are zero weights
    w = np.empty((padded_n, k), np.float32)
    shim.aff_dequant_host(P(pay), P(sc), P(of), bits, group, bf16, k, padded_n, P(w))
    ew = A.weights(dtype, blocks, n, k, bool(bf16))
    assert np.array_equal(w[:n], ew) and not w[n:].any()
    # and the packed weights stay within one 16-bit rounding of scale and offset of the exact dequantised ones
    if dtype in oracle.F16_FIELDS:
        ref = oracle.dequantize(dtype, blocks).reshape(n, k)
        qmax = (1 << bits) - 1
        tol = (2.0 ** (-8 if bf16 else -11)) * (np.abs(ref).max() * 2 + 1e-6) * 2 * max(1, qmax / 8)
        assert np.abs(w[:n] - ref).max() <= tol


def test_unsupported_format_and_shapes(shim):
    assert shim.aff_spec(1, (ctypes.c_int * 4)()) == -1 and shim.aff_spec(30, (ctypes.c_int * 4)()) == -1
    z = np.zeros(4096, np.uint8)
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert shim.aff_repack_host(12, P(z), P(z), P(z), P(z), 128, 1, 64, 0) == -1     # K must hold whole 256-blocks


def test_format_specs_cover_all_quantized_gguf_types():   # packed_affine.rs:912-947
    from mistralrs_b200 import packed_affine as PA
    got = [(t,) + PA.AffineFormatSpec.for_dtype(t)[:] for t in PA.AFFINE_DTYPES]
    M = PA.GGUF_AFFINE_MIN_BATCH
    assert got == [("q4_0", 2, 32, 4, 32, M), ("q4_1", 3, 32, 4, 32, M), ("q5_0", 6, 32, 8, 32, 16), ("q5_1", 7, 32, 8, 32, 128),
                   ("q8_0", 8, 32, 8, 32, M), ("q8_1", 9, 32, 8, 32, 1), ("q2_k", 10, 256, 4, 16, M), ("q3_k", 11, 256, 4, 16, M),
                   ("q4_k", 12, 256, 4, 32, M), ("q5_k", 13, 256, 8, 32, M), ("q6_k", 14, 256, 8, 16, 128), ("q8_k", 15, 256, 8, 32, 1)]
    for t in ("f32", "f16", "bf16"):
        assert PA.AffineFormatSpec.for_dtype(t) is None
    assert PA.minimum_batch("q6_k") == 128 and PA.minimum_batch("f16") is None


def test_marlin_shape_filter_matches_available_tiles():   # packed_affine.rs:949-963
    from mistralrs_b200 import packed_affine as PA
    assert PA.supports_marlin_shape(64, 128) and PA.supports_marlin_shape(128, 64) and PA.supports_marlin_shape(256, 64)
    for n, k in ((64, 64), (64, 96), (96, 128), (192, 64)):
        assert not PA.supports_marlin_shape(n, k)
    assert PA.padded_n_for_shape(96, 256) == 128 and PA.padded_n_for_shape(64, 64) == 128
    assert PA.padded_n_for_shape(129, 128) == 192 and PA.padded_n_for_shape(64, 32) is None


def test_plan_sizes():   # packed_affine.rs:94-135
    from mistralrs_b200 import packed_affine as PA
    plan = PA.PackedAffinePlan.new("q4_k", 96, 256)
    assert (plan.n, plan.padded_n, plan.k) == (96, 128, 256)
    assert plan.payload_bytes == 128 * 256 // 2 and plan.metadata_values == 256 // 32 * 128 and plan.metadata_bytes == plan.metadata_values * 2
    assert plan.workspace_len == 128 // 64 * 16
    assert plan.total_bytes == plan.payload_bytes + 2 * plan.metadata_bytes + plan.workspace_len * 4
    assert PA.PackedAffinePlan.new("q4_k", 64, 128) is None            # K must hold whole source blocks
    assert PA.PackedAffinePlan.new("q6_k", 64, 256).metadata_values == 256 // 16 * 64
    assert PA.PackedAffinePlan.new("f16", 64, 256) is None and PA.PackedAffinePlan.new("q4_0", 0, 256) is None


def test_dispatch_rule(monkeypatch):   # packed_affine.rs `dispatch_switches_to_packed_at_minimum_batch`, `format_specific_minimum_batches_are_enforced`
    import torch
    from mistralrs_b200 import packed_affine as PA
    monkeypatch.delenv(PA.BACKEND_ENV, raising=False)
    assert not PA.enabled() and not PA.should_dispatch("q4_k", (128, 256), 64, torch.bfloat16, "cuda")     # off by default
    monkeypatch.setenv(PA.BACKEND_ENV, "off")
    assert not PA.enabled()
    for v in ("on", "auto"):
        monkeypatch.setenv(PA.BACKEND_ENV, v)
        assert PA.enabled()
    M = PA.GGUF_AFFINE_MIN_BATCH
    assert not PA.should_dispatch("q4_k", (128, 256), M - 1, torch.bfloat16, "cuda")
    assert PA.should_dispatch("q4_k", (128, 256), M, torch.bfloat16, "cuda")
    for t, mb in (("q5_0", 16), ("q5_1", 128), ("q6_k", 128)):
        assert not PA.should_dispatch(t, (128, 256), mb - 1, torch.bfloat16, "cuda")
        assert PA.should_dispatch(t, (128, 256), mb, torch.float16, "cuda")
    assert PA.should_dispatch("q8_k", (128, 256), 1, torch.bfloat16, "cuda")                   # affine-only formats from batch 1
    assert not PA.should_dispatch("q4_k", (128, 256), 64, torch.float32, "cuda")               # 16-bit activations only
    assert not PA.should_dispatch("q4_k", (128, 256), 64, torch.bfloat16, "cpu")
    assert not PA.should_dispatch("q4_k", (128, 96), 64, torch.bfloat16, "cuda")               # `unsupported_k_tile_uses_canonical_dispatch`
    assert PA.should_dispatch("q4_0", (96, 256), 64, torch.bfloat16, "cuda")                   # `unaligned_width_uses_padded_packed_dispatch`
