"""Pins the CPU oracle (no GPU needed):
  * block decoders vs gguf-py 0.19.0 (tests/golden/gguf_dequant.npz, make_gguf_golden.py) — bit-exact;
  * Q8_1 quantiser, MMVQ arithmetic, fused GLU, RoPE, add_rms_norm, KV-cache scatter and paged
    attention vs OUTPUTS OF THE UNMODIFIED REFERENCE KERNELS compiled from /root/reference and
    run on a B200 (tests/golden/ref_golden.npz, make_ref_golden.py)."""
import os

import numpy as np
import pytest

import oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TYPES = ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0", "q2_k", "q3_k", "q4_k", "q5_k", "q6_k"]


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(G, "ref_golden.npz"))


def bf16_ulp(x):
    x = np.abs(np.asarray(x, dtype=np.float32)).clip(1e-30)
    return np.exp2(np.floor(np.log2(x)) - 7)


@pytest.mark.parametrize("t", TYPES)
def test_block_decoders_match_gguf_py(t):
    z = np.load(os.path.join(G, "gguf_dequant.npz"))
    got = oracle.dequantize(t, z[f"{t}_blocks"])
    assert np.array_equal(got, z[f"{t}_deq"])
    try:  # live cross-check when the package is importable
        from gguf import GGMLQuantizationType as T, quants
        live = quants.dequantize(z[f"{t}_blocks"], getattr(T, t.upper())).reshape(-1).astype(np.float32)
        assert np.array_equal(got, live)
    except ImportError:
        pass


def test_q8_1_quantiser_vs_reference_kernel(ref):
    want = ref["q8_1_bytes"].reshape(-1, 36)
    got, _ = oracle.quantize_q8_1(ref["mmvq_x"], 1024)
    got = got.reshape(-1, 36)
    # d and the butterfly sum (half2 header): bit-exact
    assert np.array_equal(got[:, :4], want[:, :4])
    dq = np.abs(got[:, 4:].view(np.int8).astype(int) - want[:, 4:].view(np.int8).astype(int))
    # the reference divides with --use_fast_math: an exact .5 tie may round the other way
    assert dq.max() <= 1 and (dq != 0).mean() < 1e-3


@pytest.mark.parametrize("t", TYPES)
def test_mmvq_arithmetic_vs_reference_kernel(ref, t):
    K, N, B = 1024, 24, 2
    want = ref[f"mmvq_{t}_y"]
    got = oracle.mmvq_q8_1(t, ref[f"mmvq_{t}_w"], ref["q8_1_bytes"], K, N, K // 32, B)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 3e-6 * scale, (t, np.abs(got - want).max() / scale)
    # fused GLU output (bf16): act(bf16(gate)) * bf16(up)
    up = oracle.mmvq_q8_1(t, ref[f"mmvq_{t}_up"], ref["q8_1_bytes"], K, N, K // 32, B)
    mine = oracle.fused_glu(oracle.round_dtype(got.astype(np.float32), "bf16"), oracle.round_dtype(up.astype(np.float32), "bf16"), 0, "bf16")
    w = ref[f"mmvq_{t}_glu"]
    assert (np.abs(mine - w) <= 2 * bf16_ulp(w) + 1e-30).all()
    assert (mine == w).mean() > 0.9


def test_fused_glu_vs_reference_kernel(ref):
    for act in range(5):
        got = oracle.fused_glu(ref["glu_a"], ref["glu_b"], act, "bf16")
        want = ref[f"glu_out_{act}"]
        # the reference is built with --use_fast_math (approximate exp / div / tanh): allow two
        # ulps of the product, measured at the magnitude of |a*b| where the result cancels to ~0
        tol = 2 * bf16_ulp(np.maximum(np.abs(want), 1e-2 * np.abs(ref["glu_a"] * ref["glu_b"])))
        assert (np.abs(got - want) <= tol + 1e-30).all(), act
        assert (got == want).mean() > 0.9, act


def test_rotary_vs_reference_kernel(ref):
    for neox in (1, 0):
        q, k = oracle.rotary(ref["rope_q"], ref["rope_k"], ref["rope_cos"], ref["rope_sin"], ref["rope_pos"], bool(neox),
                             128, 64, 4, 2, "bf16")
        assert np.array_equal(q, ref[f"rope_q_out_{neox}"]) and np.array_equal(k, ref[f"rope_k_out_{neox}"])


def test_add_rms_norm_vs_reference_kernel(ref):
    s, n = oracle.add_rms_norm(ref["rms_x"], ref["rms_res"], ref["rms_w"], 1e-5, "bf16")
    assert np.array_equal(s, ref["rms_sum"])
    assert (np.abs(n - ref["rms_norm"]) <= bf16_ulp(ref["rms_norm"])).all() and (n == ref["rms_norm"]).mean() > 0.98


def _u16(a):
    return oracle.round_dtype(a, "bf16").view(np.uint32).__rshift__(16).astype(np.uint16)


def test_cache_scatter_vs_reference_kernels(ref):
    KVH, D, BS, NB = 2, 128, 16, 9
    k, v = _u16(ref["pa_k"]), _u16(ref["pa_v"])
    for layout, names in ((0, ("cache_k_vllm", "cache_v_vllm")), (1, ("cache_k_hnd", "cache_v_hnd"))):
        kc = np.zeros(NB * KVH * D * BS, dtype=np.uint16); vc = np.zeros_like(kc)
        oracle.reshape_and_cache(k, v, kc, vc, ref["pa_slots"], KVH, D, BS, 8, layout)
        assert np.array_equal(kc, ref[names[0]]) and np.array_equal(vc, ref[names[1]])


def test_paged_attention_vs_reference_kernels(ref):
    KVH, D, BS = 2, 128, 16
    scale = 1.0 / np.sqrt(D)
    for layout, kn, vn, on in ((0, "cache_k_vllm", "cache_v_vllm", "pa_out_v1"), (1, "cache_k_hnd", "cache_v_hnd", "fi_out")):
        got = oracle.paged_attention(ref["pa_q"], ref[kn], ref[vn], ref["pa_tables"], ref["pa_ctx"], KVH, D, BS, scale, layout, "bf16")
        want = ref[on]
        assert np.abs(got - want).max() <= 2.5 * 2.0 ** -8 * np.abs(want).max(), (on, np.abs(got - want).max())
