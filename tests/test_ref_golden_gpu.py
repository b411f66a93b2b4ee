"""Product kernels vs the stored outputs of the UNMODIFIED reference kernels
(tests/golden/ref_golden.npz) on the same inputs, through the C ABI."""
import ctypes
import os

import numpy as np
import pytest
import torch

from mistralrs_b200 import kv_index, lib, ops, paged_attn, quant
from util import to_dev

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TYPES = ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0", "q2_k", "q3_k", "q4_k", "q5_k", "q6_k"]


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(G, "ref_golden.npz"))


def bf16_ulp(x):
    x = np.abs(np.asarray(x, dtype=np.float32)).clip(1e-30)
    return np.exp2(np.floor(np.log2(x)) - 7)


def test_q8_1_bytes_identical(cuda, ref):
    got = quant.quantize_q8_1(to_dev(ref["mmvq_x"], cuda, "bf16"), 1024).cpu().numpy()
    assert np.array_equal(got, ref["q8_1_bytes"])   # same approximate divisions -> bit-identical


@pytest.mark.parametrize("t", TYPES)
def test_mmvq_plain_and_glu(cuda, ref, t):
    K, N, B = 1024, 24, 2
    w = quant.QTensor(to_dev(ref[f"mmvq_{t}_w"].reshape(-1), cuda), t, (N, K))
    up = quant.QTensor(to_dev(ref[f"mmvq_{t}_up"].reshape(-1), cuda), t, (N, K))
    y = quant.plain(w, to_dev(ref["mmvq_x"], cuda, "f32") if False else to_dev(ref["mmvq_x"], cuda, "bf16").float()).cpu().numpy()
    want = ref[f"mmvq_{t}_y"]
    assert np.abs(y - want).max() <= 3e-6 * np.abs(want).max()
    g = quant.fused_glu(w, up, to_dev(ref["mmvq_x"], cuda, "bf16"), quant.GluActivationType.Silu).float().cpu().numpy()
    wg = ref[f"mmvq_{t}_glu"]
    assert (np.abs(g - wg) <= 2 * bf16_ulp(wg) + 1e-30).all() and (g == wg).mean() > 0.95


def test_fused_glu_elementwise(cuda, ref):
    for act in range(5):
        got = ops.fused_glu(to_dev(ref["glu_a"], cuda, "bf16"), to_dev(ref["glu_b"], cuda, "bf16"), act).float().cpu().numpy()
        want = ref[f"glu_out_{act}"]
        # same fast-math intrinsics as the reference build -> bit-identical (a stray ulp is tolerated)
        assert (np.abs(got - want) <= bf16_ulp(np.maximum(np.abs(want), 1e-2 * np.abs(ref["glu_a"] * ref["glu_b"])))).all(), act
        assert (got == want).mean() > 0.999, act


def test_rotary_bit_identical(cuda, ref):
    for neox in (1, 0):
        q = to_dev(ref["rope_q"], cuda, "bf16").reshape(5, 4, 128).clone()
        k = to_dev(ref["rope_k"], cuda, "bf16").reshape(5, 2, 128).clone()
        ops.apply_rotary_qk(q, k, to_dev(ref["rope_cos"], cuda, "bf16"), to_dev(ref["rope_sin"], cuda, "bf16"),
                            torch.from_numpy(ref["rope_pos"].astype(np.int32)).to(cuda), is_neox=bool(neox))
        assert np.array_equal(q.float().cpu().numpy().reshape(5, -1), ref[f"rope_q_out_{neox}"])
        assert np.array_equal(k.float().cpu().numpy().reshape(5, -1), ref[f"rope_k_out_{neox}"])


def test_add_rms_norm(cuda, ref):
    s, n = ops.add_rms_norm(to_dev(ref["rms_x"], cuda, "bf16"), to_dev(ref["rms_res"], cuda, "bf16"), to_dev(ref["rms_w"], cuda, "bf16"), 1e-5)
    assert np.array_equal(s.float().cpu().numpy(), ref["rms_sum"])
    n = n.float().cpu().numpy()
    assert (np.abs(n - ref["rms_norm"]) <= bf16_ulp(ref["rms_norm"])).all() and (n == ref["rms_norm"]).mean() > 0.98


def test_cache_and_attention(cuda, ref):
    KVH, D, BS, NB, H, S = 2, 128, 16, 9, 8, 2
    k = to_dev(ref["pa_k"], cuda, "bf16").reshape(-1, KVH, D)
    v = to_dev(ref["pa_v"], cuda, "bf16").reshape(-1, KVH, D)
    slots = torch.from_numpy(ref["pa_slots"]).to(cuda)
    kc = torch.zeros(NB, KVH, D // 8, BS, 8, dtype=torch.bfloat16, device=cuda)
    vc = torch.zeros(NB, KVH, D, BS, dtype=torch.bfloat16, device=cuda)
    paged_attn.reshape_and_cache(k, v, None, None, kc, vc, slots)
    u16 = lambda t: t.view(torch.int16).cpu().numpy().view(np.uint16).reshape(-1)
    assert np.array_equal(u16(kc), ref["cache_k_vllm"]) and np.array_equal(u16(vc), ref["cache_v_vllm"])
    kh = torch.zeros(NB, KVH, BS, D, dtype=torch.bfloat16, device=cuda)
    vh = torch.zeros_like(kh)
    paged_attn.reshape_and_cache_flashinfer(k, v, kh, vh, slots)
    assert np.array_equal(u16(kh), ref["cache_k_hnd"]) and np.array_equal(u16(vh), ref["cache_v_hnd"])
    q = to_dev(ref["pa_q"], cuda, "bf16")
    scale = 1.0 / np.sqrt(D)
    ctx = ref["pa_ctx"].tolist()
    o1 = paged_attn.paged_attention(q, None, None, kc, vc, to_dev(ref["pa_tables"], cuda), to_dev(ref["pa_ctx"], cuda), None, max(ctx), scale)
    want = ref["pa_out_v1"]
    assert np.abs(o1.float().cpu().numpy() - want).max() <= 2.5 * 2.0 ** -8 * np.abs(want).max()
    tables = [ref["pa_tables"][0][:3].tolist(), ref["pa_tables"][1][:5].tolist()]
    indptr, indices, last = kv_index.make_paged_kv_tensors(tables, ctx, BS, 8)
    req, tile, oind, chunk, mask = kv_index.make_paged_kv_decode_tensors(tables, ctx, BS, None, 2)
    d = lambda a: to_dev(np.ascontiguousarray(a), cuda)
    o2 = paged_attn.flashinfer_decode(q, kh, vh, d(indptr), d(indices), d(last), d(req), d(tile), d(oind), d(chunk), d(mask), scale)
    want = ref["fi_out"]
    assert np.abs(o2.float().cpu().numpy() - want).max() <= 2.5 * 2.0 ** -8 * np.abs(want).max()
