"""Decode GEMV parity: CUDA (through the reference-shaped C ABI) vs the CPU oracle's Q8_1
arithmetic on the same seeded inputs.  Tolerances: the oracle value is the infinitely-precise
result of the reference's integer-dot arithmetic; the kernel may differ by f32 accumulation
order (<= 2e-6 of the output scale) plus one rounding to the output dtype."""
import numpy as np
import pytest
import torch

import oracle
from mistralrs_b200 import quant
from util import ALL_TYPES, make_acts, make_weight, to_dev, ulp_report

pytestmark = pytest.mark.gpu


def _oracle_gemv(dtype, wb, x_np, K, N):
    xq, stride = oracle.quantize_q8_1(x_np)
    return oracle.mmvq_q8_1(dtype, wb, xq, K, N, stride, x_np.shape[0])


@pytest.mark.parametrize("dtype", ALL_TYPES)
@pytest.mark.parametrize("dt", ["bf16", "f32", "f16"])
def test_plain_batch1(cuda, dtype, dt):
    K, N = 2048, 70  # ragged N: not a multiple of the 16-row pass
    wb = make_weight(dtype, N, K, 1)
    x = make_acts(1, K, 2, dt)
    w = quant.QTensor(to_dev(wb.reshape(-1), cuda), dtype, (N, K))
    y = quant.plain(w, to_dev(x, cuda, dt)).float().cpu().numpy()
    ref = _oracle_gemv(dtype, wb, x, K, N)
    rel, frac = ulp_report(y, ref, dt)
    assert frac >= 0.995 and rel < 4e-3, (dtype, dt, rel, frac)


@pytest.mark.parametrize("dtype", ["q4_k", "q6_k", "q8_0", "q5_k", "q2_k", "q3_k", "q4_0", "q5_1"])
@pytest.mark.parametrize("batch", [2, 3, 4, 5, 8])
def test_plain_batched(cuda, dtype, batch):
    K, N = 1536, 33  # K = 1.5 segments (ragged K), odd N
    wb = make_weight(dtype, N, K, 3)
    x = make_acts(batch, K, 4, "f32")
    w = quant.QTensor(to_dev(wb.reshape(-1), cuda), dtype, (N, K))
    y = quant.plain(w, to_dev(x, cuda, "f32")).cpu().numpy()
    ref = _oracle_gemv(dtype, wb, x, K, N)
    rel, frac = ulp_report(y, ref, "f32")
    assert rel < 5e-6, (dtype, batch, rel, frac)


@pytest.mark.parametrize("dtype", ["q4_k", "q6_k", "q8_0", "q5_0"])
def test_llama_shapes_f32(cuda, dtype):
    # the real Llama-3-8B projection shapes, f32 output so the check is tight
    for (N, K) in [(4096, 4096), (1024, 4096), (256, 14336)]:
        wb = make_weight(dtype, N, K, 5)
        x = make_acts(1, K, 6, "f32")
        w = quant.QTensor(to_dev(wb.reshape(-1), cuda), dtype, (N, K))
        y = quant.plain(w, to_dev(x, cuda, "f32")).cpu().numpy()
        ref = _oracle_gemv(dtype, wb, x, K, N)
        rel, _ = ulp_report(y, ref, "f32")
        assert rel < 5e-6, (dtype, N, K, rel)


@pytest.mark.parametrize("dtype", ALL_TYPES)
@pytest.mark.parametrize("batch", [1, 3])
def test_fused_qkv_matches_plain(cuda, dtype, batch):
    # the reference's own self-consistency test (fast_mmq.rs:1583-1703): fused == unfused
    K = 1024
    nq, nk, nv = 96, 24, 24
    ws = [quant.QTensor(to_dev(make_weight(dtype, n, K, 10 + i).reshape(-1), cuda), dtype, (n, K))
          for i, n in enumerate((nq, nk, nv))]
    x = to_dev(make_acts(batch, K, 7, "bf16"), cuda, "bf16")
    q, k, v = quant.fused_qkv(ws[0], ws[1], ws[2], x)
    for got, w in zip((q, k, v), ws):
        assert torch.equal(got, quant.plain(w, x))


@pytest.mark.parametrize("dtype", ["q4_k", "q6_k", "q8_0", "q4_1", "q3_k"])
@pytest.mark.parametrize("act", [0, 1, 2, 3, 4])
def test_fused_glu(cuda, dtype, act):
    K, N = 1024, 41
    g = quant.QTensor(to_dev(make_weight(dtype, N, K, 20).reshape(-1), cuda), dtype, (N, K))
    u = quant.QTensor(to_dev(make_weight(dtype, N, K, 21).reshape(-1), cuda), dtype, (N, K))
    x = to_dev(make_acts(2, K, 8, "bf16"), cuda, "bf16")
    got = quant.fused_glu(g, u, x, quant.GluActivationType(act)).float().cpu().numpy()
    gate = quant.plain(g, x).float().cpu().numpy()
    up = quant.plain(u, x).float().cpu().numpy()
    want = oracle.fused_glu(gate, up, act, "bf16")
    # activation uses the GPU's fast exp/div/tanh (the reference builds with --use_fast_math): one
    # bf16 ulp on the activated value, plus tanh.approx's 2^-10.99 absolute error carried through
    # 0.5*gate*(1+tanh)*up where 1+tanh cancels
    tol = 2.0 ** -7 * np.abs(want) + 2.0 ** -10 * np.abs(gate * up) + 1e-6
    assert (np.abs(got - want) <= tol).all(), (dtype, act, np.abs(got - want).max())


def test_quantize_q8_1_vs_oracle(cuda):
    # bit-level: d, sum and the int8 quants (approximate GPU division may move a tie by 1)
    for dt in ("bf16", "f16", "f32"):
        x = make_acts(3, 1000, 9, dt)  # ragged K -> zero padding to 1024
        got = quant.quantize_q8_1(to_dev(x, cuda, dt)).cpu().numpy().reshape(3, -1, 36)
        want, stride = oracle.quantize_q8_1(x)
        want = want.reshape(3, -1, 36)
        assert got.shape == want.shape
        dq = np.abs(got[..., 4:].view(np.int8).astype(int) - want[..., 4:].view(np.int8).astype(int))
        assert dq.max() <= 1 and (dq != 0).mean() < 2e-3, (dt, dq.max(), (dq != 0).mean())
        dd = np.abs(got[..., :4].view(np.float16).astype(np.float32) - want[..., :4].view(np.float16).astype(np.float32))
        assert (dd <= 2e-3 * np.abs(want[..., :4].view(np.float16).astype(np.float32)) + 1e-7).all()


def test_fused_prologue_matches_two_step(cuda):
    # mrs_mmvq_fused (RMSNorm + Q8_1 + GEMV + residual in one launch) == unfused chain
    K, N = 2048, 64
    for dtype in ("q4_k", "q6_k"):
        w = quant.QTensor(to_dev(make_weight(dtype, N, K, 30).reshape(-1), cuda), dtype, (N, K))
        x = to_dev(make_acts(1, K, 31, "bf16"), cuda, "bf16")
        nw = to_dev(1.0 + 0.1 * make_acts(1, K, 32, "bf16")[0], cuda, "bf16")
        res = to_dev(make_acts(1, N, 33, "bf16"), cuda, "bf16")
        normed = oracle.rms_norm(x.float().cpu().numpy(), nw.float().cpu().numpy(), 1e-5, "bf16")
        want = quant.plain(w, to_dev(normed, cuda, "bf16"))
        want = (want.float() + res.float()).to(torch.bfloat16)
        got = quant.mmvq_fused(w, x, norm_w=nw, eps=1e-5, residual=res)
        diff = (got.float() - want.float()).abs().max().item()
        assert diff <= 2.0 ** -6 * want.float().abs().max().item(), (dtype, diff)


@pytest.mark.parametrize("dtype", ["q4_k", "q6_k"])
def test_long_segment_variant_agrees(cuda, dtype):
    # lm_head-sized streams take K segments twice as long (UPL x2): same arithmetic in a different
    # consumption order of the activation image -> results must agree bit for bit with the short
    # segments, with pre-quantised activations and with the fused RMSNorm prologue
    import ctypes
    from mistralrs_b200 import lib
    K, N = 4096, 700
    w = quant.QTensor(to_dev(make_weight(dtype, N, K, 40).reshape(-1), cuda), dtype, (N, K))
    nw = to_dev(1.0 + 0.1 * make_acts(1, K, 42, "bf16")[0], cuda, "bf16")
    x = to_dev(make_acts(1, K, 44, "bf16"), cuda, "bf16")
    outs = []
    try:
        for flags in (8 | (1 << 8), 1 << 8):      # long variant off / on from 1 MiB of weights
            lib().mrs_set_mmvq_flags(ctypes.c_int(flags))
            outs.append([quant.plain(w, x), quant.mmvq_fused(w, x, norm_w=nw, eps=1e-5)])
    finally:
        lib().mrs_set_mmvq_flags(ctypes.c_int(128 << 8))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("pair", [("q4_k", "q6_k"), ("q5_k", "q6_k"), ("q4_k", "q5_k"), ("q4_0", "q8_0")])
@pytest.mark.parametrize("batch", [1, 3])
def test_fused_qkv_mixed_matches_separate_launches(cuda, pair, batch):
    # two-type QKV grid (attn_v in its own ggml type) == mode-2 q||k launch + plain v launch, bit for
    # bit; unsupported pairs / batch > 1 take the two-launch fallback inside the same entry point
    tq, tv = pair
    K, nq, nk, nv = 2048, 320, 64, 64
    wq, wk = (quant.QTensor(to_dev(make_weight(tq, n, K, 50 + i).reshape(-1), cuda), tq, (n, K)) for i, n in enumerate((nq, nk)))
    wv = quant.QTensor(to_dev(make_weight(tv, nv, K, 52).reshape(-1), cuda), tv, (nv, K))
    x = to_dev(make_acts(batch, K, 53, "bf16"), cuda, "bf16")
    nw = to_dev(1.0 + 0.1 * make_acts(1, K, 54, "bf16")[0], cuda, "bf16")
    q, k, v = quant.fused_qkv_mixed(wq, wk, wv, x, norm_w=nw, eps=1e-5)
    q2, k2 = quant.mmvq_fused(wq, x, mode=2, w1=wk, norm_w=nw, eps=1e-5)[:2]
    v2 = quant.mmvq_fused(wv, x, norm_w=nw, eps=1e-5)
    assert torch.equal(q, q2) and torch.equal(k, k2) and torch.equal(v, v2)


def test_argument_errors(cuda):
    w = quant.QTensor(torch.zeros(144 * 4, dtype=torch.uint8, device=cuda), "q4_k", (4, 256))
    with pytest.raises(ValueError, match="batch size 9"):
        quant.plain(w, torch.zeros(9, 256, device=cuda, dtype=torch.bfloat16))
    with pytest.raises(ValueError, match="shape mismatch"):
        quant.plain(w, torch.zeros(1, 512, device=cuda, dtype=torch.bfloat16))
