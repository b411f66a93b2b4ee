"""Fused elementwise / norm / rope / cache ops: CUDA (C ABI) vs the CPU oracle."""
import numpy as np
import pytest
import torch

import oracle
from mistralrs_b200 import ops, paged_attn
from util import TORCH_DT, make_acts, to_dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("act", [0, 1, 2, 3, 4])
def test_fused_glu(cuda, dt, act):
    a = oracle.round_dtype(make_acts(3, 1000, 1, dt) * 3, dt)
    b = make_acts(3, 1000, 2, dt)
    got = ops.fused_glu(to_dev(a, cuda, dt), to_dev(b, cuda, dt), act).float().cpu().numpy()
    want = oracle.fused_glu(a, b, act, dt)
    eps = {"bf16": 2.0 ** -6, "f16": 2.0 ** -9, "f32": 2e-5}[dt]  # fast exp/div may flip the activation by 1 ulp; product rounds again
    tol = eps * np.abs(want) + 1e-6
    if act == 1:
        # tanh-form GELU: the reference builds with --use_fast_math, so tanhf is tanh.approx.f32
        # (abs error ~2^-11, tests/golden/ref_golden.npz pins the kernel against the reference's
        # own output); it reaches the result through 0.5*a*(1+tanh)*b where 1+tanh cancels
        tol = tol + 2.0 ** -10 * np.abs(a * b)
    assert (np.abs(got - want) <= tol).all()
    x = np.concatenate([a, b], axis=1)
    got2 = ops.fused_split_glu(to_dev(x, cuda, dt), act).float().cpu().numpy()
    assert np.array_equal(got2, got)


@pytest.mark.parametrize("dt", ["bf16", "f16", "f32"])
def test_rms_norm_and_add(cuda, dt):
    rows, cols = 5, 4096
    x = make_acts(rows, cols, 3, dt)
    r = make_acts(rows, cols, 4, dt)
    w = oracle.round_dtype(1.0 + 0.1 * make_acts(1, cols, 5, "f32")[0], dt)
    got = ops.rms_norm(to_dev(x, cuda, dt), to_dev(w, cuda, dt), 1e-5).float().cpu().numpy()
    want = oracle.rms_norm(x, w, 1e-5, dt)
    # rsqrtf / summation order may move a value across one rounding boundary: <= 1 ulp, mostly exact
    ulp = {"bf16": 2.0 ** -7, "f16": 2.0 ** -10, "f32": 1e-6}[dt]
    assert (np.abs(got - want) <= ulp * np.abs(want) * 1.01 + 1e-7).all()
    if dt != "f32":
        assert (got == want).mean() > 0.98
    s, n = ops.add_rms_norm(to_dev(x, cuda, dt), to_dev(r, cuda, dt), to_dev(w, cuda, dt), 1e-5)
    ws, wn = oracle.add_rms_norm(x, r, w, 1e-5, dt)
    assert np.array_equal(s.float().cpu().numpy(), ws)  # the residual sum is bit-exact
    assert (np.abs(n.float().cpu().numpy() - wn) <= ulp * np.abs(wn) * 1.01 + 1e-7).all()


@pytest.mark.parametrize("dt", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("neox", [True, False])
def test_rotary(cuda, dt, neox):
    T, H, KVH, D = 7, 8, 2, 128
    cos, sin = oracle.llama3_rope_table(64, D, 500000.0, dict(factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                                             original_max_position_embeddings=8192))
    cos, sin = oracle.round_dtype(cos, dt), oracle.round_dtype(sin, dt)
    q = make_acts(T, H * D, 6, dt)
    k = make_acts(T, KVH * D, 7, dt)
    pos = np.array([0, 1, 5, 9, 33, 63, 2], dtype=np.uint32)
    wq, wk = oracle.rotary(q, k, cos, sin, pos, neox, D, D // 2, H, KVH, dt)
    gq = to_dev(q, cuda, dt).reshape(T, H, D).clone()
    gk = to_dev(k, cuda, dt).reshape(T, KVH, D).clone()
    ops.apply_rotary_qk(gq, gk, to_dev(cos, cuda, dt), to_dev(sin, cuda, dt),
                        torch.from_numpy(pos.astype(np.int32)).to(cuda), is_neox=neox)
    if dt == "f32":  # FMA contraction may differ in the last bit
        assert np.allclose(gq.cpu().numpy().reshape(T, -1), wq, rtol=0, atol=1e-6)
        assert np.allclose(gk.cpu().numpy().reshape(T, -1), wk, rtol=0, atol=1e-6)
    else:            # per-op rounding in the dtype: bit-exact
        assert np.array_equal(gq.float().cpu().numpy().reshape(T, -1), wq)
        assert np.array_equal(gk.float().cpu().numpy().reshape(T, -1), wk)
    # sequential-position variant (rotary_embedding): row t uses table row t
    wq2, wk2 = oracle.rotary(q, k, cos, sin, None, neox, D, D // 2, H, KVH, dt)
    gq2 = to_dev(q, cuda, dt).reshape(T, H, D).clone()
    gk2 = to_dev(k, cuda, dt).reshape(T, KVH, D).clone()
    ops.apply_rotary_qk(gq2, gk2, to_dev(cos, cuda, dt), to_dev(sin, cuda, dt), None, is_neox=neox)
    if dt != "f32":
        assert np.array_equal(gq2.float().cpu().numpy().reshape(T, -1), wq2)


@pytest.mark.parametrize("layout", ["vllm", "hnd"])
@pytest.mark.parametrize("dt", ["bf16", "f16"])
def test_reshape_and_cache_bit_exact(cuda, layout, dt):
    T, KVH, D, BS, NB, x = 9, 4, 128, 16, 12, 8
    rng = np.random.default_rng(11)
    key = oracle.round_dtype(rng.standard_normal((T, KVH * D)).astype(np.float32), dt)
    val = oracle.round_dtype(rng.standard_normal((T, KVH * D)).astype(np.float32), dt)
    slots = np.array([35, 0, -1, 17, 191, 64, -1, 100, 3], dtype=np.int64)  # -1 = padding token
    tk, tv = to_dev(key, cuda, dt).reshape(T, KVH, D), to_dev(val, cuda, dt).reshape(T, KVH, D)
    ku, vu = tk.view(torch.int16).cpu().numpy().view(np.uint16).reshape(T, -1), tv.view(torch.int16).cpu().numpy().view(np.uint16).reshape(T, -1)
    n = NB * KVH * D * BS
    okc, ovc = np.zeros(n, dtype=np.uint16), np.zeros(n, dtype=np.uint16)
    oracle.reshape_and_cache(ku, vu, okc, ovc, slots, KVH, D, BS, x, 0 if layout == "vllm" else 1)
    sm = torch.from_numpy(slots).to(cuda)
    if layout == "vllm":
        kc = torch.zeros(NB, KVH, D // x, BS, x, dtype=TORCH_DT[dt], device=cuda)
        vc = torch.zeros(NB, KVH, D, BS, dtype=TORCH_DT[dt], device=cuda)
        paged_attn.reshape_and_cache(tk, tv, None, None, kc, vc, sm)
    else:
        kc = torch.zeros(NB, KVH, BS, D, dtype=TORCH_DT[dt], device=cuda)
        vc = torch.zeros_like(kc)
        paged_attn.reshape_and_cache_flashinfer(tk, tv, kc, vc, sm)
    assert np.array_equal(kc.view(torch.int16).cpu().numpy().view(np.uint16).reshape(-1), okc)
    assert np.array_equal(vc.view(torch.int16).cpu().numpy().view(np.uint16).reshape(-1), ovc)


def test_cache_write_gather_round_trip(cuda):
    # the reference's own round-trip test shape (backend/flashinfer.rs:742-781)
    KVH, D, BS, NB = 2, 64, 8, 10
    lens = [5, 13]
    tables = torch.tensor([[3, 0, 0], [7, 2, 0]], dtype=torch.int32, device=cuda)
    T = sum(lens)
    key = ((torch.arange(T * KVH * D, device=cuda) * 37 % 251).float() * 0.071).sin().to(torch.bfloat16).reshape(T, KVH, D)
    val = -key
    slots = []
    for s, L in enumerate(lens):
        slots += [int(tables[s, i // BS]) * BS + i % BS for i in range(L)]
    kc = torch.zeros(NB, KVH, BS, D, dtype=torch.bfloat16, device=cuda)
    vc = torch.zeros_like(kc)
    paged_attn.reshape_and_cache_flashinfer(key, val, kc, vc, torch.tensor(slots, dtype=torch.int64, device=cuda))
    cu = torch.tensor([0, lens[0], T], dtype=torch.int32, device=cuda)
    k_out, v_out = paged_attn.gather_kv_cache_flashinfer(kc, vc, tables, cu, T, torch.bfloat16)
    assert torch.equal(k_out, key) and torch.equal(v_out, val)
    # copy-on-write block copy
    keep = paged_attn.copy_blocks([kc], [vc], [(3, 9), (7, 8)])
    assert torch.equal(kc[9], kc[3]) and torch.equal(vc[8], vc[7])
    del keep


@pytest.mark.parametrize("dt", ["bf16", "f16", "f32"])
def test_kv_scale_update(cuda, dt):
    # update_kv_scales_*: scale = max(old, absmax/240); exact up to the reference's fast-math division
    rng = np.random.default_rng(21)
    for n in (1, 7, 4096 + 3, 300_001):            # ragged tails, unaligned heads (offset slice below)
        k = oracle.round_dtype((rng.standard_normal(n + 1) * 3).astype(np.float32), dt)
        v = oracle.round_dtype((rng.standard_normal(n + 1) * 0.02).astype(np.float32), dt)
        tk, tv = to_dev(k, cuda, dt)[1:], to_dev(v, cuda, dt)[1:]      # element-offset views: not 16-byte aligned
        ks = torch.tensor([0.001], dtype=torch.float32, device=cuda)   # below the candidate: replaced
        vs = torch.tensor([5.0], dtype=torch.float32, device=cuda)     # above the candidate: kept
        paged_attn.kv_scale_update(tk, tv, ks, vs)
        want_k = max(np.float32(0.001), np.float32(np.abs(k[1:]).max()) / np.float32(240.0))
        assert abs(ks.item() - want_k) <= 2.0 ** -22 * want_k
        assert vs.item() == 5.0
    z = torch.zeros(64, dtype=TORCH_DT[dt], device=cuda)
    ks = torch.zeros(1, dtype=torch.float32, device=cuda)
    paged_attn.kv_scale_update(z, z, ks, ks)                            # all-zero input leaves the scale alone
    assert ks.item() == 0.0


def test_swap_blocks(cuda):
    # device<->device and device<->pinned host block copies by (src, dst) mapping; untouched blocks stay
    NB, KVH, BS, D = 12, 2, 16, 64
    src = torch.randn(NB, KVH, BS, D, device=cuda).to(torch.bfloat16)
    dst = torch.zeros_like(src)
    mapping = {3: 0, 7: 11, 1: 5}
    paged_attn.swap_blocks(src, dst, mapping)
    torch.cuda.synchronize()
    for s, d in mapping.items():
        assert torch.equal(dst[d], src[s])
    untouched = [i for i in range(NB) if i not in mapping.values()]
    assert not dst[untouched].any()
    host = torch.zeros(NB, KVH, BS, D, dtype=torch.bfloat16).pin_memory()
    paged_attn.swap_blocks(src, host, [(2, 9), (4, 4)])                  # swap out
    torch.cuda.synchronize()
    assert torch.equal(host[9], src[2].cpu()) and torch.equal(host[4], src[4].cpu())
    back = torch.zeros_like(src)
    paged_attn.swap_blocks(host, back, [(9, 1)])                         # swap in
    torch.cuda.synchronize()
    assert torch.equal(back[1], src[2])
    with pytest.raises(IndexError):
        paged_attn.swap_blocks(src, dst, [(NB, 0)])


@pytest.mark.parametrize("dt", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("D", [64, 128, 80])
def test_rms_norm_strided_4d(cuda, dt, D):
    # per-head RMSNorm read through a transposed ([B,S,H,D] -> [B,H,S,D]) and a sliced (q of a fused
    # qkv row) view; the output is contiguous [B,H,S,D]
    B, H, S = 2, 3, 5
    w = oracle.round_dtype(1.0 + 0.1 * make_acts(1, D, 6, "f32")[0], dt)
    ulp = {"bf16": 2.0 ** -7, "f16": 2.0 ** -10, "f32": 1e-6}[dt]
    base = make_acts(B * S, H * D * 3, 7, dt).reshape(B, S, 3 * H * D)
    t = to_dev(base, cuda, dt)
    view = t[:, :, H * D:2 * H * D].reshape(B, S, H, D).transpose(1, 2)       # strides (S*3HD, D, 3HD, 1)
    assert not view.is_contiguous()
    got = ops.rms_norm_strided_4d(view, to_dev(w, cuda, dt), 1e-6).float().cpu().numpy()
    rows = base[:, :, H * D:2 * H * D].reshape(B, S, H, D).transpose(0, 2, 1, 3).reshape(-1, D)
    want = oracle.rms_norm(rows, w, 1e-6, dt).reshape(B, H, S, D)
    assert got.shape == want.shape
    assert (np.abs(got - want) <= ulp * np.abs(want) * 1.01 + 1e-7).all()


@pytest.mark.parametrize("dtype", ["q6_k", "q8_0", "q4_k", "q5_k", "q3_k", "q2_k", "q4_0", "q4_1", "q5_0", "q5_1"])
def test_embedding_gather_matches_dequantized_rows(cuda, dtype):
    # the reference's `assert_embedding_matches_dequantized_gather` (gguf/mod.rs:813-845: ids [2,3],
    # gather == dequantize().index_select(), <= 1e-6) for every block type, f32 and bf16 outputs
    from mistralrs_b200 import quant
    from util import make_weight
    vocab, cols = 40, 512
    wb = make_weight(dtype, vocab, cols, 77)
    w = quant.QTensor(to_dev(wb.reshape(-1), cuda), dtype, (vocab, cols))
    ids = torch.tensor([[3, 0, 39], [3, 17, 8]], dtype=torch.int32, device=cuda)
    full = oracle.dequantize(dtype, wb).reshape(vocab, cols)
    want = full[ids.cpu().numpy().reshape(-1)].reshape(2, 3, cols)
    got = ops.embedding_gather(w, ids, torch.float32).cpu().numpy()
    assert got.shape == (2, 3, cols)
    assert np.abs(got - want).max() <= 1e-6 * max(1.0, np.abs(want).max())
    got16 = ops.embedding_gather(w, ids, torch.bfloat16).float().cpu().numpy()
    assert (np.abs(got16 - oracle.round_dtype(want, "bf16")) <= 2.0 ** -8 * np.abs(want) + 1e-9).all()
