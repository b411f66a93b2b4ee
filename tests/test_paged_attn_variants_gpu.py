"""Decode attention over both cache layouts at the boundary's less common corners: f32 tensors
(`paged_attention_v{1,2}_f32`), the head sizes the reference admits besides 64/128/256 (80, 96, 112, 192;
512 on the HND path) and the FP8-E4M3 KV cache (cache_dtype 3) — against a float64 numpy restatement of
REF pagedattention.cuh:110-485 / flashinfer_decode.cu (softmax(scale * q.k) v over the block table; FP8
bytes dequantised as float(e4m3) * scale, flashinfer: scales folded into logits / output)."""
import numpy as np
import pytest
import torch

from mistralrs_b200 import kv_index, paged_attn

pytestmark = pytest.mark.gpu


def _ref(q, k, v, ctx, scale):
    S, H, D = q.shape
    KVH = k[0].shape[1]
    out = np.zeros((S, H, D))
    for s in range(S):
        kk, vv = k[s][:ctx[s]].astype(np.float64), v[s][:ctx[s]].astype(np.float64)
        for h in range(H):
            kvh = h // (H // KVH)
            sc = (kk[:, kvh] @ q[s, h].astype(np.float64)) * scale
            p = np.exp(sc - sc.max()); p /= p.sum()
            out[s, h] = p @ vv[:, kvh]
    return out


def _setup(cuda, dt, cache_dt, D, ctx, H=8, KVH=2, BS=16, seed=0):
    gen = torch.Generator(device="cpu").manual_seed(seed + D)
    S = len(ctx)
    max_blocks = -(-max(ctx) // BS)
    q = torch.randn(S, H, D, generator=gen).to(dt)
    ks = [torch.randn(c, KVH, D, generator=gen).to(dt) for c in ctx]
    vs = [torch.randn(c, KVH, D, generator=gen).to(dt) for c in ctx]
    k_scale = v_scale = 1.0
    if cache_dt == torch.float8_e4m3fn:
        k_scale, v_scale = 0.05, 0.04
        kq = [(t.float() / k_scale).to(torch.float8_e4m3fn) for t in ks]
        vq = [(t.float() / v_scale).to(torch.float8_e4m3fn) for t in vs]
        kd = [t.float().numpy() * k_scale for t in kq]; vd = [t.float().numpy() * v_scale for t in vq]
    else:
        kq, vq = ks, vs
        kd = [t.float().numpy() for t in ks]; vd = [t.float().numpy() for t in vs]
    NB = S * max_blocks + 1
    tables = [[1 + s * max_blocks + i for i in range(max_blocks)] for s in range(S)]
    return q, kq, vq, kd, vd, tables, NB, max_blocks, k_scale, v_scale


def _fill_vllm(kq, vq, tables, NB, KVH, D, BS, cache_dt, cuda):
    x = 16 // torch.empty(0, dtype=cache_dt).element_size()
    kc = torch.zeros(NB, KVH, D // x, BS, x, dtype=cache_dt if cache_dt != torch.float8_e4m3fn else torch.uint8)
    vc = torch.zeros(NB, KVH, D, BS, dtype=kc.dtype)
    for s, (k, v) in enumerate(zip(kq, vq)):
        kb = k.view(torch.uint8) if cache_dt == torch.float8_e4m3fn else k
        vb = v.view(torch.uint8) if cache_dt == torch.float8_e4m3fn else v
        for t in range(k.shape[0]):
            blk, off = tables[s][t // BS], t % BS
            kc[blk, :, :, off, :] = kb[t].reshape(KVH, D // x, x)
            vc[blk, :, :, off] = vb[t]
    if cache_dt == torch.float8_e4m3fn:
        kc, vc = kc.view(torch.float8_e4m3fn), vc.view(torch.float8_e4m3fn)
    return kc.to(cuda), vc.to(cuda)


def _fill_hnd(kq, vq, tables, NB, KVH, D, BS, cache_dt, cuda):
    raw = torch.uint8 if cache_dt == torch.float8_e4m3fn else cache_dt
    kc = torch.zeros(NB, KVH, BS, D, dtype=raw); vc = torch.zeros(NB, KVH, BS, D, dtype=raw)
    for s, (k, v) in enumerate(zip(kq, vq)):
        kb = k.view(torch.uint8) if cache_dt == torch.float8_e4m3fn else k
        vb = v.view(torch.uint8) if cache_dt == torch.float8_e4m3fn else v
        for t in range(k.shape[0]):
            kc[tables[s][t // BS], :, t % BS] = kb[t]; vc[tables[s][t // BS], :, t % BS] = vb[t]
    if cache_dt == torch.float8_e4m3fn:
        kc, vc = kc.view(torch.float8_e4m3fn), vc.view(torch.float8_e4m3fn)
    return kc.to(cuda), vc.to(cuda)


TOL = {torch.float16: 2.0 ** -10, torch.bfloat16: 2.5 * 2.0 ** -8, torch.float32: 2e-5}


@pytest.mark.parametrize("dt,D", [(torch.float32, 128), (torch.float32, 64), (torch.bfloat16, 80), (torch.float16, 96),
                                  (torch.bfloat16, 112), (torch.bfloat16, 192), (torch.float32, 96)])
def test_vllm_layout_dtypes_and_head_sizes(cuda, dt, D):
    ctx = [37, 530]       # the second sequence goes through v2 (two partitions of 512)
    q, kq, vq, kd, vd, tables, NB, max_blocks, _, _ = _setup(cuda, dt, dt, D, ctx)
    kc, vc = _fill_vllm(kq, vq, tables, NB, 2, D, 16, dt, cuda)
    bt = torch.tensor(tables, dtype=torch.int32, device=cuda); cl = torch.tensor(ctx, dtype=torch.int32, device=cuda)
    scale = D ** -0.5
    want = _ref(q.float().numpy(), kd, vd, ctx, scale)
    for max_ctx in (max(ctx), 16 * max_blocks):
        got = paged_attn.paged_attention(q.to(cuda), None, None, kc, vc, bt, cl, None, max_ctx, scale).float().cpu().numpy()
        assert np.abs(got - want).max() <= TOL[dt] * np.abs(want).max(), (dt, D)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_vllm_layout_fp8_cache(cuda, dt):
    D, ctx = 128, [70, 300]
    q, kq, vq, kd, vd, tables, NB, max_blocks, ksc, vsc = _setup(cuda, dt, torch.float8_e4m3fn, D, ctx)
    kc, vc = _fill_vllm(kq, vq, tables, NB, 2, D, 16, torch.float8_e4m3fn, cuda)
    bt = torch.tensor(tables, dtype=torch.int32, device=cuda); cl = torch.tensor(ctx, dtype=torch.int32, device=cuda)
    ks = torch.tensor([ksc], dtype=torch.float32, device=cuda); vs = torch.tensor([vsc], dtype=torch.float32, device=cuda)
    scale = D ** -0.5
    want = _ref(q.float().numpy(), kd, vd, ctx, scale)
    got = paged_attn.paged_attention(q.to(cuda), ks, vs, kc, vc, bt, cl, None, max(ctx), scale).float().cpu().numpy()
    assert np.abs(got - want).max() <= TOL[dt] * np.abs(want).max()


@pytest.mark.parametrize("dt,cache,D", [(torch.bfloat16, torch.float8_e4m3fn, 128), (torch.float16, torch.float8_e4m3fn, 64),
                                        (torch.bfloat16, torch.bfloat16, 512), (torch.float32, torch.float32, 128),
                                        (torch.float16, torch.float16, 256)])
def test_hnd_layout_fp8_f32_and_512(cuda, dt, cache, D):
    ctx = [45, 200]
    BS = 16
    q, kq, vq, kd, vd, tables, NB, max_blocks, ksc, vsc = _setup(cuda, dt, cache, D, ctx)
    kc, vc = _fill_hnd(kq, vq, tables, NB, 2, D, BS, cache, cuda)
    used = [t[:-(-c // BS)] for t, c in zip(tables, ctx)]
    indptr, indices, last = kv_index.make_paged_kv_tensors(used, ctx, BS, sum(len(u) for u in used))
    scale = D ** -0.5
    want = _ref(q.float().numpy(), kd, vd, ctx, scale)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    for split in (None, 4):      # unsplit, and 64-token split-KV tiles
        padded = 2 if split is None else sum(-(-len(u) // split) for u in used)
        req, tile, oind, chunk, mask = kv_index.make_paged_kv_decode_tensors(used, ctx, BS, split, padded)
        got = paged_attn.flashinfer_decode(q.to(cuda), kc, vc, d(indptr), d(indices), d(last), d(req), d(tile), d(oind), d(chunk),
                                           d(mask), scale, k_scale=ksc, v_scale=vsc).float().cpu().numpy()
        assert np.abs(got - want).max() <= TOL[dt] * np.abs(want).max(), (dt, cache, D, split)
