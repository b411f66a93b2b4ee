"""Paged decode attention (both cache layouts, split and unsplit) vs the CPU oracle (f64
softmax over the block table).  Tolerance: outputs are bf16/f16, inputs N(0,1); the kernel
accumulates in f32 with fast exp — 2 ulp of the output dtype on the output scale."""
import numpy as np
import pytest
import torch

import oracle
from mistralrs_b200 import kv_index, paged_attn
from util import TORCH_DT, to_dev

pytestmark = pytest.mark.gpu


def _setup(cuda, S, H, KVH, D, BS, ctx, layout, dt, seed=0):
    rng = np.random.default_rng(seed)
    max_blocks = max(-(-c // BS) for c in ctx) + 1
    NB = sum(-(-c // BS) for c in ctx) + 3
    pool = kv_index.BlockPool(NB + 1)
    tables = []
    for c in ctx:
        tables.append(pool.get_new_blocks(-(-c // BS)) or [])
    # shuffle physical blocks between sequences to exercise the indirection
    q = oracle.round_dtype(rng.standard_normal((S, H, D)).astype(np.float32), dt)
    n = (NB + 1) * KVH * D * BS
    kc = oracle.round_dtype(rng.standard_normal(n).astype(np.float32), dt)
    vc = oracle.round_dtype(rng.standard_normal(n).astype(np.float32), dt)
    tdt = TORCH_DT[dt]
    if layout == "vllm":
        kct = to_dev(kc, cuda, dt).reshape(NB + 1, KVH, D // 8, BS, 8)
        vct = to_dev(vc, cuda, dt).reshape(NB + 1, KVH, D, BS)
    else:
        kct = to_dev(kc, cuda, dt).reshape(NB + 1, KVH, BS, D)
        vct = to_dev(vc, cuda, dt).reshape(NB + 1, KVH, BS, D)
    ku = kct.view(torch.int16).cpu().numpy().view(np.uint16).reshape(-1)
    vu = vct.view(torch.int16).cpu().numpy().view(np.uint16).reshape(-1)
    bt = np.zeros((S, max_blocks), dtype=np.int32)
    for s, t in enumerate(tables):
        bt[s, :len(t)] = t
    return q, kct, vct, ku, vu, bt, tables


def _check(got, want, dt):
    ulp = {"bf16": 2.0 ** -8, "f16": 2.0 ** -11}[dt]
    scale = np.abs(want).max()
    err = np.abs(got - want).max()
    assert err <= 2.5 * ulp * scale + 1e-4, (err, scale)


@pytest.mark.parametrize("dt", ["bf16", "f16"])
@pytest.mark.parametrize("cfg", [(1, 32, 8, 128, 16, [300]), (3, 32, 8, 128, 32, [1, 129, 640]),
                                 (2, 32, 4, 64, 16, [77, 400]), (2, 8, 8, 128, 8, [33, 5]),
                                 (1, 16, 2, 256, 16, [150])])
def test_vllm_layout_v1_v2(cuda, dt, cfg):
    S, H, KVH, D, BS, ctx = cfg
    q, kct, vct, ku, vu, bt, _ = _setup(cuda, S, H, KVH, D, BS, ctx, "vllm", dt)
    scale = 1.0 / np.sqrt(D)
    want = oracle.paged_attention(q, ku, vu, bt, ctx, KVH, D, BS, scale, 0, dt)
    for max_ctx in (max(ctx), 4096):  # small -> v1, large -> v2 (512-token partitions)
        out = paged_attn.paged_attention(to_dev(q, cuda, dt), None, None, kct, vct, to_dev(bt, cuda),
                                         to_dev(np.array(ctx, dtype=np.int32), cuda), None, max_ctx, scale)
        _check(out.float().cpu().numpy(), want, dt)


@pytest.mark.parametrize("dt", ["bf16", "f16"])
@pytest.mark.parametrize("cfg", [(1, 32, 8, 128, 16, [384]), (4, 32, 8, 128, 32, [1, 31, 700, 2049]),
                                 (32, 32, 8, 128, 16, list(range(200, 232))), (2, 14, 2, 64, 16, [513, 90])])
@pytest.mark.parametrize("split", [False, True])
def test_flashinfer_hnd(cuda, dt, cfg, split):
    S, H, KVH, D, BS, ctx = cfg
    q, kct, vct, ku, vu, bt, tables = _setup(cuda, S, H, KVH, D, BS, ctx, "hnd", dt, seed=1)
    scale = 1.0 / np.sqrt(D)
    want = oracle.paged_attention(q, ku, vu, bt, ctx, KVH, D, BS, scale, 1, dt)
    indptr, indices, last = kv_index.make_paged_kv_tensors(tables, ctx, BS, bt.size)
    sp = kv_index.decode_split_pages(BS, S, KVH, max(ctx)) if split else None
    ntiles = sum(1 if not sp else -(-max(-(-c // BS), 1) // sp) for c in ctx)
    padded = ntiles + (3 if split else 0)  # graph padding tiles with mask 0
    req, tile, o_indptr, chunk, mask = kv_index.make_paged_kv_decode_tensors(tables, ctx, BS, sp, padded)
    d = lambda a: to_dev(np.ascontiguousarray(a), cuda)
    out = paged_attn.flashinfer_decode(to_dev(q, cuda, dt), kct, vct, d(indptr), d(indices), d(last), d(req), d(tile),
                                       d(o_indptr), d(chunk), d(mask), scale)
    _check(out.float().cpu().numpy(), want, dt)


def test_softcap_window_and_errors(cuda):
    S, H, KVH, D, BS, ctx = 1, 8, 2, 128, 16, [200]
    dt = "bf16"
    q, kct, vct, ku, vu, bt, tables = _setup(cuda, S, H, KVH, D, BS, ctx, "hnd", dt, seed=2)
    scale = 1.0 / np.sqrt(D)
    indptr, indices, last = kv_index.make_paged_kv_tensors(tables, ctx, BS, bt.size)
    req, tile, o_indptr, chunk, mask = kv_index.make_paged_kv_decode_tensors(tables, ctx, BS, None, 1)
    d = lambda a: to_dev(np.ascontiguousarray(a), cuda)
    out = paged_attn.flashinfer_decode(to_dev(q, cuda, dt), kct, vct, d(indptr), d(indices), d(last), d(req), d(tile),
                                       d(o_indptr), d(chunk), d(mask), scale, logits_soft_cap=30.0)
    want = oracle.paged_attention(q, ku, vu, bt, ctx, KVH, D, BS, scale, 1, dt, softcap=30.0)
    _check(out.float().cpu().numpy(), want, dt)
    # sliding window == attention over the last window_left+1 tokens only
    w = 63
    out_w = paged_attn.flashinfer_decode(to_dev(q, cuda, dt), kct, vct, d(indptr), d(indices), d(last), d(req), d(tile),
                                         d(o_indptr), d(chunk), d(mask), scale, window_left=w)
    # oracle: same cache, but pretend the context starts at 200-64 by shifting block table / offsets is awkward
    # -> compare against a dense torch reference instead
    kd, vd = paged_attn.gather_kv_cache_flashinfer(kct, vct, d(bt), d(np.array([0, 200], dtype=np.int32)), 200, torch.bfloat16)
    kd, vd = kd.float()[200 - 64:], vd.float()[200 - 64:]
    qf = to_dev(q, cuda, dt).float()[0]
    ref = torch.empty(H, D, device=cuda)
    for h in range(H):
        p = torch.softmax((kd[:, h // (H // KVH)] @ qf[h]) * scale, dim=0)
        ref[h] = p @ vd[:, h // (H // KVH)]
    _check(out_w.float().cpu().numpy()[0], ref.cpu().numpy(), dt)
    with pytest.raises(ValueError, match="i32"):
        paged_attn.flashinfer_decode(to_dev(q, cuda, dt), kct, vct, d(indptr).long(), d(indices), d(last), d(req), d(tile),
                                     d(o_indptr), d(chunk), d(mask), scale)
