"""KV block / slot / CSR / tile-plan integers: C++ host layer vs the reference's own unit-test
expectations (golden values transcribed from block_pool.rs:558-787, metadata.rs tests) and vs
the independent Python oracle on random traces.  Integers: bit-exact."""
import numpy as np
import pytest

from mistralrs_b200 import kv_index
from oracle import kv_index as okv


def test_basic_allocation_golden():  # block_pool.rs `test_basic_allocation`
    pool = kv_index.BlockPool(4)
    assert pool.num_free_blocks() == 3
    blocks = pool.get_new_blocks(2)
    assert len(blocks) == 2 and pool.num_free_blocks() == 1
    assert all(pool.block_ref_cnt(b) == 1 for b in blocks)
    assert pool.null_block_id() == 0 and blocks == [1, 2]  # fresh pool -> 1..n


def test_free_returns_to_pool_golden():  # `test_free_returns_to_pool`
    pool = kv_index.BlockPool(4)
    blocks = pool.get_new_blocks(3)
    assert pool.num_free_blocks() == 0
    pool.free_blocks(blocks)
    assert pool.num_free_blocks() == 3
    assert all(pool.block_ref_cnt(b) == 0 for b in blocks)


def test_allocation_fails_when_exhausted_golden():  # `test_allocation_fails_when_exhausted`
    pool = kv_index.BlockPool(2)
    assert pool.num_free_blocks() == 1
    assert pool.get_new_blocks(1) is not None
    assert pool.num_free_blocks() == 0
    assert pool.get_new_blocks(1) is None


def test_touch_ref_cnt_management_golden():  # `test_touch_ref_cnt_management`
    pool = kv_index.BlockPool(8)
    b = pool.get_new_blocks(1)
    pool.touch(b)
    assert pool.block_ref_cnt(b[0]) == 2
    pool.free_blocks(b)
    assert pool.block_ref_cnt(b[0]) == 1
    pool.free_blocks(b)
    assert pool.block_ref_cnt(b[0]) == 0


def test_fifo_order_frees_append_to_tail():
    pool = kv_index.BlockPool(6)
    a = pool.get_new_blocks(3)          # 1,2,3
    pool.free_blocks(list(reversed(a)))  # tail gets 3,2,1 after 4,5
    assert pool.get_new_blocks(5) == [4, 5, 3, 2, 1]


def test_random_trace_vs_oracle():
    rng = np.random.default_rng(7)
    pool, ref = kv_index.BlockPool(64), okv.BlockPool(64)
    live = []
    for _ in range(2000):
        op = rng.integers(0, 3)
        if op == 0 or not live:
            n = int(rng.integers(1, 6))
            a, b = pool.get_new_blocks(n), ref.get_new_blocks(n)
            assert a == b
            if a:
                live.append(a)
        elif op == 1:
            blocks = live.pop(int(rng.integers(0, len(live))))
            order = list(reversed(blocks))
            pool.free_blocks(order); ref.free_blocks(order)
        else:
            blocks = live[int(rng.integers(0, len(live)))]
            pool.touch(blocks); ref.touch(blocks)
            pool.free_blocks(blocks); ref.free_blocks(blocks)
        assert pool.num_free_blocks() == ref.num_free_blocks()


def test_slot_mapping_and_pad():
    table = [5, 9, 2]
    got = kv_index.slot_mapping(table, 16, 3, 40)
    assert got.tolist() == okv.slot_mapping(table, 16, 3, 40)
    assert got[0] == 5 * 16 + 3 and got[-1] == 2 * 16 + 7
    with pytest.raises(IndexError):
        kv_index.slot_mapping(table, 16, 0, 49)
    assert kv_index.PAD_SLOT_ID == -1


def test_fresh_pool_single_sequence_blocks():
    # fresh pool, one sequence of length L -> blocks 1..ceil(L/BS)
    for L, bs in ((128, 32), (129, 32), (384, 16)):
        pool = kv_index.BlockPool(64)
        nb = -(-L // bs)
        assert pool.get_new_blocks(nb) == list(range(1, nb + 1))


def test_paged_kv_csr_vs_oracle():
    rng = np.random.default_rng(3)
    for bs in (8, 16, 32):
        ctx = [int(c) for c in rng.integers(0, 300, size=5)]
        tables = [[int(v) for v in rng.integers(1, 500, size=-(-c // bs) + int(rng.integers(0, 3)))] for c in ctx]
        padded = sum(-(-c // bs) for c in ctx) + 4
        indptr, indices, last = kv_index.make_paged_kv_tensors(tables, ctx, bs, padded)
        o = okv.make_paged_kv(tables, ctx, bs, padded)
        assert indptr.tolist() == o[0] and indices.tolist() == o[1] and last.tolist() == o[2]
    with pytest.raises(IndexError):
        kv_index.make_paged_kv_tensors([[1]], [40], 16, 8)


def test_decode_split_golden_and_oracle():
    # metadata.rs:74-86: pow2 floor of clamp(ctx / ceil(2*SMs/(batch*KVH)), 256, 2048)
    assert kv_index.decode_split_pages(16, 1, 8, 384, sm_count=148) == 256 // 16
    assert kv_index.decode_split_pages(32, 1, 8, 16384, sm_count=148) == 256 // 32   # 16384/37 = 442 -> 256
    assert kv_index.decode_split_pages(32, 64, 8, 16384, sm_count=148) == 2048 // 32  # grid already full
    rng = np.random.default_rng(5)
    for _ in range(200):
        bs = int(rng.choice([8, 16, 32])); b = int(rng.integers(1, 65)); kvh = int(rng.choice([1, 2, 4, 8]))
        sm = int(rng.choice([64, 132, 148])); ctx = int(rng.integers(1, 40000))
        assert kv_index.decode_split_pages(bs, b, kvh, ctx, sm_count=sm) == okv.decode_split_pages(bs, b, kvh, sm, ctx)


def test_decode_tiles_vs_oracle():
    ctx = [700, 16, 0, 1300]
    bs, split = 16, 16
    tables = [list(range(1, -(-c // bs) + 1)) for c in ctx]
    req, tile, o_indptr, chunk, mask = kv_index.make_paged_kv_decode_tensors(tables, ctx, bs, split, 12)
    o = okv.make_decode_tiles(tables, ctx, bs, split, 12)
    assert req.tolist() == o[0] and tile.tolist() == o[1] and o_indptr.tolist() == o[2]
    assert int(chunk[0]) == o[3] == 256 and mask.tolist() == o[4]
    assert o_indptr.tolist() == [0, 3, 4, 5, 11]
    r2 = kv_index.make_paged_kv_decode_tensors(tables, ctx, bs, None, 4)
    assert r2[0].tolist() == [0, 1, 2, 3] and int(r2[3][0]) == 16 and r2[4].tolist() == [1, 1, 1, 1]


def test_runner_split_policy_fills_the_sms():
    # whole-token runner's own split-KV plan (model.runner_split_pages): chunks of >= 64 tokens, and
    # never more tiles than one attention CTA per SM for the (batch x kv-head) grid
    from mistralrs_b200 import model as M
    for bs in (8, 16, 32):
        for batch in (1, 2, 8, 64):
            for kvh in (1, 8, 32):
                for ctx in (64, 400, 4096, 32768):
                    pages = M.runner_split_pages(bs, batch, kvh, ctx)
                    assert pages >= 1 and pages * bs >= 64
                    tiles = -(-(-(-ctx // bs)) // pages)
                    assert tiles * batch * kvh <= max(148, batch * kvh)
    assert M.runner_split_pages(16, 1, 8, 400) == 4          # bench shape: 64-token chunks, 7 tiles
    assert M.runner_split_pages(16, 64, 8, 400) == 25         # large batch: one chunk per request


# ---- prefix cache: the reference's own unit tests restated (block_pool.rs:606-787, block_hash.rs tests) ----
def _chain(*blocks):
    out, parent = [], None
    for b in blocks:
        parent = kv_index.hash_block_tokens(parent, b)
        out.append(parent)
    return out


def test_block_hash_chain_properties():  # block_hash.rs: deterministic, parent- and token-sensitive, extra keys matter
    h0 = kv_index.hash_block_tokens(None, [1, 2, 3, 4])
    assert h0 == kv_index.hash_block_tokens(None, [1, 2, 3, 4])
    assert h0 != kv_index.hash_block_tokens(None, [1, 2, 3, 5])
    assert kv_index.hash_block_tokens(h0, [5, 6, 7, 8]) != kv_index.hash_block_tokens(None, [5, 6, 7, 8])
    assert kv_index.hash_block_tokens(h0, [5, 6, 7, 8]) != kv_index.hash_block_tokens(h0 ^ 1, [5, 6, 7, 8])
    assert kv_index.hash_block_tokens(None, [1, 2, 3, 4], extra_keys=[77]) != h0
    toks = list(range(1, 14))                                   # 13 tokens, block 4: three full blocks, the tail is not hashed
    hs = kv_index.compute_block_hashes(toks, 4)
    assert hs == _chain(toks[0:4], toks[4:8], toks[8:12])
    assert kv_index.compute_block_hashes(toks[:3], 4) == []
    # incremental: reusing the first two hashes gives the same third (block_hash.rs compute_new_block_hashes)
    assert kv_index.compute_block_hashes(toks, 4, prev=hs[:2]) == hs
    # no collisions over a few thousand distinct short prefixes
    rng = np.random.default_rng(0)
    seen = {kv_index.hash_block_tokens(None, rng.integers(0, 32000, 16)) for _ in range(4000)}
    assert len(seen) == 4000


def test_prefix_cache_basic_golden():  # `test_prefix_cache_basic`
    pool = kv_index.BlockPool(8, True, 4)
    ids = pool.get_new_blocks(3)
    hs = _chain([1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12])
    pool.cache_full_blocks(ids, hs, 0, 3, 0)
    assert pool.num_cached_blocks() == 3
    assert pool.get_cached_block(hs[0], [0]) == [ids[0]]
    with pytest.raises(ValueError):
        pool.cache_full_blocks(ids, hs[:1], 0, 3, 0)


def test_prefix_cache_reuse_after_free_golden():  # `test_prefix_cache_reuse_after_free`
    pool = kv_index.BlockPool(8, True, 4)
    ids = pool.get_new_blocks(2)
    hs = _chain([1, 2, 3, 4], [5, 6, 7, 8])
    pool.cache_full_blocks(ids, hs, 0, 2, 0)
    pool.free_blocks(ids)
    assert pool.num_free_blocks() == 7
    cached = pool.get_cached_block(hs[0], [0])
    assert cached is not None
    pool.touch(cached)
    assert pool.block_ref_cnt(cached[0]) == 1
    assert pool.num_free_blocks() == 6


def test_eviction_on_reallocation_golden():  # `test_eviction_on_reallocation`
    pool = kv_index.BlockPool(4, True, 4)
    ids = pool.get_new_blocks(3)
    h0 = kv_index.hash_block_tokens(None, [1, 2, 3, 4])
    pool.cache_full_blocks(ids, [h0, h0, h0], 0, 1, 0)
    pool.free_blocks(ids)
    assert pool.get_cached_block(h0, [0]) == [ids[0]]            # still there while merely free
    assert len(pool.get_new_blocks(3)) == 3
    assert pool.get_cached_block(h0, [0]) is None and pool.num_cached_blocks() == 0


def test_caching_disabled_is_inert():
    pool = kv_index.BlockPool(8, False, 4)
    ids = pool.get_new_blocks(2)
    hs = _chain([1, 2, 3, 4], [5, 6, 7, 8])
    pool.cache_full_blocks(ids, hs, 0, 2, 0)
    assert pool.num_cached_blocks() == 0 and pool.get_cached_block(hs[0], [0]) is None
    assert pool.computed_blocks(hs, 9, 4) == []


def test_null_block_never_freed_and_usage_golden():  # `test_null_block_never_freed`, `test_usage`
    pool = kv_index.BlockPool(4, False, 16)
    null = pool.null_block_id()
    pool.touch([null])                                          # ref 1 without going through the free list
    pool.free_blocks([null])
    assert pool.block_ref_cnt(null) == 0 and pool.num_free_blocks() == 3
    assert pool.usage() < 0.01
    pool.get_new_blocks(3)
    assert abs(pool.usage() - 1.0) < 0.01
    cached = kv_index.BlockPool(4, True, 4)
    cached.cache_full_blocks([cached.null_block_id()], [123], 0, 1, 0)   # the null block is never published
    assert cached.num_cached_blocks() == 0


def test_get_cached_block_multiple_groups_golden():  # `test_get_cached_block_multiple_groups`
    pool = kv_index.BlockPool(8, True, 4)
    g0, g1 = pool.get_new_blocks(1), pool.get_new_blocks(1)
    h0 = kv_index.hash_block_tokens(None, [1, 2, 3, 4])
    pool.cache_full_blocks(g0, [h0], 0, 1, 0)
    pool.cache_full_blocks(g1, [h0], 0, 1, 1)
    assert pool.get_cached_block(h0, [0, 1]) == [g0[0], g1[0]]
    assert pool.get_cached_block(h0, [0, 2]) is None


def test_same_block_can_cache_multiple_groups_golden():  # `test_same_block_can_cache_multiple_groups`
    pool = kv_index.BlockPool(8, True, 4)
    ids = pool.get_new_blocks(1)
    h0 = kv_index.hash_block_tokens(None, [1, 2, 3, 4])
    pool.cache_full_blocks(ids, [h0], 0, 1, 0)
    pool.cache_full_blocks(ids, [h0], 0, 1, 1)
    pool.cache_full_blocks(ids, [h0], 0, 1, 1)                  # a repeat is not recorded twice
    assert pool.get_cached_block(h0, [0, 1]) == [ids[0], ids[0]]
    assert pool.num_block_hashes(ids[0]) == 2
    pool.free_blocks(ids)
    pool.get_new_blocks(pool.num_free_blocks())
    assert pool.get_cached_block(h0, [0]) is None and pool.get_cached_block(h0, [1]) is None


def test_reset_prefix_cache_golden():  # `test_reset_prefix_cache`
    pool = kv_index.BlockPool(4, True, 4)
    ids = pool.get_new_blocks(2)
    h0 = kv_index.hash_block_tokens(None, [1, 2, 3, 4])
    pool.cache_full_blocks(ids, [h0, h0], 0, 1, 0)
    assert not pool.reset_prefix_cache()
    pool.free_blocks(ids)
    assert pool.reset_prefix_cache()
    assert pool.num_cached_blocks() == 0 and pool.num_block_hashes(ids[0]) == 0


def test_computed_blocks_longest_prefix():  # kv_cache_manager.rs get_computed_blocks: never covers the last token
    bs = 4
    pool = kv_index.BlockPool(16, True, bs)
    a = list(range(100, 112))                                   # 3 full blocks
    ids = pool.get_new_blocks(3)
    ha = kv_index.compute_block_hashes(a, bs)
    pool.cache_full_blocks(ids, ha, 0, 3, 0)
    b = a[:8] + [7, 7, 7, 7, 9]                                 # shares two blocks, then diverges
    hb = kv_index.compute_block_hashes(b, bs)
    assert pool.computed_blocks(hb, len(b), bs) == ids[:2]
    assert pool.computed_blocks(ha, 12, bs) == ids[:2]          # the same 12 tokens again: the last block is recomputed
    assert pool.computed_blocks(ha, 13, bs) == ids[:3]
    assert pool.computed_blocks(kv_index.compute_block_hashes([1] * 8, bs), 8, bs) == []
    # a hit is reused by touching it; the slot mapping of the suffix starts after the cached tokens
    hit = pool.computed_blocks(hb, len(b), bs)
    pool.touch(hit)
    assert [pool.block_ref_cnt(i) for i in hit] == [2, 2]
    table = hit + pool.get_new_blocks(2)
    slots = kv_index.slot_mapping(table, bs, len(hit) * bs, len(b))
    assert list(slots) == [table[2] * bs + i for i in range(4)] + [table[3] * bs]


def test_prefix_cache_random_trace_vs_oracle():
    rng = np.random.default_rng(11)
    bs = 4
    pool, ref = kv_index.BlockPool(48, True, bs), okv.BlockPool(48, True, bs)
    prefixes = [list(rng.integers(0, 50, 12)) for _ in range(6)]
    live = []
    for step in range(1500):
        op = rng.integers(0, 4)
        if op <= 1 or not live:
            toks = prefixes[int(rng.integers(0, 6))][: int(rng.integers(1, 4)) * bs] + list(rng.integers(0, 50, int(rng.integers(1, 9))))
            hs = kv_index.compute_block_hashes(toks, bs)
            grp = [0] if step % 3 else [0, 1]
            hit, rhit = pool.computed_blocks(hs, len(toks), bs, grp), ref.computed_blocks(hs, len(toks), bs, grp)
            assert hit == rhit
            need = -(-len(toks) // bs) - len(hit)
            new, rnew = pool.get_new_blocks(need), ref.get_new_blocks(need)
            assert new == rnew
            if new is None:
                continue
            pool.touch(hit); ref.touch(hit)
            table = hit + new
            for g in grp:
                pool.cache_full_blocks(table, hs, len(hit), len(toks) // bs, g)
                ref.cache_full_blocks(table, hs, len(hit), len(toks) // bs, g)
            live.append(table)
        elif op == 2:
            t = live.pop(int(rng.integers(0, len(live))))
            order = list(reversed(t))                            # tail blocks are evicted first
            pool.free_blocks(order); ref.free_blocks(order)
        else:
            if not live:
                continue
        assert pool.num_free_blocks() == ref.num_free_blocks()
        assert pool.num_cached_blocks() == ref.num_cached_blocks()
    for t in live:
        pool.free_blocks(t); ref.free_blocks(t)
    assert pool.reset_prefix_cache() and ref.reset_prefix_cache()
