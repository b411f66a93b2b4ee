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
