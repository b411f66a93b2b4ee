"""Per-request block tables (C++ host/kv_cache_manager.hpp) against the reference's own manager unit tests
(kv_cache_manager.rs:442-679, restated one for one) and against the Python oracle on a random serving trace.
Integers: bit-exact."""
import numpy as np
import pytest

from mistralrs_b200 import kv_index
from mistralrs_b200.kv_index import KVCacheManager, compute_block_hashes
from oracle import kv_index as okv


def _hashes(n):
    return compute_block_hashes(list(range(1, n + 1)), 4)


def test_basic_allocation_golden():
    mgr = KVCacheManager(16, 4, False, [0])
    assert len(mgr.allocate_slots(1, 10)) == 3
    assert mgr.num_blocks_for_request(1) == 3
    assert mgr.get_block_ids(1) == [1, 2, 3]                      # fresh pool: blocks 1.. in order (block 0 is the null block)
    assert mgr.num_usable_blocks() == 15


def test_running_request_extends_golden():
    mgr = KVCacheManager(16, 4, False, [0])
    mgr.allocate_slots(1, 8)
    assert mgr.num_blocks_for_request(1) == 2
    assert len(mgr.allocate_slots(1, 12)) == 1
    assert mgr.num_blocks_for_request(1) == 3
    assert mgr.allocate_slots(1, 12) == []                        # nothing to add


def test_allocation_fails_when_full_golden():
    mgr = KVCacheManager(4, 4, False, [0])
    mgr.allocate_slots(1, 12)
    assert mgr.allocate_slots(2, 4) is None
    assert not mgr.has_request(2)
    assert mgr.allocate_slots(1, 13) is None and mgr.num_blocks_for_request(1) == 3   # a running request cannot grow either


def test_free_returns_blocks_golden():
    mgr = KVCacheManager(8, 4, False, [0])
    mgr.allocate_slots(1, 12)
    assert mgr.num_free_blocks() == 4
    mgr.free(1)
    assert mgr.num_free_blocks() == 7
    assert not mgr.has_request(1)
    mgr.free(1)                                                   # unknown id: no effect
    assert mgr.num_free_blocks() == 7


def test_prefix_cache_hit_golden():
    mgr = KVCacheManager(16, 4, True, [0])
    hashes = _hashes(8)
    mgr.allocate_slots(1, 8)
    first = mgr.get_block_ids(1)
    mgr.cache_blocks(1, hashes, 8)
    mgr.free(1)
    computed = mgr.get_computed_blocks(hashes, 12)
    assert computed.num_computed_tokens == 8 and computed.block_ids == first
    new = mgr.allocate_slots(2, 12, computed.block_ids)
    assert len(new) == 1
    assert mgr.num_blocks_for_request(2) == 3
    assert mgr.get_block_ids(2) == first + new
    assert mgr.num_cached_blocks(2) == 2


def test_prefix_cache_partial_hit_golden():
    mgr = KVCacheManager(16, 4, True, [0])
    mgr.allocate_slots(1, 8)
    mgr.cache_blocks(1, _hashes(8), 8)
    mgr.free(1)
    assert mgr.get_computed_blocks(_hashes(12), 12).num_computed_tokens == 8


def test_prefix_cache_hit_with_group_aliases_golden():
    mgr = KVCacheManager(16, 4, True, [0, 1])
    hashes = _hashes(8)
    mgr.allocate_slots(1, 8)
    mgr.cache_blocks(1, hashes, 8)
    mgr.free(1)
    computed = mgr.get_computed_blocks(hashes, 12)
    assert computed.num_computed_tokens == 8 and len(computed.block_ids) == 2


def test_cache_blocks_incremental_golden():
    mgr = KVCacheManager(16, 4, True, [0])
    hashes = _hashes(16)
    mgr.allocate_slots(1, 16)
    mgr.cache_blocks(1, hashes, 8)
    assert mgr.num_cached_blocks(1) == 2
    mgr.cache_blocks(1, hashes, 16)
    assert mgr.num_cached_blocks(1) == 4
    mgr.cache_blocks(1, hashes, 400)                              # token counts ahead of the allocation are clamped
    assert mgr.num_cached_blocks(1) == 4


def test_slot_mapping_golden():
    mgr = KVCacheManager(16, 4, False, [0])
    mgr.allocate_slots(1, 8)
    ids = mgr.get_block_ids(1)
    slots = mgr.get_slot_mapping(1, 0, 8)
    assert list(slots) == [ids[0] * 4 + i for i in range(4)] + [ids[1] * 4 + i for i in range(4)]
    assert list(mgr.get_slot_mapping(1, 6, 4)) == [ids[1] * 4 + 2, ids[1] * 4 + 3, kv_index.PAD_SLOT_ID, kv_index.PAD_SLOT_ID]
    assert mgr.get_slot_mapping(9, 0, 1) is None


def test_slot_mapping_skip_cached_golden():
    mgr = KVCacheManager(16, 4, True, [0])
    hashes = _hashes(8)
    mgr.allocate_slots(1, 8)
    mgr.cache_blocks(1, hashes, 8)
    mgr.free(1)
    computed = mgr.get_computed_blocks(hashes, 12)
    new = mgr.allocate_slots(2, 12, computed.block_ids)
    slots = mgr.get_slot_mapping(2, 8, 4)
    assert list(slots) == [new[0] * 4 + i for i in range(4)]


def test_block_table_golden():
    mgr = KVCacheManager(16, 4, False, [0])
    mgr.allocate_slots(1, 8)
    table = mgr.get_block_table(1, 5)
    assert table.dtype == np.int32 and list(table) == mgr.get_block_ids(1) + [0, 0, 0]
    assert mgr.get_block_table(2, 5) is None


def test_trim_request_allocation_golden():
    mgr = KVCacheManager(8, 4, False, [0])
    mgr.allocate_slots(1, 12)
    assert mgr.num_blocks_for_request(1) == 3 and mgr.num_free_blocks() == 4
    mgr.trim_request_to_num_tokens(1, 8)
    assert mgr.num_blocks_for_request(1) == 2 and mgr.num_free_blocks() == 5
    mgr.trim_request_to_num_tokens(1, 100)                        # never grows
    assert mgr.num_blocks_for_request(1) == 2


def test_trim_clamps_cached_blocks_golden():
    mgr = KVCacheManager(16, 4, True, [0])
    mgr.allocate_slots(1, 16)
    mgr.cache_blocks(1, _hashes(16), 16)
    assert mgr.num_cached_blocks(1) == 4
    mgr.trim_request_to_num_tokens(1, 8)
    assert mgr.num_blocks_for_request(1) == 2 and mgr.num_cached_blocks(1) == 2


def test_get_computed_blocks_caps_at_prompt_minus_one_golden():
    mgr = KVCacheManager(16, 4, True, [0])
    hashes = _hashes(8)
    mgr.allocate_slots(1, 8)
    mgr.cache_blocks(1, hashes, 8)
    mgr.free(1)
    computed = mgr.get_computed_blocks(hashes, 8)
    assert computed.num_computed_tokens == 4 and len(computed.block_ids) == 1


def test_reset_prefix_cache_golden():
    mgr = KVCacheManager(8, 4, True, [0])
    hashes = _hashes(4)
    mgr.allocate_slots(1, 4)
    mgr.cache_blocks(1, hashes, 4)
    assert not mgr.reset_prefix_cache()
    mgr.free(1)
    assert mgr.reset_prefix_cache()
    assert mgr.get_computed_blocks(hashes, 8).num_computed_tokens == 0


def test_hits_in_the_free_list_count_against_capacity():   # kv_cache_manager.rs:226-238
    mgr = KVCacheManager(4, 4, True, [0])                  # 3 usable blocks
    hashes = _hashes(8)
    mgr.allocate_slots(1, 8)
    mgr.cache_blocks(1, hashes, 8)
    mgr.free(1)                                            # both cached blocks are now free AND findable
    mgr.allocate_slots(2, 4)                               # takes the third block; 2 free, both of them the cached ones
    computed = mgr.get_computed_blocks(hashes, 13)
    assert len(computed.block_ids) == 2
    # 13 tokens need 4 blocks: 2 hits (each removes a free block) + 2 fresh > 2 free
    assert mgr.allocate_slots(3, 13, computed.block_ids) is None
    assert mgr.num_free_blocks() == 2 and not mgr.has_request(3)
    assert len(mgr.allocate_slots(3, 8, computed.block_ids)) == 0 and mgr.num_free_blocks() == 0


def test_decode_step_fills_staging_arrays():
    mgr = KVCacheManager(64, 16, False, [0])
    lens = [5, 16, 33]
    for rid, n in zip((10, 11, 12), lens):
        mgr.allocate_slots(rid, n)
    tables = np.full((3, 6), -7, dtype=np.int32)
    slots = np.full(3, -7, dtype=np.int64)
    # every request appends one token: 6, 17 (new block), 34
    mgr.decode_step([10, 11, 12], [n + 1 for n in lens], 6, tables, slots)
    for b, rid in enumerate((10, 11, 12)):
        ids = mgr.get_block_ids(rid)
        assert list(tables[b]) == ids + [0] * (6 - len(ids))
        assert slots[b] == ids[lens[b] // 16] * 16 + lens[b] % 16
    assert mgr.num_blocks_for_request(11) == 2
    # the same arrays feed the CSR builder the attention kernel reads
    indptr, indices, last = kv_index.make_paged_kv_tensors([mgr.get_block_ids(r) for r in (10, 11, 12)], [6, 17, 34], 16, 8)
    assert list(indptr) == [0, 1, 3, 6] and list(last) == [6, 1, 2]
    with pytest.raises(MemoryError):
        mgr.decode_step([10, 99], [7, 1], 6)
    small = KVCacheManager(3, 16, False, [0])                      # 2 usable blocks
    small.allocate_slots(1, 16); small.allocate_slots(2, 16)
    with pytest.raises(MemoryError):
        small.decode_step([1, 2], [17, 17], 4)


def test_random_serving_trace_vs_oracle():
    rng = np.random.default_rng(5)
    bs = 4
    mgr, ref = KVCacheManager(40, bs, True, [0]), okv.KVCacheManager(40, bs, True, [0])
    system = [list(rng.integers(0, 99, 16)) for _ in range(4)]     # shared system prompts
    live, toks, next_id, hits = [], {}, 0, 0
    for _ in range(3000):
        op = rng.integers(0, 5)
        if op == 0 or not live:                                    # admit
            t = system[int(rng.integers(0, 4))][: int(rng.integers(1, 5)) * bs] + list(rng.integers(0, 99, int(rng.integers(1, 7))))
            h = compute_block_hashes(t, bs)
            c, rc = mgr.get_computed_blocks(h, len(t)), ref.get_computed_blocks(h, len(t))
            assert c.block_ids == rc
            a, ra = mgr.allocate_slots(next_id, len(t), c.block_ids), ref.allocate_slots(next_id, len(t), rc)
            assert a == ra
            if a is not None:
                hits += len(rc)
                mgr.cache_blocks(next_id, h, len(t)); ref.cache_blocks(next_id, h, len(t))
                live.append(next_id); toks[next_id] = t
            next_id += 1
        elif op in (1, 2):                                         # decode one token
            rid = live[int(rng.integers(0, len(live)))]
            toks[rid].append(int(rng.integers(0, 99)))
            a, ra = mgr.allocate_slots(rid, len(toks[rid])), ref.allocate_slots(rid, len(toks[rid]))
            assert a == ra
            if a is None:                                          # preempt
                mgr.free(rid); ref.free(rid); live.remove(rid)
                continue
            h = compute_block_hashes(toks[rid], bs)
            mgr.cache_blocks(rid, h, len(toks[rid])); ref.cache_blocks(rid, h, len(toks[rid]))
        elif op == 3:                                              # finish
            rid = live.pop(int(rng.integers(0, len(live))))
            mgr.free(rid); ref.free(rid)
        else:                                                      # speculative over-allocation rolled back
            rid = live[int(rng.integers(0, len(live)))]
            a, ra = mgr.allocate_slots(rid, len(toks[rid]) + 9), ref.allocate_slots(rid, len(toks[rid]) + 9)
            assert a == ra
            mgr.trim_request_to_num_tokens(rid, len(toks[rid])); ref.trim(rid, len(toks[rid]))
        assert mgr.num_free_blocks() == ref.pool.num_free_blocks()
        for rid in live:
            assert mgr.get_block_ids(rid) == ref.reqs[rid][0] and mgr.num_cached_blocks(rid) == ref.reqs[rid][1]
    assert hits > 100                                              # the trace does exercise the cache
