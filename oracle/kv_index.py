"""ORACLE (test infrastructure): pure-Python restatement of the reference's KV index producers,
written independently of the C++ host layer so the two can be diffed on random traces.
REF: mistralrs-core/src/paged_attention/block_pool.rs:60-170,290-442 (+ prefix cache :182-280,355-372,454-527);
     mistralrs-core/src/paged_attention/kv_cache_manager.rs:129-174 (longest cached prefix);
     mistralrs-core/src/pipeline/inputs_processor.rs:896-923;
     mistralrs-core/src/flashinfer/metadata.rs:61-216."""
from collections import OrderedDict


class BlockPool:
    def __init__(self, num_gpu_blocks, enable_caching=False, hash_block_size=16):
        assert num_gpu_blocks > 0
        self.n = num_gpu_blocks
        self.free = OrderedDict((i, None) for i in range(num_gpu_blocks))  # FIFO free list
        self.ref = [0] * num_gpu_blocks
        self.null = self.free.popitem(last=False)[0]  # block 0 becomes the null block
        self.caching, self.hash_block_size = enable_caching, hash_block_size
        self.keys = [[] for _ in range(num_gpu_blocks)]   # per block: (hash, group) it is published under
        self.cache = {}                                   # (hash, group) -> [block ids], oldest first

    # ---- prefix cache (hashes are opaque integers here; the chain lives in the host layer) ----
    def num_cached_blocks(self):
        return len(self.cache)

    def get_cached_block(self, h, groups):
        out = []
        for g in groups:
            ids = self.cache.get((h, g))
            if not ids:
                return None
            out.append(ids[0])
        return out

    def cache_full_blocks(self, block_ids, hashes, num_cached, num_full, group):
        if not self.caching or num_cached >= num_full:
            return
        assert len(hashes) >= num_full
        for i in range(num_cached, num_full):
            b, k = block_ids[i], (hashes[i], group)
            if b == self.null or k in self.keys[b]:
                continue
            self.keys[b].append(k)
            self.cache.setdefault(k, []).append(b)

    def reset_prefix_cache(self):
        if self.n - len(self.free) != 1:
            return False
        self.cache.clear()
        self.keys = [[] for _ in range(self.n)]
        return True

    def computed_blocks(self, hashes, num_tokens, block_size, groups=(0,)):
        """Longest cached prefix, capped so the last token is always recomputed."""
        if not self.caching:
            return []
        out = []
        for i, h in enumerate(hashes):
            if i >= max(num_tokens - 1, 0) // block_size:
                break
            ids = self.get_cached_block(h, groups)
            if ids is None or any(b != ids[0] for b in ids):
                break
            out.append(ids[0])
        return out

    def num_free_blocks(self):
        return len(self.free)

    def get_new_blocks(self, n):
        if n > len(self.free):
            return None
        out = []
        for _ in range(n):
            b = self.free.popitem(last=False)[0]
            for k in self.keys[b]:                        # evicted only now, on reallocation
                self.cache[k].remove(b)
                if not self.cache[k]:
                    del self.cache[k]
            self.keys[b] = []
            self.ref[b] = 1
            out.append(b)
        return out

    def free_blocks(self, ordered):
        for b in ordered:
            self.ref[b] = max(self.ref[b] - 1, 0)
        for b in ordered:
            if self.ref[b] == 0 and b != self.null and b not in self.free:
                self.free[b] = None

    def touch(self, ids):
        for b in ids:
            if self.ref[b] == 0 and b != self.null:
                del self.free[b]
            self.ref[b] += 1


def slot_mapping(table, block_size, start, end):
    return [table[i // block_size] * block_size + i % block_size for i in range(start, end)]


def make_paged_kv(tables, context_lens, block_size, padded):
    indptr, indices, last = [0], [], []
    for t, c in zip(tables, context_lens):
        nb = -(-c // block_size)
        indptr.append(indptr[-1] + nb)
        indices += list(t[:nb])
        last.append(0 if nb == 0 else c - (nb - 1) * block_size)
    return indptr, indices + [0] * (padded - len(indices)), last


def decode_split_tokens(batch, kvh, sm_count, max_ctx):
    unsplit = max(batch * kvh, 1)
    chunks = max(-(-2 * sm_count // unsplit), 1)
    tokens = min(max(max_ctx // chunks, 256), 2048)
    return 1 << (tokens.bit_length() - 1)


def decode_split_pages(block_size, batch, kvh, sm_count, max_ctx):
    return max(-(-decode_split_tokens(batch, kvh, sm_count, max_ctx) // block_size), 1)


def make_decode_tiles(tables, context_lens, block_size, split_pages, padded):
    req, tile, o_indptr = [], [], [0]
    for b, (t, c) in enumerate(zip(tables, context_lens)):
        nb = -(-c // block_size)
        chunks = 1 if not split_pages else -(-max(nb, 1) // split_pages)
        for k in range(chunks):
            req.append(b)
            tile.append(k)
        o_indptr.append(len(req))
    valid = len(req)
    mask = [1] * valid + [0] * (padded - valid)
    return req + [0] * (padded - valid), tile + [0] * (padded - valid), o_indptr, (split_pages or 1) * block_size, mask


class KVCacheManager:
    """Pure-Python restatement of the reference's per-request manager (REF kv_cache_manager.rs:188-350)."""

    def __init__(self, num_gpu_blocks, block_size, enable_caching, groups=(0,)):
        self.pool = BlockPool(num_gpu_blocks, enable_caching, block_size)
        self.bs, self.caching, self.groups = block_size, enable_caching, list(groups)
        self.reqs = {}   # id -> [block ids, cached count]

    def get_computed_blocks(self, hashes, num_tokens):
        return self.pool.computed_blocks(hashes, num_tokens, self.bs, self.groups) if hashes else []

    def allocate_slots(self, rid, num_tokens, computed=()):
        need = -(-num_tokens // self.bs)
        if rid in self.reqs:
            ids = self.reqs[rid][0]
            if need <= len(ids):
                return []
            new = self.pool.get_new_blocks(need - len(ids))
            if new is None:
                return None
            ids += new
            return new
        computed = list(computed)
        n_new = max(need - len(computed), 0)
        evictable = sum(self.pool.ref[b] == 0 for b in computed) if self.caching else 0
        if n_new + evictable > self.pool.num_free_blocks():
            return None
        if computed and self.caching:
            self.pool.touch(computed)
        new = self.pool.get_new_blocks(n_new) if n_new else []
        self.reqs[rid] = [computed + new, len(computed)]
        return new

    def free(self, rid):
        if rid in self.reqs:
            self.pool.free_blocks(list(reversed(self.reqs.pop(rid)[0])))

    def trim(self, rid, num_tokens):
        if rid not in self.reqs:
            return
        r = self.reqs[rid]
        need = -(-num_tokens // self.bs)
        if need < len(r[0]):
            gone = r[0][need:]
            del r[0][need:]
            self.pool.free_blocks(list(reversed(gone)))
        r[1] = min(r[1], len(r[0]))

    def cache_blocks(self, rid, hashes, num_computed_tokens):
        if not self.caching or rid not in self.reqs:
            return
        r = self.reqs[rid]
        full = min(num_computed_tokens // self.bs, len(r[0]))
        if r[1] >= full:
            return
        for g in self.groups:
            self.pool.cache_full_blocks(r[0], hashes, r[1], full, g)
        r[1] = full
