// kv_index.hpp — host-side producers of the paged-KV integer metadata, mirroring the
// reference's scheduler-side logic so block ids / slots / CSR arrays are bit-identical:
//   BlockPool           REF mistralrs-core/src/paged_attention/block_pool.rs:60-170 (free queue),
//                           :290-324 (new: block 0 popped as the null block), :395-411 (free_blocks),
//                           :419-442 (get_new_blocks), :375-393 (touch); prefix cache: :355-372 (get_cached_block),
//                           :454-503 (cache_full_blocks), :505-512 (eviction on reallocation), :514-527 (reset);
//                           block hash chain REF paged_attention/block_hash.rs:126-150
//   slot_mapping        REF mistralrs-core/src/pipeline/inputs_processor.rs:896-923
//   paged-KV CSR        REF mistralrs-core/src/flashinfer/metadata.rs:88-150 (make_paged_kv_tensors)
//   decode tile plan    REF mistralrs-core/src/flashinfer/metadata.rs:61-86,152-216
#pragma once
#include <cstdint>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace mrs {

constexpr int64_t PAD_SLOT_ID = -1;  // REF paged_attention/mod.rs:26

// Hash of one full block of tokens, chained to the hash of the block before it (or a fixed seed for the first block), so
// equal hashes mean equal PREFIXES.  The reference feeds Rust's DefaultHasher (SipHash); the value never leaves the
// scheduler, only equality matters, so this is a 64-bit mix of our own (splitmix-style finaliser over a running state).
inline uint64_t hash_block_tokens(bool has_parent, uint64_t parent, const uint32_t *tokens, size_t n, const uint64_t *extra = nullptr,
                                  size_t n_extra = 0) {
  auto mix = [](uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 27; h *= 0x94D049BB133111EBull; h ^= h >> 31;
    return h;
  };
  uint64_t h = mix(0x6D72735F62323030ull, has_parent ? parent : 0x4E4F4E455F484153ull);   // "NONE_HAS": the first block's seed
  h = mix(h, (uint64_t)n);
  for (size_t i = 0; i < n; i++) h = mix(h, tokens[i]);
  for (size_t i = 0; i < n_extra; i++) h = mix(h, extra[i] ^ 0xA5A5A5A5A5A5A5A5ull);
  return h;
}

class BlockPool {
 public:
  explicit BlockPool(size_t num_gpu_blocks, bool enable_caching = false, size_t hash_block_size = 16)
      : n_(num_gpu_blocks), caching_(enable_caching), hash_block_size_(hash_block_size) {
    if (n_ == 0) throw std::invalid_argument("Must have at least 1 GPU block");
    const size_t head = n_, tail = n_ + 1;
    prev_.assign(n_ + 2, NO_LINK); next_.assign(n_ + 2, NO_LINK);
    ref_.assign(n_ + 2, 0); is_null_.assign(n_ + 2, 0);
    hashes_.assign(n_ + 2, {});
    // doubly linked free list seeded 0..n-1 between the two sentinels
    size_t p = head;
    for (size_t i = 0; i < n_; i++) { next_[p] = i; prev_[i] = p; p = i; }
    next_[p] = tail; prev_[tail] = p;
    head_ = head; tail_ = tail; free_ = n_;
    null_id_ = popleft();  // the first block is the null block: placeholder, never freed
    is_null_[null_id_] = 1;
  }
  size_t null_block_id() const { return null_id_; }
  size_t num_free_blocks() const { return free_; }
  size_t num_gpu_blocks() const { return n_; }
  uint32_t block_ref_cnt(size_t id) const { return ref_.at(id); }
  double usage() const {
    const size_t total = n_ - 1;
    return total == 0 ? 0.0 : 1.0 - (double)free_ / (double)total;
  }
  // pops from the head of the free list; empty result when not enough blocks are free
  bool get_new_blocks(size_t num, std::vector<size_t> &out) {
    out.clear();
    if (num > free_) return false;
    for (size_t i = 0; i < num; i++) {
      const size_t id = popleft();
      if (caching_) evict(id);   // a freed block keeps its hashes until it is handed out again
      ref_[id] = 1;
      out.push_back(id);
    }
    return true;
  }
  // ---- prefix cache ----
  bool caching_enabled() const { return caching_; }
  size_t hash_block_size() const { return hash_block_size_; }
  size_t num_cached_blocks() const { return cache_.size(); }   // distinct (hash, group) keys, as the reference counts
  size_t num_block_hashes(size_t id) const { return hashes_.at(id).size(); }
  // all groups must hit: one cached block id per group, or false
  bool get_cached_block(uint64_t hash, const std::vector<uint32_t> &groups, std::vector<size_t> &out) const {
    out.clear();
    for (uint32_t g : groups) {
      auto it = cache_.find(Key{hash, g});
      if (it == cache_.end() || it->second.empty()) { out.clear(); return false; }
      out.push_back(it->second.front());
    }
    return true;
  }
  // give blocks [num_cached, num_full) of a request their hashes and publish them
  void cache_full_blocks(const std::vector<size_t> &block_ids, const std::vector<uint64_t> &block_hashes, size_t num_cached,
                         size_t num_full, uint32_t group) {
    if (!caching_ || num_cached >= num_full) return;
    if (block_hashes.size() < num_full || block_ids.size() < num_full) throw std::invalid_argument("Not enough block hashes for the full blocks");
    for (size_t i = num_cached; i < num_full; i++) {
      const size_t id = block_ids[i];
      if (is_null_.at(id)) continue;
      const Key k{block_hashes[i], group};
      bool have = false;
      for (const Key &e : hashes_[id]) have |= (e == k);
      if (have) continue;
      hashes_[id].push_back(k);
      cache_[k].push_back(id);
    }
  }
  // only when nothing but the null block is in use
  bool reset_prefix_cache() {
    if (n_ - free_ != 1) return false;
    cache_.clear();
    for (auto &h : hashes_) h.clear();
    return true;
  }
  // two passes like the reference: decrement everything, then append newly-free blocks in order
  void free_blocks(const std::vector<size_t> &ordered) {
    for (size_t id : ordered) if (ref_.at(id) > 0) ref_[id]--;
    for (size_t id : ordered) if (ref_[id] == 0 && !is_null_[id] && !in_free_list(id)) append(id);
  }
  void touch(const std::vector<size_t> &ids) {
    for (size_t id : ids) {
      if (ref_.at(id) == 0 && !is_null_[id]) remove(id);
      ref_[id]++;
    }
  }

 private:
  struct Key {
    uint64_t hash; uint32_t group;
    bool operator==(const Key &o) const { return hash == o.hash && group == o.group; }
  };
  struct KeyHash { size_t operator()(const Key &k) const { return (size_t)(k.hash ^ ((uint64_t)k.group * 0x9E3779B97F4A7C15ull)); } };
  void evict(size_t id) {
    for (const Key &k : hashes_[id]) {
      auto it = cache_.find(k);
      if (it == cache_.end()) continue;
      auto &v = it->second;
      for (size_t j = 0; j < v.size(); j++) if (v[j] == id) { v.erase(v.begin() + (long)j); break; }
      if (v.empty()) cache_.erase(it);
    }
    hashes_[id].clear();
  }
  static constexpr size_t NO_LINK = (size_t)-1;
  bool in_free_list(size_t id) const { return prev_[id] != NO_LINK; }
  size_t popleft() {
    const size_t id = next_[head_];
    if (id == tail_) throw std::runtime_error("free list empty");
    remove(id);
    return id;
  }
  void remove(size_t id) {
    const size_t p = prev_[id], nx = next_[id];
    if (p == NO_LINK || nx == NO_LINK) throw std::runtime_error("remove() called on block not in free list");
    next_[p] = nx; prev_[nx] = p; prev_[id] = NO_LINK; next_[id] = NO_LINK; free_--;
  }
  void append(size_t id) {
    const size_t last = prev_[tail_];
    next_[last] = id; prev_[id] = last; next_[id] = tail_; prev_[tail_] = id; free_++;
  }
  size_t n_, head_, tail_, free_, null_id_;
  std::vector<size_t> prev_, next_;
  std::vector<uint32_t> ref_;
  std::vector<uint8_t> is_null_;
  bool caching_;
  size_t hash_block_size_;
  std::vector<std::vector<Key>> hashes_;                                  // per block: the (hash, group) keys it is published under
  std::unordered_map<Key, std::vector<size_t>, KeyHash> cache_;           // key -> blocks holding that prefix, oldest first
};

// slot = table[i / BS] * BS + i % BS for i in [start, end)
inline std::vector<int64_t> slot_mapping(const std::vector<size_t> &table, size_t block_size, size_t start, size_t end) {
  std::vector<int64_t> out;
  for (size_t i = start; i < end; i++) {
    if (i / block_size >= table.size()) throw std::out_of_range("Block table is too small (prompt)!");
    out.push_back((int64_t)(table[i / block_size] * block_size + i % block_size));
  }
  return out;
}

struct PagedKv {
  std::vector<int32_t> indptr, indices, last_page_len;
};

inline PagedKv make_paged_kv(const std::vector<std::vector<size_t>> &tables, const std::vector<size_t> &context_lens,
                             size_t block_size, size_t padded_indices_len) {
  PagedKv r;
  r.indptr.push_back(0);
  int32_t nnz = 0;
  for (size_t b = 0; b < tables.size(); b++) {
    const size_t nb = (context_lens[b] + block_size - 1) / block_size;
    if (nb > tables[b].size()) throw std::out_of_range("paged kv block table is too small");
    nnz += (int32_t)nb;
    r.indptr.push_back(nnz);
    for (size_t i = 0; i < nb; i++) r.indices.push_back((int32_t)tables[b][i]);
    r.last_page_len.push_back(nb == 0 ? 0 : (int32_t)(context_lens[b] - (nb - 1) * block_size));
  }
  if (r.indices.size() > padded_indices_len) throw std::out_of_range("paged kv indices exceed padded length");
  r.indices.resize(padded_indices_len, 0);
  return r;
}

// pow2 floor of clamp(ctx / ceil(2*SMs / (batch*KVH)), 256, 2048) — metadata.rs:61-86
inline size_t decode_split_tokens(size_t batch, size_t kv_heads, size_t sm_count, size_t max_ctx) {
  const size_t unsplit = batch * kv_heads > 0 ? batch * kv_heads : 1;
  size_t chunks = (2 * sm_count + unsplit - 1) / unsplit;
  if (chunks < 1) chunks = 1;
  size_t tokens = max_ctx / chunks;
  if (tokens < 256) tokens = 256;
  if (tokens > 2048) tokens = 2048;
  size_t p = 1;
  while (p * 2 <= tokens) p *= 2;
  return p;
}
inline size_t decode_split_pages(size_t block_size, size_t batch, size_t kv_heads, size_t sm_count, size_t max_ctx) {
  const size_t t = decode_split_tokens(batch, kv_heads, sm_count, max_ctx);
  const size_t pages = (t + block_size - 1) / block_size;
  return pages < 1 ? 1 : pages;
}

struct DecodeTiles {
  std::vector<int32_t> request_indices, kv_tile_indices, o_indptr;
  int32_t kv_chunk_size;
  std::vector<uint8_t> block_valid_mask;
};

// split_pages == 0 means "no split" (Option::None in the reference)
inline DecodeTiles make_decode_tiles(const std::vector<size_t> &table_lens, const std::vector<size_t> &context_lens,
                                     size_t block_size, size_t split_pages, size_t padded_tiles_len) {
  DecodeTiles r;
  const size_t chunk_pages = split_pages == 0 ? (size_t)-1 : split_pages;
  r.o_indptr.push_back(0);
  for (size_t b = 0; b < context_lens.size(); b++) {
    const size_t nb = (context_lens[b] + block_size - 1) / block_size;
    if (nb > table_lens[b]) throw std::out_of_range("paged kv decode block table is too small");
    const size_t nbm = nb < 1 ? 1 : nb;
    const size_t chunks = chunk_pages == (size_t)-1 ? 1 : (nbm + chunk_pages - 1) / chunk_pages;
    for (size_t t = 0; t < chunks; t++) { r.request_indices.push_back((int32_t)b); r.kv_tile_indices.push_back((int32_t)t); }
    r.o_indptr.push_back((int32_t)r.request_indices.size());
  }
  if (r.request_indices.size() > padded_tiles_len) throw std::out_of_range("paged kv decode tiles exceed padded length");
  const size_t valid = r.request_indices.size();
  r.request_indices.resize(padded_tiles_len, 0);
  r.kv_tile_indices.resize(padded_tiles_len, 0);
  r.block_valid_mask.assign(valid, 1);
  r.block_valid_mask.resize(padded_tiles_len, 0);
  r.kv_chunk_size = (int32_t)((split_pages == 0 ? 1 : split_pages) * block_size);
  return r;
}

}  // namespace mrs
