// Per-request block tables over a BlockPool: admission (with prefix-cache hits), growth during decode, trim, release,
// publishing of full blocks, and the integer arrays the attention kernels consume (slot mapping, padded block tables).
// Host-only.  Behaviour follows the reference's manager so block ids come out identical for identical request traces:
//   REF mistralrs-core/src/paged_attention/kv_cache_manager.rs:62-76 (new), :129-174 (get_computed_blocks),
//       :188-264 (allocate_slots: running request grows; new request = hits + fresh blocks, hits that sit in the free
//       list count against capacity), :269-275 (free, tail first), :281-309 (trim), :316-350 (cache_blocks),
//       :394-420 (slot mapping, pad slot -1 past the table), :422-435 (block table padded with 0)
#pragma once
#include <algorithm>
#include <cstdint>
#include <unordered_map>
#include <vector>

#include "kv_index.hpp"

namespace mrs {

class KvCacheManager {
 public:
  KvCacheManager(size_t num_gpu_blocks, size_t block_size, bool enable_caching, std::vector<uint32_t> groups)
      : pool_(num_gpu_blocks, enable_caching, block_size), bs_(block_size), caching_(enable_caching), groups_(std::move(groups)) {
    if (block_size == 0) throw std::invalid_argument("block_size must be positive");
  }
  BlockPool &pool() { return pool_; }
  const BlockPool &pool() const { return pool_; }
  size_t block_size() const { return bs_; }
  bool caching_enabled() const { return caching_; }
  size_t num_usable_blocks() const { return pool_.num_gpu_blocks() > 0 ? pool_.num_gpu_blocks() - 1 : 0; }

  // longest cached prefix; never covers the last token (its logits are still needed)
  void computed_blocks(const uint64_t *hashes, size_t n_hashes, size_t num_tokens, std::vector<size_t> &out) const {
    out.clear();
    if (!caching_ || n_hashes == 0) return;
    const size_t cap = (num_tokens > 0 ? num_tokens - 1 : 0) / bs_;
    std::vector<size_t> hit;
    for (size_t i = 0; i < n_hashes && i < cap; i++) {
      if (!pool_.get_cached_block(hashes[i], groups_, hit) || hit.empty()) break;
      bool same = true;
      for (size_t id : hit) same &= (id == hit[0]);
      if (!same) break;
      out.push_back(hit[0]);
    }
  }

  // true + the newly allocated ids, or false when the pool cannot cover the request (nothing changes then)
  bool allocate_slots(uint64_t req, size_t num_tokens, const std::vector<size_t> &computed, std::vector<size_t> &fresh) {
    fresh.clear();
    const size_t need = (num_tokens + bs_ - 1) / bs_;
    auto it = reqs_.find(req);
    if (it != reqs_.end()) {   // running: only the tail grows
      const size_t have = it->second.ids.size();
      if (need <= have) return true;
      if (!pool_.get_new_blocks(need - have, fresh)) return false;
      it->second.ids.insert(it->second.ids.end(), fresh.begin(), fresh.end());
      return true;
    }
    const size_t n_new = need > computed.size() ? need - computed.size() : 0;
    size_t in_free_list = 0;   // touching these takes them out of the free list, so they use capacity too
    if (caching_) for (size_t id : computed) in_free_list += (pool_.block_ref_cnt(id) == 0);
    if (n_new + in_free_list > pool_.num_free_blocks()) return false;
    if (caching_ && !computed.empty()) pool_.touch(computed);
    if (n_new > 0) pool_.get_new_blocks(n_new, fresh);
    Req r;
    r.ids = computed;
    r.ids.insert(r.ids.end(), fresh.begin(), fresh.end());
    r.cached = computed.size();
    reqs_.emplace(req, std::move(r));
    return true;
  }

  void free(uint64_t req) {
    auto it = reqs_.find(req);
    if (it == reqs_.end()) return;
    std::vector<size_t> rev(it->second.ids.rbegin(), it->second.ids.rend());   // tail blocks become eviction candidates first
    reqs_.erase(it);
    pool_.free_blocks(rev);
  }

  void trim(uint64_t req, size_t num_tokens) {
    auto it = reqs_.find(req);
    if (it == reqs_.end()) return;
    Req &r = it->second;
    const size_t need = (num_tokens + bs_ - 1) / bs_;
    if (need < r.ids.size()) {
      std::vector<size_t> rev(r.ids.rbegin(), r.ids.rend() - (long)need);
      r.ids.resize(need);
      pool_.free_blocks(rev);
    }
    r.cached = std::min(r.cached, r.ids.size());
  }

  void cache_blocks(uint64_t req, const uint64_t *hashes, size_t n_hashes, size_t num_computed_tokens) {
    if (!caching_) return;
    auto it = reqs_.find(req);
    if (it == reqs_.end()) return;
    Req &r = it->second;
    const size_t full = std::min(num_computed_tokens / bs_, r.ids.size());
    if (r.cached >= full) return;
    std::vector<uint64_t> h(hashes, hashes + n_hashes);
    for (uint32_t g : groups_) pool_.cache_full_blocks(r.ids, h, r.cached, full, g);
    r.cached = full;
  }

  bool has(uint64_t req) const { return reqs_.count(req) != 0; }
  const std::vector<size_t> *block_ids(uint64_t req) const {
    auto it = reqs_.find(req);
    return it == reqs_.end() ? nullptr : &it->second.ids;
  }
  size_t num_cached_blocks(uint64_t req) const {
    auto it = reqs_.find(req);
    return it == reqs_.end() ? 0 : it->second.cached;
  }
  size_t num_requests() const { return reqs_.size(); }

  bool slot_mapping(uint64_t req, size_t start, size_t n, int64_t *out) const {
    const std::vector<size_t> *ids = block_ids(req);
    if (!ids) return false;
    for (size_t t = start; t < start + n; t++) {
      const size_t b = t / bs_;
      out[t - start] = b < ids->size() ? (int64_t)((*ids)[b] * bs_ + t % bs_) : PAD_SLOT_ID;
    }
    return true;
  }
  bool block_table(uint64_t req, size_t max_blocks, int32_t *out) const {
    const std::vector<size_t> *ids = block_ids(req);
    if (!ids) return false;
    for (size_t i = 0; i < max_blocks; i++) out[i] = i < ids->size() ? (int32_t)(*ids)[i] : 0;
    return true;
  }

  // One decode step for a batch, straight into the (pinned) staging arrays the runner copies to the device: grows every
  // request to cover context_lens[b] tokens, then writes its padded table row and the slot of its last token.
  // Returns the index of the first request that could not grow (earlier ones keep their growth), or -1.
  int64_t decode_step(const uint64_t *req_ids, const int64_t *context_lens, size_t batch, size_t max_blocks, int32_t *tables,
                      int64_t *slots) {
    std::vector<size_t> none, fresh;
    for (size_t b = 0; b < batch; b++) {
      if (!has(req_ids[b]) || context_lens[b] <= 0) return (int64_t)b;
      if (!allocate_slots(req_ids[b], (size_t)context_lens[b], none, fresh)) return (int64_t)b;
      block_table(req_ids[b], max_blocks, tables + b * max_blocks);
      slot_mapping(req_ids[b], (size_t)context_lens[b] - 1, 1, slots + b);
    }
    return -1;
  }

 private:
  struct Req {
    std::vector<size_t> ids;
    size_t cached = 0;   // leading blocks already published (or taken from the cache)
  };
  BlockPool pool_;
  size_t bs_;
  bool caching_;
  std::vector<uint32_t> groups_;
  std::unordered_map<uint64_t, Req> reqs_;
};

}  // namespace mrs
