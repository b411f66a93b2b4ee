// kv_index.hpp — host-side producers of the paged-KV integer metadata, mirroring the
// reference's scheduler-side logic so block ids / slots / CSR arrays are bit-identical:
//   BlockPool           REF mistralrs-core/src/paged_attention/block_pool.rs:60-170 (free queue),
//                           :290-324 (new: block 0 popped as the null block), :395-411 (free_blocks),
//                           :419-442 (get_new_blocks), :375-393 (touch)
//   slot_mapping        REF mistralrs-core/src/pipeline/inputs_processor.rs:896-923
//   paged-KV CSR        REF mistralrs-core/src/flashinfer/metadata.rs:88-150 (make_paged_kv_tensors)
//   decode tile plan    REF mistralrs-core/src/flashinfer/metadata.rs:61-86,152-216
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace mrs {

constexpr int64_t PAD_SLOT_ID = -1;  // REF paged_attention/mod.rs:26

class BlockPool {
 public:
  explicit BlockPool(size_t num_gpu_blocks) : n_(num_gpu_blocks) {
    if (n_ == 0) throw std::invalid_argument("Must have at least 1 GPU block");
    const size_t head = n_, tail = n_ + 1;
    prev_.assign(n_ + 2, NO_LINK); next_.assign(n_ + 2, NO_LINK);
    ref_.assign(n_ + 2, 0); is_null_.assign(n_ + 2, 0);
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
    for (size_t i = 0; i < num; i++) { const size_t id = popleft(); ref_[id] = 1; out.push_back(id); }
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
