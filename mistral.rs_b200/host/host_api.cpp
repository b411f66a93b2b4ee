// host_api.cpp — C API over the C++ host layer (kv_index.hpp) for the Python test/bench harness.
#include "kv_index.hpp"

#include <cstring>

using namespace mrs;

extern "C" {

void *mrs_block_pool_new(int64_t num_gpu_blocks) {
  try { return new BlockPool((size_t)num_gpu_blocks); } catch (...) { return nullptr; }
}
void mrs_block_pool_free(void *p) { delete (BlockPool *)p; }
int64_t mrs_block_pool_null_block_id(void *p) { return (int64_t)((BlockPool *)p)->null_block_id(); }
int64_t mrs_block_pool_num_free_blocks(void *p) { return (int64_t)((BlockPool *)p)->num_free_blocks(); }
int64_t mrs_block_pool_ref_cnt(void *p, int64_t id) { return (int64_t)((BlockPool *)p)->block_ref_cnt((size_t)id); }
// returns 1 and fills out[num] on success, 0 when not enough blocks are free
int mrs_block_pool_get_new_blocks(void *p, int64_t num, int64_t *out) {
  std::vector<size_t> v;
  if (!((BlockPool *)p)->get_new_blocks((size_t)num, v)) return 0;
  for (size_t i = 0; i < v.size(); i++) out[i] = (int64_t)v[i];
  return 1;
}
void mrs_block_pool_free_blocks(void *p, const int64_t *ids, int64_t n) {
  std::vector<size_t> v(ids, ids + n);
  ((BlockPool *)p)->free_blocks(v);
}
void mrs_block_pool_touch(void *p, const int64_t *ids, int64_t n) {
  std::vector<size_t> v(ids, ids + n);
  ((BlockPool *)p)->touch(v);
}

// slot mapping for tokens [start, end) of one sequence; returns 0 ok, -1 table too small
int mrs_slot_mapping(const int64_t *table, int64_t table_len, int64_t block_size, int64_t start, int64_t end, int64_t *out) {
  try {
    std::vector<size_t> t(table, table + table_len);
    auto s = slot_mapping(t, (size_t)block_size, (size_t)start, (size_t)end);
    std::memcpy(out, s.data(), s.size() * sizeof(int64_t));
    return 0;
  } catch (...) { return -1; }
}

// CSR builder over a dense [batch, max_blocks] table (row b uses its first ceil(ctx/bs) entries)
int mrs_make_paged_kv(const int64_t *tables, int64_t batch, int64_t max_blocks, const int64_t *context_lens,
                      int64_t block_size, int64_t padded_indices_len, int32_t *indptr, int32_t *indices,
                      int32_t *last_page_len) {
  try {
    std::vector<std::vector<size_t>> t(batch);
    std::vector<size_t> cl(context_lens, context_lens + batch);
    for (int64_t b = 0; b < batch; b++) t[b].assign(tables + b * max_blocks, tables + (b + 1) * max_blocks);
    auto r = make_paged_kv(t, cl, (size_t)block_size, (size_t)padded_indices_len);
    std::memcpy(indptr, r.indptr.data(), r.indptr.size() * 4);
    std::memcpy(indices, r.indices.data(), r.indices.size() * 4);
    std::memcpy(last_page_len, r.last_page_len.data(), r.last_page_len.size() * 4);
    return 0;
  } catch (...) { return -1; }
}

int64_t mrs_decode_split_pages(int64_t block_size, int64_t batch, int64_t kv_heads, int64_t sm_count, int64_t max_ctx) {
  return (int64_t)decode_split_pages((size_t)block_size, (size_t)batch, (size_t)kv_heads, (size_t)sm_count, (size_t)max_ctx);
}

// returns number of valid tiles, or -1 on error
int64_t mrs_make_decode_tiles(const int64_t *table_lens, const int64_t *context_lens, int64_t batch, int64_t block_size,
                              int64_t split_pages, int64_t padded_tiles_len, int32_t *request_indices,
                              int32_t *kv_tile_indices, int32_t *o_indptr, int32_t *kv_chunk_size, uint8_t *mask) {
  try {
    std::vector<size_t> tl(table_lens, table_lens + batch), cl(context_lens, context_lens + batch);
    auto r = make_decode_tiles(tl, cl, (size_t)block_size, (size_t)split_pages, (size_t)padded_tiles_len);
    std::memcpy(request_indices, r.request_indices.data(), r.request_indices.size() * 4);
    std::memcpy(kv_tile_indices, r.kv_tile_indices.data(), r.kv_tile_indices.size() * 4);
    std::memcpy(o_indptr, r.o_indptr.data(), r.o_indptr.size() * 4);
    std::memcpy(mask, r.block_valid_mask.data(), r.block_valid_mask.size());
    *kv_chunk_size = r.kv_chunk_size;
    return (int64_t)r.o_indptr.back();
  } catch (...) { return -1; }
}

}  // extern "C"
