// gguf_reader.hpp — GGUF v2/v3 archive reader (mmap), the data-format side of the hot path:
// real weights arrive as ggml blocks inside GGUF files and are uploaded to HBM *as stored* (the
// kernels consume the block layout directly).
//
// Mirrors the behaviour of the reference's archive layer — REF mistralrs-quant/src/gguf/archive.rs:
//   * header / metadata / tensor catalogue parsing            (:351-611, parse_shards)
//   * `general.alignment` (default 32), data section = align_up(end of tensor infos)
//   * split shards: `split.no`, `split.count`, `split.tensors.count`; shards ordered by split.no,
//     duplicate tensor names rejected, declared tensor total checked   (:380-445)
//   * exact byte length per tensor = n_elements / block_elems * block_bytes, must lie inside the file
//   * big-endian files are rejected (the reference refuses to load their tensors too)
// Written from the public GGUF specification; no code shared with the reference.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace mrs {

enum GgufValueType : uint32_t {
  GV_U8 = 0, GV_I8 = 1, GV_U16 = 2, GV_I16 = 3, GV_U32 = 4, GV_I32 = 5, GV_F32 = 6, GV_BOOL = 7, GV_STR = 8,
  GV_ARR = 9, GV_U64 = 10, GV_I64 = 11, GV_F64 = 12
};

// ggml type -> (elements per block, bytes per block); 0/0 when the size is not known here
inline void ggml_type_geometry(uint32_t t, int64_t &elems, int64_t &bytes) {
  switch (t) {
  case 0: elems = 1; bytes = 4; break;       // F32
  case 1: elems = 1; bytes = 2; break;       // F16
  case 2: elems = 32; bytes = 18; break;     // Q4_0
  case 3: elems = 32; bytes = 20; break;     // Q4_1
  case 6: elems = 32; bytes = 22; break;     // Q5_0
  case 7: elems = 32; bytes = 24; break;     // Q5_1
  case 8: elems = 32; bytes = 34; break;     // Q8_0
  case 9: elems = 32; bytes = 36; break;     // Q8_1
  case 10: elems = 256; bytes = 84; break;   // Q2_K
  case 11: elems = 256; bytes = 110; break;  // Q3_K
  case 12: elems = 256; bytes = 144; break;  // Q4_K
  case 13: elems = 256; bytes = 176; break;  // Q5_K
  case 14: elems = 256; bytes = 210; break;  // Q6_K
  case 15: elems = 256; bytes = 292; break;  // Q8_K
  case 24: elems = 1; bytes = 1; break;      // I8
  case 25: elems = 1; bytes = 2; break;      // I16
  case 26: elems = 1; bytes = 4; break;      // I32
  case 27: elems = 1; bytes = 8; break;      // I64
  case 28: elems = 1; bytes = 8; break;      // F64
  case 30: elems = 1; bytes = 2; break;      // BF16
  default: elems = 0; bytes = 0; break;
  }
}

struct GgufValue {
  uint32_t type = 0;       // GgufValueType
  uint32_t arr_type = 0;   // element type when type == GV_ARR
  uint64_t u = 0;          // unsigned / bool scalars
  int64_t i = 0;           // signed scalars
  double f = 0;            // float scalars
  std::string s;           // strings
  std::vector<double> arr_num;        // numeric arrays (ints are exact up to 2^53; u64/i64 kept below too)
  std::vector<int64_t> arr_int;       // integer arrays, exact
  std::vector<std::string> arr_str;   // string arrays
  bool is_int() const { return type <= GV_I32 || type == GV_BOOL || type == GV_U64 || type == GV_I64; }
  int64_t as_int() const { return (type == GV_I8 || type == GV_I16 || type == GV_I32 || type == GV_I64) ? i : (int64_t)u; }
  uint64_t arr_len() const { return arr_type == GV_STR ? arr_str.size() : arr_num.size(); }
};

struct GgufTensor {
  std::string name;
  uint32_t ggml_type = 0;
  std::vector<int64_t> dims;  // ggml order: dims[0] is the innermost (row length K)
  int shard = 0;
  uint64_t offset = 0;        // absolute byte offset inside the shard's file
  int64_t nbytes = -1;        // exact length, -1 when the type's block size is unknown
};

class GgufArchive {
 public:
  explicit GgufArchive(const std::vector<std::string> &paths) {
    if (paths.empty()) throw std::runtime_error("at least one GGUF file is required");
    try {
      std::vector<Shard> parsed;
      for (const auto &p : paths) parsed.push_back(parse_shard(p));
      order_splits(parsed);
      alignment_ = 0;
      for (auto &sh : parsed) {
        const uint64_t a = shard_alignment(sh);
        if (alignment_ == 0) alignment_ = a;
        else if (alignment_ != a) throw std::runtime_error("GGUF shards disagree on general.alignment");
      }
      int64_t declared_total = -1;
      for (auto &sh : parsed) {
        auto it = sh.meta.find("split.tensors.count");
        if (it != sh.meta.end() && it->second.is_int()) declared_total = it->second.as_int();
      }
      for (size_t si = 0; si < parsed.size(); si++) {
        Shard &sh = parsed[si];
        for (auto &kv : sh.meta) {
          auto it = meta_.find(kv.first);
          if (it == meta_.end()) { meta_keys_.push_back(kv.first); meta_.emplace(kv.first, kv.second); }
          // later shards repeat general.* / split.* keys; first occurrence wins (shard 0 carries the model metadata)
        }
        const uint64_t data_start = (sh.infos_end + alignment_ - 1) / alignment_ * alignment_;
        for (auto &t : sh.tensors) {
          if (t.offset % alignment_ != 0)
            throw std::runtime_error("GGUF tensor `" + t.name + "` is not aligned to general.alignment");
          t.offset += data_start;
          t.shard = (int)si;
          int64_t be, bb;
          ggml_type_geometry(t.ggml_type, be, bb);
          if (be > 0) {
            uint64_t n = 1;
            for (int64_t d : t.dims) {
              if (d < 0 || (d != 0 && n > UINT64_MAX / (uint64_t)d)) throw std::runtime_error("GGUF tensor element count overflow");
              n *= (uint64_t)d;
            }
            if (n % (uint64_t)be != 0)
              throw std::runtime_error("GGUF tensor `" + t.name + "` has " + std::to_string(n) + " elements, not divisible by its block size " + std::to_string(be));
            t.nbytes = (int64_t)(n / (uint64_t)be * (uint64_t)bb);
            if (t.offset + (uint64_t)t.nbytes > sh.size)
              throw std::runtime_error("GGUF tensor `" + t.name + "` extends past the end of the file");
          } else if (t.offset > sh.size) {
            throw std::runtime_error("GGUF tensor `" + t.name + "` starts past the end of the file");
          }
          if (index_.count(t.name)) throw std::runtime_error("GGUF tensor `" + t.name + "` is duplicated across shards");
          index_[t.name] = tensors_.size();
          tensors_.push_back(t);
        }
        maps_.push_back(Mapping{sh.base, sh.size});
        sh.base = nullptr;  // ownership moved
      }
      if (declared_total >= 0 && (int64_t)tensors_.size() != declared_total)
        throw std::runtime_error("GGUF split metadata declares " + std::to_string(declared_total) + " tensors, but " +
                                 std::to_string(tensors_.size()) + " were cataloged");
    } catch (...) {
      release();
      throw;
    }
  }
  ~GgufArchive() { release(); }
  GgufArchive(const GgufArchive &) = delete;
  GgufArchive &operator=(const GgufArchive &) = delete;

  uint64_t alignment() const { return alignment_; }
  const std::vector<GgufTensor> &tensors() const { return tensors_; }
  const std::vector<std::string> &metadata_keys() const { return meta_keys_; }
  const GgufValue *metadata(const std::string &key) const {
    auto it = meta_.find(key);
    return it == meta_.end() ? nullptr : &it->second;
  }
  int64_t find_tensor(const std::string &name) const {
    auto it = index_.find(name);
    return it == index_.end() ? -1 : (int64_t)it->second;
  }
  const uint8_t *tensor_data(size_t i) const { return maps_[tensors_[i].shard].base + tensors_[i].offset; }

 private:
  struct Mapping { uint8_t *base; uint64_t size; };
  struct Shard {
    std::string path;
    uint8_t *base = nullptr;
    uint64_t size = 0, infos_end = 0;
    std::map<std::string, GgufValue> meta;
    std::vector<GgufTensor> tensors;
    ~Shard() { if (base) munmap(base, size); }
    Shard() = default;
    Shard(Shard &&o) noexcept { *this = std::move(o); }
    Shard &operator=(Shard &&o) noexcept {
      path = std::move(o.path); base = o.base; size = o.size; infos_end = o.infos_end;
      meta = std::move(o.meta); tensors = std::move(o.tensors); o.base = nullptr;
      return *this;
    }
  };

  // bounds-checked little-endian cursor
  struct Cur {
    const uint8_t *p; uint64_t n, at;
    void need(uint64_t k) const { if (k > n - at) throw std::runtime_error("GGUF file is truncated"); }
    template <typename T> T get() { need(sizeof(T)); T v; memcpy(&v, p + at, sizeof(T)); at += sizeof(T); return v; }
    std::string str() {
      const uint64_t len = get<uint64_t>();
      need(len);
      std::string s((const char *)p + at, (size_t)len);
      at += len;
      return s;
    }
  };

  static void read_scalar(Cur &c, uint32_t type, GgufValue &v) {
    switch (type) {
    case GV_U8: v.u = c.get<uint8_t>(); break;
    case GV_I8: v.i = c.get<int8_t>(); break;
    case GV_U16: v.u = c.get<uint16_t>(); break;
    case GV_I16: v.i = c.get<int16_t>(); break;
    case GV_U32: v.u = c.get<uint32_t>(); break;
    case GV_I32: v.i = c.get<int32_t>(); break;
    case GV_F32: v.f = c.get<float>(); break;
    case GV_BOOL: v.u = c.get<uint8_t>() != 0; break;
    case GV_STR: v.s = c.str(); break;
    case GV_U64: v.u = c.get<uint64_t>(); break;
    case GV_I64: v.i = c.get<int64_t>(); break;
    case GV_F64: v.f = c.get<double>(); break;
    default: throw std::runtime_error("GGUF metadata value has unknown type " + std::to_string(type));
    }
  }

  static GgufValue read_value(Cur &c) {
    GgufValue v;
    v.type = c.get<uint32_t>();
    if (v.type != GV_ARR) { read_scalar(c, v.type, v); return v; }
    v.arr_type = c.get<uint32_t>();
    const uint64_t count = c.get<uint64_t>();
    if (v.arr_type == GV_ARR) throw std::runtime_error("nested GGUF metadata arrays are not supported");
    if (v.arr_type == GV_STR) {
      c.need(count * 8 > count ? count * 8 : count);  // each string carries at least its u64 length
      v.arr_str.reserve((size_t)count);
      for (uint64_t k = 0; k < count; k++) v.arr_str.push_back(c.str());
    } else {
      v.arr_num.reserve((size_t)count);
      v.arr_int.reserve((size_t)count);
      for (uint64_t k = 0; k < count; k++) {
        GgufValue e;
        e.type = v.arr_type;
        read_scalar(c, v.arr_type, e);
        if (e.type == GV_F32 || e.type == GV_F64) { v.arr_num.push_back(e.f); v.arr_int.push_back((int64_t)e.f); }
        else { v.arr_num.push_back((double)e.as_int()); v.arr_int.push_back(e.as_int()); }
      }
    }
    return v;
  }

  static Shard parse_shard(const std::string &path) {
    Shard sh;
    sh.path = path;
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open GGUF file `" + path + "`");
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 24) { ::close(fd); throw std::runtime_error("`" + path + "` is too small to be a GGUF file"); }
    sh.size = (uint64_t)st.st_size;
    void *m = mmap(nullptr, sh.size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (m == MAP_FAILED) throw std::runtime_error("cannot map GGUF file `" + path + "`");
    sh.base = (uint8_t *)m;
    Cur c{sh.base, sh.size, 0};
    const uint32_t magic = c.get<uint32_t>();
    if (magic == 0x47475546u) throw std::runtime_error("big-endian GGUF files are not supported");
    if (magic != 0x46554747u) throw std::runtime_error("`" + path + "` is not a GGUF file (bad magic)");
    const uint32_t version = c.get<uint32_t>();
    if (version != 2 && version != 3) throw std::runtime_error("unsupported GGUF version " + std::to_string(version));
    const uint64_t n_tensors = c.get<uint64_t>(), n_kv = c.get<uint64_t>();
    if (n_tensors > sh.size || n_kv > sh.size) throw std::runtime_error("GGUF header counts exceed the file size");
    for (uint64_t k = 0; k < n_kv; k++) {
      std::string key = c.str();
      GgufValue v = read_value(c);
      if (!sh.meta.emplace(std::move(key), std::move(v)).second) throw std::runtime_error("duplicate GGUF metadata key in `" + path + "`");
    }
    sh.tensors.reserve((size_t)n_tensors);
    for (uint64_t k = 0; k < n_tensors; k++) {
      GgufTensor t;
      t.name = c.str();
      const uint32_t nd = c.get<uint32_t>();
      if (nd > 8) throw std::runtime_error("GGUF tensor `" + t.name + "` has too many dimensions");
      for (uint32_t d = 0; d < nd; d++) t.dims.push_back((int64_t)c.get<uint64_t>());
      t.ggml_type = c.get<uint32_t>();
      t.offset = c.get<uint64_t>();
      sh.tensors.push_back(std::move(t));
    }
    sh.infos_end = c.at;
    return sh;
  }

  static uint64_t shard_alignment(const Shard &sh) {
    auto it = sh.meta.find("general.alignment");
    if (it == sh.meta.end()) return 32;
    if (!it->second.is_int() || it->second.as_int() <= 0) throw std::runtime_error("GGUF general.alignment must be a positive integer");
    return (uint64_t)it->second.as_int();
  }

  static void order_splits(std::vector<Shard> &parsed) {
    if (parsed.size() == 1) {
      auto it = parsed[0].meta.find("split.count");
      if (it != parsed[0].meta.end() && it->second.is_int() && it->second.as_int() > 1)
        throw std::runtime_error("GGUF split metadata declares " + std::to_string(it->second.as_int()) + " shards, but 1 was given");
      return;
    }
    std::vector<int64_t> no(parsed.size());
    for (size_t i = 0; i < parsed.size(); i++) {
      auto c = parsed[i].meta.find("split.count"), n = parsed[i].meta.find("split.no");
      if (c == parsed[i].meta.end() || n == parsed[i].meta.end() || !c->second.is_int() || !n->second.is_int())
        throw std::runtime_error("GGUF shard `" + parsed[i].path + "` is missing split.no / split.count");
      if (c->second.as_int() != (int64_t)parsed.size())
        throw std::runtime_error("GGUF split metadata declares " + std::to_string(c->second.as_int()) + " shards, but " +
                                 std::to_string(parsed.size()) + " were given");
      no[i] = n->second.as_int();
      if (no[i] < 0 || no[i] >= (int64_t)parsed.size()) throw std::runtime_error("GGUF split.no out of range");
    }
    std::vector<Shard> ordered(parsed.size());
    std::vector<bool> seen(parsed.size(), false);
    for (size_t i = 0; i < parsed.size(); i++) {
      if (seen[no[i]]) throw std::runtime_error("GGUF split.no " + std::to_string(no[i]) + " appears twice");
      seen[no[i]] = true;
      ordered[no[i]] = std::move(parsed[i]);
    }
    parsed = std::move(ordered);
  }

  void release() {
    for (auto &m : maps_) if (m.base) munmap(m.base, m.size);
    maps_.clear();
  }

  uint64_t alignment_ = 32;
  std::vector<Mapping> maps_;
  std::vector<GgufTensor> tensors_;
  std::map<std::string, size_t> index_;
  std::map<std::string, GgufValue> meta_;
  std::vector<std::string> meta_keys_;
};

}  // namespace mrs
