// safetensors_reader.hpp — safetensors container reader (mmap) for UQFF shards.
//
// UQFF, the reference's native quantized format, is "a standard safetensors file with named
// entries" (REF docs/src/content/docs/reference/uqff-format.md; reader
// mistralrs-quant/src/uqff/reader.rs).  Container layout (public safetensors spec):
//   u64 LE header length N | N bytes of JSON | tensor data
//   JSON: { "<name>": {"dtype": "U8", "shape": [..], "data_offsets": [begin, end]}, ...,
//           "__metadata__": {"k": "v", ...} }      offsets are relative to the end of the header.
// Checks follow the safetensors crate's validation: offsets inside the file, begin <= end,
// end - begin == prod(shape) * sizeof(dtype), no gaps or overlaps between tensors.
// Written from the format specification; no code shared with the reference.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace mrs {

// ---- a small JSON value + recursive-descent parser (enough for safetensors headers / config.json)
struct Json {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;
  bool is_int = false;
  int64_t inum = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;  // insertion order kept
  const Json *get(const std::string &k) const {
    for (auto &kv : obj) if (kv.first == k) return &kv.second;
    return nullptr;
  }
};

class JsonParser {
 public:
  JsonParser(const char *p, size_t n) : p_(p), n_(n) {}
  Json parse() {
    Json v = value(0);
    ws();
    if (at_ != n_) fail("trailing characters after JSON value");
    return v;
  }

 private:
  const char *p_; size_t n_, at_ = 0;
  [[noreturn]] void fail(const std::string &m) const { throw std::runtime_error("JSON: " + m + " at byte " + std::to_string(at_)); }
  void ws() { while (at_ < n_ && (p_[at_] == ' ' || p_[at_] == '\n' || p_[at_] == '\t' || p_[at_] == '\r')) at_++; }
  bool lit(const char *s) {
    const size_t l = strlen(s);
    if (n_ - at_ >= l && memcmp(p_ + at_, s, l) == 0) { at_ += l; return true; }
    return false;
  }
  static void utf8(std::string &o, uint32_t c) {
    if (c < 0x80) o += (char)c;
    else if (c < 0x800) { o += (char)(0xC0 | (c >> 6)); o += (char)(0x80 | (c & 0x3F)); }
    else if (c < 0x10000) { o += (char)(0xE0 | (c >> 12)); o += (char)(0x80 | ((c >> 6) & 0x3F)); o += (char)(0x80 | (c & 0x3F)); }
    else { o += (char)(0xF0 | (c >> 18)); o += (char)(0x80 | ((c >> 12) & 0x3F)); o += (char)(0x80 | ((c >> 6) & 0x3F)); o += (char)(0x80 | (c & 0x3F)); }
  }
  uint32_t hex4() {
    if (n_ - at_ < 4) fail("truncated \\u escape");
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) {
      const char c = p_[at_++];
      v <<= 4;
      if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
      else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
      else fail("bad \\u escape");
    }
    return v;
  }
  std::string string() {
    if (at_ >= n_ || p_[at_] != '"') fail("expected string");
    at_++;
    std::string o;
    while (true) {
      if (at_ >= n_) fail("unterminated string");
      const char c = p_[at_++];
      if (c == '"') break;
      if (c != '\\') { o += c; continue; }
      if (at_ >= n_) fail("unterminated escape");
      const char e = p_[at_++];
      switch (e) {
      case '"': o += '"'; break; case '\\': o += '\\'; break; case '/': o += '/'; break;
      case 'b': o += '\b'; break; case 'f': o += '\f'; break; case 'n': o += '\n'; break;
      case 'r': o += '\r'; break; case 't': o += '\t'; break;
      case 'u': {
        uint32_t c1 = hex4();
        if (c1 >= 0xD800 && c1 < 0xDC00 && n_ - at_ >= 6 && p_[at_] == '\\' && p_[at_ + 1] == 'u') {
          at_ += 2;
          const uint32_t c2 = hex4();
          c1 = 0x10000 + ((c1 - 0xD800) << 10) + (c2 - 0xDC00);
        }
        utf8(o, c1);
        break;
      }
      default: fail("unknown escape");
      }
    }
    return o;
  }
  Json value(int depth) {
    if (depth > 64) fail("nesting too deep");
    ws();
    if (at_ >= n_) fail("unexpected end");
    Json v;
    const char c = p_[at_];
    if (c == '{') {
      at_++;
      v.kind = Json::Obj;
      ws();
      if (at_ < n_ && p_[at_] == '}') { at_++; return v; }
      while (true) {
        ws();
        std::string k = string();
        ws();
        if (at_ >= n_ || p_[at_] != ':') fail("expected ':'");
        at_++;
        v.obj.emplace_back(std::move(k), value(depth + 1));
        ws();
        if (at_ < n_ && p_[at_] == ',') { at_++; continue; }
        if (at_ < n_ && p_[at_] == '}') { at_++; break; }
        fail("expected ',' or '}'");
      }
    } else if (c == '[') {
      at_++;
      v.kind = Json::Arr;
      ws();
      if (at_ < n_ && p_[at_] == ']') { at_++; return v; }
      while (true) {
        v.arr.push_back(value(depth + 1));
        ws();
        if (at_ < n_ && p_[at_] == ',') { at_++; continue; }
        if (at_ < n_ && p_[at_] == ']') { at_++; break; }
        fail("expected ',' or ']'");
      }
    } else if (c == '"') {
      v.kind = Json::Str;
      v.str = string();
    } else if (lit("true")) { v.kind = Json::Bool; v.b = true; }
    else if (lit("false")) { v.kind = Json::Bool; v.b = false; }
    else if (lit("null")) { v.kind = Json::Null; }
    else {
      const size_t s = at_;
      if (at_ < n_ && (p_[at_] == '-' || p_[at_] == '+')) at_++;
      bool integral = true;
      while (at_ < n_ && ((p_[at_] >= '0' && p_[at_] <= '9') || p_[at_] == '.' || p_[at_] == 'e' || p_[at_] == 'E' || p_[at_] == '-' || p_[at_] == '+')) {
        if (p_[at_] == '.' || p_[at_] == 'e' || p_[at_] == 'E') integral = false;
        at_++;
      }
      if (at_ == s) fail("unexpected character");
      const std::string t(p_ + s, at_ - s);
      v.kind = Json::Num;
      try {
        v.num = std::stod(t);
        if (integral) { v.inum = std::stoll(t); v.is_int = true; }
      } catch (...) { fail("bad number"); }
    }
    return v;
  }
};

inline int64_t safetensors_dtype_size(const std::string &d) {
  if (d == "BOOL" || d == "U8" || d == "I8" || d == "F8_E4M3" || d == "F8_E5M2") return 1;
  if (d == "U16" || d == "I16" || d == "F16" || d == "BF16") return 2;
  if (d == "U32" || d == "I32" || d == "F32") return 4;
  if (d == "U64" || d == "I64" || d == "F64") return 8;
  return 0;
}

struct StTensor {
  std::string name, dtype;
  std::vector<int64_t> shape;
  uint64_t begin = 0, end = 0;  // absolute byte offsets inside the file
};

class SafetensorsFile {
 public:
  explicit SafetensorsFile(const std::string &path) : path_(path) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open safetensors file `" + path + "`");
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 8) { ::close(fd); throw std::runtime_error("`" + path + "` is too small to be a safetensors file"); }
    size_ = (uint64_t)st.st_size;
    void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (m == MAP_FAILED) throw std::runtime_error("cannot map safetensors file `" + path + "`");
    base_ = (uint8_t *)m;
    try {
      uint64_t hlen;
      memcpy(&hlen, base_, 8);
      if (hlen > size_ - 8) throw std::runtime_error("safetensors header length exceeds the file size in `" + path + "`");
      if (hlen > (100u << 20)) throw std::runtime_error("safetensors header is too large in `" + path + "`");
      const uint64_t data0 = 8 + hlen;
      Json root = JsonParser((const char *)base_ + 8, (size_t)hlen).parse();
      if (root.kind != Json::Obj) throw std::runtime_error("safetensors header is not a JSON object in `" + path + "`");
      for (auto &kv : root.obj) {
        if (kv.first == "__metadata__") {
          if (kv.second.kind != Json::Obj) throw std::runtime_error("safetensors __metadata__ must be an object");
          for (auto &m2 : kv.second.obj) {
            if (m2.second.kind != Json::Str) throw std::runtime_error("safetensors __metadata__ values must be strings");
            meta_.emplace_back(m2.first, m2.second.str);
          }
          continue;
        }
        const Json *dt = kv.second.get("dtype"), *sh = kv.second.get("shape"), *of = kv.second.get("data_offsets");
        if (kv.second.kind != Json::Obj || !dt || !sh || !of || dt->kind != Json::Str || sh->kind != Json::Arr ||
            of->kind != Json::Arr || of->arr.size() != 2)
          throw std::runtime_error("safetensors entry `" + kv.first + "` is malformed");
        StTensor t;
        t.name = kv.first;
        t.dtype = dt->str;
        uint64_t n = 1;
        for (auto &d : sh->arr) {
          if (d.kind != Json::Num || !d.is_int || d.inum < 0) throw std::runtime_error("safetensors entry `" + kv.first + "` has a bad shape");
          if (d.inum != 0 && n > UINT64_MAX / (uint64_t)d.inum) throw std::runtime_error("safetensors element count overflow");
          n *= (uint64_t)d.inum;
          t.shape.push_back(d.inum);
        }
        if (!of->arr[0].is_int || !of->arr[1].is_int || of->arr[0].inum < 0 || of->arr[1].inum < of->arr[0].inum)
          throw std::runtime_error("safetensors entry `" + kv.first + "` has bad data_offsets");
        t.begin = data0 + (uint64_t)of->arr[0].inum;
        t.end = data0 + (uint64_t)of->arr[1].inum;
        if (t.end > size_) throw std::runtime_error("safetensors entry `" + kv.first + "` extends past the end of the file");
        const int64_t es = safetensors_dtype_size(t.dtype);
        if (es == 0) throw std::runtime_error("safetensors entry `" + kv.first + "` has unknown dtype " + t.dtype);
        if (t.end - t.begin != n * (uint64_t)es)
          throw std::runtime_error("safetensors entry `" + kv.first + "` holds " + std::to_string(t.end - t.begin) +
                                   " bytes, its shape needs " + std::to_string(n * (uint64_t)es));
        if (index_.count(t.name)) throw std::runtime_error("safetensors entry `" + t.name + "` is duplicated");
        index_[t.name] = tensors_.size();
        tensors_.push_back(std::move(t));
      }
      // the data section must be covered exactly once (no holes, no overlaps), as the crate requires
      std::vector<std::pair<uint64_t, uint64_t>> spans;
      for (auto &t : tensors_) spans.emplace_back(t.begin, t.end);
      std::sort(spans.begin(), spans.end());
      uint64_t cur = data0;
      for (auto &s : spans) {
        if (s.first != cur) throw std::runtime_error("safetensors data offsets leave a hole or overlap in `" + path + "`");
        cur = s.second;
      }
      if (cur != size_) throw std::runtime_error("safetensors data section has trailing bytes in `" + path + "`");
    } catch (...) {
      munmap(base_, size_);
      base_ = nullptr;
      throw;
    }
  }
  ~SafetensorsFile() { if (base_) munmap(base_, size_); }
  SafetensorsFile(const SafetensorsFile &) = delete;
  SafetensorsFile &operator=(const SafetensorsFile &) = delete;

  const std::string &path() const { return path_; }
  const std::vector<StTensor> &tensors() const { return tensors_; }
  const std::vector<std::pair<std::string, std::string>> &metadata() const { return meta_; }
  int64_t find(const std::string &name) const {
    auto it = index_.find(name);
    return it == index_.end() ? -1 : (int64_t)it->second;
  }
  const uint8_t *data(size_t i) const { return base_ + tensors_[i].begin; }

 private:
  std::string path_;
  uint8_t *base_ = nullptr;
  uint64_t size_ = 0;
  std::vector<StTensor> tensors_;
  std::map<std::string, size_t> index_;
  std::vector<std::pair<std::string, std::string>> meta_;
};

}  // namespace mrs
