// Native .safetensors reader (host code only): header parse + read-only mmap, no Python / torch dependency.
//
// SURVEY.md §8(f) n2 — replaces, for the engine's loader, the reference's
//   safe_open(path, framework="pt") / f.get_tensor(key) loop and load_state_dict   (/root/reference/sdmatte_nodes.py:298-323)
// which materialises every tensor as a torch CPU tensor before it is copied again: here the engine's weight repacking reads
// straight from the file mapping.
// Format (https://github.com/huggingface/safetensors): u64 little-endian header length N, N bytes of JSON
//   { "<name>": {"dtype": "F32"|"F16"|"BF16"|..., "shape": [..], "data_offsets": [begin, end]}, ..., "__metadata__": {str: str} }
// followed by the byte buffer the offsets refer to.
#include "common.cuh"
#include "kernels.h"
#include "sdmatte_b200.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace sdm {

struct StEntry {
  std::string name;
  std::string dtype;
  int ndim = 0;
  int64_t shape[8] = {0};
  uint64_t begin = 0, end = 0;
  std::vector<uint8_t> aligned;  // copy of the data when the mapping is not aligned to the element size
};

struct SafeTensorsFile {
  int fd = -1;
  uint8_t* map = nullptr;
  size_t size = 0;
  size_t data_start = 0;
  std::vector<StEntry> entries;
  ~SafeTensorsFile() {
    if (map) munmap(map, size);
    if (fd >= 0) close(fd);
  }
};

namespace {
// minimal JSON reader for the safetensors header subset
struct Json {
  const char* p;
  const char* e;
  [[noreturn]] void fail(const char* what) const { throw Error{std::string("safetensors header: ") + what}; }
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool peek(char c) { ws(); return p < e && *p == c; }
  void expect(char c) { ws(); if (p >= e || *p != c) fail("unexpected character"); ++p; }
  std::string str() {
    expect('"');
    std::string s;
    while (p < e && *p != '"') {
      if (*p == '\\') {
        if (++p >= e) fail("bad escape");
        switch (*p) {
          case 'n': s += '\n'; break;
          case 't': s += '\t'; break;
          case 'r': s += '\r'; break;
          case 'b': s += '\b'; break;
          case 'f': s += '\f'; break;
          case 'u': {  // \uXXXX: keep ASCII, replace the rest (tensor names are ASCII)
            if (e - p < 5) fail("bad \\u escape");
            unsigned v = 0;
            for (int i = 1; i <= 4; ++i) {
              const char c = p[i];
              v = v * 16 + (c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : 0);
            }
            s += v < 128 ? (char)v : '?';
            p += 4;
            break;
          }
          default: s += *p;  // \" \\ \/
        }
        ++p;
      } else {
        s += *p++;
      }
    }
    if (p >= e) fail("unterminated string");
    ++p;
    return s;
  }
  uint64_t uint() {
    ws();
    if (p >= e || *p < '0' || *p > '9') fail("expected an unsigned integer");
    uint64_t v = 0;
    while (p < e && *p >= '0' && *p <= '9') {
      const uint64_t d = (uint64_t)(*p++ - '0');
      if (v > (UINT64_MAX - d) / 10) fail("integer does not fit 64 bits");
      v = v * 10 + d;
    }
    return v;
  }
  // unknown per-tensor keys: strings, numbers, literals, nested objects / arrays — the file is untrusted, so the nesting depth
  // is capped (a header of 100 MB of '[' must not overflow the stack of the ComfyUI process)
  static constexpr int kMaxDepth = 4;
  void skip_value(int depth = 0) {
    ws();
    if (p >= e) fail("truncated");
    if (*p == '"') { str(); return; }
    if (*p == '{' || *p == '[') {
      if (depth >= kMaxDepth) fail("nesting too deep");
      const char close = *p == '{' ? '}' : ']';
      ++p;
      if (peek(close)) { ++p; return; }
      for (;;) {
        if (close == '}') { str(); expect(':'); }
        skip_value(depth + 1);
        if (peek(',')) { ++p; continue; }
        expect(close);
        return;
      }
    }
    while (p < e && *p != ',' && *p != '}' && *p != ']') ++p;
  }
  // "__metadata__" must be a flat string -> string object (the format's rule; the safetensors package rejects anything else)
  void metadata() {
    expect('{');
    if (peek('}')) { ++p; return; }
    for (;;) {
      str();
      expect(':');
      ws();
      if (p >= e || *p != '"') fail("__metadata__ must map strings to strings");
      str();
      if (peek(',')) { ++p; continue; }
      expect('}');
      return;
    }
  }
};
// numel = prod(shape) with overflow detection; false if it does not fit an int64
static bool checked_numel(const int64_t* shape, int ndim, int64_t* out) {
  uint64_t n = 1;
  for (int d = 0; d < ndim; ++d) {
    const uint64_t s = (uint64_t)shape[d];
    if (shape[d] < 0) return false;
    if (s != 0 && n > (uint64_t)INT64_MAX / s) return false;
    n *= s;
  }
  *out = (int64_t)n;
  return true;
}
}  // namespace

std::unique_ptr<SafeTensorsFile> safetensors_open(const char* path) {
  auto f = std::make_unique<SafeTensorsFile>();
  f->fd = open(path, O_RDONLY);
  if (f->fd < 0) throw Error{std::string("safetensors: cannot open '") + path + "'"};
  struct stat st;
  if (fstat(f->fd, &st) != 0 || st.st_size < 8) throw Error{std::string("safetensors: '") + path + "' is too small"};
  f->size = (size_t)st.st_size;
  void* m = mmap(nullptr, f->size, PROT_READ, MAP_PRIVATE, f->fd, 0);
  if (m == MAP_FAILED) throw Error{std::string("safetensors: mmap of '") + path + "' failed"};
  f->map = (uint8_t*)m;
  uint64_t hlen = 0;
  for (int i = 7; i >= 0; --i) hlen = (hlen << 8) | f->map[i];  // little endian
  if (hlen > f->size - 8 || hlen > (100ull << 20)) throw Error{"safetensors: implausible header length"};
  f->data_start = 8 + (size_t)hlen;
  const size_t data_bytes = f->size - f->data_start;
  Json j{(const char*)f->map + 8, (const char*)f->map + 8 + hlen};
  j.expect('{');
  if (!j.peek('}')) {
    for (;;) {
      std::string name = j.str();
      j.expect(':');
      if (name == "__metadata__") {
        j.metadata();
      } else {
        StEntry en;
        en.name = std::move(name);
        bool have_off = false;
        j.expect('{');
        for (;;) {
          const std::string key = j.str();
          j.expect(':');
          if (key == "dtype") en.dtype = j.str();
          else if (key == "shape") {
            j.expect('[');
            if (!j.peek(']')) {
              for (;;) {
                if (en.ndim >= 8) j.fail("more than 8 dimensions");
                const uint64_t ext = j.uint();
                if (ext > (uint64_t)INT64_MAX) j.fail("shape extent does not fit 63 bits");
                en.shape[en.ndim++] = (int64_t)ext;
                if (j.peek(',')) { ++j.p; continue; }
                break;
              }
            }
            j.expect(']');
          } else if (key == "data_offsets") {
            j.expect('[');
            en.begin = j.uint();
            j.expect(',');
            en.end = j.uint();
            j.expect(']');
            have_off = true;
          } else j.skip_value();
          if (j.peek(',')) { ++j.p; continue; }
          j.expect('}');
          break;
        }
        if (!have_off || en.dtype.empty()) j.fail("tensor entry without dtype / data_offsets");
        if (en.begin > en.end || en.end > data_bytes) throw Error{"safetensors: data_offsets of '" + en.name + "' outside the file"};
        int64_t numel = 0;
        if (!checked_numel(en.shape, en.ndim, &numel)) throw Error{"safetensors: shape of '" + en.name + "' overflows"};
        f->entries.push_back(std::move(en));
      }
      if (j.peek(',')) { ++j.p; continue; }
      j.expect('}');
      break;
    }
  } else {
    ++j.p;
  }
  // the package the reference reads checkpoints with rejects duplicate names and any layout in which the tensors do not tile the
  // byte buffer exactly (sorted by offset: first begins at 0, each begins where the previous ended, the last ends at the end of
  // the file): same rules here, so that a file is either valid for both readers or for neither
  {
    std::vector<const StEntry*> by_name, by_off;
    for (auto& en : f->entries) { by_name.push_back(&en); by_off.push_back(&en); }
    std::sort(by_name.begin(), by_name.end(), [](const StEntry* a, const StEntry* b) { return a->name < b->name; });
    for (size_t i = 1; i < by_name.size(); ++i)
      if (by_name[i]->name == by_name[i - 1]->name) throw Error{"safetensors: duplicate tensor name '" + by_name[i]->name + "'"};
    std::sort(by_off.begin(), by_off.end(), [](const StEntry* a, const StEntry* b) { return a->begin != b->begin ? a->begin < b->begin : a->end < b->end; });
    uint64_t pos = 0;
    for (const StEntry* en : by_off) {
      if (en->begin != pos) throw Error{"safetensors: tensor '" + en->name + "' " + (en->begin < pos ? "overlaps the previous tensor" : "leaves a gap in the byte buffer")};
      pos = en->end;
    }
    if (pos != data_bytes) throw Error{"safetensors: the tensors do not cover the byte buffer (incomplete metadata or trailing bytes)"};
  }
  return f;
}

static int st_dtype_code(const std::string& d, int* elem) {
  if (d == "F32") { *elem = 4; return 0; }
  if (d == "F16") { *elem = 2; return 1; }
  if (d == "BF16") { *elem = 2; return 2; }
  *elem = 0;
  return -1;  // not a floating-point weight type the engine reads (I64 step counters etc.)
}

}  // namespace sdm

// ------------------------------------------------------------------------------------------------ C ABI
struct sdm_safetensors {
  std::unique_ptr<sdm::SafeTensorsFile> f;
};

extern "C" {

int sdm_safetensors_open(const char* path, sdm_safetensors** out) {
  try {
    if (!path || !out) throw sdm::Error{"sdm_safetensors_open: null argument"};
    auto h = new sdm_safetensors;
    try {
      h->f = sdm::safetensors_open(path);
    } catch (...) {
      delete h;
      throw;
    }
    *out = h;
    return 0;
  } catch (const sdm::Error& e) {
    sdm::set_last_error(e.msg);
    return 1;
  } catch (const std::exception& e) {
    sdm::set_last_error(e.what());
    return 1;
  }
}

int sdm_safetensors_count(const sdm_safetensors* h) { return h && h->f ? (int)h->f->entries.size() : -1; }

int sdm_safetensors_entry(sdm_safetensors* h, int i, sdm_tensor_desc* out) {
  try {
    if (!h || !h->f || !out || i < 0 || i >= (int)h->f->entries.size()) throw sdm::Error{"sdm_safetensors_entry: bad argument"};
    sdm::StEntry& en = h->f->entries[(size_t)i];
    int elem = 0;
    out->name = en.name.c_str();
    out->dtype = sdm::st_dtype_code(en.dtype, &elem);
    out->ndim = en.ndim;  // may exceed 4: such tensors are not weights of this model (shape[] holds the first four)
    int64_t numel = 1;
    for (int d = 0; d < 4; ++d) out->shape[d] = d < en.ndim ? en.shape[d] : 0;
    if (!sdm::checked_numel(en.shape, en.ndim, &numel)) throw sdm::Error{"safetensors: shape of '" + en.name + "' overflows"};
    const uint8_t* p = h->f->map + h->f->data_start + en.begin;
    if (elem > 0) {
      if ((uint64_t)numel > UINT64_MAX / (uint64_t)elem || (uint64_t)numel * (uint64_t)elem != en.end - en.begin) throw sdm::Error{"safetensors: byte size of '" + en.name + "' does not match its shape"};
      if ((reinterpret_cast<uintptr_t>(p) % (uintptr_t)elem) != 0) {  // header length not a multiple of the element size
        if (en.aligned.empty()) en.aligned.assign(p, p + (en.end - en.begin));
        p = en.aligned.data();
      }
    }
    out->data = p;
    return 0;
  } catch (const sdm::Error& e) {
    sdm::set_last_error(e.msg);
    return 1;
  }
}

void sdm_safetensors_close(sdm_safetensors* h) { delete h; }

}  // extern "C"
