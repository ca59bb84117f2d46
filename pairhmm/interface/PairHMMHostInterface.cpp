// PairHMMHostInterface.cpp -- see the header.  A cursor over a byte buffer does all the work; the std::string
// overloads size the string first and reuse the buffer versions.
#include "PairHMMHostInterface.h"

#include <cstring>
#include <stdexcept>

namespace {

struct Writer {
  char* p;
  void i32(int v) { memcpy(p, &v, sizeof v); p += sizeof v; }
  void bytes(const char* s, int n) { if (n > 0) { memcpy(p, s, (size_t)n); p += n; } }
};

struct Reader {
  const char* p;
  const char* end;      // nullptr = unbounded (raw-pointer overloads trust the caller like the reference does)
  int i32() {
    need(sizeof(int));
    int v; memcpy(&v, p, sizeof v); p += sizeof v; return v;
  }
  char* dup(int n) {    // len + 1 bytes, NUL-terminated
    if (n < 0) throw std::runtime_error("negative length in serialized PairHMM block");
    need((size_t)n);
    char* s = static_cast<char*>(malloc((size_t)n + 1));
    if (!s) throw std::bad_alloc();
    memcpy(s, p, (size_t)n); s[n] = '\0'; p += n;
    return s;
  }
  void need(size_t n) const {
    if (end && (size_t)(end - p) < n) throw std::runtime_error("truncated serialized PairHMM block");
  }
};

template <typename E> E* alloc_array(int num) {
  if (num < 0) throw std::runtime_error("negative count in serialized PairHMM block");
  return static_cast<E*>(calloc(num > 0 ? (size_t)num : 1, sizeof(E)));
}

int read_reads(Reader rd, read_t*& reads) {
  const int num = rd.i32();
  reads = alloc_array<read_t>(num);
  int done = 0;
  try {
    for (; done < num; ++done) {
      read_t& r = reads[done];
      r.len = rd.i32();
      r._b = rd.dup(r.len); r._q = rd.dup(r.len); r._i = rd.dup(r.len); r._d = rd.dup(r.len); r._c = rd.dup(r.len);
    }
  } catch (...) {
    free_reads(reads, done + 1 <= num ? done + 1 : num);   // calloc'ed: untouched pointers are NULL
    reads = nullptr;
    throw;
  }
  return num;
}

int read_haps(Reader rd, hap_t*& haps) {
  const int num = rd.i32();
  haps = alloc_array<hap_t>(num);
  int done = 0;
  try {
    for (; done < num; ++done) {
      haps[done].len = rd.i32();
      haps[done]._b = rd.dup(haps[done].len);
    }
  } catch (...) {
    free_haps(haps, done + 1 <= num ? done + 1 : num);
    haps = nullptr;
    throw;
  }
  return num;
}

}  // namespace

uint64_t serialized_size(const read_t* reads, int num) {
  uint64_t n = sizeof(int);
  for (int k = 0; k < num; ++k) n += sizeof(int) + 5ull * (uint64_t)(reads[k].len > 0 ? reads[k].len : 0);
  return n;
}

uint64_t serialized_size(const hap_t* haps, int num) {
  uint64_t n = sizeof(int);
  for (int k = 0; k < num; ++k) n += sizeof(int) + (uint64_t)(haps[k].len > 0 ? haps[k].len : 0);
  return n;
}

uint64_t serialize(void* buf, const read_t* reads, int num) {
  Writer w{static_cast<char*>(buf)};
  w.i32(num);
  for (int k = 0; k < num; ++k) {
    const read_t& r = reads[k];
    w.i32(r.len);
    w.bytes(r._b, r.len); w.bytes(r._q, r.len); w.bytes(r._i, r.len); w.bytes(r._d, r.len); w.bytes(r._c, r.len);
  }
  return (uint64_t)(w.p - static_cast<char*>(buf));
}

uint64_t serialize(void* buf, const hap_t* haps, int num) {
  Writer w{static_cast<char*>(buf)};
  w.i32(num);
  for (int k = 0; k < num; ++k) { w.i32(haps[k].len); w.bytes(haps[k]._b, haps[k].len); }
  return (uint64_t)(w.p - static_cast<char*>(buf));
}

int deserialize(const void* buf, read_t*& reads) { return read_reads(Reader{static_cast<const char*>(buf), nullptr}, reads); }
int deserialize(const void* buf, hap_t*& haps) { return read_haps(Reader{static_cast<const char*>(buf), nullptr}, haps); }

std::string serialize(const read_t* reads, int num) {
  std::string s(serialized_size(reads, num), '\0');
  serialize(&s[0], reads, num);
  return s;
}

std::string serialize(const hap_t* haps, int num) {
  std::string s(serialized_size(haps, num), '\0');
  serialize(&s[0], haps, num);
  return s;
}

int deserialize(const std::string& data, read_t*& reads) {
  return read_reads(Reader{data.data(), data.data() + data.size()}, reads);
}

int deserialize(const std::string& data, hap_t*& haps) {
  return read_haps(Reader{data.data(), data.data() + data.size()}, haps);
}
