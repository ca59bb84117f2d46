// PairHMMHostInterface.h -- the batch element types and the wire format of the PairHMM path, with the names and
// signatures of the reference's interface (/root/reference/pairhmm/interface/PairHMMHostInterface.h:27-83) so that
// callers of the reference compile against this header unchanged.
//
// Wire format (PairHMMHostInterface.cpp:175-207 there): native-endian int32 count; per read an int32 length followed
// by `length` bytes of each of _b, _q, _i, _d, _c; per haplotype an int32 length and `length` bytes of _b.  No
// padding, no checksum.  deserialize() returns malloc'ed arrays whose strings are NUL-terminated (len + 1 bytes)
// like the reference's; release them with free_reads / free_haps.
#ifndef PAIRHMM_HOST_INTERFACE_H
#define PAIRHMM_HOST_INTERFACE_H
#include <cstdint>
#include <cstdlib>
#include <string>

typedef struct {
  int   len;
  char* _b;   // bases (ASCII)
  char* _q;   // base qualities
  char* _i;   // insertion gap-open qualities
  char* _d;   // deletion gap-open qualities
  char* _c;   // gap-continuation qualities
} read_t;

typedef struct {
  int   len;
  char* _b;
} hap_t;

static inline void alloc_data(read_t* v, int len) {
  v->len = len;
  char** tracks[5] = {&v->_b, &v->_q, &v->_i, &v->_d, &v->_c};
  for (int t = 0; t < 5; ++t) *tracks[t] = static_cast<char*>(malloc(len > 0 ? len : 1));
}

static inline void alloc_data(hap_t* v, int len) {
  v->len = len;
  v->_b = static_cast<char*>(malloc(len > 0 ? len : 1));
}

static inline void free_reads(read_t* r, int n) {
  if (!r) return;
  for (int k = 0; k < n; ++k) { free(r[k]._b); free(r[k]._q); free(r[k]._i); free(r[k]._d); free(r[k]._c); }
  free(r);
}

static inline void free_haps(hap_t* h, int n) {
  if (!h) return;
  for (int k = 0; k < n; ++k) free(h[k]._b);
  free(h);
}

// Bytes serialize() will write (not in the reference, which sizes its blocks by the FPGA limits instead).
uint64_t serialized_size(const read_t* reads, int num);
uint64_t serialized_size(const hap_t* haps, int num);

// Into a caller buffer of at least serialized_size() bytes; returns the bytes written.
uint64_t serialize(void* buf, const read_t* reads, int num);
uint64_t serialize(void* buf, const hap_t* haps, int num);

// From a buffer; returns the element count and hands out a malloc'ed array.
int deserialize(const void* buf, read_t*& reads);
int deserialize(const void* buf, hap_t*& haps);

std::string serialize(const read_t* reads, int num);
std::string serialize(const hap_t* haps, int num);

int deserialize(const std::string& data, read_t*& reads);
int deserialize(const std::string& data, hap_t*& haps);

#endif
