// hostif_ref_shim.cpp -- C entry points around the REFERENCE's own wire-format code
// (/root/reference/pairhmm/interface/PairHMMHostInterface.cpp:175-340, compiled from where it lies into
// oracle/_ref/libhostif_ref.so by oracle/Makefile).  TEST INFRASTRUCTURE ONLY: tests/test_wire_format_ref.py pins the
// repo's serialize()/deserialize() (pairhmm/interface) and the Python writer (acc_genomics_b200/batch.py) to these.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "PairHMMHostInterface.h"      // the reference's header (-I$(REF)/pairhmm/interface)

namespace {
std::vector<read_t> view_reads(int num, const int* off, const char* b, const char* q, const char* i, const char* d, const char* c)
{
    std::vector<read_t> r(num);
    for (int k = 0; k < num; ++k) {
        r[k].len = off[k + 1] - off[k];
        r[k]._b = const_cast<char*>(b + off[k]); r[k]._q = const_cast<char*>(q + off[k]); r[k]._i = const_cast<char*>(i + off[k]);
        r[k]._d = const_cast<char*>(d + off[k]); r[k]._c = const_cast<char*>(c + off[k]);
    }
    return r;
}
std::vector<hap_t> view_haps(int num, const int* off, const char* b)
{
    std::vector<hap_t> h(num);
    for (int k = 0; k < num; ++k) { h[k].len = off[k + 1] - off[k]; h[k]._b = const_cast<char*>(b + off[k]); }
    return h;
}
}  // namespace

extern "C" {

// serialize(void*, const read_t*, int) / serialize(void*, const hap_t*, int): bytes written
uint64_t refif_serialize_reads(void* buf, int num, const int* off, const char* b, const char* q, const char* i, const char* d, const char* c)
{
    std::vector<read_t> r = view_reads(num, off, b, q, i, d, c);
    return serialize(buf, r.data(), num);
}
uint64_t refif_serialize_haps(void* buf, int num, const int* off, const char* b)
{
    std::vector<hap_t> h = view_haps(num, off, b);
    return serialize(buf, h.data(), num);
}
// the std::string overloads
uint64_t refif_serialize_reads_str(void* buf, uint64_t cap, int num, const int* off, const char* b, const char* q, const char* i, const char* d, const char* c)
{
    std::vector<read_t> r = view_reads(num, off, b, q, i, d, c);
    const std::string s = serialize(r.data(), num);
    if (s.size() <= cap) memcpy(buf, s.data(), s.size());
    return s.size();
}
uint64_t refif_serialize_haps_str(void* buf, uint64_t cap, int num, const int* off, const char* b)
{
    std::vector<hap_t> h = view_haps(num, off, b);
    const std::string s = serialize(h.data(), num);
    if (s.size() <= cap) memcpy(buf, s.data(), s.size());
    return s.size();
}
// deserialize(const void*, read_t*&) followed by serialize(void*, ...): the reference reading a foreign writer's bytes
// and writing them back.  which = 0: raw-pointer overloads, 1: std::string overloads.  Returns bytes written; *num_out = count.
uint64_t refif_reserialize_reads(const void* in, uint64_t in_bytes, void* out, int which, int* num_out)
{
    read_t* r = nullptr;
    const int n = which ? deserialize(std::string(static_cast<const char*>(in), in_bytes), r) : deserialize(in, r);
    uint64_t w;
    if (which) { const std::string s = serialize(r, n); memcpy(out, s.data(), s.size()); w = s.size(); }
    else w = serialize(out, r, n);
    if (num_out) *num_out = n;
    if (n > 0) free_reads(r, n); else free(r);
    return w;
}
uint64_t refif_reserialize_haps(const void* in, uint64_t in_bytes, void* out, int which, int* num_out)
{
    hap_t* h = nullptr;
    const int n = which ? deserialize(std::string(static_cast<const char*>(in), in_bytes), h) : deserialize(in, h);
    uint64_t w;
    if (which) { const std::string s = serialize(h, n); memcpy(out, s.data(), s.size()); w = s.size(); }
    else w = serialize(out, h, n);
    if (num_out) *num_out = n;
    if (n > 0) free_haps(h, n); else free(h);
    return w;
}

}  // extern "C"
