// fixture_io.h -- the on-disk batch format of the reference's test bench (a folder of input<i> / output<i> text
// files dumped from GATK runs; reader in /root/reference/pairhmm/host/main.cpp:67-159):
//
//   input<i>   line 1      four tokens, the 2nd is num_read and the 4th num_hap ("readListSize 3 numHaplotypes 2")
//              per read    a line with the read length, then for each of _b, _q, _i, _d, _c a caption line (ignored)
//                          followed by a line of `length` decimal byte values
//              one line    ignored (blank)
//              per hap     a line with the length, a caption line (ignored), a line with the bases as characters
//   output<i>  per pair    two tokens: the log10 likelihood in decimal and its IEEE-754 bit pattern as a signed
//                          64-bit integer; the bit pattern is what is compared
//
// acc_genomics_b200/fixtures.py writes and reads the same files from Python.
#ifndef PAIRHMM_FIXTURE_IO_H
#define PAIRHMM_FIXTURE_IO_H
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "PairHMMHostInterface.h"

namespace fixture {

inline std::vector<std::string> split_line(std::istream& in) {
  std::string line;
  std::vector<std::string> tok;
  if (!std::getline(in, line)) return tok;
  std::istringstream ss(line);
  for (std::string t; ss >> t;) tok.push_back(t);
  return tok;
}

inline int to_int(const std::string& s, const char* what) {
  try { return std::stoi(s); } catch (...) { throw std::runtime_error(std::string("fixture: bad ") + what + ": " + s); }
}

// reads / haps are malloc'ed like deserialize() does; release with free_reads / free_haps
inline void read_input(const std::string& path, int& num_read, int& num_hap, read_t*& reads, hap_t*& haps) {
  std::ifstream in(path.c_str());
  if (!in.good()) throw std::runtime_error("fixture: cannot open " + path);
  std::vector<std::string> head = split_line(in);
  if (head.size() != 4) throw std::runtime_error("fixture: bad header in " + path);
  num_read = to_int(head[1], "read count");
  num_hap = to_int(head[3], "haplotype count");
  if (num_read < 0 || num_hap < 0) throw std::runtime_error("fixture: negative count in " + path);
  reads = static_cast<read_t*>(calloc(num_read ? num_read : 1, sizeof(read_t)));
  haps = static_cast<hap_t*>(calloc(num_hap ? num_hap : 1, sizeof(hap_t)));
  std::string skip;
  for (int r = 0; r < num_read; ++r) {
    std::vector<std::string> t = split_line(in);
    if (t.size() != 1) throw std::runtime_error("fixture: expected a read length in " + path);
    const int len = to_int(t[0], "read length");
    alloc_data(&reads[r], len);
    char* track[5] = {reads[r]._b, reads[r]._q, reads[r]._i, reads[r]._d, reads[r]._c};
    for (int k = 0; k < 5; ++k) {
      std::getline(in, skip);                                   // caption
      std::vector<std::string> v = split_line(in);
      if ((int)v.size() != len) throw std::runtime_error("fixture: track length mismatch in " + path);
      for (int p = 0; p < len; ++p) track[k][p] = (char)to_int(v[p], "byte value");
    }
  }
  std::getline(in, skip);                                       // separator line
  for (int h = 0; h < num_hap; ++h) {
    std::vector<std::string> t = split_line(in);
    if (t.size() != 1) throw std::runtime_error("fixture: expected a haplotype length in " + path);
    const int len = to_int(t[0], "haplotype length");
    std::getline(in, skip);                                     // caption
    std::string bases;
    std::getline(in, bases);
    if ((int)bases.size() != len) throw std::runtime_error("fixture: haplotype length mismatch in " + path);
    alloc_data(&haps[h], len);
    memcpy(haps[h]._b, bases.data(), (size_t)len);
  }
}

inline void read_output(const std::string& path, double* likelihood, int size) {
  std::ifstream in(path.c_str());
  if (!in.good()) throw std::runtime_error("fixture: cannot open " + path);
  for (int k = 0; k < size; ++k) {
    double text; long long bits;
    if (!(in >> text >> bits)) throw std::runtime_error("fixture: truncated " + path);
    memcpy(&likelihood[k], &bits, sizeof(double));
  }
}

}  // namespace fixture
#endif
