// FalconSWGpu.cpp -- see FalconSWGpu.h.  One lazily created aligner context per process, like the reference's global
// OpenCL state (/root/reference/htc-sw/host/smithWatermanHost.cpp); not thread-safe, like the reference.
#include "FalconSWGpu.h"

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "smithwaterman_cuda.h"

namespace {
sw_ctx* g_sw = nullptr;
int g_device = 0;

sw_ctx* context() {
  if (!g_sw && sw_create(g_device, &g_sw) != SW_OK)
    throw std::runtime_error(std::string("Smith-Waterman CUDA aligner unavailable: ") + sw_last_error(nullptr));
  return g_sw;
}

// one reference, batchSize alternates stored MAX_SEQ_LENGTH apart
double run(const char* ref, int refLength, const char* alts, size_t alt_pitch, const int* altLengths, int batchSize, int strategy,
           int w_match, int w_mismatch, int w_open, int w_extend, struct Cigar* cigarResults, int* alignmentOffsets) {
  if (batchSize <= 0) return 0.0;
  sw_ctx* c = context();
  std::vector<uint32_t> s1(batchSize, 0), l1(batchSize, (uint32_t)refLength), s2(batchSize), l2(batchSize);
  std::vector<uint8_t> blob2;
  for (int k = 0; k < batchSize; ++k) {
    s2[k] = (uint32_t)blob2.size(); l2[k] = (uint32_t)altLengths[k];
    blob2.insert(blob2.end(), alts + (size_t)k * alt_pitch, alts + (size_t)k * alt_pitch + altLengths[k]);
  }
  std::vector<sw_cigar_elem_t> elems((size_t)batchSize * MAX_SEQ_LENGTH);
  std::vector<int32_t> n(batchSize), off(batchSize);
  if (sw_align_batch(c, (uint32_t)batchSize, reinterpret_cast<const uint8_t*>(ref), s1.data(), l1.data(), blob2.data(), s2.data(),
                     l2.data(), w_match, w_mismatch, w_open, w_extend, strategy, MAX_SEQ_LENGTH, elems.data(), n.data(), off.data(),
                     nullptr) != SW_OK)
    throw std::runtime_error(std::string("sw_align_batch: ") + sw_last_error(c));
  for (int k = 0; k < batchSize; ++k) {
    if (n[k] > MAX_SEQ_LENGTH) throw std::runtime_error("CIGAR longer than struct Cigar holds");
    cigarResults[k].CigarElementNum = n[k];
    for (int e = 0; e < n[k]; ++e) {
      cigarResults[k].cigarElements[e].length = elems[(size_t)k * MAX_SEQ_LENGTH + e].length;
      cigarResults[k].cigarElements[e].state = elems[(size_t)k * MAX_SEQ_LENGTH + e].state;
    }
    alignmentOffsets[k] = off[k];
  }
  sw_stats_t st;
  sw_get_stats(c, &st);
  return (double)st.ms_kernel * 1e6;
}
}  // namespace

int SWPairwiseAlignmentMultiBatch(char* ref, int refLength, char alts[][MAX_SEQ_LENGTH], int batchSize, int* altLengths,
                                  struct Cigar* cigarResults, int* alignmentOffsets, int overhang_strategy, int /*option*/) {
  run(ref, refLength, &alts[0][0], MAX_SEQ_LENGTH, altLengths, batchSize, overhang_strategy, W_MATCH, W_MISMATCH, W_OPEN, W_EXTEND,
      cigarResults, alignmentOffsets);
  return 0;
}

void FalconSWFPGA_init(char* conf) {
  g_device = (conf && !strncmp(conf, "cuda:", 5)) ? atoi(conf + 5) : 0;
  context();
}

double FalconSWFPGA_run(char* ref, int refLength, char alts[][MAX_SEQ_LENGTH], int* altLengths, int batchSize, int overhang_strategy,
                        int w_match, int w_mismatch, int w_open, int w_extend, struct Cigar* cigarResults, int* alignmentOffsets,
                        bool /*unused*/) {
  return run(ref, refLength, &alts[0][0], MAX_SEQ_LENGTH, altLengths, batchSize, overhang_strategy, w_match, w_mismatch, w_open,
             w_extend, cigarResults, alignmentOffsets);
}

void FalconSWFPGA_release() {
  if (g_sw) { sw_destroy(g_sw); g_sw = nullptr; }
}

int32_t runSWOnePairBT_gpu(int32_t match, int32_t mismatch, int32_t open, int32_t extend, uint8_t* seq1, uint8_t* seq2, int32_t len1,
                           int32_t len2, int8_t overhangStrategy, struct Cigar* cigarRet) {
  int off = 0;
  run(reinterpret_cast<const char*>(seq1), len1, reinterpret_cast<const char*>(seq2), 0, &len2, 1, overhangStrategy, match, mismatch,
      open, extend, cigarRet, &off);
  return off;
}
