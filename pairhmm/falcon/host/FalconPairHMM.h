// FalconPairHMM.h -- the standalone batch entry of the PairHMM path (what Falcon's GATK fork links when it does not go
// through Blaze), with the class and method names of /root/reference/pairhmm/xlnx/host/FalconPairHMM.h:17-65, over
// the engine's C ABI (include/pairhmm_cuda.h).
//
//   FalconPairHMM f;  or  FalconPairHMM f("cuda:1");      the string stands where the reference takes a bitstream path
//   bool used = false;
//   f.computePairhmm(&input, &output, used);              output.likelihoodData[r * haps + h] = log10 likelihood
//
// Results are those of the reference's computePairhmmAVX (FalconPairHMM.cpp:69-95): float pass, `< 1e-28f` test,
// double re-run, log10f / log10 of the host libm -- bit for bit.  (The reference's *FPGA* branch takes the double
// log10 of the float result instead of log10f, FalconPairHMM.cpp:656; SURVEY.md says to follow the AVX form, which
// is also what the Blaze worker does.)
//
// Differences by design: there is no CPU compute path in this product, so computePairhmmAVX / computePairhmmBaseline
// throw std::runtime_error, the routing test worthFPGA() has nothing to route to and only rejects empty batches, and
// the FPGA limits (192-base reads, 1024-base haplotypes, batch sizes) do not exist.  Not thread-safe per object, like
// the reference; use one object per host thread.
#ifndef FALCONPAIRHMM_H
#define FALCONPAIRHMM_H
#include <cstdint>
#include <vector>

#include "host_type.h"

struct pmm_ctx;

class FalconPairHMM {
 public:
  FalconPairHMM();                                  // GPU 0 (or the CUDA current device)
  explicit FalconPairHMM(const char* conf);         // NULL, "", "-" = default device; "cuda:N" = GPU N
  ~FalconPairHMM();
  FalconPairHMM(const FalconPairHMM&) = delete;
  FalconPairHMM& operator=(const FalconPairHMM&) = delete;

  // usedFPGA is set to true when the accelerator computed the batch (always, for a non-empty batch)
  void computePairhmm(pairhmmInput* input, pairhmmOutput* output, bool& usedFPGA);
  int computePairhmmFalcon(pairhmmInput* input, pairhmmOutput* output, bool& usedFPGA);
  int computePairhmmAVX(pairhmmInput* input, pairhmmOutput* output, bool use_double);        // throws: no CPU path
  int computePairhmmBaseline(pairhmmInput* input, pairhmmOutput* output, bool use_double);   // throws: no CPU path
  double get_kernel_time();                         // accumulated device time of the kernels, nanoseconds (:1180)

  // not in the reference: counters of the last batch
  uint64_t last_fallback_pairs() const { return last_fallback_; }
  double peak_kernel_gcups() const { return peak_kernel_gcups_; }

 private:
  void open(const char* conf);
  pmm_ctx* ctx_ = nullptr;
  std::vector<char> scratch_;                       // pmm_read_t / pmm_hap_t views of the caller's strings
  double kernel_time_ = 0;
  double peak_kernel_gcups_ = 0;
  uint64_t last_fallback_ = 0;
};

// Cells of the batch as the reference counts them (FalconPairHMM.cpp:97-110); no length limits to violate here.
double countCell(pairhmmInput* input, short maxCols, bool& violate);
// The reference weighs FPGA time against AVX time (FalconPairHMM.cpp:112-139).  With no CPU path the answer is
// "yes" for every batch that has at least one pair.
bool worthFPGA(pairhmmInput* input, short maxCols, double cellNum);

#endif
