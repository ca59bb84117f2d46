// FalconPairHMM.cpp -- see FalconPairHMM.h.
#include "FalconPairHMM.h"

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "pairhmm_cuda.h"

namespace {
[[noreturn]] void no_cpu_path(const char* who) {
  throw std::runtime_error(std::string(who) + ": this build has no CPU compute path; use computePairhmm()");
}
}  // namespace

void FalconPairHMM::open(const char* conf) {
  int device = -1;
  if (conf && *conf && strcmp(conf, "-") != 0) {
    if (strncmp(conf, "cuda:", 5) != 0 || !conf[5]) throw std::runtime_error(std::string("FalconPairHMM: expected \"cuda:N\", got ") + conf);
    char* end = nullptr;
    device = (int)strtol(conf + 5, &end, 10);
    if (*end || device < 0) throw std::runtime_error(std::string("FalconPairHMM: bad device in ") + conf);
  }
  const int rc = pmm_create(device, &ctx_);
  if (rc != PMM_OK) throw std::runtime_error(std::string("FalconPairHMM: ") + pmm_last_error(nullptr));
}

FalconPairHMM::FalconPairHMM() { open(nullptr); }
FalconPairHMM::FalconPairHMM(const char* conf) { open(conf); }
FalconPairHMM::~FalconPairHMM() { pmm_destroy(ctx_); }

double FalconPairHMM::get_kernel_time() { return kernel_time_; }

void FalconPairHMM::computePairhmm(pairhmmInput* input, pairhmmOutput* output, bool& usedFPGA) {
  computePairhmmFalcon(input, output, usedFPGA);
}

int FalconPairHMM::computePairhmmFalcon(pairhmmInput* input, pairhmmOutput* output, bool& usedFPGA) {
  if (!input || !output) throw std::invalid_argument("FalconPairHMM: null batch");
  output->likelihoodData.clear();
  last_fallback_ = 0;
  const size_t nr = input->reads.size(), nh = input->haps.size();
  if (nr == 0 || nh == 0) { usedFPGA = false; return 0; }       // nothing to compute (the reference's loops do not run either)
  // views of the caller's strings; nothing is copied on this side of the C ABI
  scratch_.resize(nr * sizeof(pmm_read_t) + nh * sizeof(pmm_hap_t));
  pmm_read_t* r = reinterpret_cast<pmm_read_t*>(scratch_.data());
  pmm_hap_t* h = reinterpret_cast<pmm_hap_t*>(scratch_.data() + nr * sizeof(pmm_read_t));
  for (size_t k = 0; k < nr; ++k) {
    Read& s = input->reads[k];
    const size_t len = s.bases.size();
    if (s._q.size() != len || s._i.size() != len || s._d.size() != len || s._c.size() != len)
      throw std::invalid_argument("FalconPairHMM: read " + std::to_string(k) + " has tracks of different lengths");
    r[k].len = (int)len;
    r[k]._b = &s.bases[0]; r[k]._q = &s._q[0]; r[k]._i = &s._i[0]; r[k]._d = &s._d[0]; r[k]._c = &s._c[0];
  }
  for (size_t k = 0; k < nh; ++k) { h[k].len = (int)input->haps[k].bases.size(); h[k]._b = &input->haps[k].bases[0]; }
  output->likelihoodData.resize(nr * nh);
  uint64_t nfb = 0;
  const int rc = pmm_forward_log10(ctx_, r, (int)nr, h, (int)nh, output->likelihoodData.data(), &nfb);
  if (rc != PMM_OK) {
    output->likelihoodData.clear();
    throw std::runtime_error(std::string("FalconPairHMM: ") + pmm_last_error(ctx_));
  }
  last_fallback_ = nfb;
  usedFPGA = true;
  pmm_stats_t st;
  if (pmm_get_stats(ctx_, &st) == PMM_OK) {
    const double ns = ((double)st.ms_f32 + (double)st.ms_fallback) * 1e6;
    kernel_time_ += ns;
    if (ns > 0 && (double)st.cells / ns > peak_kernel_gcups_) peak_kernel_gcups_ = (double)st.cells / ns;
  }
  return 0;
}

int FalconPairHMM::computePairhmmAVX(pairhmmInput*, pairhmmOutput*, bool) { no_cpu_path("computePairhmmAVX"); }
int FalconPairHMM::computePairhmmBaseline(pairhmmInput*, pairhmmOutput*, bool) { no_cpu_path("computePairhmmBaseline"); }

double countCell(pairhmmInput* input, short, bool& violate) {
  violate = false;
  double reads = 0, haps = 0;
  for (const Read& r : input->reads) reads += (double)r.bases.size();
  for (const Hap& h : input->haps) haps += (double)h.bases.size();
  return reads * haps;                                            // = sum over pairs of read_len * hap_len
}

bool worthFPGA(pairhmmInput* input, short, double) { return input && !input->reads.empty() && !input->haps.empty(); }
