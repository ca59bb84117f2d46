// PairHMMWorker.cpp -- see PairHMMWorker.h.
#include "PairHMMWorker.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>

#include "pairhmm_cuda.h"

namespace {

// threshold of the float result below which the double result is used (MIN_ACCEPTED, client/PairHMMWorker.cpp:176)
const float kMinAccepted = 1e-28f;

}  // namespace

PairHMMWorker::PairHMMWorker(PairHMMClient* client, int num_read, int num_hap, read_t* reads, hap_t* haps)
    : client_(client), num_read_(num_read), num_hap_(num_hap), host_reads_(reads), host_haps_(haps), ran_(false) {
  if (!client) throw std::invalid_argument("PairHMMWorker: null client");
  if (num_read < 0 || num_hap < 0) throw std::invalid_argument("PairHMMWorker: negative batch size");
  output_.resize((size_t)num_read * (size_t)num_hap);
}

PairHMMWorker::~PairHMMWorker() {}

void PairHMMWorker::compute() {
  throw std::runtime_error("PairHMMWorker::compute(): this build has no CPU compute path");
}

void PairHMMWorker::run() {
  fallback_index_.clear(); fallback_value_.clear();
  ran_ = true;
  if (num_read_ == 0 || num_hap_ == 0) return;

  // Rows (reads) per accelerator call: everything, unless that would cross the engine's job limits.
  uint64_t hap_bytes = 0, max_read = 1;
  for (int j = 0; j < num_hap_; ++j) hap_bytes += (uint64_t)host_haps_[j].len + 4;
  for (int i = 0; i < num_read_; ++i) max_read = std::max<uint64_t>(max_read, (uint64_t)host_reads_[i].len);
  const uint64_t byte_budget = (1ull << 30);                        // half of the 2 GiB limit, for headroom
  if (hap_bytes >= byte_budget) throw std::runtime_error("PairHMMWorker: haplotypes of one batch exceed 1 GiB");
  uint64_t rows = std::min<uint64_t>((byte_budget - hap_bytes) / (5 * max_read + 4), (1ull << 30) / (uint64_t)num_hap_);
  rows = std::max<uint64_t>(1, std::min<uint64_t>(rows, (uint64_t)num_read_));

  for (int row = 0; row < num_read_; row += (int)rows) {
    const int n = std::min<int>((int)rows, num_read_ - row);
    client_->setup(&host_reads_[row], n, host_haps_, num_hap_);
    client_->start();

    const float* results = static_cast<const float*>(client_->getOutputPtr(0));
    memcpy(&output_[(size_t)row * num_hap_], results, sizeof(float) * (size_t)n * num_hap_);

    if (client_->getNumOutputs() > 1 && client_->getOutputSize(1) >= sizeof(uint64_t)) {
      const char* p = static_cast<const char*>(client_->getOutputPtr(1));
      uint64_t nfb; memcpy(&nfb, p, sizeof nfb);
      const uint32_t* idx = reinterpret_cast<const uint32_t*>(p + sizeof(uint64_t));
      const double* val = reinterpret_cast<const double*>(p + sizeof(uint64_t) + (nfb * sizeof(uint32_t) + 7) / 8 * 8);
      for (uint64_t k = 0; k < nfb; ++k) {
        fallback_index_.push_back((uint32_t)((uint64_t)row * num_hap_ + idx[k]));
        fallback_value_.push_back(val[k]);
      }
    }
  }
}

void PairHMMWorker::getOutput(double* output) {
  if (!ran_) throw std::runtime_error("PairHMMWorker::getOutput() before run()");
  const size_t total = (size_t)num_read_ * (size_t)num_hap_;
  // every result below the threshold must have come back with a double re-run (there is no CPU re-run in this build)
  size_t under = 0;
  for (size_t p = 0; p < total; ++p) under += output_[p] < kMinAccepted;
  if (under != fallback_index_.size())
    throw std::runtime_error("PairHMMWorker: float results underflowed but the task returned no double results "
                             "(no CPU re-run in this build)");
  // log10(d) - log10(2^1020) in double for those (client/PairHMMWorker.cpp:184), (double)(log10f(v) - log10f(2^120))
  // with a float subtraction for the rest (:190); host libm, on the engine's host worker threads
  if (pmm_host_finish_log10(output_.data(), total, fallback_index_.data(), fallback_value_.data(),
                            fallback_index_.size(), output) != PMM_OK)
    throw std::runtime_error("pmm_host_finish_log10 failed");
}
