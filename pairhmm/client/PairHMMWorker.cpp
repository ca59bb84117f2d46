// PairHMMWorker.cpp -- see PairHMMWorker.h.
#include "PairHMMWorker.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
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

// How many tiles (slices of reads) a batch is cut into.  Up to blaze::Client::kMaxInFlight tasks are in flight, each on
// its own accelerator slot (engine context, stream, staging buffers), so all tiles of a typical batch are submitted at
// once: their serialization, packing and host-to-device copies run under the kernels of the first tile, and the
// copy-back and log10 of a finished tile under the kernels of the next.  Only the first tile's way in and the last
// tile's way out stay exposed, which is why the first tile is half the size of the others.  More tiles shrink the
// exposed part but cost GPU efficiency (fewer warp-tasks per launch, more launches), and a small batch gains nothing.
// PAIRHMM_WORKER_TILES overrides (tuning).
// Tried (round 2, tools/plugin_threads.py): one tile per batch while three or more workers of the process are inside run() --
// their batches fill the gaps tiles exist for, and an uncut batch is a more efficient launch.  One process, three / four
// caller threads: 2 126 / 2 131 GCUPS against 2 039 / 2 041 with four tiles.  Eight processes of three threads on a 32-core
// host (one per GPU): 1 600-2 040 per process and noisy, against a steady 1 970-2 030 with four tiles -- an uncut batch
// puts 0.3 ms of serialization and packing on one thread.  The steady one stays.
static int pick_tiles(uint64_t cells, int num_read) {
  if (const char* e = getenv("PAIRHMM_WORKER_TILES")) { const int v = atoi(e); if (v >= 1) return std::min(v, std::max(1, num_read)); }
  if (cells < 1200000000ull) return 1;
  const int t = (int)std::min<uint64_t>(6, std::max<uint64_t>(3, cells / 1100000000ull));
  return std::min(t, std::max(1, num_read));
}

void PairHMMWorker::consume(int row, int n) {
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
  // the task's final doubles (output block 2): nothing is left to do in getOutput()
  if (client_->getNumOutputs() > 2 && client_->getOutputSize(2) >= sizeof(double) * (size_t)n * num_hap_) {
    if (final_.empty()) final_.resize((size_t)num_read_ * num_hap_);
    memcpy(&final_[(size_t)row * num_hap_], client_->getOutputPtr(2), sizeof(double) * (size_t)n * num_hap_);
    final_rows_ += n;
  }
}

void PairHMMWorker::run() {
  fallback_index_.clear(); fallback_value_.clear(); final_.clear(); final_rows_ = 0;
  ran_ = true;
  if (num_read_ == 0 || num_hap_ == 0) return;

  // Rows (reads) per accelerator call.  Upper bound: the engine's job limits.
  uint64_t hap_bytes = 0, hap_bases = 0, read_bases = 0, max_read = 1;
  for (int j = 0; j < num_hap_; ++j) { hap_bytes += (uint64_t)host_haps_[j].len + 4; hap_bases += (uint64_t)host_haps_[j].len; }
  for (int i = 0; i < num_read_; ++i) { max_read = std::max<uint64_t>(max_read, (uint64_t)host_reads_[i].len); read_bases += (uint64_t)host_reads_[i].len; }
  const uint64_t byte_budget = (1ull << 30);                        // half of the 2 GiB limit, for headroom
  if (hap_bytes >= byte_budget) throw std::runtime_error("PairHMMWorker: haplotypes of one batch exceed 1 GiB");
  uint64_t rows = std::min<uint64_t>((byte_budget - hap_bytes) / (5 * max_read + 4), (1ull << 30) / (uint64_t)num_hap_);
  rows = std::max<uint64_t>(1, std::min<uint64_t>(rows, (uint64_t)num_read_));
  const int tiles = pick_tiles(read_bases * hap_bases, num_read_);
  // tile sizes: the first one half a share (2 tiles - 1 half-shares in all), never above the engine's limit
  const uint64_t half_share = std::max<uint64_t>(1, ((uint64_t)num_read_ + 2 * tiles - 2) / (2 * tiles - 1));
  const uint64_t limit = rows;

  // two tasks in flight: submit tile k+1 (and k+2) before waiting for tile k
  std::deque<std::pair<int, int> > flying;                          // (first row, rows) of the tasks in flight, oldest first
  int next = 0;
  static const bool trace = getenv("PAIRHMM_TRACE") != nullptr;     // one line per step on stderr, microseconds since run()
  const auto t0 = std::chrono::steady_clock::now();
  auto us = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
  try {
    while (next < num_read_ || !flying.empty()) {
      while (next < num_read_ && client_->inFlight() < blaze::Client::kMaxInFlight) {
        const uint64_t want = tiles == 1 ? (uint64_t)num_read_ : (next == 0 ? half_share : 2 * half_share);
        const int n = (int)std::min<uint64_t>(std::min<uint64_t>(want, limit), (uint64_t)(num_read_ - next));
        const double a = us();
        client_->setup(&host_reads_[next], n, host_haps_, num_hap_);
        const double b = us();
        client_->startAsync();
        if (trace) fprintf(stderr, "[worker] tile rows %d+%d: setup %.0f..%.0f us, started %.0f\n", next, n, a, b, us());
        flying.emplace_back(next, n);
        next += n;
      }
      const double a = us();
      client_->wait();
      const double b = us();
      consume(flying.front().first, flying.front().second);
      if (trace) fprintf(stderr, "[worker] tile rows %d: waited %.0f..%.0f us, consumed %.0f\n", flying.front().first, a, b, us());
      flying.pop_front();
    }
  } catch (...) {
    // leave nothing in flight behind an exception: the tasks borrow the caller's reads through their input blocks
    while (client_->inFlight() > 0) { try { client_->wait(); } catch (...) {} }
    throw;
  }
}

void PairHMMWorker::getOutput(double* output) {
  if (!ran_) throw std::runtime_error("PairHMMWorker::getOutput() before run()");
  const size_t total = (size_t)num_read_ * (size_t)num_hap_;
  if (final_rows_ == num_read_ && final_.size() == total) {        // every tile came back with the task's final doubles
    memcpy(output, final_.data(), sizeof(double) * total);
    return;
  }
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
