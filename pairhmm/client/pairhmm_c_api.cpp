// pairhmm_c_api.cpp -- see pairhmm_c_api.h.
#include "pairhmm_c_api.h"

#include <cstdio>
#include <memory>
#include <vector>

#include "PairHMMManager.h"
#include "PairHMMWorker.h"

static std::unique_ptr<PairHMMClient>& thread_client() {
  static thread_local std::unique_ptr<PairHMMClient> client;
  return client;
}

extern "C" int pairhmm_worker_forward(int num_read, const int32_t* read_off, const char* bases, const char* q, const char* i,
                                      const char* d, const char* c, int num_hap, const int32_t* hap_off, const char* hap,
                                      double* out, int* num_recalc, char* err, int err_capacity) {
  try {
    if (num_read <= 0 || num_hap <= 0 || !read_off || !hap_off || !bases || !q || !i || !d || !c || !hap || !out)
      throw std::invalid_argument("pairhmm_worker_forward: null or empty input");
    if (!blaze::AppCommManager::lookup(1027)) pairhmm_default_manager();
    std::unique_ptr<PairHMMClient>& client = thread_client();
    if (!client) client.reset(new PairHMMClient());
    // read_t / hap_t views of the caller's arrays: nothing is copied until PairHMMClient::setup serializes a tile
    static thread_local std::vector<read_t> reads;
    static thread_local std::vector<hap_t> haps;
    reads.resize(num_read); haps.resize(num_hap);
    for (int k = 0; k < num_read; ++k) {
      const int32_t o = read_off[k];
      reads[k].len = read_off[k + 1] - o;
      reads[k]._b = const_cast<char*>(bases + o); reads[k]._q = const_cast<char*>(q + o); reads[k]._i = const_cast<char*>(i + o);
      reads[k]._d = const_cast<char*>(d + o); reads[k]._c = const_cast<char*>(c + o);
    }
    for (int k = 0; k < num_hap; ++k) { haps[k].len = hap_off[k + 1] - hap_off[k]; haps[k]._b = const_cast<char*>(hap + hap_off[k]); }
    PairHMMWorker worker(client.get(), num_read, num_hap, reads.data(), haps.data());
    worker.run();
    worker.getOutput(out);
    if (num_recalc) *num_recalc = worker.numRecalculated();
    return 0;
  } catch (const std::exception& e) {
    if (err && err_capacity > 0) snprintf(err, (size_t)err_capacity, "%s", e.what());
    return 1;
  }
}

extern "C" void pairhmm_worker_shutdown(void) {
  thread_client().reset();            // the calling thread's client (other threads' clients end with their threads)
  pairhmm_shutdown_manager();
}
