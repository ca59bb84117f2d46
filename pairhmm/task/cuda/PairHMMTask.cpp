// PairHMMTask.cpp -- see PairHMMTask.h.
#include "PairHMMTask.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

PairHMMEngine::PairHMMEngine(int device) : ctx(nullptr) {
  const int rc = pmm_create(device, &ctx);
  if (rc != PMM_OK)
    throw std::runtime_error(std::string("PairHMM CUDA engine unavailable: ") + pmm_last_error(nullptr));
}

PairHMMEngine::~PairHMMEngine() { pmm_destroy(ctx); }

PairHMM::PairHMM() : blaze::Task(3), env(nullptr), num_cell(0), num_read(0), num_hap(0) {}

PairHMM::~PairHMM() {
  if (env) {
    // hand the engine and the output block back for the next task on this device
    env->putScratch("engine", engine_);
    env->putScratch("output", output_);
  }
}

bool PairHMM::conf_flag(const std::string& key, bool dflt) {
  std::string v;
  if (!get_conf(key, v)) return dflt;
  return !(v == "0" || v == "false" || v == "no" || v == "off");
}

void PairHMM::check(int rc, const char* what) {
  if (rc == PMM_OK) return;
  const std::string msg = std::string(what) + ": " + pmm_last_error(engine_ ? engine_->ctx : nullptr);
  if (rc == PMM_ERR_INVALID) throw blaze::invalidParam(msg);
  throw std::runtime_error(msg);
}

void PairHMM::prepare() {
  env = dynamic_cast<blaze::CudaEnv*>(getEnv());
  if (!env) throw blaze::invalidParam("PairHMM task needs a CudaEnv");

  num_cell = *static_cast<uint64_t*>(getInput(0));

  if (!env->getScratch("engine", engine_)) {
    engine_.reset(new PairHMMEngine(env->getDevice()));
    // Slots of one GPU are handed out in order, so the tiles of a batch (client/PairHMMWorker.cpp) land on slots 0, 1, 2, ...:
    // the earlier tile's kernels go first wherever two tiles compete for thread-block slots, and the tiles finish in order
    // instead of all at once -- the worker consumes tile k while tile k+1 is still on the GPU.
    // How the task waits for the GPU: several tiles per client and one client per caller thread mean many waiting threads;
    // the first few of the process spin, the rest sleep, unless told otherwise (conf "sync" or $PAIRHMM_SYNC: spin | block |
    // hybrid | auto).
    std::string sync = "auto";
    if (const char* e = getenv("PAIRHMM_SYNC")) sync = e;
    get_conf("sync", sync);
    check(pmm_set_option(engine_->ctx, "sync", sync.c_str()), "sync");
    if (conf_flag("slot_priority", true))
      check(pmm_set_option(engine_->ctx, "priority", std::to_string(env->getSlot()).c_str()), "priority");
  }

  std::string v;
  if (get_conf("tasks_per_warp", v)) check(pmm_set_option(engine_->ctx, "tasks_per_warp", v.c_str()), "tasks_per_warp");

  // parse the two wire-format blocks, pack into pinned memory, copy to the GPU, build the haplotype stream
  check(pmm_stage_serialized(engine_->ctx, getInput(1), getInputLength(1), getInput(2), getInputLength(2),
                             &num_read, &num_hap), "stage");

  const uint64_t pairs = (uint64_t)num_read * (uint64_t)num_hap;
  if (!env->getScratch("output", output_) || output_->getSize() < pairs * sizeof(float)) {
    // grow-only; a recycled block may be larger than this batch needs, exactly like the reference's fixed
    // 2048 x 128 block (task/xlnx/PairHMMTask.cpp:69-76)
    const uint64_t cap = pairs + pairs / 4 + 1024;
    output_ = env->create_block(1, (int)std::min<uint64_t>(cap, 0x7fffffff), cap * sizeof(float), 4096, blaze::DataBlock::OWNED);
  }
  setOutput(0, output_);
}

void PairHMM::compute() {
  if (!engine_) throw std::runtime_error("PairHMM::compute() before prepare()");
  const uint64_t pairs = (uint64_t)num_read * (uint64_t)num_hap;
  static const bool trace = getenv("PAIRHMM_TRACE") != nullptr;
  const uint64_t t0 = blaze::getUs();

  check(pmm_launch(engine_->ctx), "launch");
  // Output block 2 (extension, see PairHMMTask.h): the final log10 doubles.  Taken first: the engine takes log10f of the
  // raw floats on the host while the double re-run is still on the GPU, so the task's wait ends with the kernels.
  // The three fetches below share one set of device-to-host copies.
  if (conf_flag("emit_log10", true)) {
    blaze::DataBlock_ptr lg = env->create_block(1, (int)std::min<uint64_t>(pairs, 0x7fffffff), pairs * sizeof(double), 64, blaze::DataBlock::OWNED);
    check(pmm_fetch_log10(engine_->ctx, reinterpret_cast<double*>(lg->getData()), pairs, nullptr), "fetch log10");
    setOutput(2, lg);
  }
  check(pmm_fetch_raw(engine_->ctx, reinterpret_cast<float*>(output_->getData()), output_->getSize() / sizeof(float)), "fetch");

  if (conf_flag("emit_fallback", true)) {
    uint64_t n = 0;
    check(pmm_fetch_fallback(engine_->ctx, nullptr, nullptr, 0, &n), "fallback count");
    const size_t idx_bytes = (n * sizeof(uint32_t) + 7) / 8 * 8;
    const size_t bytes = sizeof(uint64_t) + idx_bytes + n * sizeof(double);
    blaze::DataBlock_ptr fb = env->create_block(1, (int)std::min<size_t>(bytes, 0x7fffffff), bytes, 8, blaze::DataBlock::OWNED);
    char* p = fb->getData();
    memcpy(p, &n, sizeof n);
    if (n) {
      check(pmm_fetch_fallback(engine_->ctx, reinterpret_cast<uint32_t*>(p + sizeof(uint64_t)),
                               reinterpret_cast<double*>(p + sizeof(uint64_t) + idx_bytes), n, &n), "fallback list");
    }
    setOutput(1, fb);
  }
  if (trace) {
    pmm_stats_t st; pmm_get_stats(engine_->ctx, &st);
    fprintf(stderr, "[task] %d x %d: compute() %llu us (stage %.0f us before it; kernels f32 %.0f + f64 %.0f us)\n", num_read, num_hap,
            (unsigned long long)(blaze::getUs() - t0), st.ms_stage * 1e3, st.ms_f32 * 1e3, st.ms_fallback * 1e3);
  }
}

extern "C" blaze::Task* create() { return new PairHMM(); }

extern "C" void destroy(blaze::Task* p) { delete p; }
