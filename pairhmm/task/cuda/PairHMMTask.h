// PairHMMTask.h -- the Blaze task plugin of the PairHMM path on CUDA.
//
// Drop-in for the reference's FPGA task (/root/reference/pairhmm/task/xlnx/PairHMMTask.h:58-96,
// task/xlnx/PairHMMTask.cpp:14-143): same class name, same three input blocks -- in0 `uint64 num_cell`, in1
// serialized reads, in2 serialized haplotypes -- same output block 0 of `float[num_read * num_hap]` read-major raw
// likelihoods scaled by 2^120, same prepare()/compute() split (host work + H2D, then kernels + wait), same
// create()/destroy() entry points, same recycling of per-device buffers through the TaskEnv's scratch table.
// Where the reference packs an FpgaInputBundle and enqueues an OpenCL kernel, this task hands the two serialized
// blocks to the C ABI of include/pairhmm_cuda.h.  There are no length or batch-size limits.
//
// Extension: output block 1 (fallback list).  The GPU also re-runs every pair whose float result is below 1e-28f in
// double precision (what the reference's client does on the CPU afterwards, client/PairHMMWorker.cpp:176-184); block 1
// carries `uint64 n`, then n x `uint32 index` (position in block 0), padding to 8 bytes, then n x `double` likelihoods
// scaled by 2^1020.  Output block 2 (conf "emit_log10", default on): `double[num_read * num_hap]`, the final log10
// likelihoods of PairHMMWorker::getOutput (client/PairHMMWorker.cpp:157-197) -- log10f of the floats is taken by the
// host libm while the double re-run is still on the GPU.  A client that only reads block 0 behaves exactly like the
// reference's.
#ifndef PAIRHMM_TASK_H
#define PAIRHMM_TASK_H

#include <memory>
#include <string>

#include "blaze/Block.h"
#include "blaze/Task.h"
#include "blaze/TaskEnv.h"
#include "pairhmm_cuda.h"

// One engine context (device arenas, pinned staging buffers, stream) -- the object a task borrows from its
// TaskEnv and puts back when it is destroyed, like the reference's PairHMMInput (task/xlnx/PairHMMTask.h:15-53).
class PairHMMEngine {
 public:
  explicit PairHMMEngine(int device);
  ~PairHMMEngine();
  PairHMMEngine(const PairHMMEngine&) = delete;
  PairHMMEngine& operator=(const PairHMMEngine&) = delete;
  pmm_ctx* ctx;
};
typedef std::shared_ptr<PairHMMEngine> PairHMMEngine_ptr;

class PairHMM : public blaze::Task {
 public:
  PairHMM();
  virtual ~PairHMM();

  virtual uint64_t estimateClientTime() { return 0; }
  virtual uint64_t estimateTaskTime() { return 0; }

  virtual void prepare();
  virtual void compute();

 private:
  bool conf_flag(const std::string& key, bool dflt);
  void check(int rc, const char* what);

  blaze::CudaEnv* env;
  uint64_t num_cell;
  int      num_read;
  int      num_hap;
  PairHMMEngine_ptr    engine_;
  blaze::DataBlock_ptr output_;
};

extern "C" blaze::Task* create();
extern "C" void destroy(blaze::Task* p);

#endif
