// PairHMMClient.h -- client side of the PairHMM accelerator, with the class and method names of the reference
// (/root/reference/pairhmm/client/PairHMMClient.h:9-30): construct once per thread, setup() a batch, start(),
// read the raw float likelihoods from output block 0.  reads/haps are borrowed, not copied: keep them alive until the
// results have been read.  Not thread-safe; use one client per thread (several may run concurrently and are spread
// over the GPUs by the manager).
//
// Differences from the reference, all on purpose:
//   * no batch-size or length caps (the reference sizes its blocks for 2048 x 192-base reads and 128 x 1024-base
//     haplotypes, client/PairHMMClient.cpp:15-16);
//   * three output blocks: block 1 is the task's fallback list, block 2 its final log10 doubles (task/cuda/PairHMMTask.h);
//   * compute() -- the CPU path Blaze falls back to -- does not exist in this build.  It throws, carrying the
//     accelerator's error message: the B200 engine has no CPU fallback by design.
#ifndef PairHMMCLIENT_H
#define PairHMMCLIENT_H
#include <stdexcept>

#include "blaze/Client.h"
#include "PairHMMHostInterface.h"

class PairHMMClient : public blaze::Client {
 public:
  PairHMMClient();

  void setup(read_t* reads, int num_read,
             hap_t*  haps,  int num_hap);

  void compute();

  int      numRead() const { return num_read_; }
  int      numHap() const { return num_hap_; }
  uint64_t numCell() const { return num_cell_; }

 private:
  int      num_read_;
  int      num_hap_;
  uint64_t num_cell_;
  read_t*  reads_;
  hap_t*   haps_;
};

#endif
