// PairHMMWorker.h -- one batch of reads x haplotypes through a PairHMMClient, returning the final log10
// likelihoods.  Same class and method names as the reference (/root/reference/pairhmm/client/PairHMMWorker.h:9-49):
// construct per batch, run(), getOutput(double*).
//
// What changed against the reference's worker (client/PairHMMWorker.cpp):
//   * there is no CPU routing: the reference keeps small batches (< 32 reads, < 2 haplotypes or < 5e6 cells,
//     :57-66) and everything beyond the FPGA's length limits (:70-93) on the host's AVX code and runs it on a side
//     thread (:214); the GPU has no limits and this build has no CPU compute path at all;
//   * a batch is tiled to pipeline it, not to fit device limits (2048 x 128 in the reference, :217-221): three to six
//     slices of reads with two tasks in flight, so that the host work and copies of one tile run under the kernels of
//     its neighbours (blaze::Client::startAsync / wait); a small batch is one tile;
//   * the double-precision re-run of underflowed pairs already happened on the GPU, and the task delivers the final
//     doubles (output block 2: the same two log10 formulas, host libm, :184,:190); getOutput() copies them.  Against a
//     task that only returns blocks 0 and 1, getOutput() applies the formulas itself to the raw floats and the
//     fallback list instead of calling compute_fp_avxd (:176-184).
#ifndef PAIRHMMWORKER_H
#define PAIRHMMWORKER_H

#include <cstdint>
#include <vector>

#include "PairHMMClient.h"

class PairHMMWorker {
 public:
  PairHMMWorker(PairHMMClient* client,
      int num_read, int num_hap,
      read_t* reads, hap_t* haps);

  ~PairHMMWorker();

  // perform all computation
  void run();

  // NOTE: output must hold num_read * num_hap doubles, read-major
  void getOutput(double* output);

  // the reference's CPU fallback entry; throws in this build
  void compute();

  int numRecalculated() const { return (int)fallback_index_.size(); }

 private:
  PairHMMClient* client_;
  int num_read_;
  int num_hap_;
  read_t* host_reads_;
  hap_t*  host_haps_;
  bool ran_;

  void consume(int row, int n);                     // take the output blocks of the tile that just finished

  std::vector<double> final_;                       // final log10 likelihoods when the task delivers them (output block 2)
  int final_rows_ = 0;
  std::vector<float> output_;                       // raw float likelihoods, read-major
  std::vector<uint32_t> fallback_index_;            // positions in output_ that took the double re-run on the GPU
  std::vector<double>   fallback_value_;            // their double likelihoods (scaled by 2^1020)
};

#endif
