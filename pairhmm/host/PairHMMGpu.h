// PairHMMGpu.h -- direct dispatch of one serialized batch to the GPU, the entry the reference's test bench uses
// when it is built without the Blaze client (/root/reference/pairhmm/host/PairHMMFpga.h:16-22, host/main.cpp:322-334).
//
//   float* compute_gpu(conf, read_data, hap_data, num_cell)
// takes the two strings produced by serialize() and returns the raw float likelihoods (scaled by 2^120), read-major,
// in a buffer owned by this library that stays valid until the next call (the reference returns its static ret_buf,
// host/PairHMMFpga.cpp:153-161).  Like the reference it is single-threaded and non-reentrant: one lazily created
// global engine, released by cleanup() at process exit.  `conf` stands where the reference takes the bitstream path:
// NULL, "" or "-" selects GPU 0, "cuda:N" selects GPU N.  Errors throw std::runtime_error.
//
// compute_fpga() is kept as an alias with the reference's exact signature so that its main.cpp links unchanged.
#ifndef PAIRHMM_GPU_H
#define PAIRHMM_GPU_H

#include <cstdint>
#include <string>

#include "PairHMMHostInterface.h"

extern double peak_kernel_gcups;   // best and latest kernel-only GCUPS seen by compute_gpu (host/PairHMMFpga.h:13-14)
extern double curr_kernel_gcups;

float* compute_gpu(const char* conf, std::string read_data, std::string hap_data, uint64_t num_cell);

float* compute_fpga(const char* bit_path, std::string read_data, std::string hap_data, uint64_t num_cell);

// The fallback list of the last compute_gpu call: pairs whose float result is below 1e-28f and their
// double-precision likelihoods (scaled by 2^1020), computed on the GPU in the same call.
uint64_t last_fallback(const uint32_t** index, const double** value);

void __attribute__((destructor)) cleanup();

#endif
