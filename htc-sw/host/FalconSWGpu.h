// FalconSWGpu.h -- the Smith-Waterman host entry points of the reference, served by the B200 aligner.
//
//   SWPairwiseAlignmentMultiBatch   one reference against a batch of alternates, the call GATK's haplotype-to-reference
//                                   alignment makes (/root/reference/htc-sw/host/FalconSW_AVX.cpp:304-313); `option`
//                                   selected among the reference's CPU variants and is ignored here
//   FalconSWFPGA_init / _run / _release   the accelerator dispatch the reference's bench calls (host/sw_host.cpp:13-15,
//                                   :255-259); init takes "cuda:N" (or anything else for GPU 0) where the reference took
//                                   a bitstream path; run returns the kernel time in nanoseconds like the reference's
//   runSWOnePairBT_gpu              one pair with explicit weights, the signature of runSWOnePairBT_fp_avx2
//                                   (intel_avx/avx2_impl.h:6); for tests -- a single pair cannot fill a GPU
// All of them return the same alignment offsets and CIGARs as the reference.  Sequences longer than MAX_SEQ_LENGTH are
// accepted up to 4095 bases (the CIGAR struct still holds at most MAX_SEQ_LENGTH elements).  Errors throw
// std::runtime_error; there is no CPU fallback.
#ifndef FALCON_SW_GPU_H
#define FALCON_SW_GPU_H
#include <cstdint>

#include "common.h"

int SWPairwiseAlignmentMultiBatch(char* ref, int refLength, char alts[][MAX_SEQ_LENGTH], int batchSize, int* altLengths,
                                  struct Cigar* cigarResults, int* alignmentOffsets, int overhang_strategy, int option);

void   FalconSWFPGA_init(char* conf);
double FalconSWFPGA_run(char* ref, int refLength, char alts[][MAX_SEQ_LENGTH], int* altLengths, int batchSize,
                        int overhang_strategy, int w_match, int w_mismatch, int w_open, int w_extend,
                        struct Cigar* cigarResults, int* alignmentOffsets, bool unused);
void   FalconSWFPGA_release();

int32_t runSWOnePairBT_gpu(int32_t match, int32_t mismatch, int32_t open, int32_t extend, uint8_t* seq1, uint8_t* seq2,
                           int32_t len1, int32_t len2, int8_t overhangStrategy, struct Cigar* cigarRet);

#endif
