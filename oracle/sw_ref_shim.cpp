// sw_ref_shim.cpp -- C entry points around the REFERENCE's own Smith-Waterman sources.  TEST INFRASTRUCTURE ONLY.
//
// Compiled by oracle/Makefile together with /root/reference/htc-sw/intel_avx/avx2_impl.cc (the AVX2 kernel with
// backtrack, runSWOnePairBT_fp_avx2) and /root/reference/htc-sw/host/FalconSW_AVX.cpp (Falcon's own
// SWPairwiseAlignmentOneBatch, the golden side of the reference's test, host/sw_host.cpp:253-264), from where those
// files lie, into oracle/_ref/libsw_ref.so.  This file adds the globals FalconSW_AVX.cpp declares extern and thin
// extern "C" wrappers that flatten struct Cigar.
#include <cstdint>
#include <cstring>

#include "host/common.h"
#include "intel_avx/avx2_impl.h"

struct Cigar* retCigarBatch = nullptr;
long SWPairwiseAlignment_C_time = 0, isSWFailure_C_time = 0, trimCigarByBases_C_time = 0, getReferenceLength_C_time = 0,
     leftAlignCigar_C_time = 0, calMatrix_C_time = 0, calCigar_C_time = 0, malloc_time = 0, SW_complexity = 0,
     lastKernel_C_time = 0;

struct timespec diff_time(struct timespec start, struct timespec end)
{
    struct timespec t;
    t.tv_sec = end.tv_sec - start.tv_sec; t.tv_nsec = end.tv_nsec - start.tv_nsec;
    if (t.tv_nsec < 0) { t.tv_sec -= 1; t.tv_nsec += 1000000000L; }
    return t;
}

int addCigarElement(struct Cigar* cigar, int length, int state)       // host/sw_host.cpp:18-27 (needed by FalconSW_AVX.cpp)
{
    if (cigar->CigarElementNum < 0) return -1;
    if (length > 0) {
        cigar->cigarElements[cigar->CigarElementNum].length = length;
        cigar->cigarElements[cigar->CigarElementNum].state = state;
        cigar->CigarElementNum++;
    }
    return 0;
}

static int flatten(const struct Cigar& c, int* len, int* state, int cap)
{
    for (int k = 0; k < c.CigarElementNum && k < cap; ++k) { len[k] = c.cigarElements[k].length; state[k] = c.cigarElements[k].state; }
    return c.CigarElementNum;
}

extern "C" int ref_sw_gkl(int match, int mismatch, int open, int extend, const uint8_t* seq1, int len1, const uint8_t* seq2,
                          int len2, int strategy, int* cigar_len, int* cigar_state, int cap, int* n_elem)
{
    static thread_local struct Cigar c;
    c.CigarElementNum = 0;
    const int off = runSWOnePairBT_fp_avx2(match, mismatch, open, extend, const_cast<uint8_t*>(seq1), const_cast<uint8_t*>(seq2),
                                           len1, len2, (int8_t)strategy, &c);
    *n_elem = flatten(c, cigar_len, cigar_state, cap);
    return off;
}

// Falcon's implementation uses the compile-time weights of host/common.h:15-18 (200, -150, -260, -11).
extern "C" int ref_sw_falcon(const uint8_t* ref, int ref_len, const uint8_t* alt, int alt_len, int strategy, int option,
                             int* cigar_len, int* cigar_state, int cap, int* n_elem, int* rc)
{
    static thread_local struct Cigar c;
    c.CigarElementNum = 0;
    int off = 0;
    *rc = SWPairwiseAlignmentOneBatch(reinterpret_cast<char*>(const_cast<uint8_t*>(ref)), reinterpret_cast<char*>(const_cast<uint8_t*>(alt)),
                                      ref_len, alt_len, &c, &off, strategy, option);
    *n_elem = flatten(c, cigar_len, cigar_state, cap);
    return off;
}
