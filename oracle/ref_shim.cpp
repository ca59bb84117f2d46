// ref_shim.cpp -- C entry points around the REFERENCE's own PairHMM sources.  TEST INFRASTRUCTURE ONLY.
//
// Compiled by oracle/Makefile together with /root/reference/pairhmm/xlnx/host/{avx_impl,baseline_impl}.cpp,
// from where those files lie (nothing is copied), into oracle/_ref/libpairhmm_ref.so.  This file adds only:
//   * the one definition the reference keeps in a translation unit that needs OpenCL headers
//     (uint8_t ConvertChar::conversionTable[255], FalconPairHMM.cpp:17),
//   * extern "C" wrappers so that tests and bench.py can call the reference through ctypes,
//   * the batch loop of FalconPairHMM::computePairhmmAVX (FalconPairHMM.cpp:69-95) spread over host threads.
// Flush-to-zero is set in every calling thread, as the reference's callers do (pairhmm/host/main.cpp:248).
#include <cmath>
#include <thread>
#include <vector>
#include <xmmintrin.h>

#include "host/avx_impl.h"
#include "host/Context.h"
#include "host/baseline_impl.h"

uint8_t ConvertChar::conversionTable[255];

extern template float  compute_full_prob_baseline<float>(testcase*, float*);   // instantiated in baseline_impl.cpp
extern template double compute_full_prob_baseline<double>(testcase*, double*);

static Context<float>*  g_ctxf = nullptr;
static Context<double>* g_ctxd = nullptr;

static testcase make_tc(int R, int C, const char* rs, const char* q, const char* i, const char* d,
                        const char* c, const char* hap)
{
    testcase tc;
    tc.rslen = R; tc.haplen = C; tc.rs = rs; tc.q = q; tc.i = i; tc.d = d; tc.c = c; tc.hap = hap;
    return tc;
}

struct FtzGuard {
    unsigned old;
    FtzGuard() : old(_mm_getcsr()) { _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON); }
    ~FtzGuard() { _mm_setcsr(old); }
};

extern "C" {

void ref_init()
{
    if (!g_ctxf) {
        ConvertChar::init();
        g_ctxf = new Context<float>();
        g_ctxd = new Context<double>();
    }
}

float ref_avxs(int R, int C, const char* rs, const char* q, const char* i, const char* d, const char* c, const char* hap)
{
    ref_init(); FtzGuard g; testcase tc = make_tc(R, C, rs, q, i, d, c, hap); return compute_fp_avxs(&tc);
}
double ref_avxd(int R, int C, const char* rs, const char* q, const char* i, const char* d, const char* c, const char* hap)
{
    ref_init(); FtzGuard g; testcase tc = make_tc(R, C, rs, q, i, d, c, hap); return compute_fp_avxd(&tc);
}
float ref_baseline_f32(int R, int C, const char* rs, const char* q, const char* i, const char* d, const char* c, const char* hap)
{
    ref_init(); FtzGuard g; testcase tc = make_tc(R, C, rs, q, i, d, c, hap); return compute_full_prob_baseline<float>(&tc, nullptr);
}
double ref_baseline_f64(int R, int C, const char* rs, const char* q, const char* i, const char* d, const char* c, const char* hap)
{
    ref_init(); FtzGuard g; testcase tc = make_tc(R, C, rs, q, i, d, c, hap); return compute_full_prob_baseline<double>(&tc, nullptr);
}

const float*  ref_ph2pr_f32() { ref_init(); return Context<float>::ph2pr; }
const double* ref_ph2pr_f64() { ref_init(); return Context<double>::ph2pr; }
const float*  ref_m2m_f32()   { ref_init(); return Context<float>::matchToMatchProb; }
const double* ref_m2m_f64()   { ref_init(); return Context<double>::matchToMatchProb; }
float  ref_log10_ic_f32() { ref_init(); return Context<float>::LOG10_INITIAL_CONSTANT; }
double ref_log10_ic_f64() { ref_init(); return Context<double>::LOG10_INITIAL_CONSTANT; }

// Same signature as pmm_oracle_batch (oracle/pairhmm_oracle.c).
void ref_batch(int num_read, const int* read_off, const char* rs, const char* q, const char* ins,
               const char* del, const char* gcp, int num_hap, const int* hap_off, const char* hap,
               float* raw, double* out, unsigned char* fb, int float_only, int nthreads)
{
    ref_init();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > num_read) nthreads = num_read > 0 ? num_read : 1;
    auto work = [&](int r0, int r1) {
        FtzGuard g;
        for (int i = r0; i < r1; i++) {
            const int ro = read_off[i], R = read_off[i + 1] - ro;
            for (int h = 0; h < num_hap; h++) {
                const int ho = hap_off[h], C = hap_off[h + 1] - ho;
                const size_t k = (size_t)i * num_hap + h;
                testcase tc = make_tc(R, C, rs + ro, q + ro, ins + ro, del + ro, gcp + ro, hap + ho);
                float f = compute_fp_avxs(&tc);
                if (raw) raw[k] = f;
                bool low = f < MIN_ACCEPTED;
                if (fb) fb[k] = low ? 1 : 0;
                if (float_only || !out) continue;
                if (low) {
                    double dres = compute_fp_avxd(&tc);
                    out[k] = log10(dres) - g_ctxd->LOG10_INITIAL_CONSTANT;
                } else {
                    out[k] = (double)(log10f(f) - g_ctxf->LOG10_INITIAL_CONSTANT);
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t + 1 < nthreads; t++)
        th.emplace_back(work, (int)((long long)num_read * t / nthreads), (int)((long long)num_read * (t + 1) / nthreads));
    work((int)((long long)num_read * (nthreads - 1) / nthreads), num_read);
    for (auto& t : th) t.join();
}

}  // extern "C"
