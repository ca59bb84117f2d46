/*
 * pairhmm_oracle.c -- CPU restatement of the reference PairHMM forward path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The product path (acc_genomics_b200/csrc) never links, imports or calls it.
 *
 * What it restates (all paths relative to /root/reference/pairhmm/xlnx/host unless noted):
 *   - probability tables                        Context.h:42-61 (jacobian, matchToMatch), :67-90, :105-110, :145-151
 *   - per-row transition parameters             avx-pairhmm-template.h:83-128 (initializeVectors), :155-158
 *   - base classes and the match rule           host_type.h:123-143 (ConvertChar), avx-pairhmm-template.h:3-35, :70-75
 *   - boundary conditions                       avx-pairhmm-template.h:136-177 (stripeINITIALIZATION)
 *   - the cell update, operation order          avx-pairhmm-template.h:183-198 (computeMXY)
 *   - the final reduction                       avx-pairhmm-template.h:308-343
 *   - threshold, double re-run and log10        FalconPairHMM.cpp:69-95 == ../../client/PairHMMWorker.cpp:171-193
 *
 * The restatement is scalar and row-major, but reproduces the *arithmetic* of the AVX implementation bit for
 * bit.  One oddity of the AVX striping needs no special case: for every stripe after the first,
 * stripeINITIALIZATION sets M_t_1_y = M_t_1 = {shiftOutM[AVX_LENGTH], 0, ...} (avx-pairhmm-template.h:171-176),
 * so the first row of the stripe uses M[r-1][1] instead of M[r][0] as its "left M" in column 1 -- but M[r-1][1]
 * is exactly 0 for r-1 >= 2 (column 0 of every row below the first is all zeros), so nothing changes
 * (tests/test_oracle.py::test_stripe_artefact_is_void).
 *
 * Pinning: oracle/_ref/libpairhmm_ref.so is the reference's own source compiled with -O3 -mavx
 * -ffp-contract=off (see oracle/Makefile); tests/test_oracle.py checks this file against it bit for bit, and
 * against the committed golden vectors in tests/golden/ (generated from that library by
 * tests/golden/make_golden.py).  The reference ships no golden vectors of its own.
 *
 * x86 flush-to-zero is switched on around every entry point, as the reference's callers do
 * (../../host/main.cpp:248, FalconPairHMM.cpp:850).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>

#define MAX_QUAL 254
#define JAC_SIZE 80001                 /* (int)(8.0 / 0.0001) + 1            Context.h:8-11 */
#define M2M_SIZE (((MAX_QUAL + 1) * (MAX_QUAL + 2)) >> 1)

static float  ph2pr_f[128], jac_f[JAC_SIZE], m2m_f[M2M_SIZE], IC_f, LIC_f;
static double ph2pr_d[128], jac_d[JAC_SIZE], m2m_d[M2M_SIZE], IC_d, LIC_d;
static unsigned char cls_tab[256];
static int g_init = 0;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

/* Context.h:63-65 */
static int fast_round_f(float d)  { return (d > 0.0f) ? (int)(d + 0.5f) : (int)(d - 0.5f); }
static int fast_round_d(double d) { return (d > 0.0)  ? (int)(d + 0.5)  : (int)(d - 0.5); }

/* Context.h:67-90, NUMBER = float.  Arguments arrive already narrowed to float. */
static float approx_log10_sum_f(float small, float big)
{
    if (small > big) { float t = big; big = small; small = t; }
    if (isinf(small) == -1 || isinf(big) == -1) return big;
    float diff = big - small;
    if (diff >= 8.0f) return big;
    int ind = fast_round_f((float)(diff * 10000.0f));
    return big + jac_f[ind];
}
static double approx_log10_sum_d(double small, double big)
{
    if (small > big) { double t = big; big = small; small = t; }
    if (isinf(small) == -1 || isinf(big) == -1) return big;
    double diff = big - small;
    if (diff >= 8.0) return big;
    int ind = fast_round_d(diff * 10000.0);
    return big + jac_d[ind];
}

void pmm_oracle_init(void)
{
    pthread_mutex_lock(&g_lock);
    if (!g_init) {
        /* Context.h:42-47 */
        for (int k = 0; k < JAC_SIZE; k++) {
            double v = log10(1.0 + pow(10.0, -((double)k) * 0.0001));
            jac_f[k] = (float)v;
            jac_d[k] = v;
        }
        /* Context.h:50-61 */
        double INV_LN10 = 1.0 / log(10);
        for (int i = 0, offset = 0; i <= MAX_QUAL; offset += ++i)
            for (int j = 0; j <= i; j++) {
                double sf = approx_log10_sum_f((float)(-0.1 * i), (float)(-0.1 * j));
                double sd = approx_log10_sum_d(-0.1 * i, -0.1 * j);
                double mf = log1p(-fmin(1.0, pow(10, sf))) * INV_LN10;
                double md = log1p(-fmin(1.0, pow(10, sd))) * INV_LN10;
                m2m_f[offset + j] = (float)pow(10, mf);
                m2m_d[offset + j] = pow(10, md);
            }
        /* Context.h:105-110, :145-151 */
        for (int x = 0; x < 128; x++) {
            ph2pr_f[x] = powf(10.f, -((float)x) / 10.f);
            ph2pr_d[x] = pow(10.0, -((double)x) / 10.0);
        }
        IC_f = ldexpf(1.f, 120);   LIC_f = log10f(IC_f);
        IC_d = ldexp(1.0, 1020);   LIC_d = log10(IC_d);
        /* host_type.h:123-143: a zero-initialised table, so every other byte is class 0 ('A') */
        memset(cls_tab, 0, sizeof cls_tab);
        cls_tab['A'] = 0; cls_tab['C'] = 1; cls_tab['T'] = 2; cls_tab['G'] = 3; cls_tab['N'] = 4;
        g_init = 1;
    }
    pthread_mutex_unlock(&g_lock);
}

const float*  pmm_oracle_ph2pr_f32(void) { pmm_oracle_init(); return ph2pr_f; }
const double* pmm_oracle_ph2pr_f64(void) { pmm_oracle_init(); return ph2pr_d; }
const float*  pmm_oracle_m2m_f32(void)   { pmm_oracle_init(); return m2m_f; }
const double* pmm_oracle_m2m_f64(void)   { pmm_oracle_init(); return m2m_d; }
int           pmm_oracle_m2m_size(void)  { return M2M_SIZE; }
float         pmm_oracle_log10_ic_f32(void) { pmm_oracle_init(); return LIC_f; }
double        pmm_oracle_log10_ic_f64(void) { pmm_oracle_init(); return LIC_d; }

static unsigned ftz_on(void)  { unsigned old = _mm_getcsr(); _mm_setcsr(old | 0x8000u); return old; }
static void ftz_restore(unsigned old) { _mm_setcsr(old); }

/* Context.h:123-134 / :163-174 (the MAX_QUAL < maxQual branch is unreachable: quals are masked to 0..127) */
static int m2m_index(int ins, int del)
{
    int mn = del, mx = ins;
    if (ins <= del) { mn = ins; mx = del; }
    return ((mx * (mx + 1)) >> 1) + mn;
}

#define DEFINE_FORWARD(NAME, T, PH2PR, M2M, IC)                                                               \
static T NAME(int R, int C, const char* rs, const char* q, const char* ins, const char* del,                  \
              const char* gcp, const char* hap)                                                               \
{                                                                                                             \
    if (R <= 0) return (T)0;   /* the reference reads garbage for an empty read; callers never send one */     \
    T* buf = (T*)malloc(sizeof(T) * 6 * (size_t)(C + 1));                                                     \
    T *pM = buf, *pX = pM + (C + 1), *pY = pX + (C + 1), *cM = pY + (C + 1), *cX = cM + (C + 1),              \
      *cY = cX + (C + 1);                                                                                     \
    unsigned char* hc = (unsigned char*)malloc((size_t)C + 1);                                                \
    for (int c = 0; c < C; c++) hc[c] = cls_tab[(unsigned char)hap[c]];                                       \
    const T init_Y = IC / (T)C;                    /* avx-pairhmm-template.h:86,151 */                        \
    for (int c = 0; c <= C; c++) { pM[c] = (T)0; pX[c] = (T)0; pY[c] = init_Y; }                              \
    for (int r = 1; r <= R; r++) {                                                                            \
        const int _i = ins[r - 1] & 127, _d = del[r - 1] & 127, _c = gcp[r - 1] & 127, _q = q[r - 1] & 127;   \
        const T pMM = M2M[m2m_index(_i, _d)];      /* avx-pairhmm-template.h:114 */                           \
        const T pGAPM = (T)1.0 - PH2PR[_c];        /* :115 */                                                 \
        const T pMX = PH2PR[_i];                   /* :116 */                                                 \
        const T pXX = PH2PR[_c];                   /* :117 */                                                 \
        const T pMY = PH2PR[_d];                   /* :118 */                                                 \
        const T pYY = PH2PR[_c];                   /* :119 */                                                 \
        const T dm = PH2PR[_q];                    /* :126 */                                                 \
        const T w_match = (T)1.0 - dm;             /* :156 */                                                 \
        const T w_mis = dm / (T)3.0;               /* :158 */                                                 \
        const unsigned char rc = cls_tab[(unsigned char)rs[r - 1]];                                           \
        cM[0] = (T)0; cX[0] = (T)0; cY[0] = (T)0;                                                             \
        for (int c = 1; c <= C; c++) {                                                                        \
            const int match = (rc == hc[c - 1]) || rc == 4 || hc[c - 1] == 4;                                 \
            const T w = match ? w_match : w_mis;                                                              \
            /* :188  M = ((Md*pMM + Xd*pGAPM) + Yd*pGAPM) * w */                                              \
            T t1 = pM[c - 1] * pMM;                                                                           \
            T t2 = pX[c - 1] * pGAPM;                                                                         \
            T t3 = t1 + t2;                                                                                   \
            T t4 = pY[c - 1] * pGAPM;                                                                         \
            T t5 = t3 + t4;                                                                                   \
            cM[c] = t5 * w;                                                                                   \
            /* :194  X = Mup*pMX + Xup*pXX */                                                                 \
            T x1 = pM[c] * pMX;                                                                               \
            T x2 = pX[c] * pXX;                                                                               \
            cX[c] = x1 + x2;                                                                                  \
            /* :197  Y = Mleft*pMY + Yleft*pYY */                                                             \
            T y1 = cM[c - 1] * pMY;                                                                           \
            T y2 = cY[c - 1] * pYY;                                                                           \
            cY[c] = y1 + y2;                                                                                  \
        }                                                                                                     \
        T* t;                                                                                                 \
        t = pM; pM = cM; cM = t;  t = pX; pX = cX; cX = t;  t = pY; pY = cY; cY = t;                          \
    }                                                                                                         \
    /* :308-343  two running sums, left to right, added once */                                               \
    T sumM = (T)0, sumX = (T)0;                                                                               \
    for (int c = 1; c <= C; c++) { sumM = sumM + pM[c]; sumX = sumX + pX[c]; }                                \
    T res = sumM + sumX;                                                                                      \
    free(buf); free(hc);                                                                                      \
    return res;                                                                                               \
}

DEFINE_FORWARD(forward_f32, float, ph2pr_f, m2m_f, IC_f)
DEFINE_FORWARD(forward_f64, double, ph2pr_d, m2m_d, IC_d)

float pmm_oracle_f32(int R, int C, const char* rs, const char* q, const char* ins, const char* del,
                     const char* gcp, const char* hap)
{
    pmm_oracle_init();
    unsigned old = ftz_on();
    float v = forward_f32(R, C, rs, q, ins, del, gcp, hap);
    ftz_restore(old);
    return v;
}

double pmm_oracle_f64(int R, int C, const char* rs, const char* q, const char* ins, const char* del,
                      const char* gcp, const char* hap)
{
    pmm_oracle_init();
    unsigned old = ftz_on();
    double v = forward_f64(R, C, rs, q, ins, del, gcp, hap);
    ftz_restore(old);
    return v;
}

/* Batch contract of FalconPairHMM::computePairhmmAVX (FalconPairHMM.cpp:69-95): read-major output,
 * f < 1e-28f -> double re-run -> log10(d) - log10(2^1020), else (double)(log10f(f) - log10f(2^120)). */
typedef struct {
    int num_read, num_hap;
    const int *read_off, *hap_off;
    const char *rs, *q, *ins, *del, *gcp, *hap;
    float* raw; double* out; unsigned char* fb;
    int float_only;
    int r0, r1;
} batch_job;

static void* batch_worker(void* p)
{
    batch_job* j = (batch_job*)p;
    unsigned old = ftz_on();
    for (int i = j->r0; i < j->r1; i++) {
        const int ro = j->read_off[i], R = j->read_off[i + 1] - ro;
        for (int h = 0; h < j->num_hap; h++) {
            const int ho = j->hap_off[h], C = j->hap_off[h + 1] - ho;
            const size_t k = (size_t)i * j->num_hap + h;
            float f = forward_f32(R, C, j->rs + ro, j->q + ro, j->ins + ro, j->del + ro, j->gcp + ro, j->hap + ho);
            if (j->raw) j->raw[k] = f;
            int fb = f < 1e-28f;
            if (j->fb) j->fb[k] = (unsigned char)fb;
            if (j->float_only || !j->out) continue;
            if (fb) {
                double d = forward_f64(R, C, j->rs + ro, j->q + ro, j->ins + ro, j->del + ro, j->gcp + ro, j->hap + ho);
                j->out[k] = log10(d) - LIC_d;
            } else {
                j->out[k] = (double)(log10f(f) - LIC_f);
            }
        }
    }
    ftz_restore(old);
    return NULL;
}

/* reads: five parallel byte arrays indexed by read_off[0..num_read]; haps: one byte array indexed by hap_off.
 * raw / out / fb may be NULL.  float_only skips the double re-run and the log10 (the "float pass" timing). */
void pmm_oracle_batch(int num_read, const int* read_off, const char* rs, const char* q, const char* ins,
                      const char* del, const char* gcp, int num_hap, const int* hap_off, const char* hap,
                      float* raw, double* out, unsigned char* fb, int float_only, int nthreads)
{
    pmm_oracle_init();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > num_read) nthreads = num_read > 0 ? num_read : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    batch_job* jobs = (batch_job*)malloc(sizeof(batch_job) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        batch_job b = { num_read, num_hap, read_off, hap_off, rs, q, ins, del, gcp, hap, raw, out, fb, float_only,
                        (int)((long long)num_read * t / nthreads), (int)((long long)num_read * (t + 1) / nthreads) };
        jobs[t] = b;
        if (t + 1 < nthreads) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    batch_worker(&jobs[nthreads - 1]);
    for (int t = 0; t + 1 < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}
