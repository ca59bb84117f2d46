/* pairhmm_c_api.h -- the client path (PairHMMClient + PairHMMWorker over the task plugin) behind one C call, for hosts
 * that are not C++: a JNI or ctypes binding passes flat arrays and gets the final log10 likelihoods, exactly what
 * Falcon's GATK fork obtains from PairHMMWorker::run() + getOutput() (/root/reference/pairhmm/client/PairHMMWorker.h:9-29,
 * PairHMMWorker.cpp:157-271).  bench.py measures `e2e.plugin_value` through it. */
#ifndef PAIRHMM_C_API_H
#define PAIRHMM_C_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* One batch: num_read reads (five parallel byte tracks indexed by read_off[0..num_read]) against num_hap haplotypes
 * (hap indexed by hap_off[0..num_hap]); out receives num_read * num_hap doubles, read-major.  The calling thread gets
 * its own PairHMMClient (kept for its lifetime); the accelerator manager is the process's default one (every visible
 * GPU, the task plugin next to this library) unless the host program published its own.  Returns 0, or 1 with the
 * message in err.  num_recalc (may be NULL): pairs that took the double-precision re-run. */
int pairhmm_worker_forward(int num_read, const int32_t* read_off, const char* bases, const char* q, const char* i,
                           const char* d, const char* c, int num_hap, const int32_t* hap_off, const char* hap,
                           double* out, int* num_recalc, char* err, int err_capacity);
/* Drops the process's default manager (and with it the task plugin and its GPU contexts). */
void pairhmm_worker_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif
