// sw_host.cpp -- self-checking bench of the Smith-Waterman host interface, after the reference's
// /root/reference/htc-sw/host/sw_host.cpp:150-331: random reference windows and alternates (GenInputs, :145-181 there),
// batch sizes 1..256, all four overhang strategies; the accelerator path (FalconSWFPGA_run) against the batch entry
// (SWPairwiseAlignmentMultiBatch) and, pair by pair, runSWOnePairBT_gpu.
//   usage: sw_host [cuda:N]                self-check on generated batches, prints failures and kernel GCUPs
//          sw_host cuda:N <file>           every line of <file> is "strategy reference alternate"; prints "offset cigar" per
//                                          line (tests/test_gpu_sw.py compares that with the oracle)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "FalconSWGpu.h"

static unsigned long long g_rng = 88172645463325252ull;
static unsigned rnd() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return (unsigned)(g_rng >> 11); }

static void gen_inputs(std::vector<char>& ref, int& refLength, std::vector<char>& alts, int batchSize, std::vector<int>& altLengths) {
  static const char kBase[4] = {'A', 'T', 'C', 'G'};
  refLength = 60 + rnd() % 450;
  ref.assign(MAX_SEQ_LENGTH, 0);
  for (int i = 0; i < refLength; ++i) ref[i] = kBase[rnd() % 4];
  alts.assign((size_t)batchSize * MAX_SEQ_LENGTH, 0);
  altLengths.assign(batchSize, 0);
  for (int b = 0; b < batchSize; ++b) {
    int len = refLength - 10 + (int)(rnd() % 21);
    if (len < 1) len = 1;
    if (len > MAX_SEQ_LENGTH - 2) len = MAX_SEQ_LENGTH - 2;
    altLengths[b] = len;
    for (int j = 0; j < len; ++j)
      alts[(size_t)b * MAX_SEQ_LENGTH + j] = (rnd() % 10 == 0 || j >= refLength) ? kBase[rnd() % 4] : ref[j];
  }
}

static bool same(const Cigar& a, const Cigar& b) {
  if (a.CigarElementNum != b.CigarElementNum) return false;
  for (int k = 0; k < a.CigarElementNum; ++k)
    if (a.cigarElements[k].length != b.cigarElements[k].length || a.cigarElements[k].state != b.cigarElements[k].state) return false;
  return true;
}

int main(int argc, char** argv) {
  try {
    std::string dev = argc > 1 ? argv[1] : "cuda:0";
    FalconSWFPGA_init(const_cast<char*>(dev.c_str()));
    if (argc > 2) {
      FILE* f = fopen(argv[2], "r");
      if (!f) throw std::runtime_error("cannot open pair file");
      static char ref[MAX_SEQ_LENGTH + 2], alt[1][MAX_SEQ_LENGTH];
      static char altbuf[MAX_SEQ_LENGTH + 2];
      int strategy;
      while (fscanf(f, "%d %1537s %1537s", &strategy, ref, altbuf) == 3) {
        int rl = (int)strlen(ref), al = (int)strlen(altbuf), off = 0;
        if (rl > MAX_SEQ_LENGTH || al > MAX_SEQ_LENGTH) throw std::runtime_error("sequence longer than MAX_SEQ_LENGTH");
        memcpy(alt[0], altbuf, (size_t)al);
        static Cigar c;
        SWPairwiseAlignmentMultiBatch(ref, rl, alt, 1, &al, &c, &off, strategy, 0);
        printf("%d ", off);
        for (int e = 0; e < c.CigarElementNum; ++e) printf("%d%c", c.cigarElements[e].length, "MID?S"[c.cigarElements[e].state]);
        printf("\n");
      }
      fclose(f);
      FalconSWFPGA_release();
      return 0;
    }
    int failures = 0;
    double cells = 0, kernel_ns = 0;
    for (int batchSize = 1; batchSize <= 256; batchSize *= 2)
      for (int strategy = 0; strategy < 4; ++strategy) {
        std::vector<char> ref, alts; std::vector<int> lens; int refLength;
        gen_inputs(ref, refLength, alts, batchSize, lens);
        auto* alt2d = reinterpret_cast<char(*)[MAX_SEQ_LENGTH]>(alts.data());
        std::vector<Cigar> a(batchSize), b(batchSize);
        std::vector<int> offa(batchSize), offb(batchSize);
        kernel_ns += FalconSWFPGA_run(ref.data(), refLength, alt2d, lens.data(), batchSize, strategy, W_MATCH, W_MISMATCH, W_OPEN,
                                      W_EXTEND, a.data(), offa.data(), true);
        SWPairwiseAlignmentMultiBatch(ref.data(), refLength, alt2d, batchSize, lens.data(), b.data(), offb.data(), strategy, 0);
        for (int k = 0; k < batchSize; ++k) {
          cells += (double)refLength * lens[k];
          bool ok = offa[k] == offb[k] && same(a[k], b[k]);
          if (ok && k < 2) {                          // and the single-pair entry on a couple of them
            Cigar c; c.CigarElementNum = 0;
            const int off = runSWOnePairBT_gpu(W_MATCH, W_MISMATCH, W_OPEN, W_EXTEND, reinterpret_cast<uint8_t*>(ref.data()),
                                               reinterpret_cast<uint8_t*>(alt2d[k]), refLength, lens[k], (int8_t)strategy, &c);
            ok = off == offa[k] && same(c, a[k]);
          }
          // every alternate base is covered by M, I or S
          int q = 0;
          for (int e = 0; e < a[k].CigarElementNum; ++e) if (a[k].cigarElements[e].state != STATE_DELETION) q += a[k].cigarElements[e].length;
          if (!ok || q != lens[k]) { ++failures; printf("batch %d strategy %d pair %d: mismatch\n", batchSize, strategy, k); }
        }
      }
    printf("%d failures\n", failures);
    printf("kernel GCUPs is %lf\n", kernel_ns > 0 ? cells / kernel_ns : 0.0);
    FalconSWFPGA_release();
    return failures ? 1 : 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "sw_host: %s\n", e.what());
    return 2;
  }
}
