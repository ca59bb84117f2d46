// pairhmm_test.cpp -- test bench of the standalone entry (FalconPairHMM), after the reference's
// /root/reference/pairhmm/xlnx/pairhmm_test.cpp:
//
//   pairhmm_test <conf> --real <folder>            folder of input<i> / output<i> files (GATK dumps or minted fixtures,
//                                                  format in ../host/fixture_io.h); compares like the reference's cmp()
//                                                  (:238-268): relative error bar 5e-3, plus the count of bit-identical
//                                                  results, which is the bar of this repository
//   pairhmm_test <conf> --syn <n> [--dump <dir>]   n synthetic batches of the reference's shapes (:62-83: 16*(i+1) reads
//                                                  of up to 192 bases, i+1 haplotypes of up to 1024 bases, Q ~ N(30,5) >= 6,
//                                                  ins/del ~ N(40,1) >= 1, GCP 10).  The reference mints the expected values
//                                                  with its CPU path; this product has none, so the bench checks what can
//                                                  be checked without one (results are finite or -inf, never NaN or
//                                                  positive; the same pairs in reversed batch order give the same bits)
//                                                  and --dump writes input<i> / output<i> for an external checker
//                                                  (tests/test_host_layer.py runs the oracle over them).
//   <conf>: "-" or "cuda:N".
//
// Unlike the reference this bench clears the input between cases (its GetInputs appends to the previous batch) and
// seeds one std::mt19937_64 instead of default-constructing an engine per call (which yields the same number every time).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <random>
#include <stdexcept>
#include <string>

#include "fixture_io.h"
#include "host/FalconPairHMM.h"

namespace {

constexpr int kMaxReadLen = 192, kMaxHapLen = 1024;        // the reference's synthetic shapes (xlnx/common/common.h:3-4)

double now_ns() {
  timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec * 1e9 + t.tv_nsec;
}

void gen_inputs(std::mt19937_64& rng, pairhmmInput* in, int size) {
  static const char kBase[4] = {'A', 'T', 'C', 'G'};
  std::uniform_int_distribution<int> base(0, 3), rlen(kMaxReadLen / 4, kMaxReadLen), hlen(kMaxHapLen / 4, kMaxHapLen);
  std::normal_distribution<double> qual(30.0, 5.0), indel(40.0, 1.0);
  in->reads.clear(); in->haps.clear();
  in->reads.resize(16 * (size + 1));
  in->haps.resize(size + 1);
  for (Hap& h : in->haps) {
    const int n = hlen(rng);
    for (int j = 0; j < n; ++j) h.bases.push_back(kBase[base(rng)]);
  }
  for (Read& r : in->reads) {
    // a read is a noisy copy of a stretch of one haplotype, so that likelihoods are not all vanishing
    const Hap& src = in->haps[rng() % in->haps.size()];
    const int n = std::min<int>(rlen(rng), (int)src.bases.size());
    const int off = (int)(rng() % (src.bases.size() - n + 1));
    for (int j = 0; j < n; ++j) {
      const int q = std::max(6, (int)qual(rng));
      const bool err = std::generate_canonical<double, 30>(rng) < std::pow(10.0, -q / 10.0);
      r.bases.push_back(err ? kBase[base(rng)] : src.bases[off + j]);
      r._q.push_back((char)q);
      r._i.push_back((char)std::max(1, (int)indel(rng)));
      r._d.push_back((char)std::max(1, (int)indel(rng)));
      r._c.push_back((char)10);
    }
  }
}

void load_case(const std::string& path, pairhmmInput* in) {
  int nr = 0, nh = 0; read_t* reads = nullptr; hap_t* haps = nullptr;
  fixture::read_input(path, nr, nh, reads, haps);
  in->reads.assign(nr, Read()); in->haps.assign(nh, Hap());
  for (int k = 0; k < nr; ++k) {
    const size_t n = (size_t)reads[k].len;
    in->reads[k].bases.assign(reads[k]._b, n); in->reads[k]._q.assign(reads[k]._q, n); in->reads[k]._i.assign(reads[k]._i, n);
    in->reads[k]._d.assign(reads[k]._d, n); in->reads[k]._c.assign(reads[k]._c, n);
  }
  for (int k = 0; k < nh; ++k) in->haps[k].bases.assign(haps[k]._b, (size_t)haps[k].len);
  free_reads(reads, nr); free_haps(haps, nh);
}

void dump_case(const std::string& dir, int id, const pairhmmInput& in, const pairhmmOutput& out) {
  std::ofstream f((dir + "/input" + std::to_string(id)).c_str());
  f << "readListSize " << in.reads.size() << " numHaplotypes " << in.haps.size() << "\n";
  static const char* kCaption[5] = {"readBases", "readQuals", "insertionGOP", "deletionGOP", "overallGCP"};
  for (size_t k = 0; k < in.reads.size(); ++k) {
    const Read& r = in.reads[k];
    const std::string* tr[5] = {&r.bases, &r._q, &r._i, &r._d, &r._c};
    f << r.bases.size() << "\n";
    for (int t = 0; t < 5; ++t) {
      f << "readDataArray[" << k << "]." << kCaption[t] << "[" << r.bases.size() << "]: \n";
      for (size_t j = 0; j < r.bases.size(); ++j) f << (int)(unsigned char)(*tr[t])[j] << " ";
      f << "\n";
    }
  }
  f << "\n";
  for (size_t k = 0; k < in.haps.size(); ++k) {
    f << in.haps[k].bases.size() << "\n" << "mHaplotypeDataArray[" << k << "].haplotypeBases[" << in.haps[k].bases.size() << "]: \n";
    f << in.haps[k].bases << "\n";
  }
  std::ofstream g((dir + "/output" + std::to_string(id)).c_str());
  char buf[96];
  for (double v : out.likelihoodData) {
    long long bits; memcpy(&bits, &v, sizeof bits);
    snprintf(buf, sizeof buf, "%.17g %lld\n", v, bits);
    g << buf;
  }
}

// the reference's cmp(), plus the bit-identical count
void cmp(const pairhmmOutput& target, const pairhmmOutput& golden, int test_id, double& errors, double& largest, uint64_t& same_bits) {
  int bad = 0;
  for (size_t i = 0; i < golden.likelihoodData.size(); ++i) {
    const double t = target.likelihoodData[i], g = golden.likelihoodData[i];
    if (memcmp(&t, &g, sizeof t) == 0) { ++same_bits; continue; }
    if (std::isnan(t)) { printf("error, target is nan\n"); ++bad; continue; }
    const double e = std::fabs((t - g) / g);
    if (e > largest) largest = e;
    if (e > 5e-3) { printf("%dth test: %zuth result has significant error, golden=%f, target=%f\n", test_id, i, g, t); ++bad; }
  }
  if (bad) printf("%d out of %zu have significant error\n", bad, golden.likelihoodData.size());
  errors += bad;
}

void usage() {
  printf("pairhmm_test <conf> --real <real cases folder>\npairhmm_test <conf> --syn <syn cases number> [--dump <folder>]\n"
         "  <conf>: - or cuda:N\n");
}

}  // namespace

int main(int argc, char* argv[]) {
  if (argc == 2 && (!strcmp(argv[1], "-h") || !strcmp(argv[1], "--help"))) { usage(); return 0; }
  if (argc == 2 && (!strcmp(argv[1], "-v") || !strcmp(argv[1], "--version"))) { printf("Host code for NVIDIA B200 (sm_100a), C ABI pairhmm_cuda.h\n"); return 0; }
  if (argc != 4 && argc != 6) { printf("Invalid argument list\n"); usage(); return EXIT_FAILURE; }
  const bool synthetic = !strcmp(argv[2], "--syn");
  if (!synthetic && strcmp(argv[2], "--real")) { usage(); return EXIT_FAILURE; }
  std::string dump;
  if (argc == 6) { if (strcmp(argv[4], "--dump")) { usage(); return EXIT_FAILURE; } dump = argv[5]; }
  int test_num = 0;
  std::string folder;
  if (synthetic) {
    test_num = atoi(argv[3]);
    if (test_num <= 0) { printf("Invalid synthetic cases number %s\n", argv[3]); return EXIT_FAILURE; }
  } else {
    folder = argv[3];
    if (!folder.empty() && folder.back() != '/') folder += '/';
    while (std::ifstream((folder + "input" + std::to_string(test_num)).c_str()).good()) ++test_num;
    printf("find %d cases in dir %s\n", test_num, folder.c_str());
    if (!test_num) return EXIT_FAILURE;
  }
  try {
    FalconPairHMM falcon(argv[1]);
    std::mt19937_64 rng(20260);
    pairhmmInput input, reversed; pairhmmOutput golden, target, check;
    double total_cells = 0, total_time = 0, total_results = 0, errors = 0, largest = 0, peak = 0;
    uint64_t same_bits = 0, fallback = 0, invariant_failures = 0;
    for (int i = 0; i < test_num; ++i) {
      if (synthetic) gen_inputs(rng, &input, i);
      else {
        load_case(folder + "input" + std::to_string(i), &input);
        golden.likelihoodData.assign(input.reads.size() * input.haps.size(), 0.0);
        fixture::read_output(folder + "output" + std::to_string(i), golden.likelihoodData.data(), (int)golden.likelihoodData.size());
      }
      bool violate = false, used = false;
      const double cells = countCell(&input, 0, violate);
      const double t0 = now_ns();
      falcon.computePairhmm(&input, &target, used);
      const double dt = now_ns() - t0;
      if (!used) { printf("%dth test: batch was not computed\n", i); ++invariant_failures; continue; }
      total_cells += cells; total_time += dt; total_results += (double)target.likelihoodData.size();
      fallback += falcon.last_fallback_pairs();
      if (cells / dt > peak) peak = cells / dt;
      if (synthetic) {
        for (double v : target.likelihoodData) if (std::isnan(v) || v > 0.0) ++invariant_failures;
        reversed.reads.assign(input.reads.rbegin(), input.reads.rend());
        reversed.haps.assign(input.haps.rbegin(), input.haps.rend());
        falcon.computePairhmm(&reversed, &check, used);
        const size_t nr = input.reads.size(), nh = input.haps.size();
        for (size_t r = 0; r < nr; ++r)
          for (size_t h = 0; h < nh; ++h)
            if (memcmp(&target.likelihoodData[r * nh + h], &check.likelihoodData[(nr - 1 - r) * nh + (nh - 1 - h)], sizeof(double))) ++invariant_failures;
        if (!dump.empty()) dump_case(dump, i, input, target);
      } else {
        cmp(target, golden, i, errors, largest, same_bits);
      }
    }
    printf("%d cases, %.0f results, %.3e cells, %llu pairs re-run in double\n", test_num, total_results, total_cells, (unsigned long long)fallback);
    printf("end-to-end %.3f GCUPS, best case %.3f GCUPS, kernels only %.3f GCUPS (peak %.3f)\n", total_cells / total_time, peak,
           falcon.get_kernel_time() > 0 ? total_cells * (synthetic ? 2 : 1) / falcon.get_kernel_time() : 0.0, falcon.peak_kernel_gcups());
    if (synthetic) {
      printf("invariant failures: %llu\n", (unsigned long long)invariant_failures);
      return invariant_failures ? EXIT_FAILURE : 0;
    }
    printf("largest relative error %.3e, %.0f results with significant error\n", largest, errors);
    printf("bit-identical results: %llu of %.0f\n", (unsigned long long)same_bits, total_results);
    return errors > 0 || invariant_failures ? EXIT_FAILURE : 0;
  } catch (const std::exception& e) {
    printf("error: %s\n", e.what());
    return 2;
  }
}
