// main.cpp -- host test bench of the PairHMM path: runs every input<i> of a fixture folder through the GPU and
// compares with output<i>.  Same job as the reference's bench (/root/reference/pairhmm/host/main.cpp:230-425) and
// the same positional arguments `<conf> <folder>`; what it calls differs:
//   default    direct dispatch, compute_gpu() on the serialized batch + the GPU's fallback list (PairHMMGpu.h)
//   --client   PairHMMClient + PairHMMWorker through the in-process accelerator manager (what a GATK host does)
// A result passes when |target - golden| / |golden| <= 5e-3 like the reference's check (:381); on top of that the
// bench counts how many results are bit-identical to the golden file, which is the bar this repo holds itself to.
#include <dirent.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "PairHMMClient.h"
#include "PairHMMGpu.h"
#include "PairHMMHostInterface.h"
#include "PairHMMManager.h"
#include "PairHMMWorker.h"
#include "fixture_io.h"
#include "pairhmm_cuda.h"

static int count_batches(const char* folder) {
  DIR* d = opendir(folder);
  if (!d) throw std::runtime_error("input folder is not valid");
  int n = 0;
  while (dirent* e = readdir(d)) if (!strncmp(e->d_name, "input", 5)) ++n;
  closedir(d);
  return n;
}

int main(int argc, char** argv) {
  bool client_mode = false, compare = true;
  std::vector<const char*> pos;
  for (int a = 1; a < argc; ++a) {
    if (!strcmp(argv[a], "--client")) client_mode = true;
    else if (!strcmp(argv[a], "--no-compare")) compare = false;
    else pos.push_back(argv[a]);
  }
  if (pos.size() != 2) {
    fprintf(stderr, "usage: %s [--client] [--no-compare] <conf | cuda:N | -> <folder with input<i>/output<i>>\n", argv[0]);
    return 1;
  }
  const char* conf = pos[0];
  const char* folder = pos[1];

  try {
    const int test_num = count_batches(folder);

    PairHMMClient* client = nullptr;
    if (client_mode) {
      const size_t n = strlen(conf);
      if (n > 5 && !strcmp(conf + n - 5, ".conf")) pairhmm_manager_from_conf(conf);
      else pairhmm_default_manager();
      client = new PairHMMClient();
    }

    printf("test id, num_read, num_haps, kernel gcups, host gcups\n");
    int fail_num = 0, test_run_num = 0;
    uint64_t total_num_cell = 0, total_results = 0, total_bit_equal = 0;
    double total_s = 0;

    for (int i = 0; i < test_num; ++i) {
      const std::string fin = std::string(folder) + "/input" + std::to_string(i);
      const std::string fout = std::string(folder) + "/output" + std::to_string(i);
      read_t* reads = nullptr; hap_t* haps = nullptr;
      int num_read = 0, num_hap = 0;
      fixture::read_input(fin, num_read, num_hap, reads, haps);
      const int output_size = num_read * num_hap;

      uint64_t total_rl = 0, total_hl = 0;
      for (int r = 0; r < num_read; ++r) total_rl += reads[r].len;
      for (int h = 0; h < num_hap; ++h) total_hl += haps[h].len;
      const uint64_t num_cell = total_rl * total_hl;
      total_num_cell += num_cell;

      std::vector<double> target(output_size);
      int recal_count = 0;
      const auto t0 = std::chrono::steady_clock::now();
      if (client_mode) {
        PairHMMWorker worker(client, num_read, num_hap, reads, haps);
        worker.run();
        worker.getOutput(target.data());
        recal_count = worker.numRecalculated();
      } else {
        const std::string read_data = serialize(reads, num_read);
        const std::string hap_data = serialize(haps, num_hap);
        const float* output = compute_gpu(conf, read_data, hap_data, num_cell);
        const uint32_t* fb_idx; const double* fb_val;
        const uint64_t nfb = last_fallback(&fb_idx, &fb_val);
        // the reference's bench does this tail itself (host/main.cpp:361-371); same formulas, host libm
        if (pmm_host_finish_log10(output, (uint64_t)output_size, fb_idx, fb_val, nfb, target.data()) != PMM_OK)
          throw std::runtime_error("pmm_host_finish_log10 failed");
        recal_count = (int)nfb;
      }
      const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      total_s += secs;
      printf("%d, %d, %d, %.3f, %.3f\n", i, num_read, num_hap, client_mode ? 0.0 : curr_kernel_gcups,
             (double)num_cell / secs * 1e-9);

      if (compare) {
        std::vector<double> golden(output_size);
        fixture::read_output(fout, golden.data(), output_size);
        int error_count = 0;
        for (int k = 0; k < output_size; ++k) {
          if (std::isnan(target[k])) { printf("target is nan\n"); ++error_count; continue; }
          if (!memcmp(&target[k], &golden[k], sizeof(double))) { ++total_bit_equal; continue; }
          if (std::isinf(golden[k]) && std::isinf(target[k]) && (golden[k] < 0) == (target[k] < 0)) continue;
          const double error = fabs((target[k] - golden[k]) / golden[k]);
          if (!(error <= 5e-3)) {
            printf("result has significant error, golden=%f, target=%f\n", golden[k], target[k]);
            ++error_count;
          }
        }
        total_results += output_size;
        if (error_count > 0) { ++fail_num; printf("Test #%d: %d errors\n", i, error_count); }
        else if (recal_count == 0) printf("Test #%d: Pass (no recalc)\n", i);
        else printf("Test #%d: Pass (recalc %d/%d)\n", i, recal_count, output_size);
      }
      ++test_run_num;
      free_reads(reads, num_read);
      free_haps(haps, num_hap);
    }

    delete client;
    if (client_mode) pairhmm_shutdown_manager();
    printf("%d out of %d failed test\n", fail_num, test_run_num);
    if (compare) printf("bit-identical results: %llu of %llu\n", (unsigned long long)total_bit_equal, (unsigned long long)total_results);
    if (total_s > 0) printf("Average host GCUPS: %.3f (peak kernel GCUPS %.3f)\n", (double)total_num_cell / total_s * 1e-9, peak_kernel_gcups);
    return fail_num ? 1 : 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "host_tb: %s\n", e.what());
    return 2;
  }
}
