// PairHMMGpu.cpp -- see PairHMMGpu.h.
#include "PairHMMGpu.h"

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "pairhmm_cuda.h"

double peak_kernel_gcups = 0;
double curr_kernel_gcups = 0;

namespace {

pmm_ctx* g_ctx = nullptr;            // like the reference's global `env` (host/PairHMMFpga.cpp:8-10)
int      g_device = -1;
std::vector<float>    g_ret;         // callee-owned result buffer, grown on demand
std::vector<uint32_t> g_fb_index;
std::vector<double>   g_fb_value;

int parse_device(const char* conf) {
  if (!conf || !*conf || !strcmp(conf, "-")) return 0;
  if (!strncmp(conf, "cuda:", 5)) return atoi(conf + 5);
  return 0;                          // a path (the reference's argv[1]) means "default device"
}

void fail(const char* what) {
  throw std::runtime_error(std::string(what) + ": " + pmm_last_error(g_ctx));
}

}  // namespace

float* compute_gpu(const char* conf, std::string read_data, std::string hap_data, uint64_t num_cell) {
  const int device = parse_device(conf);
  if (g_ctx && device != g_device) { pmm_destroy(g_ctx); g_ctx = nullptr; }
  if (!g_ctx) {
    if (pmm_create(device, &g_ctx) != PMM_OK)
      throw std::runtime_error(std::string("PairHMM CUDA engine unavailable: ") + pmm_last_error(nullptr));
    g_device = device;
  }

  int num_read = 0, num_hap = 0;
  if (pmm_stage_serialized(g_ctx, read_data.data(), read_data.size(), hap_data.data(), hap_data.size(),
                           &num_read, &num_hap) != PMM_OK) fail("stage");
  const uint64_t pairs = (uint64_t)num_read * (uint64_t)num_hap;
  if (g_ret.size() < pairs) g_ret.resize(pairs + pairs / 4);

  if (pmm_launch(g_ctx) != PMM_OK) fail("launch");
  if (pmm_fetch_raw(g_ctx, g_ret.data(), g_ret.size()) != PMM_OK) fail("fetch");

  uint64_t nfb = 0;
  if (pmm_fetch_fallback(g_ctx, nullptr, nullptr, 0, &nfb) != PMM_OK) fail("fallback count");
  g_fb_index.resize(nfb); g_fb_value.resize(nfb);
  if (nfb && pmm_fetch_fallback(g_ctx, g_fb_index.data(), g_fb_value.data(), nfb, &nfb) != PMM_OK) fail("fallback list");

  pmm_stats_t st;
  if (pmm_get_stats(g_ctx, &st) == PMM_OK && st.ms_f32 + st.ms_fallback > 0) {
    const uint64_t cells = num_cell ? num_cell : st.cells;
    curr_kernel_gcups = (double)cells / ((st.ms_f32 + st.ms_fallback) * 1e-3) * 1e-9;
    if (curr_kernel_gcups > peak_kernel_gcups) peak_kernel_gcups = curr_kernel_gcups;
  }
  return g_ret.data();
}

float* compute_fpga(const char* bit_path, std::string read_data, std::string hap_data, uint64_t num_cell) {
  return compute_gpu(bit_path, std::move(read_data), std::move(hap_data), num_cell);
}

uint64_t last_fallback(const uint32_t** index, const double** value) {
  if (index) *index = g_fb_index.data();
  if (value) *value = g_fb_value.data();
  return g_fb_index.size();
}

void cleanup() {
  if (g_ctx) { pmm_destroy(g_ctx); g_ctx = nullptr; }
}
