// selftest.cpp -- checks of the C++ host layer, driven by tests/test_host_layer.py.  The first four modes need no GPU.
//   reserialize <reads.bin> <haps.bin> <reads.out> <haps.out>   wire format: parse and write back (both overload sets)
//   conf <file>                                                  manager conf parser: print what was understood
//   plugin <libPairHMMTask.so>                                   plugin exports create()/destroy(), task has 3 inputs
//   nofallback                                                   a client without accelerator must throw, never compute
//   threads <fixture-dir> <n>                                    (GPU) n client threads through one manager, all batches;
//                                                                results must be bit-identical across threads
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

#include "PairHMMClient.h"
#include "PairHMMHostInterface.h"
#include "PairHMMManager.h"
#include "PairHMMWorker.h"
#include "blaze/Task.h"
#include "fixture_io.h"

static std::string slurp(const char* path) {
  std::ifstream in(path, std::ios::binary);
  if (!in.good()) throw std::runtime_error(std::string("cannot open ") + path);
  std::stringstream ss; ss << in.rdbuf();
  return ss.str();
}
static void spit(const char* path, const std::string& s) {
  std::ofstream out(path, std::ios::binary);
  out.write(s.data(), (std::streamsize)s.size());
}

static int reserialize(char** a) {
  const std::string rin = slurp(a[0]), hin = slurp(a[1]);
  read_t* reads = nullptr; hap_t* haps = nullptr;
  const int nr = deserialize(rin, reads), nh = deserialize(hin, haps);
  for (int k = 0; k < nr; ++k) {
    const char* t[5] = {reads[k]._b, reads[k]._q, reads[k]._i, reads[k]._d, reads[k]._c};
    for (int z = 0; z < 5; ++z) if (reads[k].len > 0 && t[z][reads[k].len] != '\0') { puts("read track not NUL-terminated"); return 1; }
  }
  for (int k = 0; k < nh; ++k) if (haps[k].len > 0 && haps[k]._b[haps[k].len] != '\0') { puts("hap not NUL-terminated"); return 1; }
  // string overloads
  const std::string r1 = serialize(reads, nr), h1 = serialize(haps, nh);
  // buffer overloads, sized by serialized_size, parsed again through the raw-pointer deserialize
  std::string r2(serialized_size(reads, nr), '\0'), h2(serialized_size(haps, nh), '\0');
  if (serialize(&r2[0], reads, nr) != r2.size() || serialize(&h2[0], haps, nh) != h2.size()) { puts("size mismatch"); return 1; }
  read_t* reads2 = nullptr; hap_t* haps2 = nullptr;
  if (deserialize(static_cast<const void*>(r2.data()), reads2) != nr || deserialize(static_cast<const void*>(h2.data()), haps2) != nh) { puts("count mismatch"); return 1; }
  const std::string r3 = serialize(reads2, nr), h3 = serialize(haps2, nh);
  free_reads(reads, nr); free_haps(haps, nh); free_reads(reads2, nr); free_haps(haps2, nh);
  if (r1 != r2 || r1 != r3 || h1 != h2 || h1 != h3) { puts("overloads disagree"); return 1; }
  spit(a[2], r1); spit(a[3], h1);
  // a truncated block must be refused by the bounded overload
  bool threw = false;
  if (rin.size() > 8) {
    try { read_t* x = nullptr; deserialize(rin.substr(0, rin.size() - 3), x); } catch (const std::exception&) { threw = true; }
    if (!threw) { puts("truncated block accepted"); return 1; }
  }
  printf("ok %d reads %d haps\n", nr, nh);
  return 0;
}

static int conf(char** a) {
  blaze::ManagerConf c; std::string err;
  if (!c.ParseFromFile(a[0], &err)) { printf("error: %s\n", err.c_str()); return 1; }
  printf("verbose=%d\n", c.verbose());
  for (auto& p : c.platform) {
    printf("platform id=%s path=%s cache_loc=%s\n", p.id.c_str(), p.path.c_str(), p.cache_loc.c_str());
    for (auto& x : p.acc) {
      printf("  acc id=%s path=%s\n", x.id.c_str(), x.path.c_str());
      for (auto& kv : x.param) printf("    %s=%s\n", kv.first.c_str(), kv.second.c_str());
    }
  }
  return 0;
}

static int plugin(char** a) {
  void* h = dlopen(a[0], RTLD_NOW | RTLD_LOCAL);
  if (!h) { printf("dlopen failed: %s\n", dlerror()); return 1; }
  auto create = reinterpret_cast<blaze::Task* (*)()>(dlsym(h, "create"));
  auto destroy = reinterpret_cast<void (*)(blaze::Task*)>(dlsym(h, "destroy"));
  if (!create || !destroy) { puts("create/destroy missing"); return 1; }
  blaze::Task* t = create();
  const int n = t->getNumInputs();
  destroy(t);
  dlclose(h);
  printf("ok inputs=%d\n", n);
  return n == 3 ? 0 : 1;
}

static int nofallback() {
  // no manager published: start() must end in compute(), which must throw instead of computing on the CPU
  char b[4] = {'A', 'C', 'G', 'T'}, q[4] = {30, 30, 30, 30}, g[4] = {40, 40, 40, 40}, c[4] = {10, 10, 10, 10};
  read_t r{4, b, q, g, g, c}; hap_t h{4, b};
  PairHMMClient client;
  client.setup(&r, 1, &h, 1);
  try { client.start(); } catch (const std::runtime_error& e) { printf("ok threw: %s\n", e.what()); return 0; }
  puts("client computed without an accelerator");
  return 1;
}

static int threads(char** a) {
  const std::string folder = a[0];
  const int nthreads = atoi(a[1]);
  pairhmm_default_manager();
  int nb = 0;
  for (;; ++nb) { std::ifstream f(folder + "/input" + std::to_string(nb)); if (!f.good()) break; }
  if (nb == 0) { puts("no batches"); return 1; }
  std::vector<std::vector<std::vector<double> > > res(nthreads, std::vector<std::vector<double> >(nb));
  std::vector<std::string> errors(nthreads);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t] {
      try {
        PairHMMClient client;
        for (int rep = 0; rep < 2; ++rep)
          for (int i = 0; i < nb; ++i) {
            read_t* reads; hap_t* haps; int nr, nh;
            fixture::read_input(folder + "/input" + std::to_string(i), nr, nh, reads, haps);
            PairHMMWorker w(&client, nr, nh, reads, haps);
            w.run();
            res[t][i].resize((size_t)nr * nh);
            w.getOutput(res[t][i].data());
            free_reads(reads, nr); free_haps(haps, nh);
          }
      } catch (const std::exception& e) { errors[t] = e.what(); }
    });
  for (auto& x : th) x.join();
  for (int t = 0; t < nthreads; ++t) if (!errors[t].empty()) { printf("thread %d: %s\n", t, errors[t].c_str()); return 1; }
  for (int t = 1; t < nthreads; ++t)
    for (int i = 0; i < nb; ++i)
      if (res[t][i].size() != res[0][i].size() || memcmp(res[t][i].data(), res[0][i].data(), res[0][i].size() * sizeof(double))) {
        printf("thread %d batch %d differs from thread 0\n", t, i); return 1;
      }
  // against the golden files
  uint64_t total = 0, equal = 0;
  for (int i = 0; i < nb; ++i) {
    std::vector<double> g(res[0][i].size());
    fixture::read_output(folder + "/output" + std::to_string(i), g.data(), (int)g.size());
    for (size_t k = 0; k < g.size(); ++k) { ++total; equal += !memcmp(&g[k], &res[0][i][k], sizeof(double)); }
  }
  blaze::Accelerator* acc = blaze::AppCommManager::lookup(1027)->find("PairHMM");
  for (int e = 0; e < acc->numEnvs(); ++e) printf("env %d device %d tasks %llu\n", e, acc->deviceOf(e), (unsigned long long)acc->tasksRunOn(e));
  pairhmm_shutdown_manager();
  printf("ok threads=%d batches=%d bit-identical %llu of %llu\n", nthreads, nb, (unsigned long long)equal, (unsigned long long)total);
  return equal == total ? 0 : 1;
}

int main(int argc, char** argv) {
  try {
    if (argc >= 6 && !strcmp(argv[1], "reserialize")) return reserialize(argv + 2);
    if (argc >= 3 && !strcmp(argv[1], "conf")) return conf(argv + 2);
    if (argc >= 3 && !strcmp(argv[1], "plugin")) return plugin(argv + 2);
    if (argc >= 2 && !strcmp(argv[1], "nofallback")) return nofallback();
    if (argc >= 4 && !strcmp(argv[1], "threads")) return threads(argv + 2);
  } catch (const std::exception& e) {
    printf("exception: %s\n", e.what());
    return 2;
  }
  fprintf(stderr, "usage: selftest reserialize|conf|plugin|nofallback|threads ...\n");
  return 64;
}
