// PairHMMManager.cpp -- see PairHMMManager.h.
#include "PairHMMManager.h"

#include <dlfcn.h>
#include <unistd.h>

#include <memory>

#include "pairhmm_cuda.h"

namespace {
// Plain pointers on purpose: the default manager lives until pairhmm_shutdown_manager() or the end of the process.  It is
// not torn down from a static destructor -- by then the CUDA runtime and the statics of other libraries may be gone.
struct Holder { blaze::PlatformManager* get() const { return p; } blaze::PlatformManager* p = nullptr; explicit operator bool() const { return p != nullptr; } };
Holder g_pm;
blaze::AppCommManager* g_comm = nullptr;
std::mutex g_mu;

void drop() {
  delete g_comm; g_comm = nullptr;
  delete g_pm.p; g_pm.p = nullptr;
}

void publish(blaze::PlatformManager* pm) {
  drop();
  g_pm.p = pm;
  g_comm = new blaze::AppCommManager(pm, "127.0.0.1", 1027);
}
}  // namespace

std::string pairhmm_default_plugin_path() {
  if (const char* e = getenv("PAIRHMM_TASK_LIB")) return e;
  Dl_info info;
  if (dladdr(reinterpret_cast<void*>(&pairhmm_default_plugin_path), &info) && info.dli_fname) {
    std::string p(info.dli_fname);
    const size_t slash = p.rfind('/');
    const std::string dir = slash == std::string::npos ? "." : p.substr(0, slash);
    const std::string cand = dir + "/libPairHMMTask.so";
    if (access(cand.c_str(), R_OK) == 0) return cand;
  }
  return "libPairHMMTask.so";
}

blaze::PlatformManager* pairhmm_default_manager(const char* plugin_path, int slots_per_device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_pm) return g_pm.get();
  const int n = pmm_device_count();
  if (n <= 0) throw std::runtime_error("no CUDA device visible: the PairHMM accelerator has no CPU fallback");
  std::map<std::string, std::string> param;
  // $PAIRHMM_DEVICES ("0,1", default every visible GPU) and $PAIRHMM_SLOTS (tasks in flight per GPU) let a host program
  // that runs one process per GPU (bench.py under torchrun) keep each process on its own device
  const char* devs = getenv("PAIRHMM_DEVICES");
  param["devices"] = devs && *devs ? devs : "all";
  if (const char* e = getenv("PAIRHMM_SLOTS")) { const int v = atoi(e); if (v >= 1 && v <= 8) slots_per_device = v; }
  param["slots_per_device"] = std::to_string(slots_per_device);
  std::unique_ptr<blaze::PlatformManager> pm(new blaze::PlatformManager());
  pm->registerAcc("PairHMM", plugin_path && *plugin_path ? plugin_path : pairhmm_default_plugin_path(), param, n);
  publish(pm.release());
  return g_pm.get();
}

blaze::PlatformManager* pairhmm_manager_from_conf(const std::string& conf_path) {
  std::lock_guard<std::mutex> lk(g_mu);
  blaze::ManagerConf conf;
  std::string err;
  if (!conf.ParseFromFile(conf_path, &err)) throw std::runtime_error("cannot parse manager conf: " + err);
  // plugin paths in the conf are relative to the conf file, like the reference's "lib/xlnx/libPairHMMTask.so"
  const size_t slash = conf_path.rfind('/');
  const std::string dir = slash == std::string::npos ? "." : conf_path.substr(0, slash);
  for (auto& p : conf.platform)
    for (auto& a : p.acc)
      if (!a.path.empty() && a.path[0] != '/') a.path = dir + "/" + a.path;
  publish(new blaze::PlatformManager(&conf, std::max(1, pmm_device_count())));
  return g_pm.get();
}

void pairhmm_shutdown_manager() {
  std::lock_guard<std::mutex> lk(g_mu);
  drop();
}
