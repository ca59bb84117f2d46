// blaze/Common.h -- minimal in-repo stand-in for the parts of the Blaze accelerator runtime the PairHMM path
// touches.  Blaze itself is an un-vendored dependency of the reference (cmake/FindBlaze.cmake:3 downloads a
// tarball); SURVEY.md appendix B lists the symbols used.  Nothing here is a port of Blaze: it is the smallest
// in-process runtime with the same names, so that PairHMMClient / PairHMMWorker / the task plugin keep the
// reference's shape while start() runs the plugin in this process instead of sending blocks over RPC.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>

namespace blaze {

// microsecond wall clock (task/intel/PairHMMTask.cpp:77 uses it for interval timers)
inline uint64_t getUs()
{
    return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
}

class invalidParam : public std::runtime_error {
 public:
    explicit invalidParam(const std::string& what) : std::runtime_error(what) {}
};

// string key/value table: per-accelerator params of the manager conf (pairhmm/xlnx.conf:9-28) and per-block conf
class ConfigTable {
 public:
    template <typename T> void write_conf(const std::string& key, const T& val)
    {
        std::ostringstream ss; ss << val;
        std::lock_guard<std::mutex> lk(mu_);
        kv_[key] = ss.str();
    }
    bool get_conf(const std::string& key, std::string& val) const
    {
        std::lock_guard<std::mutex> lk(mu_);
        auto it = kv_.find(key);
        if (it == kv_.end()) return false;
        val = it->second;
        return true;
    }
 private:
    mutable std::mutex mu_;
    std::map<std::string, std::string> kv_;
};
typedef std::shared_ptr<ConfigTable> ConfigTable_ptr;

}  // namespace blaze
