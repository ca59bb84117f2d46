// blaze/PlatformManager.h -- the in-process accelerator manager (see Common.h for scope).
//
// In the reference the manager is a separate service: clients ship their blocks to it (AppCommManager, port 1027,
// host/main.cpp:264-273), it dlopen()s the task plugin named in its conf (pairhmm/xlnx.conf:9-16), queues the task on
// the accelerator's executor and ships the output blocks back.  Here the same steps happen inside Client::start():
// the Accelerator loads the plugin once, owns one CudaEnv per (GPU, slot) and hands each incoming task to the
// least-loaded free one -- that queue is how several client threads spread over the GPUs of a box; there is no
// collective and no NCCL because read x haplotype regions are independent.
#pragma once
#include <dlfcn.h>

#include <condition_variable>
#include <fstream>
#include <vector>

#include "Task.h"

namespace blaze {

// ---- manager configuration -------------------------------------------------------------------------------------
// Text format of the reference's conf files (protobuf text: nested `name { ... }` groups and `key: value` fields).
struct AccConf {
    std::string id, path;
    std::map<std::string, std::string> param;
};
struct PlatformConf {
    std::string id, path, cache_loc;
    std::vector<AccConf> acc;
};
class ManagerConf {
 public:
    int verbose_ = 0;
    std::vector<PlatformConf> platform;
    int verbose() const { return verbose_; }

    // Returns false with a message on malformed input.
    bool ParseFromString(const std::string& text, std::string* err = nullptr)
    {
        std::vector<std::string> tok;
        if (!tokenize(text, tok, err)) return false;
        size_t i = 0;
        while (i < tok.size()) {
            if (tok[i] == "verbose:") { if (!need(tok, i + 1, err)) return false; verbose_ = atoi(tok[i + 1].c_str()); i += 2; }
            else if (tok[i] == "platform" && i + 1 < tok.size() && tok[i + 1] == "{") {
                i += 2; PlatformConf p;
                while (i < tok.size() && tok[i] != "}") {
                    if (tok[i] == "id:") { if (!need(tok, i + 1, err)) return false; p.id = tok[i + 1]; i += 2; }
                    else if (tok[i] == "path:") { if (!need(tok, i + 1, err)) return false; p.path = tok[i + 1]; i += 2; }
                    else if (tok[i] == "cache_loc:") { if (!need(tok, i + 1, err)) return false; p.cache_loc = tok[i + 1]; i += 2; }
                    else if (tok[i] == "acc" && i + 1 < tok.size() && tok[i + 1] == "{") {
                        i += 2; AccConf a;
                        while (i < tok.size() && tok[i] != "}") {
                            if (tok[i] == "id:") { if (!need(tok, i + 1, err)) return false; a.id = tok[i + 1]; i += 2; }
                            else if (tok[i] == "path:") { if (!need(tok, i + 1, err)) return false; a.path = tok[i + 1]; i += 2; }
                            else if (tok[i] == "param" && i + 1 < tok.size() && tok[i + 1] == "{") {
                                i += 2; std::string k, v;
                                while (i < tok.size() && tok[i] != "}") {
                                    if (tok[i] == "key:") { if (!need(tok, i + 1, err)) return false; k = tok[i + 1]; i += 2; }
                                    else if (tok[i] == "value:") { if (!need(tok, i + 1, err)) return false; v = tok[i + 1]; i += 2; }
                                    else return fail(err, "unexpected token in param: " + tok[i]);
                                }
                                if (i >= tok.size()) return fail(err, "unterminated param group");
                                ++i; a.param[k] = v;
                            } else return fail(err, "unexpected token in acc: " + tok[i]);
                        }
                        if (i >= tok.size()) return fail(err, "unterminated acc group");
                        ++i; p.acc.push_back(a);
                    } else return fail(err, "unexpected token in platform: " + tok[i]);
                }
                if (i >= tok.size()) return fail(err, "unterminated platform group");
                ++i; platform.push_back(p);
            } else return fail(err, "unexpected token: " + tok[i]);
        }
        return true;
    }
    bool ParseFromFile(const std::string& path, std::string* err = nullptr)
    {
        std::ifstream in(path.c_str());
        if (!in.good()) return fail(err, "cannot open " + path);
        std::stringstream ss; ss << in.rdbuf();
        return ParseFromString(ss.str(), err);
    }

 private:
    static bool fail(std::string* err, const std::string& m) { if (err) *err = m; return false; }
    static bool need(const std::vector<std::string>& t, size_t i, std::string* err) { return i < t.size() ? true : fail(err, "value missing at end of conf"); }
    static bool tokenize(const std::string& s, std::vector<std::string>& out, std::string* err)
    {
        size_t i = 0;
        while (i < s.size()) {
            const char ch = s[i];
            if (isspace((unsigned char)ch)) { ++i; continue; }
            if (ch == '#') { while (i < s.size() && s[i] != '\n') ++i; continue; }
            if (ch == '{' || ch == '}') { out.push_back(std::string(1, ch)); ++i; continue; }
            if (ch == '"') {
                size_t j = s.find('"', i + 1);
                if (j == std::string::npos) return fail(err, "unterminated string");
                out.push_back(s.substr(i + 1, j - i - 1)); i = j + 1; continue;
            }
            size_t j = i;
            while (j < s.size() && !isspace((unsigned char)s[j]) && s[j] != '{' && s[j] != '}' && s[j] != '"') {
                if (s[j] == ':') { ++j; break; }
                ++j;
            }
            out.push_back(s.substr(i, j - i)); i = j;
        }
        return true;
    }
};

// ---- one registered accelerator -------------------------------------------------------------------------------
class Accelerator {
 public:
    typedef Task* (*create_fn)();
    typedef void (*destroy_fn)(Task*);

    // devices: CUDA device indices to run on; slots_per_device tasks may be in flight on each at once
    Accelerator(const std::string& id, const std::string& plugin_path, const std::map<std::string, std::string>& param,
                const std::vector<int>& devices, int slots_per_device)
        : id_(id), conf_(new ConfigTable())
    {
        handle_ = dlopen(plugin_path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!handle_) throw invalidParam("cannot load task plugin " + plugin_path + ": " + dlerror());
        create_ = reinterpret_cast<create_fn>(dlsym(handle_, "create"));
        destroy_ = reinterpret_cast<destroy_fn>(dlsym(handle_, "destroy"));
        if (!create_ || !destroy_) throw invalidParam(plugin_path + " does not export create()/destroy()");
        for (auto& kv : param) conf_->write_conf(kv.first, kv.second);
        if (devices.empty()) throw invalidParam("accelerator " + id + " has no device");
        for (int s = 0; s < std::max(1, slots_per_device); ++s)
            for (int d : devices) { envs_.emplace_back(new CudaEnv(d, s)); busy_.push_back(false); }
        load_.assign(envs_.size(), 0);
    }
    ~Accelerator()
    {
        envs_.clear();                         // scratch objects reference code of the plugin: drop them first
        // The plugin stays loaded: output blocks it created (their deleters are its code) may outlive the manager in the
        // hands of clients.
    }
    const std::string& id() const { return id_; }
    int numEnvs() const { return (int)envs_.size(); }
    uint64_t tasksRunOn(int env) const { return load_[env]; }
    int deviceOf(int env) const { return envs_[env]->getDevice(); }

    // Run one task to completion on a free environment; blocks while all are busy.  Exceptions of the plugin pass
    // through to the caller (Client::start turns them into its fallback path).
    void run(const std::vector<DataBlock_ptr>& inputs, std::vector<DataBlock_ptr>& outputs)
    {
        const int slot = acquire();
        Task* t = nullptr;
        try {
            t = create_();
            if (!t) throw std::runtime_error("task plugin create() returned null");
            if (t->getNumInputs() != (int)inputs.size()) throw invalidParam("task expects a different number of input blocks");
            t->env_ = envs_[slot].get();
            t->conf_ = conf_;
            t->inputs_ = inputs;
            t->prepare();
            t->compute();
            outputs = t->outputs_;
            destroy_(t);
        } catch (...) {
            if (t) destroy_(t);
            release(slot);
            throw;
        }
        release(slot);
    }

 private:
    int acquire()
    {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            int best = -1;
            for (size_t k = 0; k < envs_.size(); ++k)
                if (!busy_[k] && (best < 0 || load_[k] < load_[best])) best = (int)k;
            if (best >= 0) { busy_[best] = true; load_[best]++; return best; }
            cv_.wait(lk);
        }
    }
    void release(int slot)
    {
        { std::lock_guard<std::mutex> lk(mu_); busy_[slot] = false; }
        cv_.notify_one();
    }

    std::string id_;
    ConfigTable_ptr conf_;
    void* handle_ = nullptr;
    create_fn create_ = nullptr;
    destroy_fn destroy_ = nullptr;
    std::vector<std::unique_ptr<CudaEnv> > envs_;
    std::vector<bool> busy_;
    std::vector<uint64_t> load_;
    std::mutex mu_;
    std::condition_variable cv_;
};

// ---- the manager -------------------------------------------------------------------------------------------------
class PlatformManager {
 public:
    PlatformManager() {}
    // Every `acc` of every platform in the conf is registered.  Params understood by the manager itself:
    //   "devices"          comma-separated CUDA device indices (default: "0"); "all" = every visible GPU is resolved by
    //                      the plugin's own device count, passed in through num_visible_devices
    //   "slots_per_device" tasks in flight per GPU (default 2)
    explicit PlatformManager(const ManagerConf* conf, int num_visible_devices = 1)
    {
        for (const PlatformConf& p : conf->platform)
            for (const AccConf& a : p.acc) registerAcc(a.id, a.path, a.param, num_visible_devices);
    }
    void registerAcc(const std::string& id, const std::string& plugin_path, const std::map<std::string, std::string>& param,
                     int num_visible_devices = 1)
    {
        std::vector<int> devices;
        auto it = param.find("devices");
        if (it == param.end()) devices.push_back(0);
        else if (it->second == "all") for (int d = 0; d < std::max(1, num_visible_devices); ++d) devices.push_back(d);
        else {
            std::stringstream ss(it->second); std::string item;
            while (std::getline(ss, item, ',')) if (!item.empty()) devices.push_back(atoi(item.c_str()));
        }
        int slots = 2;
        it = param.find("slots_per_device");
        if (it != param.end()) slots = atoi(it->second.c_str());
        std::lock_guard<std::mutex> lk(mu_);
        acc_[id].reset(new Accelerator(id, plugin_path, param, devices, slots));
    }
    Accelerator* find(const std::string& id)
    {
        std::lock_guard<std::mutex> lk(mu_);
        auto it = acc_.find(id);
        return it == acc_.end() ? nullptr : it->second.get();
    }

 private:
    std::mutex mu_;
    std::map<std::string, std::unique_ptr<Accelerator> > acc_;
};

// The endpoint clients of this process connect to.  Constructing one publishes the manager under (ip, port) like the
// reference's `blaze::AppCommManager comm(&platform_manager, "127.0.0.1", 1027)` (host/main.cpp:273); destroying it
// withdraws it.  No socket is opened.
class AppCommManager {
 public:
    AppCommManager(PlatformManager* pm, const std::string& ip = "127.0.0.1", int port = 1027) : port_(port)
    {
        (void)ip;
        std::lock_guard<std::mutex> lk(registry_mu());
        registry()[port_] = pm;
    }
    ~AppCommManager()
    {
        std::lock_guard<std::mutex> lk(registry_mu());
        registry().erase(port_);
    }
    static PlatformManager* lookup(int port)
    {
        std::lock_guard<std::mutex> lk(registry_mu());
        auto it = registry().find(port);
        return it == registry().end() ? nullptr : it->second;
    }

 private:
    // Never destroyed: endpoints may be withdrawn from static destructors of the host program, after function-local statics
    // of this header would already be gone.
    static std::map<int, PlatformManager*>& registry() { static std::map<int, PlatformManager*>* r = new std::map<int, PlatformManager*>(); return *r; }
    static std::mutex& registry_mu() { static std::mutex* m = new std::mutex(); return *m; }
    int port_;
};

}  // namespace blaze
