// blaze/Task.h -- base class of accelerator task plugins (see Common.h for scope).  A plugin is a shared object
// exporting `extern "C" blaze::Task* create()` and `extern "C" void destroy(blaze::Task*)`
// (task/xlnx/PairHMMTask.h:90-96); the manager fills in env, conf and input blocks, then calls prepare() and
// compute() and collects the output blocks.
#pragma once
#include <vector>

#include "Block.h"
#include "TaskEnv.h"

namespace blaze {

class Accelerator;

class Task {
 public:
    explicit Task(int num_inputs) : inputs_(num_inputs) {}
    virtual ~Task() {}

    virtual void prepare() {}
    virtual void compute() = 0;
    virtual uint64_t estimateClientTime() { return 0; }
    virtual uint64_t estimateTaskTime() { return 0; }

    int getNumInputs() const { return (int)inputs_.size(); }
    int getNumOutputs() const { return (int)outputs_.size(); }

 protected:
    TaskEnv* getEnv() { return env_; }
    // pointer to the bytes of input block idx
    void* getInput(int idx)
    {
        if (idx < 0 || idx >= (int)inputs_.size() || !inputs_[idx]) throw invalidParam("task input block missing");
        return inputs_[idx]->getData();
    }
    size_t getInputLength(int idx)
    {
        if (idx < 0 || idx >= (int)inputs_.size() || !inputs_[idx]) throw invalidParam("task input block missing");
        return inputs_[idx]->getSize();
    }
    void setOutput(int idx, DataBlock_ptr block)
    {
        if (idx < 0) throw invalidParam("negative output index");
        if (idx >= (int)outputs_.size()) outputs_.resize(idx + 1);
        outputs_[idx] = block;
    }
    bool get_conf(const std::string& key, std::string& val)
    {
        return conf_ ? conf_->get_conf(key, val) : false;
    }

 private:
    friend class Accelerator;
    TaskEnv* env_ = nullptr;
    ConfigTable_ptr conf_;
    std::vector<DataBlock_ptr> inputs_, outputs_;
};

}  // namespace blaze
