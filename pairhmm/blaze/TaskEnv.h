// blaze/TaskEnv.h -- what a task gets from getEnv(): a per-device environment with named scratch objects that
// survive from one task instance to the next (the reference recycles its pinned input bundle and its output block
// this way, task/xlnx/PairHMMTask.cpp:19-25,48,70).  CudaEnv stands where blaze::OpenCLEnv is
// (task/xlnx/PairHMMTask.h:13): instead of a cl_context / command queue / kernel it names a CUDA device.
#pragma once
#include "Block.h"

namespace blaze {

class TaskEnv {
 public:
    virtual ~TaskEnv() {}

    template <typename T> bool getScratch(const std::string& name, std::shared_ptr<T>& out)
    {
        std::lock_guard<std::mutex> lk(mu_);
        auto it = scratch_.find(name);
        if (it == scratch_.end() || !it->second) return false;
        out = std::static_pointer_cast<T>(it->second);
        scratch_.erase(it);                     // one owner at a time: the task puts it back in its destructor
        return true;
    }
    template <typename T> void putScratch(const std::string& name, const std::shared_ptr<T>& obj)
    {
        if (!obj) return;
        std::lock_guard<std::mutex> lk(mu_);
        scratch_[name] = std::static_pointer_cast<void>(obj);
    }
    DataBlock_ptr create_block(int num_items, int item_length, size_t bytes, int align = 0,
                               DataBlock::Flag flag = DataBlock::OWNED, ConfigTable_ptr conf = ConfigTable_ptr())
    {
        return DataBlock_ptr(new DataBlock(num_items, item_length, bytes, align, flag, conf));
    }

 private:
    std::mutex mu_;
    std::map<std::string, std::shared_ptr<void> > scratch_;
};

class CudaEnv : public TaskEnv {
 public:
    explicit CudaEnv(int device, int slot = 0) : device_(device), slot_(slot) {}
    int getDevice() const { return device_; }
    // which of the device's task slots this environment is (0 = the one a free accelerator hands out first)
    int getSlot() const { return slot_; }
 private:
    int device_, slot_;
};

}  // namespace blaze
