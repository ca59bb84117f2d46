// blaze/Client.h -- base class of accelerator clients (see Common.h for scope).  Same surface as the reference uses
// (client/PairHMMClient.h:9-21, client/PairHMMClient.cpp:12-17,57-65,75-76, client/PairHMMWorker.cpp:241,251):
// numbered input and output blocks, start() to run the accelerator, a virtual compute() that Blaze calls when the
// accelerator is unreachable or its task fails.
#pragma once
#include "PlatformManager.h"

namespace blaze {

class Client {
 public:
    Client(const std::string& acc_id, int num_inputs, int num_outputs, int port = 1027)
        : acc_id_(acc_id), port_(port), inputs_(num_inputs), capacity_(num_inputs, 0), outputs_(num_outputs) {}
    virtual ~Client() {}

    // Allocate (or re-declare the size of) input block idx: num_items x item_length elements of data_width bytes.
    // A block only ever grows; declaring a smaller size keeps the allocation and its contents.
    void* createInput(int idx, int num_items, int item_length, int data_width)
    {
        check_in(idx);
        const size_t bytes = (size_t)num_items * (size_t)item_length * (size_t)data_width;
        if (!inputs_[idx] || bytes > capacity_[idx]) {
            inputs_[idx].reset(new DataBlock(num_items, item_length, bytes, 4096));
            capacity_[idx] = bytes;
        } else {
            inputs_[idx]->resize_within(bytes, num_items, item_length);
        }
        return inputs_[idx]->getData();
    }
    // Copy caller data into input block idx.
    void* setInput(int idx, void* src, int num_items, int item_length, int data_width)
    {
        void* dst = createInput(idx, num_items, item_length, data_width);
        memcpy(dst, src, (size_t)num_items * (size_t)item_length * (size_t)data_width);
        return dst;
    }
    void* getInputPtr(int idx)
    {
        check_in(idx);
        if (!inputs_[idx]) throw invalidParam("input block not created");
        return inputs_[idx]->getData();
    }
    size_t getInputCapacity(int idx) { check_in(idx); return capacity_[idx]; }

    void* createOutput(int idx, int num_items, int item_length, int data_width)
    {
        check_out(idx);
        const size_t bytes = (size_t)num_items * (size_t)item_length * (size_t)data_width;
        outputs_[idx].reset(new DataBlock(num_items, item_length, bytes, 4096));
        return outputs_[idx]->getData();
    }
    void* getOutputPtr(int idx)
    {
        check_out(idx);
        if (!outputs_[idx]) throw invalidParam("output block not available");
        return outputs_[idx]->getData();
    }
    int getOutputNumItems(int idx) { check_out(idx); return outputs_[idx] ? outputs_[idx]->getNumItems() : 0; }
    int getOutputLength(int idx) { check_out(idx); return outputs_[idx] ? outputs_[idx]->getItemLength() : 0; }
    size_t getOutputSize(int idx) { check_out(idx); return outputs_[idx] ? outputs_[idx]->getSize() : 0; }
    int getNumOutputs() const { return (int)outputs_.size(); }

    // Run the accelerator on the current input blocks; on return the output blocks are readable.  If no manager
    // serves this accelerator, or its task throws, the client's own compute() runs instead (Blaze's contract).
    void start()
    {
        last_error_.clear();
        PlatformManager* pm = AppCommManager::lookup(port_);
        Accelerator* acc = pm ? pm->find(acc_id_) : nullptr;
        if (!acc) {
            last_error_ = "no accelerator manager serves \"" + acc_id_ + "\"";
            compute();
            return;
        }
        std::vector<DataBlock_ptr> out;
        try {
            acc->run(inputs_, out);
        } catch (const std::exception& e) {
            last_error_ = e.what();
            compute();
            return;
        }
        for (size_t k = 0; k < outputs_.size(); ++k) outputs_[k] = k < out.size() ? out[k] : DataBlock_ptr();
    }

    virtual void compute() = 0;

    const std::string& lastError() const { return last_error_; }

 private:
    void check_in(int idx) const { if (idx < 0 || idx >= (int)inputs_.size()) throw invalidParam("input index out of range"); }
    void check_out(int idx) const { if (idx < 0 || idx >= (int)outputs_.size()) throw invalidParam("output index out of range"); }

    std::string acc_id_;
    int port_;
    std::vector<DataBlock_ptr> inputs_;
    std::vector<size_t> capacity_;
    std::vector<DataBlock_ptr> outputs_;
    std::string last_error_;
};

}  // namespace blaze
