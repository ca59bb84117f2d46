// blaze/Client.h -- base class of accelerator clients (see Common.h for scope).  Same surface as the reference uses
// (client/PairHMMClient.h:9-21, client/PairHMMClient.cpp:12-17,57-65,75-76, client/PairHMMWorker.cpp:241,251):
// numbered input and output blocks, start() to run the accelerator, a virtual compute() that Blaze calls when the
// accelerator is unreachable or its task fails.
//
// Beyond the reference's surface: startAsync() / wait().  The reference's worker overlaps its CPU share with the
// accelerator on a side thread (client/PairHMMWorker.cpp:213-214); here there is no CPU share, and what is worth
// overlapping is the accelerator's own tiles: startAsync() hands the current input blocks to a helper thread that runs
// the task, the caller fills the next set of input blocks meanwhile, and wait() delivers the outputs of the oldest
// task in flight.  At most kMaxInFlight tasks are in flight; each needs a free slot of the accelerator (four per GPU by
// default), so the tiles of one batch really run side by side: the packing and copies of one under the kernels of the
// others.
#pragma once
#include <atomic>
#include <deque>
#include <thread>

#include "PlatformManager.h"

namespace blaze {

class Client {
 public:
    static constexpr int kMaxInFlight = 4;

    Client(const std::string& acc_id, int num_inputs, int num_outputs, int port = 1027)
        : acc_id_(acc_id), port_(port), inputs_(num_inputs), capacity_(num_inputs, 0), outputs_(num_outputs) {}
    virtual ~Client()
    {
        for (auto& l : lanes_) {
            { std::lock_guard<std::mutex> lk(l->mu); l->stop = true; }
            l->cv.notify_all();
            if (l->th.joinable()) l->th.join();
        }
    }

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

    // Start the accelerator on the current input blocks and return at once.  The blocks travel with the task; the next
    // createInput()/setInput() fills a fresh (recycled) set.  Throws if kMaxInFlight tasks are already in flight.
    void startAsync()
    {
        if ((int)flights_.size() >= kMaxInFlight) throw invalidParam("Client::startAsync: too many tasks in flight, call wait()");
        Lane* lane = nullptr;
        for (auto& l : lanes_) if (!l->busy) { lane = l.get(); break; }
        if (!lane) {
            lanes_.emplace_back(new Lane());
            lane = lanes_.back().get();
            lane->th = std::thread([lane] { lane->loop(); });
        }
        lane->inputs.swap(inputs_);              // the task owns this set now
        lane->in_capacity.swap(capacity_);
        lane->outputs.clear(); lane->error.clear(); lane->failed = false;
        lane->acc_id = acc_id_; lane->port = port_;
        // the next batch is packed into the set the oldest finished task left behind (grow-only blocks, recycled)
        inputs_.assign(lane->inputs.size(), DataBlock_ptr()); capacity_.assign(lane->inputs.size(), 0);
        if (!spare_inputs_.empty()) { inputs_.swap(spare_inputs_.back().first); capacity_.swap(spare_inputs_.back().second); spare_inputs_.pop_back(); }
        lane->done_flag.store(false, std::memory_order_relaxed);
        { std::lock_guard<std::mutex> lk(lane->mu); lane->busy = true; lane->go = true; lane->done = false; }
        lane->flag.store(true, std::memory_order_release);
        lane->cv.notify_all();
        flights_.push_back(lane);
    }
    // Wait for the oldest task in flight; afterwards getOutputPtr() etc. refer to its output blocks.  If that task
    // failed, the client's own compute() runs (Blaze's contract), like in start().
    void wait()
    {
        if (flights_.empty()) throw invalidParam("Client::wait: nothing in flight");
        Lane* lane = flights_.front(); flights_.pop_front();
        for (int spin = 0; spin < 4000 && !lane->done_flag.load(std::memory_order_acquire); ++spin) Lane::cpu_relax();
        { std::unique_lock<std::mutex> lk(lane->mu); lane->cv.wait(lk, [&] { return lane->done; }); lane->busy = false; }
        spare_inputs_.emplace_back(std::move(lane->inputs), std::move(lane->in_capacity));
        lane->inputs.clear(); lane->in_capacity.clear();
        last_error_.clear();
        if (lane->failed) { last_error_ = lane->error; compute(); return; }
        for (size_t k = 0; k < outputs_.size(); ++k) outputs_[k] = k < lane->outputs.size() ? lane->outputs[k] : DataBlock_ptr();
    }
    int inFlight() const { return (int)flights_.size(); }

    virtual void compute() = 0;

    const std::string& lastError() const { return last_error_; }

 private:
    void check_in(int idx) const { if (idx < 0 || idx >= (int)inputs_.size()) throw invalidParam("input index out of range"); }
    void check_out(int idx) const { if (idx < 0 || idx >= (int)outputs_.size()) throw invalidParam("output index out of range"); }

    // one helper thread and the task it is running
    struct Lane {
        std::thread th;
        std::mutex mu;
        std::condition_variable cv;
        bool go = false, done = false, stop = false, busy = false, failed = false;
        std::atomic<bool> flag{false}, done_flag{false};   // lock-free mirrors of go / done for the short polls
        static void cpu_relax()
        {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
        std::string acc_id, error;
        int port = 0;
        std::vector<DataBlock_ptr> inputs, outputs;
        std::vector<size_t> in_capacity;
        void loop()
        {
            for (;;) {
                // A tile follows the previous one within microseconds while a batch is in flight: poll briefly before
                // sleeping, a condition-variable wake-up costs more than a tile's host work
                for (int spin = 0; spin < 2000 && !flag.load(std::memory_order_acquire); ++spin) cpu_relax();
                { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return go || stop; }); if (stop) return; go = false; flag.store(false, std::memory_order_relaxed); }
                try {
                    PlatformManager* pm = AppCommManager::lookup(port);
                    Accelerator* acc = pm ? pm->find(acc_id) : nullptr;
                    if (!acc) { failed = true; error = "no accelerator manager serves \"" + acc_id + "\""; }
                    else acc->run(inputs, outputs);
                } catch (const std::exception& e) { failed = true; error = e.what(); }
                { std::lock_guard<std::mutex> lk(mu); done = true; }
                done_flag.store(true, std::memory_order_release);
                cv.notify_all();
            }
        }
    };

    std::string acc_id_;
    int port_;
    std::vector<DataBlock_ptr> inputs_;
    std::vector<size_t> capacity_;
    std::vector<DataBlock_ptr> outputs_;
    std::string last_error_;
    std::vector<std::unique_ptr<Lane> > lanes_;
    std::deque<Lane*> flights_;
    std::vector<std::pair<std::vector<DataBlock_ptr>, std::vector<size_t> > > spare_inputs_;
};

}  // namespace blaze
