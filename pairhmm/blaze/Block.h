// blaze/Block.h -- DataBlock: a sized host buffer travelling between client and task (see Common.h for scope).
#pragma once
#include "Common.h"

namespace blaze {

class DataBlock {
 public:
    enum Flag { OWNED, SHARED };

    // same argument order as the reference's call sites (task/xlnx/PairHMMTask.cpp:72-74)
    DataBlock(int num_items, int item_length, size_t bytes, int align = 0, Flag flag = OWNED,
              ConfigTable_ptr conf = ConfigTable_ptr())
        : num_items_(num_items), item_length_(item_length), bytes_(bytes), flag_(flag), conf_(conf)
    {
        if (flag == OWNED && bytes) {
            const size_t a = align > 0 ? (size_t)align : 64;
            const size_t padded = (bytes + a - 1) / a * a;
            data_ = static_cast<char*>(aligned_alloc(a, padded));
            if (!data_) throw std::bad_alloc();
        }
    }
    // a view of caller memory (Client::setInput)
    DataBlock(void* borrowed, int num_items, int item_length, size_t bytes)
        : num_items_(num_items), item_length_(item_length), bytes_(bytes), flag_(SHARED), data_(static_cast<char*>(borrowed)) {}
    ~DataBlock() { if (flag_ == OWNED) free(data_); }
    DataBlock(const DataBlock&) = delete;
    DataBlock& operator=(const DataBlock&) = delete;

    char* getData() { return data_; }
    const char* getData() const { return data_; }
    size_t getSize() const { return bytes_; }
    int getNumItems() const { return num_items_; }
    int getItemLength() const { return item_length_; }
    // the logical size may shrink below the allocation (Client::createInput is called again "just to update block
    // size", client/PairHMMClient.cpp:63-65)
    void resize_within(size_t bytes, int num_items, int item_length)
    {
        bytes_ = bytes; num_items_ = num_items; item_length_ = item_length;
    }
    ConfigTable_ptr conf() const { return conf_; }

 private:
    int num_items_, item_length_;
    size_t bytes_;
    Flag flag_;
    ConfigTable_ptr conf_;
    char* data_ = nullptr;
};
typedef std::shared_ptr<DataBlock> DataBlock_ptr;

}  // namespace blaze
