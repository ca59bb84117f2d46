// PairHMMClient.cpp -- see PairHMMClient.h.
#include "PairHMMClient.h"

PairHMMClient::PairHMMClient()
    : blaze::Client("PairHMM", 3, 3), num_read_(0), num_hap_(0), num_cell_(0), reads_(nullptr), haps_(nullptr) {}

void PairHMMClient::setup(read_t* reads, int num_read, hap_t* haps, int num_hap) {
  if (!reads || !haps || num_read <= 0 || num_hap <= 0) throw blaze::invalidParam("PairHMMClient::setup: empty batch");
  num_read_ = num_read;
  num_hap_  = num_hap;
  reads_    = reads;
  haps_     = haps;

  // the reference's cell count (client/PairHMMClient.cpp:47-55): total read bases x total haplotype bases
  uint64_t read_bases = 0, hap_bases = 0;
  for (int k = 0; k < num_read; ++k) read_bases += (uint64_t)reads[k].len;
  for (int k = 0; k < num_hap; ++k)  hap_bases  += (uint64_t)haps[k].len;
  num_cell_ = read_bases * hap_bases;
  setInput(0, &num_cell_, 1, 1, sizeof(uint64_t));

  // blocks 1 and 2: the wire format, written straight into the (grow-only) input blocks
  const uint64_t read_bytes = serialized_size(reads, num_read);
  const uint64_t hap_bytes  = serialized_size(haps, num_hap);
  if (read_bytes > 0x7fffffffull || hap_bytes > 0x7fffffffull)
    throw blaze::invalidParam("PairHMMClient::setup: batch larger than 2 GiB, split it (PairHMMWorker does)");
  serialize(createInput(1, 1, (int)read_bytes, 1), reads, num_read);
  serialize(createInput(2, 1, (int)hap_bytes, 1), haps, num_hap);
}

void PairHMMClient::compute() {
  throw std::runtime_error("PairHMM accelerator failed and this build has no CPU fallback: " + lastError());
}
