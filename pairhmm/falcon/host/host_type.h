// host_type.h -- batch types of the standalone (non-Blaze) entry of the PairHMM path, with the reference's names and
// field order (/root/reference/pairhmm/xlnx/host/host_type.h:97-119): a read is five equally long byte strings (bases
// and the base / insertion / deletion / gap-continuation qualities), a haplotype its bases, an input batch the cross
// product of the two vectors, the output one log10 likelihood per pair, read-major.
#ifndef FALCON_HOST_TYPE_H
#define FALCON_HOST_TYPE_H
#include <string>
#include <vector>

typedef struct {
  std::string bases;
  std::string _q;
  std::string _i;
  std::string _d;
  std::string _c;
} Read;

typedef struct {
  std::string bases;
} Hap;

typedef struct {
  std::vector<Read> reads;
  std::vector<Hap> haps;
} pairhmmInput;

typedef struct {
  std::vector<double> likelihoodData;
} pairhmmOutput;

#endif
