// pmm_types.h -- plain data shared by the host planner (pmm_plan.cpp), the engine and the kernels.  No CUDA here.
#pragma once
#include <cstdint>

namespace pmm {

constexpr int kMaxGroups = 4;           // reads per warp: 32 / W, W >= 8
constexpr int kWarpsPerCta = 4;
constexpr int kF64K = 6;                // rows per lane of the double kernel (W = 32): 191-base reads in one stripe
constexpr int kStripedK = 8;            // rows per lane of the striped float kernel (W = 32)

struct ReadDesc { uint32_t off, stride, len; };
struct HapDesc  { uint32_t off, len; };
struct RegionDesc { uint32_t read_first, nreads, hap_first, nhaps, out_first; };

// One unit of work for one warp.  Group g (lanes g*W .. g*W+W-1) owns read[g]; all groups walk the same run of
// haplotypes [hap_first, hap_first + nhaps).  The result for (read[g], hap_first + n) goes to out[out_base[g] + n].
struct Task {
    uint32_t read[kMaxGroups];
    uint32_t out_base[kMaxGroups];
    uint32_t hap_first;
    uint32_t nhaps;
    uint32_t nreads;
    uint32_t reserved;
};
static_assert(sizeof(Task) == 48, "Task layout");

// (K rows per lane, W lanes per read) instantiations of the float kernel.
#define PMM_F32_VARIANTS(X) \
    X(4, 8) X(5, 8) X(6, 8) X(7, 8) X(8, 8) X(10, 8) X(12, 8) X(13, 8) X(14, 8) X(16, 8) \
    X(4, 16) X(5, 16) X(6, 16) X(7, 16) X(8, 16) X(9, 16) X(10, 16) X(11, 16) X(12, 16) X(14, 16) X(16, 16) \
    X(4, 32) X(5, 32) X(6, 32) X(7, 32) X(8, 32) X(9, 32) X(10, 32) X(12, 32) X(14, 32) X(16, 32)

inline bool forward_f32_has_variant(int K, int W)
{
#define PMM_X(k, w) if (K == k && W == w) return true;
    PMM_F32_VARIANTS(PMM_X)
#undef PMM_X
    return false;
}

struct Variant { int K, W; bool striped; };
inline bool operator==(const Variant& a, const Variant& b) { return a.K == b.K && a.W == b.W && a.striped == b.striped; }

}  // namespace pmm
