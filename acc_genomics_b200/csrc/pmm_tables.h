// pmm_tables.h -- quality-to-probability tables of the PairHMM forward path, generated with the HOST libm.
//
// The reference builds these once per process in Context<NUMBER> (/root/reference/pairhmm/xlnx/host/Context.h:
// 42-61 jacobian + matchToMatch, :105-110 and :145-151 ph2pr and the scaling constants).  They must come from
// the same libm calls (powf / pow / log10 / log1p) to be bit-identical, so they are produced here on the host and
// uploaded; the device never recomputes them.
#pragma once
#include <cstdint>

namespace pmm {

constexpr int kPh2prSize = 128;
// quals are masked with 127 before every lookup (avx-pairhmm-template.h:110-112), so only the first
// 127*128/2 + 127 + 1 entries of the reference's 32640-entry matchToMatchProb table are reachable.
constexpr int kM2mSize = 8256;

struct HostTables {
    float  ph2pr_f[kPh2prSize];
    double ph2pr_d[kPh2prSize];
    float  m2m_f[kM2mSize];
    double m2m_d[kM2mSize];
    float  ic_f;      // 2^120   (Context.h:149)
    double ic_d;      // 2^1020  (Context.h:109)
    float  log10_ic_f;
    double log10_ic_d;
};

// Built on first use; thread-safe.
const HostTables& host_tables();

}  // namespace pmm
