// pmm_tables.cpp -- see pmm_tables.h.  Follows Context.h:42-90 (tables) and :105-110,:145-151 (constants).
#include "pmm_tables.h"

#include <cmath>
#include <mutex>
#include <vector>

namespace pmm {
namespace {

constexpr int kMaxQual = 127;              // reachable range only, see kM2mSize
constexpr double kJacTolerance = 8.0;      // MAX_JACOBIAN_TOLERANCE
constexpr double kJacStep = 0.0001;        // JACOBIAN_LOG_TABLE_STEP
constexpr int kJacSize = (int)(kJacTolerance / kJacStep) + 1;

// log10(10^a + 10^b) through the quantised Jacobian table, evaluated in NUMBER (Context.h:67-90).
template <class NUMBER>
NUMBER log10_sum(NUMBER lo, NUMBER hi, const std::vector<NUMBER>& jac)
{
    if (lo > hi) { NUMBER t = hi; hi = lo; lo = t; }
    if (std::isinf(lo) || std::isinf(hi)) return hi;       // only -inf can occur; never with finite quals
    NUMBER gap = hi - lo;
    if (gap >= (NUMBER)kJacTolerance) return hi;
    NUMBER scaled = (NUMBER)(gap * (NUMBER)(1.0 / kJacStep));
    int idx = scaled > (NUMBER)0.0 ? (int)(scaled + (NUMBER)0.5) : (int)(scaled - (NUMBER)0.5);   // Context.h:63-65
    return hi + jac[idx];
}

template <class NUMBER>
void fill_m2m(NUMBER* out)
{
    std::vector<NUMBER> jac(kJacSize);
    for (int k = 0; k < kJacSize; ++k)
        jac[k] = (NUMBER)(std::log10(1.0 + std::pow(10.0, -((double)k) * kJacStep)));          // Context.h:45
    const double inv_ln10 = 1.0 / std::log(10);
    for (int hi = 0, base = 0; hi <= kMaxQual; base += ++hi)
        for (int lo = 0; lo <= hi; ++lo) {
            // arguments are narrowed to NUMBER before the sum; the rest is double (Context.h:56-59)
            double s = log10_sum<NUMBER>((NUMBER)(-0.1 * hi), (NUMBER)(-0.1 * lo), jac);
            double l = std::log1p(-std::fmin(1.0, std::pow(10, s))) * inv_ln10;
            out[base + lo] = (NUMBER)std::pow(10, l);
        }
}

HostTables* build()
{
    HostTables* t = new HostTables();
    for (int x = 0; x < kPh2prSize; ++x) {
        t->ph2pr_f[x] = powf(10.f, -((float)x) / 10.f);                                        // Context.h:146
        t->ph2pr_d[x] = std::pow(10.0, -((double)x) / 10.0);                                   // Context.h:106
    }
    fill_m2m<float>(t->m2m_f);
    fill_m2m<double>(t->m2m_d);
    t->ic_f = ldexpf(1.f, 120);
    t->ic_d = std::ldexp(1.0, 1020);
    t->log10_ic_f = log10f(t->ic_f);
    t->log10_ic_d = std::log10(t->ic_d);
    return t;
}

}  // namespace

const HostTables& host_tables()
{
    static std::once_flag once;
    static HostTables* tables = nullptr;
    std::call_once(once, [] { tables = build(); });
    return *tables;
}

}  // namespace pmm
