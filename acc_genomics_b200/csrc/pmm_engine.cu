// pmm_engine.cu -- the C ABI of include/pairhmm_cuda.h: host-side packing, task queue construction, kernel
// sequencing, result collection.  This is the layer that stands where the reference has pack_fpga_input +
// OpenCL buffer migration + clEnqueueTask (/root/reference/pairhmm/task/xlnx/PairHMMTask.cpp:27-143,
// /root/reference/pairhmm/host/PairHMMFpga.cpp:125-162).  There is no CPU compute path in this file: without a
// GPU every entry point fails with PMM_ERR_NO_DEVICE.
#include "../../include/pairhmm_cuda.h"
#include "pmm_kernels.cuh"
#include "pmm_plan.h"
#include "pmm_tables.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

using namespace pmm;

namespace {

thread_local std::string g_create_error;      // last error of the calls that have no context (pmm_create, pmm_plan_flat), per thread

constexpr size_t kAlign = 256;
constexpr uint32_t kCtrlWords = kCtrlCursors + 32 * 104;
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = std::max(n, (size_t)4096);
        want += want / 4;                         // grow-only with slack: buffers are recycled across jobs like the
        cudaError_t e = cudaMalloc(&p, want);     // reference's scratch blocks (PairHMMTask.cpp:19-25,48,70)
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = std::max(n, (size_t)4096);
        want += want / 4;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

// What one launch produces, twice per context: launch i + 1's float pass runs while launch i's double re-run is still on
// the GPU (on its own stream), so the two must not share results, fallback lists, control words or the pinned block the
// results are copied to.  A fetch always refers to the set of the last launch.
struct ResultSet {
    DevBuf d_raw, d_fb_tasks, d_fb_idx, d_fb_hap, d_fb_rows, d_dres, d_ctrl;
    PinBuf h_out;                       // control words | raw floats | fallback indices | doubles
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};    // launch starts, float pass over, double re-run over
    cudaEvent_t ev_raw = nullptr, ev_lists = nullptr;   // the copies out of this set are over (its next user waits for them)
};

struct pmm_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t f64_stream = nullptr;  // fallback builders + double re-run: next to the following launch's float pass
    cudaStream_t copy_stream = nullptr; // D2H of the raw floats as soon as the float pass is over, under the double pass
    cudaStream_t list_stream = nullptr; // D2H of the control words and the fallback list behind the double pass
    ResultSet rs[2];
    int cur = 0;                        // set of the last launch
    ResultSet& R() { return rs[cur]; }
    const ResultSet& R() const { return rs[cur]; }
    cudaEvent_t ev_h2d = nullptr;       // recorded after the input arena's H2D copy; the next stage waits before repacking
    bool h2d_pending = false;
    cudaEvent_t ev_ref = nullptr;       // recorded and waited for in pmm_create: origin of the context's device timeline ...
    double ref_host_s = 0;              // ... and the host's steady clock at that moment (pmm_get_timeline)
    cudaEvent_t ev_probe = nullptr;     // issue-rate probes
    cudaEvent_t ev_block = nullptr;     // "sync" = "block": waits sleep on this event instead of spinning on the stream
    cudaEvent_t ev_poll = nullptr;      // "sync" = "auto": spun on or polled
    bool block_sync = false;
    int spin_us = 0;                    // "sync" = "hybrid": poll this long, then sleep on the blocking event
    bool auto_sync = false;             // "sync" = "auto": spin while few threads of the process do, else sleep
    std::string err;
    int tasks_per_warp = 16;
    bool fast = false;                  // "mode" option: fast = contracted float kernels + exact re-check near the threshold
    float guard = 0.0078125f;           // "guard" option: relative half-width of the re-check band around 1e-28f (2^-7)
    int f64_rows = kF64K;               // rows per lane of the double kernel for the staged job (pick_f64_rows)
    bool f64_striped = true;            // ... and whether its longest read needs more than one stripe of 32 x rows
    Variant force{0, 0, false};         // "force_variant" option (tuning sweeps): K,W of the float kernel
    int f64_tasks_per_warp = 6;         // "f64_tasks_per_warp": tasks the double re-run aims at per resident warp ...
    int f64_max_run = 6;                // "f64_max_run": ... and the most haplotypes it puts into one task (tools/f64_sweep.py)
    bool overlap = true;                // "overlap": consecutive launches overlap by one pass (off: a launch ends with its own re-run)

    // device-resident tables
    DevBuf tables;
    DeviceTables dtab{};

    // job state
    PinBuf h_in;  DevBuf d_in;          // one arena: read blob | descs | hap blob | descs | spos | tasks | regions
    DevBuf d_params, d_stream, d_iyf, d_iyd, d_tiny_tasks, d_scratch, d_scratch64, d_probe;
    size_t off_rblob = 0, off_rdesc = 0, off_hblob = 0, off_hdesc = 0, off_spos = 0, off_tasks = 0, off_regions = 0, off_groups = 0;
    uint32_t num_groups = 0;
    uint32_t num_read = 0, num_hap = 0, num_region = 0, num_tasks = 0, num_rows = 0;
    uint64_t pairs = 0, cells = 0;
    uint32_t max_hap_len = 0;
    std::vector<LaunchSeg> segs;
    std::vector<pmm_region_t> regions;
    bool staged = false, launched = false;
    bool have_raw = false, have_lists = false;   // what of the last launch's results is already in h_out
    uint32_t spec = 0;                           // fallback entries copied back speculatively by pmm_launch
    pmm_stats_t stats{};

    // scratch reused by the one-shot calls
    std::vector<uint32_t> tmp_roff, tmp_hoff;
    std::vector<ReadDesc> tmp_rdesc;
    std::vector<HapDesc> tmp_hdesc;
    std::vector<float> tmp_raw;

    int fail_cuda(cudaError_t e, const char* what)
    {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return PMM_ERR_CUDA;
    }
    int fail(int code, const std::string& m) { err = m; return code; }
};

#define PMM_CUDA(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return (ctx)->fail_cuda(e__, #call); } while (0)

namespace {

int ensure_tables(pmm_ctx* c)
{
    if (c->tables.p) return PMM_OK;
    const HostTables& t = host_tables();
    const size_t nf = kPh2prSize + kM2mSize;
    const size_t bytes = align_up(nf * sizeof(float)) + nf * sizeof(double);
    PMM_CUDA(c, c->tables.reserve(bytes));
    char* base = static_cast<char*>(c->tables.p);
    float* f = reinterpret_cast<float*>(base);
    double* d = reinterpret_cast<double*>(base + align_up(nf * sizeof(float)));
    PMM_CUDA(c, cudaMemcpy(f, t.ph2pr_f, sizeof t.ph2pr_f, cudaMemcpyHostToDevice));
    PMM_CUDA(c, cudaMemcpy(f + kPh2prSize, t.m2m_f, sizeof t.m2m_f, cudaMemcpyHostToDevice));
    PMM_CUDA(c, cudaMemcpy(d, t.ph2pr_d, sizeof t.ph2pr_d, cudaMemcpyHostToDevice));
    PMM_CUDA(c, cudaMemcpy(d + kPh2prSize, t.m2m_d, sizeof t.m2m_d, cudaMemcpyHostToDevice));
    c->dtab = DeviceTables{f, f + kPh2prSize, d, d + kPh2prSize};
    return PMM_OK;
}

// Where the bytes of a job come from.  The device-side blobs are the concatenation of `parts`; the descriptors give
// every read's five tracks (off, stride, len) and every haplotype (off, len) inside those blobs.  Two producers:
// five flat track arrays (stride = total bases) and the reference's wire format taken as it is (stride = len,
// interface/PairHMMHostInterface.cpp:175-207) -- the latter makes staging a serialized block one memcpy.
struct BlobPart { const uint8_t* p; size_t n; };
struct JobSource {
    std::vector<BlobPart> read_parts, hap_parts;
    const ReadDesc* rdesc = nullptr;
    const HapDesc* hdesc = nullptr;
};

// Common staging: reads are given as (offset, length) into five tracks, haplotypes likewise into one.
int stage_common(pmm_ctx* c, uint32_t num_read, const uint32_t* read_off, uint32_t num_hap, const uint32_t* hap_off,
                 const JobSource& src, uint32_t num_region, const pmm_region_t* regions)
{
    auto t0 = std::chrono::steady_clock::now();
    c->staged = false; c->launched = false;
    // the pinned input arena is about to be repacked: the previous job's H2D copy out of it must be over
    // (stage, launch, stage with no fetch in between)
    if (c->h2d_pending) { PMM_CUDA(c, cudaEventSynchronize(c->ev_h2d)); c->h2d_pending = false; }
    // ... and the device-side inputs are about to be overwritten: the double re-runs of earlier launches read them on
    // their own stream
    for (ResultSet& r : c->rs) PMM_CUDA(c, cudaStreamWaitEvent(c->stream, r.ev[2], 0));
    int rc = ensure_tables(c);
    if (rc) return rc;

    size_t sz_rblob = 0, sz_hblob = 0;
    for (const BlobPart& bp : src.read_parts) sz_rblob += bp.n;
    for (const BlobPart& bp : src.hap_parts) sz_hblob += bp.n;
    const size_t sz_rdesc = sizeof(ReadDesc) * num_read, sz_hdesc = sizeof(HapDesc) * num_hap;
    const size_t sz_spos = sizeof(uint32_t) * (num_hap + 1), sz_regions = sizeof(RegionDesc) * num_region;
    size_t sz_tasks = 0, sz_groups = 0, arena = 0;
    cudaError_t arena_err = cudaSuccess;

    // The planner calls back once it knows how many tasks and groups there are; the input arena is laid out and
    // reserved then, and the tasks are written straight into pinned memory (no intermediate copy).
    Plan plan;
    {
        std::string perr;
        rc = plan_job(num_read, read_off, num_hap, hap_off, num_region, regions, c->sm_count, c->tasks_per_warp, plan, perr,
                      c->force.K ? &c->force : nullptr, [&](const Plan& p) -> Task* {
            sz_tasks = sizeof(Task) * p.num_tasks; sz_groups = sizeof(GroupDesc) * p.groups.size();
            size_t off = 0;
            c->off_rblob = off; off = align_up(off + sz_rblob);
            c->off_rdesc = off; off = align_up(off + sz_rdesc);
            c->off_hblob = off; off = align_up(off + sz_hblob);
            c->off_hdesc = off; off = align_up(off + sz_hdesc);
            c->off_spos = off; off = align_up(off + sz_spos);
            c->off_tasks = off; off = align_up(off + sz_tasks);
            c->off_regions = off; off = align_up(off + sz_regions);
            c->off_groups = off; off = align_up(off + sz_groups);
            arena = off;
            if ((arena_err = c->h_in.reserve(arena)) != cudaSuccess) return nullptr;
            if ((arena_err = c->d_in.reserve(arena)) != cudaSuccess) return nullptr;
            return reinterpret_cast<Task*>(static_cast<char*>(c->h_in.p) + c->off_tasks);
        });
        if (arena_err != cudaSuccess) return c->fail_cuda(arena_err, "input arena");
        if (rc) return c->fail(rc, perr);
    }
    static const bool trace = getenv("PMM_TRACE_STAGE") != nullptr;       // phase times of staging on stderr
    const auto t_plan = std::chrono::steady_clock::now();
    const uint32_t max_hap = plan.max_hap_len;
    const uint64_t pairs = plan.pairs, cells = plan.cells;
    const std::vector<RegionDesc>& rdesc = plan.regions;
    c->max_hap_len = max_hap;
    c->f64_rows = pick_f64_rows(plan.max_read_len);
    c->f64_striped = plan.max_read_len + 1 > 32u * (uint32_t)c->f64_rows;
    c->segs = plan.segs;

    // ---- pack the rest of the input arena ---------------------------------------------------------------------
    char* hb = static_cast<char*>(c->h_in.p);
    {
        char* w = hb + c->off_rblob;
        for (const BlobPart& bp : src.read_parts) { memcpy(w, bp.p, bp.n); w += bp.n; }
        w = hb + c->off_hblob;
        for (const BlobPart& bp : src.hap_parts) { memcpy(w, bp.p, bp.n); w += bp.n; }
    }
    memcpy(hb + c->off_rdesc, src.rdesc, sz_rdesc);
    memcpy(hb + c->off_hdesc, src.hdesc, sz_hdesc);
    uint32_t* spos = reinterpret_cast<uint32_t*>(hb + c->off_spos);
    uint32_t pos = 0;
    for (uint32_t h = 0; h < num_hap; ++h) { spos[h] = pos; pos += src.hdesc[h].len + 1; }
    spos[num_hap] = pos;                           // final separator
    memcpy(hb + c->off_regions, rdesc.data(), sz_regions);
    memcpy(hb + c->off_groups, plan.groups.data(), sz_groups);
    c->num_groups = (uint32_t)plan.groups.size();

    const auto t_pack = std::chrono::steady_clock::now();
    // ---- device buffers -------------------------------------------------------------------------------------
    const size_t stream_bytes = kStreamFrontPad + (size_t)pos + 1 + kStreamTailPad;
    PMM_CUDA(c, c->d_stream.reserve(stream_bytes));
    PMM_CUDA(c, c->d_params.reserve(sizeof(float) * plan.param_floats));
    PMM_CUDA(c, c->d_iyf.reserve(sizeof(float) * num_hap));
    PMM_CUDA(c, c->d_iyd.reserve(sizeof(double) * num_hap));
    if (c->fast) PMM_CUDA(c, c->d_tiny_tasks.reserve(sizeof(Task) * pairs));     // re-check list of the guard band
    for (ResultSet& r : c->rs) {
        PMM_CUDA(c, r.d_raw.reserve(sizeof(float) * pairs));
        PMM_CUDA(c, r.d_fb_tasks.reserve(sizeof(Task) * pairs));
        PMM_CUDA(c, r.d_fb_idx.reserve(sizeof(uint32_t) * pairs));
        PMM_CUDA(c, r.d_fb_hap.reserve(sizeof(uint32_t) * pairs));
        PMM_CUDA(c, r.d_fb_rows.reserve(sizeof(uint32_t) * 2 * plan.rows));
        PMM_CUDA(c, r.d_dres.reserve(sizeof(double) * pairs));
        PMM_CUDA(c, r.d_ctrl.reserve(sizeof(uint32_t) * kCtrlWords));
        PMM_CUDA(c, r.h_out.reserve(256 + align_up(sizeof(float) * pairs) + align_up(sizeof(uint32_t) * pairs) + sizeof(double) * pairs + 256));
    }
    // carry rows of the striped kernels: one haplotype (+2 separators) per warp, three rows of doubles
    {
        int ctas64 = 1;
        for (int k : {4, 5, 6, 8}) ctas64 = std::max(ctas64, forward_f64_ctas_per_sm(k, true));
        const int ctas32 = std::max(std::max(forward_f32_ctas_per_sm(kStripedK, 32, true, false), forward_f32_ctas_per_sm(kStripedK, 32, true, true)),
                                    recheck_f32_ctas_per_sm());
        // the float and the double kernel of neighbouring launches can be on the GPU together: a scratch of their own each
        PMM_CUDA(c, c->d_scratch.reserve((size_t)c->sm_count * ctas32 * kWarpsPerCta * 3 * (size_t)(max_hap + 8) * sizeof(float)));
        PMM_CUDA(c, c->d_scratch64.reserve((size_t)c->sm_count * ctas64 * kWarpsPerCta * 3 * (size_t)(max_hap + 8) * sizeof(double)));
    }

    cudaStream_t s = c->stream;
    PMM_CUDA(c, cudaMemcpyAsync(c->d_in.p, c->h_in.p, arena, cudaMemcpyHostToDevice, s));
    PMM_CUDA(c, cudaEventRecord(c->ev_h2d, s));
    c->h2d_pending = true;
    PMM_CUDA(c, cudaMemsetAsync(c->d_stream.p, 0, kStreamFrontPad, s));
    PMM_CUDA(c, cudaMemsetAsync(static_cast<char*>(c->d_stream.p) + kStreamFrontPad + pos + 1, 0, kStreamTailPad, s));
    char* db = static_cast<char*>(c->d_in.p);
    const HostTables& ht = host_tables();
    PMM_CUDA(c, launch_build_stream(reinterpret_cast<uint8_t*>(db + c->off_hblob), reinterpret_cast<HapDesc*>(db + c->off_hdesc),
                                    reinterpret_cast<uint32_t*>(db + c->off_spos), num_hap,
                                    static_cast<uint8_t*>(c->d_stream.p) + kStreamFrontPad,
                                    static_cast<float*>(c->d_iyf.p), static_cast<double*>(c->d_iyd.p), ht.ic_f, ht.ic_d, s));

    c->num_read = num_read; c->num_hap = num_hap; c->num_region = num_region; c->num_tasks = (uint32_t)plan.num_tasks;
    c->num_rows = (uint32_t)plan.rows;
    c->pairs = pairs; c->cells = cells;
    c->regions.assign(regions, regions + num_region);
    c->stats = pmm_stats_t{};
    c->stats.pairs = pairs; c->stats.cells = cells; c->stats.h2d_bytes = arena; c->stats.f32_tasks = c->num_tasks;
    c->stats.ms_stage = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (trace) {
        auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        fprintf(stderr, "[stage] %u reads x %u haps, %zu tasks: plan %.0f us, pack %.0f us, reserve + enqueue %.0f us\n", num_read, num_hap,
                (size_t)plan.num_tasks, us(t0, t_plan), us(t_plan, t_pack), us(t_pack, std::chrono::steady_clock::now()));
    }
    c->staged = true;
    return PMM_OK;
}

// Walk the reference's wire format (PairHMMHostInterface.cpp:175-207) and return offsets into the blob.
bool scan_reads(const uint8_t* p, uint64_t n, std::vector<uint32_t>& off, std::vector<uint32_t>& len)
{
    if (n < 4) return false;
    int32_t num; memcpy(&num, p, 4);
    if (num < 0) return false;
    uint64_t pos = 4;
    if ((uint64_t)num > (n - 4) / 4) return false;          // every read takes at least its length word
    off.resize(num); len.resize(num);
    for (int32_t i = 0; i < num; ++i) {
        if (pos + 4 > n) return false;
        int32_t l; memcpy(&l, p + pos, 4); pos += 4;
        if (l < 0 || pos + 5ull * l > n) return false;
        off[i] = (uint32_t)pos; len[i] = (uint32_t)l; pos += 5ull * l;
    }
    return true;
}

bool scan_haps(const uint8_t* p, uint64_t n, std::vector<uint32_t>& off, std::vector<uint32_t>& len)
{
    if (n < 4) return false;
    int32_t num; memcpy(&num, p, 4);
    if (num < 0) return false;
    uint64_t pos = 4;
    if ((uint64_t)num > (n - 4) / 4) return false;
    off.resize(num); len.resize(num);
    for (int32_t i = 0; i < num; ++i) {
        if (pos + 4 > n) return false;
        int32_t l; memcpy(&l, p + pos, 4); pos += 4;
        if (l < 0 || pos + (uint64_t)l > n) return false;
        off[i] = (uint32_t)pos; len[i] = (uint32_t)l; pos += l;
    }
    return true;
}

// The wire format as the job source: lengths go to tmp_roff / tmp_hoff for the planner, the blobs are the serialized
// buffers themselves.
int source_from_serialized(pmm_ctx* c, const void* reads_ser, uint64_t reads_bytes, const void* haps_ser, uint64_t haps_bytes,
                           JobSource& src)
{
    std::vector<uint32_t> roff, rlen, hoff, hlen;
    const uint8_t* rp = static_cast<const uint8_t*>(reads_ser);
    const uint8_t* hp = static_cast<const uint8_t*>(haps_ser);
    if (!rp || !hp || !scan_reads(rp, reads_bytes, roff, rlen) || !scan_haps(hp, haps_bytes, hoff, hlen))
        return c->fail(PMM_ERR_INVALID, "malformed serialized block");
    if (reads_bytes >= (1ull << 31) || haps_bytes >= (1ull << 31)) return c->fail(PMM_ERR_INVALID, "serialized block larger than 2 GiB");
    const size_t nr = roff.size(), nh = hoff.size();
    c->tmp_roff.assign(nr + 1, 0); c->tmp_hoff.assign(nh + 1, 0);
    c->tmp_rdesc.resize(nr); c->tmp_hdesc.resize(nh);
    for (size_t i = 0; i < nr; ++i) {
        c->tmp_roff[i + 1] = c->tmp_roff[i] + rlen[i];
        c->tmp_rdesc[i] = ReadDesc{roff[i], rlen[i], rlen[i]};
    }
    for (size_t i = 0; i < nh; ++i) {
        c->tmp_hoff[i + 1] = c->tmp_hoff[i] + hlen[i];
        c->tmp_hdesc[i] = HapDesc{hoff[i], hlen[i]};
    }
    src.read_parts = {BlobPart{rp, (size_t)reads_bytes}};
    src.hap_parts = {BlobPart{hp, (size_t)haps_bytes}};
    src.rdesc = c->tmp_rdesc.data(); src.hdesc = c->tmp_hdesc.data();
    return PMM_OK;
}

// Five flat tracks + one haplotype array, (offset, length) per element.
void source_from_flat(pmm_ctx* c, uint32_t num_read, const uint32_t* read_off, const uint8_t* const tracks[5],
                      uint32_t num_hap, const uint32_t* hap_off, const uint8_t* hap_bases, JobSource& src)
{
    const uint32_t r_base = read_off[0], h_base = hap_off[0];
    const size_t total_bases = read_off[num_read] - r_base, total_hap = hap_off[num_hap] - h_base;
    c->tmp_rdesc.resize(num_read); c->tmp_hdesc.resize(num_hap);
    for (uint32_t i = 0; i < num_read; ++i)
        c->tmp_rdesc[i] = ReadDesc{read_off[i] - r_base, (uint32_t)total_bases, read_off[i + 1] - read_off[i]};
    for (uint32_t h = 0; h < num_hap; ++h) c->tmp_hdesc[h] = HapDesc{hap_off[h] - h_base, hap_off[h + 1] - hap_off[h]};
    src.read_parts.clear();
    for (int t = 0; t < 5; ++t) src.read_parts.push_back(BlobPart{tracks[t] + r_base, total_bases});
    src.hap_parts = {BlobPart{hap_bases + h_base, total_hap}};
    src.rdesc = c->tmp_rdesc.data(); src.hdesc = c->tmp_hdesc.data();
}

int stage_single_region(pmm_ctx* c, const JobSource& src)
{
    const uint32_t nr = (uint32_t)c->tmp_roff.size() - 1, nh = (uint32_t)c->tmp_hoff.size() - 1;
    if (nr == 0 || nh == 0) return c->fail(PMM_ERR_INVALID, "empty batch");
    pmm_region_t reg{0, nr, 0, nh};
    return stage_common(c, nr, c->tmp_roff.data(), nh, c->tmp_hoff.data(), src, 1, &reg);
}

// Host worker threads for the per-result log10 (the one piece of the path that stays on the host so that it uses the
// host libm like the reference).  Created once per process on first use; a fetch hands out chunks and takes part
// itself, so small jobs never wait for a wake-up.  A call that finds the workers taken runs on its own thread.
class HostWorkers {
 public:
    static HostWorkers& get() { static HostWorkers w; return w; }
    void run(uint64_t n, uint64_t grain, const std::function<void(uint64_t, uint64_t)>& f)
    {
        if (n <= grain || th_.empty()) { f(0, n); return; }
        // Workers busy with another context's results: the process is streaming jobs through several contexts, and
        // the caller's own thread is the parallelism (24 feeder threads on an 8-GPU box) -- do not queue up behind them.
        std::unique_lock<std::mutex> call(call_mu_, std::try_to_lock);
        if (!call.owns_lock()) { f(0, n); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &f; n_ = n; grain_ = grain; next_.store(0); busy_ = (int)th_.size(); ++gen_;
        }
        cv_work_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return busy_ == 0; });
        fn_ = nullptr;
    }

 private:
    HostWorkers()
    {
        const unsigned hw = std::thread::hardware_concurrency();
        const unsigned n = hw > 1 ? std::min(hw - 1, 7u) : 0;
        for (unsigned k = 0; k < n; ++k) th_.emplace_back([this] { loop(); });
    }
    ~HostWorkers()
    {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_work_.notify_all();
        for (auto& t : th_) t.join();
    }
    void work()
    {
        for (;;) {
            const uint64_t a = next_.fetch_add(grain_);
            if (a >= n_) return;
            (*fn_)(a, std::min(n_, a + grain_));
        }
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work();
            { std::lock_guard<std::mutex> lk(mu_); --busy_; }
            cv_done_.notify_one();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(uint64_t, uint64_t)>* fn_ = nullptr;
    uint64_t n_ = 0, grain_ = 1, gen_ = 0;
    std::atomic<uint64_t> next_{0};
    int busy_ = 0;
    bool stop_ = false;
};

void parallel_for(uint64_t n, uint64_t grain, const std::function<void(uint64_t, uint64_t)>& f)
{
    HostWorkers::get().run(n, std::max<uint64_t>(grain, 1), f);
}
}  // namespace

// the same workers for the other engine of this library (sw_engine.cu: staging and CIGAR scatter)
namespace pmm { void host_parallel_for(uint64_t n, uint64_t grain, const std::function<void(uint64_t, uint64_t)>& f) { parallel_for(n, grain, f); } }

// The double re-run's stream, at the priority of the float pass's.  (Tried: one level above it, so that a re-run's blocks
// take freed slots before the next float pass's -- no difference in throughput on configs 2, 3 and 4 or through the pool,
// an FP32-pipe and an FP64-pipe kernel do not co-run to any advantage: the float pass leaves too few issue slots.)
static cudaError_t create_f64_stream(cudaStream_t* out, int prio)
{
    return cudaStreamCreateWithPriority(out, cudaStreamNonBlocking, prio);
}

// No exception crosses the C boundary (include/pairhmm_cuda.h): allocation failures of the host-side vectors and
// anything else thrown below an entry point become a status code and a message.
template <class F> static int guarded(pmm_ctx* c, F&& f)
{
    try { return f(); }
    catch (const std::bad_alloc&) {
        if (c) return c->fail(PMM_ERR_INVALID, "out of host memory");
        g_create_error = "out of host memory"; return PMM_ERR_INVALID;
    }
    catch (const std::exception& e) {
        if (c) return c->fail(PMM_ERR_INVALID, e.what());
        g_create_error = e.what(); return PMM_ERR_INVALID;
    }
}

// =========================================================================================================
extern "C" {

int pmm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pmm_create(int device, pmm_ctx** out)
{
    if (!out) return PMM_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        g_create_error = "no CUDA device visible; this engine has no CPU fallback";
        return PMM_ERR_NO_DEVICE;
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= n) { g_create_error = "device index out of range"; return PMM_ERR_INVALID; }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e); return PMM_ERR_CUDA;
    }
    if (prop.major != 10) {
        g_create_error = std::string("kernels are built for sm_100a only; device is ") + prop.name;
        return PMM_ERR_NO_DEVICE;
    }
    pmm_ctx* c = new pmm_ctx();
    c->device = device; c->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e); delete c; return PMM_ERR_CUDA;
    }
    c->stream = c->own_stream;
    if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->list_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = create_f64_stream(&c->f64_stream, 0)) != cudaSuccess) {       // 0 = the default priority, like own_stream
        g_create_error = cudaGetErrorString(e);
        if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
        if (c->list_stream) cudaStreamDestroy(c->list_stream);
        cudaStreamDestroy(c->own_stream); delete c; return PMM_ERR_CUDA;
    }
    for (ResultSet& r : c->rs) {
        for (auto& ev : r.ev) cudaEventCreate(&ev);
        cudaEventCreateWithFlags(&r.ev_raw, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&r.ev_lists, cudaEventDisableTiming);
    }
    cudaEventCreate(&c->ev_probe);
    cudaEventCreateWithFlags(&c->ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_poll, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_h2d, cudaEventDisableTiming);
    cudaEventCreate(&c->ev_ref);
    cudaEventRecord(c->ev_ref, c->own_stream);
    cudaEventSynchronize(c->ev_ref);
    c->ref_host_s = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    *out = c;
    return PMM_OK;
}

void pmm_destroy(pmm_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->f64_stream) cudaStreamSynchronize(c->f64_stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->list_stream) cudaStreamSynchronize(c->list_stream);
    for (DevBuf* b : {&c->tables, &c->d_in, &c->d_params, &c->d_stream, &c->d_iyf, &c->d_iyd, &c->d_tiny_tasks, &c->d_scratch,
                      &c->d_scratch64, &c->d_probe}) b->release();
    for (ResultSet& r : c->rs) {
        for (DevBuf* b : {&r.d_raw, &r.d_fb_tasks, &r.d_fb_idx, &r.d_fb_hap, &r.d_fb_rows, &r.d_dres, &r.d_ctrl}) b->release();
        r.h_out.release();
        for (auto& ev : r.ev) if (ev) cudaEventDestroy(ev);
        if (r.ev_raw) cudaEventDestroy(r.ev_raw);
        if (r.ev_lists) cudaEventDestroy(r.ev_lists);
    }
    c->h_in.release();
    if (c->ev_probe) cudaEventDestroy(c->ev_probe);
    if (c->f64_stream) cudaStreamDestroy(c->f64_stream);
    if (c->ev_block) cudaEventDestroy(c->ev_block);
    if (c->ev_poll) cudaEventDestroy(c->ev_poll);
    if (c->list_stream) cudaStreamDestroy(c->list_stream);
    if (c->ev_h2d) cudaEventDestroy(c->ev_h2d);
    if (c->ev_ref) cudaEventDestroy(c->ev_ref);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* pmm_last_error(const pmm_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int pmm_set_option(pmm_ctx* c, const char* key, const char* value)
{
    if (!c || !key || !value) return PMM_ERR_INVALID;
    const std::string k(key);
    if (k == "stream") {
        const std::string sv(value);
        if (sv == "own") c->stream = c->own_stream;
        else if (sv == "default") c->stream = nullptr;                       // the legacy default stream
        else c->stream = reinterpret_cast<cudaStream_t>(strtoull(value, nullptr, 0));
        return PMM_OK;
    }
    if (k == "priority") {
        // rank of this context's kernels among the contexts sharing the GPU: 0 = first in line (highest CUDA stream
        // priority), larger = later.  The kernels are persistent (one launch fills the GPU), so the order only decides
        // whose blocks take the slots another kernel's tail frees -- e.g. the double re-run of tile k before the float
        // pass of tile k+2 when a batch is cut into tiles that run on several contexts (pairhmm/client/PairHMMWorker).
        const int rank = atoi(value);
        if (rank < 0 || rank > 64) return c->fail(PMM_ERR_INVALID, "priority rank out of range");
        cudaSetDevice(c->device);
        int least = 0, greatest = 0;
        PMM_CUDA(c, cudaDeviceGetStreamPriorityRange(&least, &greatest));        // numerically: greatest <= least
        const int prio = std::min(least, greatest + rank);
        cudaStream_t ns = nullptr, nd = nullptr;
        PMM_CUDA(c, cudaStreamCreateWithPriority(&ns, cudaStreamNonBlocking, prio));
        PMM_CUDA(c, create_f64_stream(&nd, prio));
        PMM_CUDA(c, cudaStreamSynchronize(c->own_stream));
        PMM_CUDA(c, cudaStreamSynchronize(c->f64_stream));
        const bool was_own = c->stream == c->own_stream;
        cudaStreamDestroy(c->own_stream); cudaStreamDestroy(c->f64_stream);
        c->own_stream = ns; c->f64_stream = nd;
        if (was_own) c->stream = ns;
        return PMM_OK;
    }
    if (k == "tasks_per_warp") {
        const int v = atoi(value);
        if (v < 1 || v > 64) return c->fail(PMM_ERR_INVALID, "tasks_per_warp out of range");
        c->tasks_per_warp = v;
        return PMM_OK;
    }
    if (k == "small_job_widening") {
        const std::string v(value);
        if (v != "0" && v != "1" && v != "on" && v != "off") return c->fail(PMM_ERR_INVALID, "small_job_widening is \"on\" or \"off\"");
        set_small_job_widening(v == "1" || v == "on");   // process-wide (the planner has no context)
        return PMM_OK;
    }
    if (k == "run_tiers") {
        int d = -1, share = kRunTierShare, top = kRunTierTop;
        if (std::sscanf(value, "%d,%d,%d", &d, &share, &top) < 1 || d < 0 || d > 64 || share < 1 || share > 100 || top < 1 || top > 64)
            return c->fail(PMM_ERR_INVALID, "run_tiers is \"depth[,share[,top]]\", depth 0..64, share 1..100, top 1..64");
        set_run_tiers(d, share, top);                // process-wide (the planner has no context)
        return PMM_OK;
    }
    if (k == "f64_tasks_per_warp" || k == "f64_max_run") {
        const int v = atoi(value);
        if (v < 1 || v > 64) return c->fail(PMM_ERR_INVALID, k + " out of range (1..64)");
        (k == "f64_max_run" ? c->f64_max_run : c->f64_tasks_per_warp) = v;
        return PMM_OK;
    }
    if (k == "overlap") {
        const std::string v(value);
        if (v != "0" && v != "1" && v != "on" && v != "off") return c->fail(PMM_ERR_INVALID, "overlap is \"on\" or \"off\"");
        c->overlap = v == "1" || v == "on";
        return PMM_OK;
    }
    if (k == "mode") {
        const std::string v(value);
        if (v != "exact" && v != "fast") return c->fail(PMM_ERR_INVALID, "mode is \"exact\" or \"fast\"");
        c->fast = v == "fast";
        return PMM_OK;
    }
    if (k == "guard") {
        const float g = (float)atof(value);
        if (!(g >= 0.0f && g < 1.0f)) return c->fail(PMM_ERR_INVALID, "guard must be in [0, 1)");
        c->guard = g;
        return PMM_OK;
    }
    if (k == "sync") {
        const std::string v = value;
        if (v != "spin" && v != "block" && v != "hybrid" && v != "auto") return c->fail(PMM_ERR_INVALID, "sync is \"spin\", \"block\", \"hybrid\" or \"auto\"");
        c->block_sync = v == "block" || v == "hybrid";
        c->spin_us = v == "hybrid" ? 60 : 0;
        c->auto_sync = v == "auto";
        return PMM_OK;
    }
    if (k == "force_variant") {
        int K = 0, W = 0;
        if (sscanf(value, "%d,%d", &K, &W) != 2 || (K > 0 && !forward_f32_has_variant(K, W)))
            return c->fail(PMM_ERR_INVALID, "force_variant wants \"K,W\" of an instantiated float kernel, \"0,0\" (planner's choice) or \"-1,0\" (no consolidation of rare variants)");
        c->force = Variant{K, W, false};
        return PMM_OK;
    }
    return c->fail(PMM_ERR_INVALID, "unknown option " + k);
}

static int pmm_stage_flat_impl(pmm_ctx* c, uint32_t num_read, const uint32_t* read_off,
                   const uint8_t* bases, const uint8_t* q, const uint8_t* i, const uint8_t* d, const uint8_t* cc,
                   uint32_t num_hap, const uint32_t* hap_off, const uint8_t* hap_bases,
                   uint32_t num_region, const pmm_region_t* regions)
{
    if (!c) return PMM_ERR_INVALID;
    if (!bases || !q || !i || !d || !cc || !hap_bases) return c->fail(PMM_ERR_INVALID, "null track");
    cudaSetDevice(c->device);
    if (!read_off || !hap_off || !num_read || !num_hap) return c->fail(PMM_ERR_INVALID, "empty job");
    if (read_off[num_read] < read_off[0] || hap_off[num_hap] < hap_off[0]) return c->fail(PMM_ERR_INVALID, "offsets not ascending");
    for (uint32_t k = 0; k < num_read; ++k) if (read_off[k + 1] < read_off[k]) return c->fail(PMM_ERR_INVALID, "offsets not ascending");
    for (uint32_t k = 0; k < num_hap; ++k) if (hap_off[k + 1] < hap_off[k]) return c->fail(PMM_ERR_INVALID, "offsets not ascending");
    const uint8_t* const tracks[5] = {bases, q, i, d, cc};
    JobSource src;
    source_from_flat(c, num_read, read_off, tracks, num_hap, hap_off, hap_bases, src);
    return stage_common(c, num_read, read_off, num_hap, hap_off, src, num_region, regions);
}

int pmm_launch(pmm_ctx* c)
{
    if (!c) return PMM_ERR_INVALID;
    if (!c->staged) return c->fail(PMM_ERR_STATE, "pmm_launch before pmm_stage_*");
    if (c->fast && c->d_tiny_tasks.cap < sizeof(Task) * c->pairs) return c->fail(PMM_ERR_STATE, "mode changed to fast after pmm_stage_*: stage the job again");
    cudaSetDevice(c->device);
    cudaStream_t s = c->stream, sd = c->f64_stream;
    char* db = static_cast<char*>(c->d_in.p);
    // This launch takes the other result set: the previous launch's double re-run may still be using its own.
    ResultSet& prev = c->rs[c->cur];
    c->cur ^= 1;
    ResultSet& R = c->rs[c->cur];
    // control words, hot ones on their own 128-byte lines: [0] fallback count, [1] flush count, [2] re-check count,
    // [3] tasks of the double re-run, [4] fallback slots handed out, [kCtrlHist ..) and [kCtrlClassCursor ..) the counting
    // sort of those tasks, [kCtrlCursors + 32 k] work-queue cursor of launch k
    uint32_t* ctrl = static_cast<uint32_t*>(R.d_ctrl.p);
    uint32_t launches = 0;
    c->launched = false; c->have_raw = false; c->have_lists = false;
    // the launch before the previous one used this set: its copies out of d_raw, d_ctrl and the fallback list (which
    // follow its double re-run) must be over (no-ops the first times)
    PMM_CUDA(c, cudaStreamWaitEvent(s, R.ev_raw, 0));
    PMM_CUDA(c, cudaStreamWaitEvent(s, R.ev_lists, 0));
    PMM_CUDA(c, cudaEventRecord(R.ev[0], s));
    PMM_CUDA(c, cudaMemsetAsync(ctrl, 0, sizeof(uint32_t) * kCtrlWords, s));

    ForwardArgs a{};
    a.read_blob = reinterpret_cast<uint8_t*>(db + c->off_rblob);
    a.reads = reinterpret_cast<ReadDesc*>(db + c->off_rdesc);
    a.stream = static_cast<uint8_t*>(c->d_stream.p) + kStreamFrontPad;
    a.spos = reinterpret_cast<uint32_t*>(db + c->off_spos);
    a.tab = c->dtab;
    a.scratch = c->d_scratch.p;
    a.scratch_stride = c->max_hap_len + 8;

    // ---- row parameters, once per read (the reference recomputes them per pair) --------------------------------
    a.params = static_cast<float*>(c->d_params.p);
    PMM_CUDA(c, launch_read_params(a.read_blob, a.reads, reinterpret_cast<GroupDesc*>(db + c->off_groups), c->num_groups,
                                   c->dtab, static_cast<float*>(c->d_params.p), s));
    ++launches;

    // ---- float pass, one launch per (K, W) variant present in the job; results below 1e-28f are appended to the
    //      fallback list by the kernel itself (PairHMMWorker.cpp:176) ----------------------------------------------
    if (c->segs.size() > 100) return c->fail(PMM_ERR_INVALID, "too many kernel variants in one job");
    // fast mode: results within the guard band around the threshold are re-run by an exact kernel before the decision
    // is taken, so the decision (and the float value of those pairs) is the reference's, bit for bit
    const float thr = 1e-28f;
    FallbackQueue fq{ctrl + 0, (uint32_t)c->pairs,
                     c->fast ? thr * (1.0f - c->guard) : thr, c->fast ? thr * (1.0f + c->guard) : thr,
                     static_cast<Task*>(c->d_tiny_tasks.p), ctrl + 2};
    uint32_t cursor = kCtrlCursors;
    for (const LaunchSeg& seg : c->segs) {
        a.inity = c->d_iyf.p;
        a.tasks = reinterpret_cast<Task*>(db + c->off_tasks) + seg.task_first;
        a.ntasks = seg.task_count; a.ntasks_dev = nullptr;
        a.counter = ctrl + cursor; cursor += 32;
        a.out = R.d_raw.p;
        const int per_sm = forward_f32_ctas_per_sm(seg.v.K, seg.v.W, seg.v.striped, c->fast);
        if (per_sm <= 0) return c->fail(PMM_ERR_INVALID, "kernel variant unavailable");
        const int ctas = (int)std::min<uint64_t>((seg.task_count + kWarpsPerCta - 1) / kWarpsPerCta, (uint64_t)c->sm_count * per_sm);
        PMM_CUDA(c, launch_forward_f32(seg.v.K, seg.v.W, seg.v.striped, c->fast, a, fq, ctas, s));
        ++launches;
    }
    if (c->fast) {
        // ---- exact re-check of the guard band (normally a handful of pairs, often none) -----------------------------
        a.tasks = static_cast<Task*>(c->d_tiny_tasks.p); a.ntasks = 0; a.ntasks_dev = ctrl + 2;
        a.counter = ctrl + cursor; cursor += 32;
        fq.lo = fq.hi = thr;
        PMM_CUDA(c, launch_recheck_f32(a, fq, c->sm_count * std::max(1, recheck_f32_ctas_per_sm()), s));
        ++launches;
    }
    PMM_CUDA(c, cudaEventRecord(R.ev[1], s));
    // The raw floats are final now: copy them back on the copy stream while the double pass runs, so that a fetch can
    // take log10f of them (host libm) under it.
    {
        char* ho = static_cast<char*>(R.h_out.p);
        PMM_CUDA(c, cudaStreamWaitEvent(c->copy_stream, R.ev[1], 0));
        PMM_CUDA(c, cudaMemcpyAsync(ho + 256, R.d_raw.p, sizeof(float) * c->pairs, cudaMemcpyDeviceToHost, c->copy_stream));
        PMM_CUDA(c, cudaEventRecord(R.ev_raw, c->copy_stream));
    }

    // ---- double re-run (PairHMMWorker.cpp:176-184): the results below the threshold become tasks, the failing haplotypes
    //      of a read together; their number is only known on the device.  On its own stream: the next launch's float pass
    //      (this stream) starts as soon as this one's is over and runs next to it ------------------------------------------
    PMM_CUDA(c, cudaStreamWaitEvent(sd, R.ev[1], 0));
    const int KD = c->f64_rows;
    const int f64_ctas = c->sm_count * std::max(1, forward_f64_ctas_per_sm(KD, c->f64_striped));
    FallbackBuild fb{};
    fb.raw = static_cast<float*>(R.d_raw.p);
    fb.regions = reinterpret_cast<RegionDesc*>(db + c->off_regions);
    fb.reads = a.reads;
    fb.num_region = c->num_region; fb.num_rows = c->num_rows;
    fb.threshold = thr;
    fb.ctrl = ctrl;
    fb.tasks = static_cast<Task*>(R.d_fb_tasks.p);
    fb.out_index = static_cast<uint32_t*>(R.d_fb_idx.p);
    fb.hap_list = static_cast<uint32_t*>(R.d_fb_hap.p);
    fb.row_slot = static_cast<uint32_t*>(R.d_fb_rows.p);
    fb.spos = a.spos;
    fb.max_hap_len = c->max_hap_len;
    fb.capacity = (uint32_t)c->pairs;
    fb.single_stripe_rows = 32u * (uint32_t)KD;
    fb.target_tasks = (uint32_t)(f64_ctas * kWarpsPerCta * c->f64_tasks_per_warp);
    fb.max_run = (uint32_t)c->f64_max_run;
    PMM_CUDA(c, launch_build_fallback(fb, c->sm_count, sd));
    launches += 2;
    a.inity = c->d_iyd.p; a.out = R.d_dres.p; a.tasks = fb.tasks; a.hap_list = fb.hap_list;
    a.ntasks = 0; a.ntasks_dev = ctrl + 3; a.counter = ctrl + cursor; cursor += 32;
    a.tiny_threshold = ldexp(1.0, -800); a.tiny_count = ctrl + 1;
    a.scratch = c->d_scratch64.p;
    PMM_CUDA(c, launch_forward_f64(KD, c->f64_striped, a, f64_ctas, sd));
    ++launches;
    PMM_CUDA(c, cudaEventRecord(R.ev[2], sd));
    // What is queued on the launch stream from here on (the next launch, an event of the caller's) comes after the
    // PREVIOUS launch's double re-run: consecutive launches overlap by one pass, no further, and a pair of events around
    // each pmm_launch of a steady sequence brackets exactly one float pass and one double re-run.
    PMM_CUDA(c, cudaStreamWaitEvent(s, (c->overlap ? prev : R).ev[2], 0));
    // The length of the fallback list is only known on the device.  Copy the control words and a first slice of the list
    // (an eighth of the pairs) back right behind the double pass, so that the common case needs no further round trip;
    // the rest, if any, follows in the fetch.
    {
        char* ho = static_cast<char*>(R.h_out.p);
        uint32_t* hidx = reinterpret_cast<uint32_t*>(ho + 256 + align_up(sizeof(float) * c->pairs));
        double* hd = reinterpret_cast<double*>(reinterpret_cast<char*>(hidx) + align_up(sizeof(uint32_t) * c->pairs));
        c->spec = (uint32_t)std::min<uint64_t>(c->pairs, std::max<uint64_t>(1024, c->pairs / 8));
        cudaStream_t ls = c->list_stream;                       // not a kernels' stream: the next job need not wait for these
        PMM_CUDA(c, cudaStreamWaitEvent(ls, R.ev[2], 0));
        PMM_CUDA(c, cudaMemcpyAsync(ho, R.d_ctrl.p, 12, cudaMemcpyDeviceToHost, ls));
        PMM_CUDA(c, cudaMemcpyAsync(hidx, R.d_fb_idx.p, sizeof(uint32_t) * c->spec, cudaMemcpyDeviceToHost, ls));
        PMM_CUDA(c, cudaMemcpyAsync(hd, R.d_dres.p, sizeof(double) * c->spec, cudaMemcpyDeviceToHost, ls));
        PMM_CUDA(c, cudaEventRecord(R.ev_lists, ls));
    }
    c->stats.kernel_launches = launches;
    c->launched = true;
    return PMM_OK;
}

int pmm_join(pmm_ctx* c)
{
    if (!c) return PMM_ERR_INVALID;
    if (!c->launched) return PMM_OK;
    cudaSetDevice(c->device);
    PMM_CUDA(c, cudaStreamWaitEvent(c->stream, c->R().ev[2], 0));
    return PMM_OK;
}

int pmm_sync(pmm_ctx* c)
{
    if (!c) return PMM_ERR_INVALID;
    cudaSetDevice(c->device);
    PMM_CUDA(c, cudaStreamSynchronize(c->stream));
    PMM_CUDA(c, cudaStreamSynchronize(c->f64_stream));
    PMM_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    PMM_CUDA(c, cudaStreamSynchronize(c->list_stream));
    if (c->launched) {
        cudaEventElapsedTime(&c->stats.ms_f32, c->R().ev[0], c->R().ev[1]);
        cudaEventElapsedTime(&c->stats.ms_fallback, c->R().ev[1], c->R().ev[2]);
    }
    return PMM_OK;
}

// Wait for everything queued on one of the context's streams.  A caller with one context per core spins (lowest latency);
// the pool, which runs several contexts per GPU from as many host threads, can sleep on a blocking event so that the
// waiting threads do not take the cores the packing and log10 work of the other contexts needs.
// "auto": a spinning wait is the fastest (no wake-up), but every spinner holds a core.  The first few waiting threads
// of the process spin, the rest sleep -- a host with one or two caller threads gets the latency, a host with dozens (or a
// box with a process per GPU) keeps its cores for the packing and log10 work.
static std::atomic<int> g_spinners{0};
static int spinner_budget()
{
    static const int b = [] { const unsigned hw = std::thread::hardware_concurrency(); return (int)std::max(1u, hw / 16); }();
    return b;
}

static cudaError_t wait_stream(pmm_ctx* c, cudaStream_t st)
{
    if (c->auto_sync) {
        // Whoever finds a spinner's place free takes it (the earliest waiter, i.e. the job that finishes first); the others
        // nap and look again -- for the event and for a free place -- so a waiter is promoted as soon as the one ahead of it
        // is done, and a sleeping thread never costs more than a wake-up of a few tens of microseconds.
        cudaError_t e = cudaEventRecord(c->ev_poll, st);
        if (e != cudaSuccess) return e;
        for (;;) {
            if (g_spinners.fetch_add(1, std::memory_order_relaxed) < spinner_budget()) {
                e = cudaEventSynchronize(c->ev_poll);
                g_spinners.fetch_sub(1, std::memory_order_relaxed);
                return e;
            }
            g_spinners.fetch_sub(1, std::memory_order_relaxed);
            e = cudaEventQuery(c->ev_poll);
            if (e != cudaErrorNotReady) return e;
            std::this_thread::sleep_for(std::chrono::microseconds(25));
        }
    }
    if (!c->block_sync) return cudaStreamSynchronize(st);
    cudaError_t e = cudaEventRecord(c->ev_block, st);
    if (e != cudaSuccess) return e;
    if (c->spin_us > 0) {
        // hybrid: what is about to finish is caught without a wake-up, what takes long does not hold a core
        const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(c->spin_us);
        do {
            e = cudaEventQuery(c->ev_block);
            if (e != cudaErrorNotReady) return e;
        } while (std::chrono::steady_clock::now() < until);
    }
    return cudaEventSynchronize(c->ev_block);
}

struct HostOut { uint32_t* ctrl; float* raw; uint32_t* idx; double* dres; };
static HostOut host_out(const pmm_ctx* c)
{
    char* ho = static_cast<char*>(c->R().h_out.p);
    HostOut o;
    o.ctrl = reinterpret_cast<uint32_t*>(ho);                                       // 256 B header
    o.raw = reinterpret_cast<float*>(ho + 256);
    o.idx = reinterpret_cast<uint32_t*>(ho + 256 + align_up(sizeof(float) * c->pairs));
    o.dres = reinterpret_cast<double*>(reinterpret_cast<char*>(o.idx) + align_up(sizeof(uint32_t) * c->pairs));
    return o;
}

// The raw floats of the last launch are in h_out (copied by pmm_launch on the copy stream; the double pass may still run).
static int ensure_raw(pmm_ctx* c)
{
    if (!c->launched) return c->fail(PMM_ERR_STATE, "fetch before pmm_launch");
    if (c->have_raw) return PMM_OK;
    cudaSetDevice(c->device);
    PMM_CUDA(c, wait_stream(c, c->copy_stream));
    c->have_raw = true;
    return PMM_OK;
}

// The control words and the whole fallback list of the last launch are in h_out.  Every fetch_* of one launch shares
// these copies: nothing is transferred twice.
static int ensure_lists(pmm_ctx* c)
{
    if (!c->launched) return c->fail(PMM_ERR_STATE, "fetch before pmm_launch");
    if (c->have_lists) return PMM_OK;
    cudaSetDevice(c->device);
    cudaStream_t s = c->list_stream;
    const HostOut o = host_out(c);
    PMM_CUDA(c, wait_stream(c, s));
    const uint32_t nfb = o.ctrl[0], spec = c->spec;
    uint64_t d2h = 12 + sizeof(float) * c->pairs + (sizeof(uint32_t) + sizeof(double)) * spec;
    if (nfb > spec) {
        PMM_CUDA(c, cudaMemcpyAsync(o.idx + spec, static_cast<uint32_t*>(c->R().d_fb_idx.p) + spec, sizeof(uint32_t) * (nfb - spec), cudaMemcpyDeviceToHost, s));
        PMM_CUDA(c, cudaMemcpyAsync(o.dres + spec, static_cast<double*>(c->R().d_dres.p) + spec, sizeof(double) * (nfb - spec), cudaMemcpyDeviceToHost, s));
        PMM_CUDA(c, cudaEventRecord(c->R().ev_lists, s));
        PMM_CUDA(c, wait_stream(c, s));
        d2h += (sizeof(uint32_t) + sizeof(double)) * (nfb - spec);
    }
    c->stats.fallback_pairs = nfb; c->stats.flush_pairs = o.ctrl[1]; c->stats.recheck_pairs = o.ctrl[2]; c->stats.d2h_bytes = d2h;
    cudaEventElapsedTime(&c->stats.ms_f32, c->R().ev[0], c->R().ev[1]);
    cudaEventElapsedTime(&c->stats.ms_fallback, c->R().ev[1], c->R().ev[2]);
    c->have_lists = true;
    return PMM_OK;
}

// (double)(log10f(v) - log10f(2^120)), float subtraction (PairHMMWorker.cpp:190); host libm on purpose
static void log10_of_raw(const float* raw, uint64_t n, double* out)
{
    const float licf = host_tables().log10_ic_f;
    parallel_for(n, 1 << 13, [&](uint64_t a, uint64_t b) { for (uint64_t k = a; k < b; ++k) out[k] = (double)(log10f(raw[k]) - licf); });
}
// log10(d) - log10(2^1020) for the pairs that fell back (PairHMMWorker.cpp:184)
static int log10_of_fallback(const uint32_t* fb_index, const double* fb_value, uint64_t n_fb, uint64_t n, double* out)
{
    const double licd = host_tables().log10_ic_d;
    for (uint64_t k = 0; k < n_fb; ++k) {
        if (fb_index[k] >= n) return PMM_ERR_INVALID;
        out[fb_index[k]] = log10(fb_value[k]) - licd;
    }
    return PMM_OK;
}

// The final doubles of the last launch: log10f of the raw floats is taken while the double pass is still running.
static int fetch_log10_common(pmm_ctx* c, double* out, uint32_t* nfb_out)
{
    int rc = ensure_raw(c);
    if (rc) return rc;
    const HostOut o = host_out(c);
    log10_of_raw(o.raw, c->pairs, out);
    if ((rc = ensure_lists(c))) return rc;
    if (log10_of_fallback(o.idx, o.dres, o.ctrl[0], c->pairs, out) != PMM_OK) return c->fail(PMM_ERR_CUDA, "fallback index out of range");
    if (nfb_out) *nfb_out = o.ctrl[0];
    return PMM_OK;
}

int pmm_fetch_raw(pmm_ctx* c, float* out_raw, uint64_t cap)
{
    if (!c || !out_raw) return PMM_ERR_INVALID;
    if (cap < c->pairs) return c->fail(PMM_ERR_INVALID, "output buffer too small");
    auto t0 = std::chrono::steady_clock::now();
    int rc = ensure_raw(c);
    if (rc) return rc;
    memcpy(out_raw, host_out(c).raw, sizeof(float) * c->pairs);
    c->stats.ms_fetch = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return PMM_OK;
}

int pmm_fetch_fallback_mask(pmm_ctx* c, uint8_t* mask, uint64_t cap)
{
    if (!c || !mask) return PMM_ERR_INVALID;
    if (cap < c->pairs) return c->fail(PMM_ERR_INVALID, "mask buffer too small");
    int rc = ensure_lists(c);
    if (rc) return rc;
    memset(mask, 0, c->pairs);
    const HostOut o = host_out(c);
    for (uint32_t k = 0; k < o.ctrl[0]; ++k) mask[o.idx[k]] = 1;
    return PMM_OK;
}

int pmm_host_finish_log10(const float* raw, uint64_t n, const uint32_t* fb_index, const double* fb_value, uint64_t n_fb, double* out)
{
    if (!out || (n && !raw) || (n_fb && (!fb_index || !fb_value))) return PMM_ERR_INVALID;
    log10_of_raw(raw, n, out);
    return log10_of_fallback(fb_index, fb_value, n_fb, n, out);
}

int pmm_fetch_log10(pmm_ctx* c, double* out, uint64_t cap, uint64_t* n_fallback)
{
    if (!c || !out) return PMM_ERR_INVALID;
    if (cap < c->pairs) return c->fail(PMM_ERR_INVALID, "output buffer too small");
    auto t0 = std::chrono::steady_clock::now();
    uint32_t nfb = 0;
    int rc = fetch_log10_common(c, out, &nfb);
    if (rc) return rc;
    if (n_fallback) *n_fallback = nfb;
    c->stats.ms_fetch = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return PMM_OK;
}

int pmm_fetch_log10_indexed(pmm_ctx* c, double* out, uint64_t cap, uint32_t* fb_index, uint64_t fb_cap, uint64_t* n_fallback)
{
    if (!c || !out || !n_fallback) return PMM_ERR_INVALID;
    if (cap < c->pairs) return c->fail(PMM_ERR_INVALID, "output buffer too small");
    auto t0 = std::chrono::steady_clock::now();
    uint32_t nfb = 0;
    int rc = fetch_log10_common(c, out, &nfb);
    if (rc) return rc;
    *n_fallback = nfb;
    if (fb_index) {
        if (fb_cap < nfb) return c->fail(PMM_ERR_INVALID, "fallback index buffer too small");
        memcpy(fb_index, host_out(c).idx, sizeof(uint32_t) * nfb);
    }
    c->stats.ms_fetch = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return PMM_OK;
}

int pmm_fetch_fallback(pmm_ctx* c, uint32_t* index, double* value, uint64_t cap, uint64_t* count)
{
    if (!c || !count) return PMM_ERR_INVALID;
    int rc = ensure_lists(c);
    if (rc) return rc;
    const HostOut o = host_out(c);
    const uint32_t nfb = o.ctrl[0];
    *count = nfb;
    if (!index && !value) return PMM_OK;
    if (!index || !value || cap < nfb) return c->fail(PMM_ERR_INVALID, "fallback buffers too small");
    memcpy(index, o.idx, sizeof(uint32_t) * nfb);
    memcpy(value, o.dres, sizeof(double) * nfb);
    return PMM_OK;
}

static int pmm_stage_serialized_impl(pmm_ctx* c, const void* reads_ser, uint64_t reads_bytes, const void* haps_ser,
                         uint64_t haps_bytes, int* num_read, int* num_hap)
{
    if (!c) return PMM_ERR_INVALID;
    cudaSetDevice(c->device);
    JobSource src;
    int rc = source_from_serialized(c, reads_ser, reads_bytes, haps_ser, haps_bytes, src);
    if (rc) return rc;
    if (num_read) *num_read = (int)c->tmp_roff.size() - 1;
    if (num_hap) *num_hap = (int)c->tmp_hoff.size() - 1;
    return stage_single_region(c, src);
}

int pmm_get_stats(const pmm_ctx* c, pmm_stats_t* out)
{
    if (!c || !out) return PMM_ERR_INVALID;
    *out = c->stats;
    return PMM_OK;
}

static int pmm_forward_raw_serialized_impl(pmm_ctx* c, const void* reads_ser, uint64_t reads_bytes, const void* haps_ser,
                               uint64_t haps_bytes, float* out_raw, uint64_t cap, int* num_read, int* num_hap)
{
    if (!c || !out_raw) return PMM_ERR_INVALID;
    cudaSetDevice(c->device);
    JobSource src;
    int rc = source_from_serialized(c, reads_ser, reads_bytes, haps_ser, haps_bytes, src);
    if (rc) return rc;
    if (num_read) *num_read = (int)c->tmp_roff.size() - 1;
    if (num_hap) *num_hap = (int)c->tmp_hoff.size() - 1;
    if ((rc = stage_single_region(c, src))) return rc;
    if ((rc = pmm_launch(c))) return rc;
    return pmm_fetch_raw(c, out_raw, cap);
}

static int pmm_forward_log10_serialized_impl(pmm_ctx* c, const void* reads_ser, uint64_t reads_bytes, const void* haps_ser,
                                 uint64_t haps_bytes, double* out, uint64_t cap, int* num_read, int* num_hap,
                                 uint64_t* n_fallback)
{
    if (!c || !out) return PMM_ERR_INVALID;
    cudaSetDevice(c->device);
    JobSource src;
    int rc = source_from_serialized(c, reads_ser, reads_bytes, haps_ser, haps_bytes, src);
    if (rc) return rc;
    if (num_read) *num_read = (int)c->tmp_roff.size() - 1;
    if (num_hap) *num_hap = (int)c->tmp_hoff.size() - 1;
    if ((rc = stage_single_region(c, src))) return rc;
    if ((rc = pmm_launch(c))) return rc;
    return pmm_fetch_log10(c, out, cap, n_fallback);
}

static int pmm_forward_log10_impl(pmm_ctx* c, const pmm_read_t* reads, int num_read, const pmm_hap_t* haps, int num_hap,
                      double* out, uint64_t* n_fallback)
{
    if (!c || !reads || !haps || !out || num_read <= 0 || num_hap <= 0) return c ? c->fail(PMM_ERR_INVALID, "bad arguments") : PMM_ERR_INVALID;
    cudaSetDevice(c->device);
    // every track of every read is one part of the blob, laid out like the wire format (five tracks of a read back to back)
    JobSource src;
    c->tmp_roff.assign(num_read + 1, 0); c->tmp_hoff.assign(num_hap + 1, 0);
    c->tmp_rdesc.resize(num_read); c->tmp_hdesc.resize(num_hap);
    src.read_parts.reserve(5 * (size_t)num_read); src.hap_parts.reserve(num_hap);
    uint64_t off = 0;
    for (int i = 0; i < num_read; ++i) {
        const pmm_read_t& r = reads[i];
        if (r.len <= 0 || !r._b || !r._q || !r._i || !r._d || !r._c) return c->fail(PMM_ERR_INVALID, "read of length 0 or null track");
        const uint32_t len = (uint32_t)r.len;
        c->tmp_roff[i + 1] = c->tmp_roff[i] + len;
        c->tmp_rdesc[i] = ReadDesc{(uint32_t)off, len, len};
        const char* tr[5] = {r._b, r._q, r._i, r._d, r._c};
        for (int t = 0; t < 5; ++t) src.read_parts.push_back(BlobPart{reinterpret_cast<const uint8_t*>(tr[t]), len});
        off += 5ull * len;
        if (off >= (1ull << 31)) return c->fail(PMM_ERR_INVALID, "batch larger than 2 GiB: split it");
    }
    off = 0;
    for (int i = 0; i < num_hap; ++i) {
        if (haps[i].len <= 0 || !haps[i]._b) return c->fail(PMM_ERR_INVALID, "haplotype of length 0 or null");
        const uint32_t len = (uint32_t)haps[i].len;
        c->tmp_hoff[i + 1] = c->tmp_hoff[i] + len;
        c->tmp_hdesc[i] = HapDesc{(uint32_t)off, len};
        src.hap_parts.push_back(BlobPart{reinterpret_cast<const uint8_t*>(haps[i]._b), len});
        off += len;
    }
    src.rdesc = c->tmp_rdesc.data(); src.hdesc = c->tmp_hdesc.data();
    int rc = stage_single_region(c, src);
    if (rc) return rc;
    if ((rc = pmm_launch(c))) return rc;
    return pmm_fetch_log10(c, out, (uint64_t)num_read * num_hap, n_fallback);
}

// A GKL-shaped batch: one testcase per pair, pointers shared between the pairs of a read and of a haplotype.  Consecutive
// testcases with the same read are a row; consecutive rows with the same haplotype sequence are a region.  The results
// of the regions, read-major, are then exactly the testcases' order.
static int pmm_forward_log10_testcases_impl(pmm_ctx* c, const pmm_testcase_t* tc, uint64_t n, double* out, uint64_t* n_fallback)
{
    if (!c) return PMM_ERR_INVALID;
    if (!tc || !out || !n) return c->fail(PMM_ERR_INVALID, "bad arguments");
    if (n >= (1ull << 31)) return c->fail(PMM_ERR_INVALID, "more than 2^31 pairs in one batch: split it");
    cudaSetDevice(c->device);
    auto same_read = [](const pmm_testcase_t& a, const pmm_testcase_t& b) {
        return a.rs == b.rs && a.rslen == b.rslen && a.q == b.q && a.i == b.i && a.d == b.d && a.c == b.c; };
    JobSource src;
    std::vector<pmm_region_t> regions;
    std::vector<uint32_t>& roff = c->tmp_roff; std::vector<uint32_t>& hoff = c->tmp_hoff;
    roff.assign(1, 0); hoff.assign(1, 0);
    c->tmp_rdesc.clear(); c->tmp_hdesc.clear();
    uint64_t rbytes = 0, hbytes = 0;
    uint64_t k = 0, prev_row = 0, prev_len = 0;        // previous row: testcases [prev_row, prev_row + prev_len)
    bool have_prev = false;
    while (k < n) {
        uint64_t e = k + 1;
        while (e < n && same_read(tc[k], tc[e])) ++e;
        const pmm_testcase_t& t = tc[k];
        if (t.rslen <= 0 || !t.rs || !t.q || !t.i || !t.d || !t.c) return c->fail(PMM_ERR_INVALID, "read of length 0 or null track");
        bool same_haps = have_prev && e - k == prev_len;
        for (uint64_t z = 0; same_haps && z < e - k; ++z)
            same_haps = tc[k + z].hap == tc[prev_row + z].hap && tc[k + z].haplen == tc[prev_row + z].haplen;
        if (!same_haps) {
            // a new region: its haplotypes are this row's
            pmm_region_t rg{(uint32_t)c->tmp_rdesc.size(), 0, (uint32_t)c->tmp_hdesc.size(), (uint32_t)(e - k)};
            for (uint64_t z = k; z < e; ++z) {
                if (tc[z].haplen <= 0 || !tc[z].hap) return c->fail(PMM_ERR_INVALID, "haplotype of length 0 or null");
                const uint32_t len = (uint32_t)tc[z].haplen;
                c->tmp_hdesc.push_back(HapDesc{(uint32_t)hbytes, len});
                hoff.push_back(hoff.back() + len);
                src.hap_parts.push_back(BlobPart{reinterpret_cast<const uint8_t*>(tc[z].hap), len});
                hbytes += len;
            }
            regions.push_back(rg);
        }
        const uint32_t len = (uint32_t)t.rslen;
        c->tmp_rdesc.push_back(ReadDesc{(uint32_t)rbytes, len, len});
        roff.push_back(roff.back() + len);
        const char* tr[5] = {t.rs, t.q, t.i, t.d, t.c};
        for (int z = 0; z < 5; ++z) src.read_parts.push_back(BlobPart{reinterpret_cast<const uint8_t*>(tr[z]), len});
        rbytes += 5ull * len;
        if (rbytes + hbytes >= (1ull << 31)) return c->fail(PMM_ERR_INVALID, "batch larger than 2 GiB: split it");
        regions.back().num_read++;
        prev_row = k; prev_len = e - k; have_prev = true;
        k = e;
    }
    src.rdesc = c->tmp_rdesc.data(); src.hdesc = c->tmp_hdesc.data();
    int rc = stage_common(c, (uint32_t)c->tmp_rdesc.size(), roff.data(), (uint32_t)c->tmp_hdesc.size(), hoff.data(), src,
                          (uint32_t)regions.size(), regions.data());
    if (rc) return rc;
    if (c->pairs != n) return c->fail(PMM_ERR_INVALID, "internal: testcase grouping lost pairs");
    if ((rc = pmm_launch(c))) return rc;
    return pmm_fetch_log10(c, out, n, n_fallback);
}

static int pmm_plan_flat_impl(uint32_t num_read, const uint32_t* read_off, uint32_t num_hap, const uint32_t* hap_off,
                  uint32_t num_region, const pmm_region_t* regions, int sm_count, int tasks_per_warp,
                  pmm_task_info_t* out, uint64_t capacity, uint64_t* num_tasks)
{
    Plan plan; std::string perr;
    int rc = plan_job(num_read, read_off, num_hap, hap_off, num_region, regions, sm_count, tasks_per_warp, plan, perr);
    if (rc) { g_create_error = perr; return rc; }
    if (num_tasks) *num_tasks = plan.tasks.size();
    if (!out) return PMM_OK;
    if (capacity < plan.tasks.size()) { g_create_error = "task buffer too small"; return PMM_ERR_INVALID; }
    for (const LaunchSeg& seg : plan.segs)
        for (uint32_t k = seg.task_first; k < seg.task_first + seg.task_count; ++k) {
            const Task& t = plan.tasks[k];
            pmm_task_info_t& o = out[k];
            for (int z = 0; z < 4; ++z) { o.read[z] = t.read[z]; o.out_base[z] = t.out_base[z]; }
            o.hap_first = t.hap_first; o.num_hap = t.nhaps; o.num_read = t.nreads;
            o.rows_per_lane = (uint32_t)seg.v.K; o.lanes_per_read = (uint32_t)seg.v.W; o.striped = seg.v.striped ? 1u : 0u;
        }
    return PMM_OK;
}

int pmm_stage_flat(pmm_ctx* c, uint32_t num_read, const uint32_t* read_off,
                   const uint8_t* bases, const uint8_t* q, const uint8_t* i, const uint8_t* d, const uint8_t* cc,
                   uint32_t num_hap, const uint32_t* hap_off, const uint8_t* hap_bases,
                   uint32_t num_region, const pmm_region_t* regions)
{
    return guarded(c, [&] { return pmm_stage_flat_impl(c, num_read, read_off, bases, q, i, d, cc, num_hap, hap_off, hap_bases, num_region, regions); });
}

int pmm_stage_serialized(pmm_ctx* c, const void* reads_ser, uint64_t reads_bytes, const void* haps_ser,
                         uint64_t haps_bytes, int* num_read, int* num_hap)
{
    return guarded(c, [&] { return pmm_stage_serialized_impl(c, reads_ser, reads_bytes, haps_ser, haps_bytes, num_read, num_hap); });
}

int pmm_forward_raw_serialized(pmm_ctx* c, const void* reads_ser, uint64_t reads_bytes, const void* haps_ser,
                               uint64_t haps_bytes, float* out_raw, uint64_t cap, int* num_read, int* num_hap)
{
    return guarded(c, [&] { return pmm_forward_raw_serialized_impl(c, reads_ser, reads_bytes, haps_ser, haps_bytes, out_raw, cap, num_read, num_hap); });
}

int pmm_forward_log10_serialized(pmm_ctx* c, const void* reads_ser, uint64_t reads_bytes, const void* haps_ser,
                                 uint64_t haps_bytes, double* out, uint64_t cap, int* num_read, int* num_hap,
                                 uint64_t* n_fallback)
{
    return guarded(c, [&] { return pmm_forward_log10_serialized_impl(c, reads_ser, reads_bytes, haps_ser, haps_bytes, out, cap, num_read, num_hap, n_fallback); });
}

int pmm_forward_log10(pmm_ctx* c, const pmm_read_t* reads, int num_read, const pmm_hap_t* haps, int num_hap,
                      double* out, uint64_t* n_fallback)
{
    return guarded(c, [&] { return pmm_forward_log10_impl(c, reads, num_read, haps, num_hap, out, n_fallback); });
}

int pmm_plan_flat(uint32_t num_read, const uint32_t* read_off, uint32_t num_hap, const uint32_t* hap_off,
                  uint32_t num_region, const pmm_region_t* regions, int sm_count, int tasks_per_warp,
                  pmm_task_info_t* out, uint64_t capacity, uint64_t* num_tasks)
{
    return guarded(nullptr, [&] { return pmm_plan_flat_impl(num_read, read_off, num_hap, hap_off, num_region, regions, sm_count, tasks_per_warp, out, capacity, num_tasks); });
}
int pmm_forward_log10_testcases(pmm_ctx* c, const pmm_testcase_t* tc, uint64_t n, double* out, uint64_t* n_fallback)
{
    return guarded(c, [&] { return pmm_forward_log10_testcases_impl(c, tc, n, out, n_fallback); });
}

int pmm_host_table(int which, void* out, uint64_t capacity_bytes)
{
    const HostTables& t = host_tables();
    const void* src = nullptr; size_t n = 0;
    switch (which) {
        case 0: src = t.ph2pr_f; n = sizeof t.ph2pr_f; break;
        case 1: src = t.m2m_f;   n = sizeof t.m2m_f;   break;
        case 2: src = t.ph2pr_d; n = sizeof t.ph2pr_d; break;
        case 3: src = t.m2m_d;   n = sizeof t.m2m_d;   break;
        case 4: src = &t.log10_ic_f; n = sizeof(float); break;
        case 5: src = &t.log10_ic_d; n = sizeof(double); break;
        default: return PMM_ERR_INVALID;
    }
    if (!out || capacity_bytes < n) return PMM_ERR_INVALID;
    memcpy(out, src, n);
    return PMM_OK;
}

int pmm_measure_fp32_peak(pmm_ctx* c, double* lane_instr_per_s, double* sm_mhz)
{
    if (!c || !lane_instr_per_s) return PMM_ERR_INVALID;
    cudaSetDevice(c->device);
    const int ctas = c->sm_count * 8, iters = 500;
    PMM_CUDA(c, c->d_probe.reserve(sizeof(float) * ctas * 256));
    cudaStream_t s = c->stream;
    PMM_CUDA(c, launch_fp32_probe(static_cast<float*>(c->d_probe.p), 25, ctas, s));
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        PMM_CUDA(c, cudaEventRecord(c->ev_probe, s));
        PMM_CUDA(c, launch_fp32_probe(static_cast<float*>(c->d_probe.p), iters, ctas, s));
        cudaEvent_t e1; cudaEventCreate(&e1);
        PMM_CUDA(c, cudaEventRecord(e1, s));
        PMM_CUDA(c, cudaEventSynchronize(e1));
        float ms = 0; cudaEventElapsedTime(&ms, c->ev_probe, e1); cudaEventDestroy(e1);
        const double ops = (double)ctas * 256 * iters * 512.0;
        best = std::max(best, ops / (ms * 1e-3));
    }
    *lane_instr_per_s = best;
    if (sm_mhz) *sm_mhz = best / ((double)c->sm_count * 128.0) * 1e-6;   // lower bound: assumes 128 lanes/clk/SM fully used
    return PMM_OK;
}

int pmm_get_timeline(pmm_ctx* c, pmm_timeline_t* out)
{
    if (!c || !out) return PMM_ERR_INVALID;
    if (!c->launched) return c->fail(PMM_ERR_STATE, "pmm_get_timeline before pmm_launch");
    cudaSetDevice(c->device);
    PMM_CUDA(c, cudaEventSynchronize(c->R().ev[2]));
    float a = 0, b = 0, d = 0;
    PMM_CUDA(c, cudaEventElapsedTime(&a, c->ev_ref, c->R().ev[0]));
    PMM_CUDA(c, cudaEventElapsedTime(&b, c->ev_ref, c->R().ev[1]));
    PMM_CUDA(c, cudaEventElapsedTime(&d, c->ev_ref, c->R().ev[2]));
    out->ref_host_s = c->ref_host_s;
    out->kernels_start_s = a * 1e-3; out->f32_end_s = b * 1e-3; out->kernels_end_s = d * 1e-3;
    return PMM_OK;
}

int pmm_measure_fp64_peak(pmm_ctx* c, double* lane_instr_per_s)
{
    if (!c || !lane_instr_per_s) return PMM_ERR_INVALID;
    cudaSetDevice(c->device);
    const int ctas = c->sm_count * 8, iters = 60;
    PMM_CUDA(c, c->d_probe.reserve(sizeof(double) * ctas * 256));
    cudaStream_t s = c->stream;
    PMM_CUDA(c, launch_fp64_probe(static_cast<double*>(c->d_probe.p), 5, ctas, s));
    double best = 0;
    cudaEvent_t e1;
    PMM_CUDA(c, cudaEventCreate(&e1));
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(c->ev_probe, s);
        cudaError_t e = launch_fp64_probe(static_cast<double*>(c->d_probe.p), iters, ctas, s);
        if (e == cudaSuccess) e = cudaEventRecord(e1, s);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaEventDestroy(e1); return c->fail_cuda(e, "fp64 probe"); }
        float ms = 0; cudaEventElapsedTime(&ms, c->ev_probe, e1);
        best = std::max(best, (double)ctas * 256 * iters * 512.0 / (ms * 1e-3));
    }
    cudaEventDestroy(e1);
    *lane_instr_per_s = best;
    return PMM_OK;
}

}  // extern "C"
