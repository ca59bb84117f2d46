// sw_kernels.cu -- sm_100a Smith-Waterman with backtrack: matrix fill, end-cell search and traceback in one kernel.
//
// The arithmetic and every tie-break are those of the reference's AVX2 kernel
// (/root/reference/htc-sw/intel_avx/PairWiseSW.h: MAIN_CODE :4-40, boundaries :218-228, end cell :233-262, getCIGAR
// :275-437), which Falcon's own SWPairwiseAlignmentOneBatch (host/FalconSW_AVX.cpp:315-411, :2303-2415) agrees with.
//
// Mapping to the machine:
//   * One warp aligns one pair.  The matrix is processed in blocks of 32 x kRowsPerLane rows: lane l owns kRowsPerLane
//     consecutive rows and sweeps the columns, skewed by one column per lane, so the lane above has always just
//     finished the column a lane is about to do (the same wavefront as the PairHMM kernel).  H and F of the row above
//     a lane's first row arrive by two __shfl_up_sync per step; H of a lane's own rows, and E, stay in registers.
//   * The last row of a block is carried to the next block through a per-warp row in shared memory (written by lane 31,
//     read 31 steps earlier by lane 0 -- one buffer is enough).  The alternate sequence, the last row / last column of H
//     (for the end-cell search) and the traceback tile live in shared memory too.
//   * Backtrack codes are packed 8 cells per 32-bit word per row (row-major), so a lane stores one word every 8 steps
//     and the traceback can fetch a 32-row x 64-column tile with one 32-byte segment per lane.
//   * The traceback is inherently serial (each step depends on the cell before): lane 0 walks, all lanes fetch tiles.
#include "sw_kernels.cuh"

#include <climits>

namespace sw {
namespace {

constexpr int kLowInit = INT32_MIN / 2;          // LOW_INIT_VALUE (smithwaterman_common.h:52)
constexpr int kMinCutoff = -100000000;           // MATRIX_MIN_CUTOFF (:51)
constexpr int kInsert = 1, kDelete = 2, kInsertExt = 4, kDeleteExt = 8;
constexpr int kStateClip = 4;                    // STATE_CLIP (host/common.h:23)
constexpr int kSoftClip = 0, kIndel = 1, kLeadingIndel = 2, kIgnore = 3;

// Integer multiply-add / high multiply as explicit PTX so that they stay IMAD / IMAD.HI (FMA pipe).
__device__ __forceinline__ int imad(int a, int b, int c)
{
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

struct WarpSmem {
    int* carryH; int* carryF; int* lastrow; int* lastcol; uint32_t* tile; uint8_t* alt;
};

__host__ __device__ inline size_t align4(size_t x) { return (x + 3) & ~(size_t)3; }

__host__ __device__ inline size_t per_warp_bytes(uint32_t max_l1, uint32_t max_l2)
{
    const size_t c1 = align4(max_l2 + 1), r1 = align4(max_l1 + 1);
    const size_t b = sizeof(int) * (3 * c1 + r1) + sizeof(uint32_t) * 32 * 8 + align4(max_l2 + 4);
    return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ WarpSmem carve(char* base, uint32_t max_l1, uint32_t max_l2)
{
    WarpSmem w;
    const size_t c1 = align4(max_l2 + 1), r1 = align4(max_l1 + 1);
    w.carryH = reinterpret_cast<int*>(base);
    w.carryF = w.carryH + c1;
    w.lastrow = w.carryF + c1;
    w.lastcol = w.lastrow + c1;
    w.tile = reinterpret_cast<uint32_t*>(w.lastcol + r1);
    w.alt = reinterpret_cast<uint8_t*>(w.tile + 32 * 8);
    return w;
}

template <int K>
__global__ void __launch_bounds__(kWarpsPerCta * 32) sw_align_kernel(const Args a)
{
    extern __shared__ int4 smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = per_warp_bytes(a.max_l1, a.max_l2);
    const WarpSmem sm = carve(reinterpret_cast<char*>(smem_raw) + warp * per_warp, a.max_l1, a.max_l2);
    const int match = a.match, mismatch = a.mismatch, open = a.open, extend = a.extend, strategy = a.strategy;
    const bool indel = strategy == kIndel || strategy == kLeadingIndel;
    // multipliers the compiler must not know (see the step function): the host passes -1, 2, 4, 8
    const int neg1 = a.k_neg1, two = a.k_two, four = a.k_four, eight = a.k_eight;
    auto sign_of = [](int x) { return (unsigned)x >> 31; };

    for (;;) {
        uint32_t p = 0;
        if (lane == 0) p = atomicAdd(a.counter, 1u);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= a.npairs) break;
        const PairDesc pd = a.pairs[p];
        const int nrow = (int)pd.l1, ncol = (int)pd.l2;
        const uint8_t* s1 = a.seq1 + pd.s1;
        const uint8_t* s2 = a.seq2 + pd.s2;
        uint32_t* B = a.bt + pd.bt_off;
        const uint32_t stride = pd.bt_stride;

        __syncwarp();
        for (int x = lane; x < ncol; x += 32) sm.alt[x] = s2[x];
        // row 0: H[0][j] (PairWiseSW.h:218-228, H[0][0] = 0 at :87), F[0][j] = low
        for (int j = lane; j <= ncol; j += 32) {
            sm.carryH[j] = (j && indel) ? open + (j - 1) * extend : 0;
            sm.carryF[j] = kLowInit;
        }
        __syncwarp();

        // ---- matrix fill -----------------------------------------------------------------------------------------
        const int nblk = (nrow + 32 * K - 1) / (32 * K);
        #pragma unroll 1
        for (int blk = 0; blk < nblk; ++blk) {
            const int ifirst = blk * 32 * K + lane * K + 1;           // 1-based row of this lane's first row
            int Hrow[K], E[K];
            uint32_t acc[K];
            int c1[K];
            uint32_t* brow[K];                                         // backtrack row of each of the lane's rows
            #pragma unroll
            for (int k = 0; k < K; ++k) {
                const int i = ifirst + k;
                c1[k] = i <= nrow ? (int)s1[i - 1] : -1;
                Hrow[k] = indel ? open + (i - 1) * extend : 0;       // H[i][0]
                E[k] = kLowInit;
                acc[k] = 0;
                brow[k] = i <= nrow ? B + (size_t)(i - 1) * stride : nullptr;
            }
            int prev_up = (ifirst - 1 == 0) ? 0 : (indel ? open + (ifirst - 2) * extend : 0);   // H[ifirst-1][0]
            int lastF = kLowInit;
            // the lane and row that hold the matrix's last row (for the end-cell search); -1 if not in this block
            const int own_k = (nrow >= ifirst && nrow < ifirst + K) ? nrow - ifirst : -1;

            // one step: column j of this lane's K rows.  MAIN_CODE (PairWiseSW.h:4-40): same comparisons, same order.
            // Pipe balance: compares, selects, min/max and shifts issue on the ALU pipe, multiply-adds on the FMA pipe, and
            // the straightforward form of this update is 29 ALU instructions per cell.  Here each tie-break bit is the
            // sign of a difference (a multiply-add by -1 that ptxas cannot see through, then one shift) instead of a
            // compare plus a select, and the four bits are packed and appended to the row's word by multiply-adds.
            // Measured on 16 640 pairs: 431 GCUPS against 406 for compare/select; moving the remaining adds and the
            // sign extraction (mul.hi) to the FMA pipe as well was slower (366): three-register IMAD forms issue at half
            // rate like three-register FFMA.  (Differences cannot overflow: |values| <= 2^30 + 2^20.)
            auto step = [&](const int j, int upH, int upF) {
                const int c2 = (int)sm.alt[j - 1];
                int diag = prev_up;
                prev_up = upH;
                int hup = upH, fup = upF;
                const int pw = 1 << (4 * ((j - 1) & 7));               // where this column's 4 bits go in the word
                const bool flush = ((j - 1) & 7) == 7 || j == ncol;
                #pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int hleft = Hrow[k];
                    const int ext_h = E[k] + extend, open_h = hleft + open;
                    const int e11 = max(ext_h, open_h);
                    const unsigned ins_open = sign_of(imad(open_h, neg1, ext_h));   // open_h > ext_h; ties extend
                    const int ext_v = fup + extend, open_v = hup + open;
                    const int f11 = max(ext_v, open_v);
                    const unsigned del_open = sign_of(imad(open_v, neg1, ext_v));
                    const int m11 = diag + (c1[k] == c2 ? match : mismatch);
                    const int h0 = max(kMinCutoff, m11);
                    const unsigned take_ins = sign_of(imad(e11, neg1, h0));          // e11 > h0
                    const int h1 = max(h0, e11);
                    const unsigned take_del = sign_of(imad(f11, neg1, h1));          // f11 > h1
                    const int h11 = max(h1, f11);
                    const int code = imad((int)take_del, eight, imad((int)take_ins, four, imad((int)del_open, two, (int)ins_open)));
                    diag = hleft;
                    Hrow[k] = h11; E[k] = e11; hup = h11; fup = f11;
                    acc[k] = (uint32_t)imad(code, pw, (int)acc[k]);
                    if (flush) {
                        if (brow[k]) brow[k][(j - 1) >> 3] = acc[k];
                        acc[k] = 0;
                    }
                }
                lastF = fup;
                if (own_k >= 0) {
                    int v = Hrow[0];
                    #pragma unroll
                    for (int k = 1; k < K; ++k) v = own_k == k ? Hrow[k] : v;
                    sm.lastrow[j] = v;
                }
                if (lane == 31) { sm.carryH[j] = Hrow[K - 1]; sm.carryF[j] = lastF; }
            };

            const int steps = ncol + 31;
            int t = 0;
            // fill: lanes join one per step
            #pragma unroll 1
            for (; t < min(31, steps); ++t) {
                int upH = __shfl_up_sync(0xffffffffu, Hrow[K - 1], 1);
                int upF = __shfl_up_sync(0xffffffffu, lastF, 1);
                const int j = t - lane + 1;
                if (j >= 1 && j <= ncol) {
                    if (lane == 0) { upH = sm.carryH[j]; upF = sm.carryF[j]; }
                    step(j, upH, upF);
                }
            }
            // steady: every lane has a column
            #pragma unroll 1
            for (; t < ncol; ++t) {
                int upH = __shfl_up_sync(0xffffffffu, Hrow[K - 1], 1);
                int upF = __shfl_up_sync(0xffffffffu, lastF, 1);
                const int j = t - lane + 1;
                if (lane == 0) { upH = sm.carryH[j]; upF = sm.carryF[j]; }
                step(j, upH, upF);
            }
            // drain
            #pragma unroll 1
            for (; t < steps; ++t) {
                int upH = __shfl_up_sync(0xffffffffu, Hrow[K - 1], 1);
                int upF = __shfl_up_sync(0xffffffffu, lastF, 1);
                const int j = t - lane + 1;
                if (j >= 1 && j <= ncol) {
                    if (lane == 0) { upH = sm.carryH[j]; upF = sm.carryF[j]; }
                    step(j, upH, upF);
                }
            }
            // H[i][ncol] of the lane's rows: still in registers, nothing touched them after the last column
            #pragma unroll
            for (int k = 0; k < K; ++k) if (ifirst + k <= nrow) sm.lastcol[ifirst + k] = Hrow[k];
            __syncwarp();
        }

        // ---- end cell (PairWiseSW.h:233-262): anti-diagonals ascending, last-row cell before last-column cell ----------
        int ti = 0, tj = 0, seg = 0, best = INT32_MIN;
        if (lane == 0) {
            int max_i = 0, max_j = 0;
            bool updated = false;
            for (int ad = 1; ad <= nrow + ncol; ++ad) {
                const int jj = ad - nrow;
                if (jj >= 1 && jj <= ncol && (strategy == kSoftClip || strategy == kIgnore)) {
                    const int s = sm.lastrow[jj];
                    if (best < s || (best == s && abs(nrow - jj) < abs(max_i - max_j))) { best = s; max_i = nrow; max_j = jj; updated = true; }
                }
                const int ii = ad - ncol;
                if (ii >= 1 && ii <= nrow) {
                    const int s = sm.lastcol[ii];
                    if (best < s || (best == s && (max_j == ncol || abs(ii - ncol) <= abs(max_i - max_j)))) { best = s; max_i = ii; max_j = ncol; updated = true; }
                }
            }
            // start of the traceback (getCIGAR, :285-314)
            if (strategy == kIndel) { ti = nrow; tj = ncol; }
            else if (strategy == kLeadingIndel) { ti = max_i; tj = ncol; }
            else { ti = max_i; tj = max_j; }
            if (strategy == kIgnore && updated && tj != ncol) { ti = nrow; seg = ncol - max_j; }
        }

        // ---- traceback: lane 0 walks, all lanes fetch 32-row x 64-column tiles of backtrack codes -----------------------
        int2* out = a.cigars + (size_t)pd.index * a.cigar_cap;
        int n = 0;                         // elements written (run-length encoded, still in backward order)
        int cur_state = -1, cur_len = 0;   // the open run
        int raw_ops = 0;                   // getCIGAR's cigarId: number of un-merged operations so far
        int state = 0;
        auto emit = [&](int st, int len) {
            if (st == cur_state) { cur_len += len; return; }
            if (cur_state >= 0) { if ((uint32_t)n < a.cigar_cap) out[n] = make_int2(cur_len, cur_state); ++n; }
            cur_state = st; cur_len = len;
        };
        if (lane == 0 && tj < ncol && strategy == kSoftClip) { emit(kStateClip, ncol - tj); ++raw_ops; }
        for (;;) {
            const int ai = __shfl_sync(0xffffffffu, ti, 0), aj = __shfl_sync(0xffffffffu, tj, 0);
            if (!(ai > 0 && aj > 0)) break;
            const int cw0 = max(0, ((aj - 1) >> 3) - 7);
            {
                const int row = ai - lane;
                #pragma unroll
                for (int w = 0; w < 8; ++w) {
                    uint32_t v = 0;
                    if (row >= 1 && (uint32_t)(cw0 + w) < stride) v = __ldcg(B + (size_t)(row - 1) * stride + cw0 + w);
                    sm.tile[w * 32 + lane] = v;
                }
            }
            __syncwarp();
            if (lane == 0) {
                while (ti > 0 && tj > 0 && ai - ti < 32 && ((tj - 1) >> 3) >= cw0) {
                    const int btr = (int)((sm.tile[(((tj - 1) >> 3) - cw0) * 32 + (ai - ti)] >> (4 * ((tj - 1) & 7))) & 15u);
                    // stored bits: 0 insertion opened (not an extension), 1 deletion opened, 2 insertion taken, 3 deletion
                    // taken (it wins over the insertion); the reference's codes are move + "was an extension" flags
                    const int ins_ext = (btr & 1) ? 0 : kInsertExt, del_ext = (btr & 2) ? 0 : kDeleteExt;
                    if (state == kInsertExt) { --tj; emit(kInsert, 1); state = ins_ext; }
                    else if (state == kDeleteExt) { --ti; emit(kDelete, 1); state = del_ext; }
                    else {
                        if (btr & 8) { --ti; emit(kDelete, 1); state = del_ext; }
                        else if (btr & 4) { --tj; emit(kInsert, 1); state = ins_ext; }
                        else { --ti; --tj; emit(0, (raw_ops == 0 && strategy == kIgnore) ? seg + 1 : 1); state = 0; }
                        ++raw_ops;
                    }
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            int off;
            if (strategy == kSoftClip) {
                if (tj > 0) emit(kStateClip, tj);
                off = ti;
            } else if (strategy == kIgnore) {
                if (tj > 0) emit(cur_state >= 0 ? cur_state : 0, tj);     // "same operation as the last one" (:391-394)
                off = ti - tj;
            } else {
                if (ti > 0) emit(kDelete, ti);
                else if (tj > 0) emit(kInsert, tj);
                off = 0;
            }
            if (cur_state >= 0) { if ((uint32_t)n < a.cigar_cap) out[n] = make_int2(cur_len, cur_state); ++n; }
            // forward order
            const int stored = min(n, (int)a.cigar_cap);
            for (int x = 0, y = stored - 1; x < y; ++x, --y) { const int2 u = out[x]; out[x] = out[y]; out[y] = u; }
            a.n_elem[pd.index] = n;
            a.offset[pd.index] = off;
            if (a.score) a.score[pd.index] = best;
        }
    }
}

}  // namespace

size_t smem_bytes_per_warp(uint32_t max_l1, uint32_t max_l2) { return per_warp_bytes(max_l1, max_l2); }

cudaError_t launch_align(const Args& a, int sm_count, cudaStream_t s, int* ctas_out)
{
    const size_t smem = kWarpsPerCta * smem_bytes_per_warp(a.max_l1, a.max_l2);
    auto kern = sw_align_kernel<kRowsPerLane>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarpsPerCta * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    const int want = (int)((a.npairs + kWarpsPerCta - 1) / kWarpsPerCta);
    const int ctas = want < sm_count * per_sm ? want : sm_count * per_sm;
    if (ctas_out) *ctas_out = ctas;
    kern<<<ctas, kWarpsPerCta * 32, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace sw
