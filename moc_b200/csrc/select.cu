// Top-J patch selection, union and ascending compaction; sorted top-J; pooled top-K.
//
// Replaces utils/patch_selection_classifier_index.py:17-87 (the four selectors), the Python set union and
// sort of main_moc.py:341-354, Tensor.topk in the selectors' return values, and topj_pooling /
// delta_*_classifier_pooling (utils/patch_selection_classifier.py:18-78) for the zero-shot path.
//
// Selection is an exact radix select on the order-preserving uint32 image of the fp32 keys: the value of rank
// J is found digit by digit, rows strictly beyond it are taken, and ties at the threshold are taken in
// ascending row order until exactly J rows are chosen (torch leaves tie order unspecified).  The batched
// selection kernel reads a long unmasked column ONCE (select_rows_sampled: a provisional threshold from a 1/16 sample,
// every key above it parked in shared memory, the top J resolved there) and otherwise scans it three times with
// 16-byte loads (range, a 4096-bin histogram of the occupied range, then marking + on-chip resolution of the
// threshold bin; see select_rows_fast).  The generic four-pass version serves the stand-alone top-J / pooling
// kernels and degenerate columns.  A column is a stored key plane or, in the compact key layout of wide class sets
// (include/moc_b200.h), L_c - lse from two planes.  The union of the 2C+2 selections of a slide is a bitmap over its
// rows, so the ascending order of the reference's sorted(set(...)) falls out of the compaction for free and nothing
// ever goes back to the host.
#include <stdlib.h>

#include "common.cuh"

namespace moc {

constexpr int SEL_THREADS = 512;
constexpr int SEL_WARPS = SEL_THREADS / 32;

struct SelShared {
    unsigned int hist[256];
    unsigned int warp_tot[SEL_WARPS];
    unsigned int prefix;
    unsigned int need;
    unsigned int n_equal;
    unsigned int taken;
    unsigned int n_kept;
};

template <bool SMALLEST>
__device__ __forceinline__ uint32_t sel_key(float v) {
    const uint32_t u = f2ord(v);
    return SMALLEST ? ~u : u;
}

// A column of selection keys: one stored plane of a slide (possibly strided: the stand-alone top-J over [N,C] logits),
// or - compact key layout, softmax selections - L_c - lse = log softmax(L)[:,c] formed from two planes on the fly: the
// same ranking of the slide's patches as the softmax values themselves (up to rounding among near-equal keys).
struct ColPlain {
    static constexpr int kScanDepth = 4;
    const float* v;
    int64_t ld;
    __device__ __forceinline__ float at(int i) const { return v[(int64_t)i * ld]; }
    __device__ __forceinline__ bool vec() const { return ld == 1; }
    __device__ __forceinline__ int head() const { return (4 - (int)((reinterpret_cast<uintptr_t>(v) >> 2) & 3)) & 3; }
    __device__ __forceinline__ float4 at4(int head, int k) const { return __ldg(reinterpret_cast<const float4*>(v + head) + k); }
};
struct ColLogSoftmax {
    static constexpr int kScanDepth = 2;
    const float *l, *lse;
    __device__ __forceinline__ float at(int i) const { return l[i] - lse[i]; }
    // 16-byte loads need both planes in the same alignment phase (key_stride a multiple of 4: ops.alloc_keys pads it)
    __device__ __forceinline__ bool vec() const {
        return ((reinterpret_cast<uintptr_t>(l) ^ reinterpret_cast<uintptr_t>(lse)) & 15) == 0;
    }
    __device__ __forceinline__ int head() const { return (4 - (int)((reinterpret_cast<uintptr_t>(l) >> 2) & 3)) & 3; }
    __device__ __forceinline__ float4 at4(int head, int k) const {
        const float4 a = __ldg(reinterpret_cast<const float4*>(l + head) + k);
        const float4 b = __ldg(reinterpret_cast<const float4*>(lse + head) + k);
        return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    }
};

// Block-wide: value (as ordered uint) of the element of rank j (1-based, largest first) among the kept
// rows of the column.  Leaves sh.prefix = threshold, sh.need = how many threshold-equal rows to take,
// sh.n_equal = how many exist.  Requires 1 <= j <= kept rows.
template <bool SMALLEST, bool HAS_MASK, typename Col>
__device__ void radix_threshold(const Col& col, const uint8_t* __restrict__ mk, int n, int j, SelShared& sh) {
    const int tid = threadIdx.x;
    uint32_t prefix = 0, known = 0;
    uint32_t need = (uint32_t)j;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) sh.hist[tid] = 0;
        __syncthreads();
        int cur_bin = -1;
        unsigned int cur_cnt = 0;
        for (int i = tid; i < n; i += SEL_THREADS) {
            if (HAS_MASK && !mk[i]) continue;
            const uint32_t u = sel_key<SMALLEST>(col.at(i));
            if ((u & known) != prefix) continue;
            const int b = (u >> shift) & 255;
            if (b == cur_bin) {
                ++cur_cnt;
            } else {
                if (cur_cnt) atomicAdd(&sh.hist[cur_bin], cur_cnt);
                cur_bin = b;
                cur_cnt = 1;
            }
        }
        if (cur_cnt) atomicAdd(&sh.hist[cur_bin], cur_cnt);
        __syncthreads();
        if (tid < 32) {
            // lane l owns bins 255-8l .. 248-8l (descending); find where the running count reaches `need`
            unsigned int loc[8], s = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                loc[k] = sh.hist[255 - (tid * 8 + k)];
                s += loc[k];
            }
            unsigned int incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(FULL, incl, o);
                if (tid >= o) incl += t;
            }
            const unsigned int excl = incl - s;
            if (excl < need && need <= incl) {
                unsigned int run = excl;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (run < need && need <= run + loc[k]) {
                        sh.prefix = prefix | ((uint32_t)(255 - (tid * 8 + k)) << shift);
                        sh.need = need - run;
                        sh.n_equal = loc[k];
                    }
                    run += loc[k];
                }
            }
        }
        __syncthreads();
        prefix = sh.prefix;
        need = sh.need;
        known |= 255u << shift;
        // (the next pass's hist reset is ordered after these reads by its own __syncthreads)
        __syncthreads();
    }
}

// Calls emit(i) for exactly the j selected rows (unordered, except that threshold ties are resolved towards
// lower row indices).  Block-wide; all threads must call.
template <bool SMALLEST, bool HAS_MASK, typename Col, typename Emit>
__device__ void select_rows(const Col& col, const uint8_t* __restrict__ mk, int n, int n_kept, int j, SelShared& sh,
                            Emit emit) {
    const int tid = threadIdx.x;
    if (j <= 0) return;
    if (j >= n_kept) {
        for (int i = tid; i < n; i += SEL_THREADS)
            if (!HAS_MASK || mk[i]) emit(i);
        return;
    }
    radix_threshold<SMALLEST, HAS_MASK>(col, mk, n, j, sh);
    const uint32_t thr = sh.prefix;
    const uint32_t need = sh.need;
    const bool all_equal_taken = (sh.n_equal == need);
    for (int i = tid; i < n; i += SEL_THREADS) {
        if (HAS_MASK && !mk[i]) continue;
        const uint32_t u = sel_key<SMALLEST>(col.at(i));
        if (u > thr || (all_equal_taken && u == thr)) emit(i);
    }
    if (all_equal_taken) return;
    // more rows sit exactly at the threshold than are needed: take the first `need` in row order
    if (tid == 0) sh.taken = 0;
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    for (int base = 0; base < n; base += SEL_THREADS) {
        const int i = base + tid;
        bool eq = false;
        if (i < n && (!HAS_MASK || mk[i])) eq = sel_key<SMALLEST>(col.at(i)) == thr;
        const unsigned int bal = __ballot_sync(FULL, eq);
        if (lane == 0) sh.warp_tot[warp] = __popc(bal);
        __syncthreads();
        unsigned int before = sh.taken, total = 0;
        for (int w = 0; w < SEL_WARPS; ++w) {
            const unsigned int t = sh.warp_tot[w];
            if (w < warp) before += t;
            total += t;
        }
        const unsigned int rank = before + __popc(bal & ((1u << lane) - 1u));
        if (eq && rank < need) emit(i);
        __syncthreads();
        if (tid == 0) sh.taken += total;
        __syncthreads();
        if (sh.taken >= need) break;
    }
}

// ---- fast path of the batched selection: 12-bit first digit + in-shared-memory candidates -----------------------
constexpr int FS_BITS = 12;
constexpr int FS_BINS = 1 << FS_BITS;       // 4096
constexpr int FS_CAND = 4096;               // candidates (keys inside the threshold bin) kept on chip
constexpr int FS_TIES = 512;                // threshold-value ties resolved on chip
struct FastShared {
    unsigned int hist[FS_BINS];
    uint2 cand[FS_CAND];                    // (ordered key, row)
};

// f(i, u) for every kept key of a contiguous column; 16-byte loads, two in flight per thread.
template <bool SMALLEST, bool HAS_MASK, typename Col, typename F>
__device__ __forceinline__ void scan_column(const Col& col, const uint8_t* __restrict__ mk, int n, F f) {
    const int tid = threadIdx.x;
    auto visit = [&](int i, float x) {
        if (HAS_MASK && !mk[i]) return;
        f(i, sel_key<SMALLEST>(x));
    };
    if (!col.vec()) {
        for (int i = tid; i < n; i += SEL_THREADS) visit(i, col.at(i));
        return;
    }
    int head = col.head();
    if (head > n) head = n;
    if (tid < head) visit(tid, col.at(tid));
    const int n4 = (n - head) >> 2;
    int k = tid;
    for (; k + SEL_THREADS < n4; k += 2 * SEL_THREADS) {
        const float4 a = col.at4(head, k), b = col.at4(head, k + SEL_THREADS);
        const int ia = head + 4 * k, ib = ia + 4 * SEL_THREADS;
        visit(ia, a.x); visit(ia + 1, a.y); visit(ia + 2, a.z); visit(ia + 3, a.w);
        visit(ib, b.x); visit(ib + 1, b.y); visit(ib + 2, b.z); visit(ib + 3, b.w);
    }
    if (k < n4) {
        const float4 a = col.at4(head, k);
        const int ia = head + 4 * k;
        visit(ia, a.x); visit(ia + 1, a.y); visit(ia + 2, a.z); visit(ia + 3, a.w);
    }
    const int t0 = head + 4 * n4;
    if (t0 + tid < n) visit(t0 + tid, col.at(t0 + tid));
}

// Block-wide: the bin (counting from the top) where the running count reaches `need`.  Leaves sh.prefix = bin,
// sh.need = rank inside the bin, sh.n_equal = population of the bin.  NB a multiple of SEL_THREADS.
template <int NB>
__device__ void find_bin_desc(const unsigned int* hist, unsigned int need, SelShared& sh) {
    constexpr int PER = NB / SEL_THREADS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned int loc[PER], s = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        loc[k] = hist[NB - 1 - (tid * PER + k)];
        s += loc[k];
    }
    unsigned int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sh.warp_tot[warp] = incl;
    __syncthreads();
    unsigned int before = 0;
    for (int w = 0; w < warp; ++w) before += sh.warp_tot[w];
    incl += before;
    unsigned int run = incl - s;
    if (run < need && need <= incl) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            if (run < need && need <= run + loc[k]) {
                sh.prefix = (unsigned int)(NB - 1 - (tid * PER + k));
                sh.need = need - run;
                sh.n_equal = loc[k];
            }
            run += loc[k];
        }
    }
    __syncthreads();
}

// Marks exactly the j selected rows of a contiguous column (1 <= j < kept rows).
//   scan 1: min / max of the ordered keys;
//   scan 2: FS_BINS-bin histogram of (u - min) >> shift, the shift chosen so the occupied range fills the bins
//           (narrow-range columns such as a 2-class softmax spread out instead of piling into a few bins) -
//           repeated on the threshold bin's own sub-range while that bin still holds more than FS_CAND keys;
//   scan 3: rows above the threshold bin are marked, the keys inside it are parked in shared memory, where
//           the remaining <= 20 bits are resolved with 10-bit digits and ties go to the lowest row indices.
// Returns false (block-uniform) for columns too degenerate for the on-chip lists (e.g. thousands of equal keys
// at the threshold); the caller then runs the generic path, which is correct on its own (marking is idempotent).
template <bool SMALLEST, bool HAS_MASK, typename Col, typename Emit>
__device__ bool select_rows_fast(const Col& v, const uint8_t* __restrict__ mk, int n, int j,
                                 SelShared& sh, FastShared& fs, Emit emit) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- scan 1: range of the kept keys
    uint32_t lo = 0xffffffffu, hi = 0u;
    scan_column<SMALLEST, HAS_MASK>(v, mk, n, [&](int, uint32_t u) { lo = min(lo, u); hi = max(hi, u); });
    lo = __reduce_min_sync(FULL, lo);
    hi = __reduce_max_sync(FULL, hi);
    if (lane == 0) { fs.hist[warp] = lo; fs.hist[SEL_WARPS + warp] = hi; }
    __syncthreads();
    for (int w = 0; w < SEL_WARPS; ++w) { lo = min(lo, fs.hist[w]); hi = max(hi, fs.hist[SEL_WARPS + w]); }
    __syncthreads();
    // ---- scan 2 (repeated while the threshold bin is over-full): histogram of the interval [base, base + 2^bits)
    uint32_t base = lo;
    int bits = 32 - __clz(hi - lo);            // keys satisfy u - base < 2^bits   (bits == 0: all keys equal)
    unsigned int need = (unsigned int)j;
    int shift;
    uint32_t bin0;
    unsigned int n_cand;
    for (int level = 0;; ++level) {
        shift = bits > FS_BITS ? bits - FS_BITS : 0;
        for (int b = tid; b < FS_BINS; b += SEL_THREADS) fs.hist[b] = 0;
        __syncthreads();
        const uint32_t span_m1 = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
        scan_column<SMALLEST, HAS_MASK>(v, mk, n, [&](int, uint32_t u) {
            const uint32_t rel = u - base;
            if (u >= base && rel <= span_m1) atomicAdd(&fs.hist[rel >> shift], 1u);
        });
        __syncthreads();
        find_bin_desc<FS_BINS>(fs.hist, need, sh);
        bin0 = sh.prefix;
        n_cand = sh.n_equal;
        need = sh.need;
        __syncthreads();
        if (n_cand <= (unsigned int)FS_CAND) break;
        if (shift == 0 || level == 2) return false;   // thousands of identical keys at the threshold
        base += bin0 << shift;
        bits = shift;
    }
    // ---- scan 3: mark above the threshold bin, park the bin's keys on chip
    const uint32_t cand_lo = base + (bin0 << shift);                  // first key value of the threshold bin
    const uint32_t cand_span_m1 = shift == 0 ? 0u : ((1u << shift) - 1u);
    if (tid == 0) sh.taken = 0;
    __syncthreads();
    scan_column<SMALLEST, HAS_MASK>(v, mk, n, [&](int i, uint32_t u) {
        if (u < cand_lo) return;
        const uint32_t rel = u - cand_lo;
        if (rel > cand_span_m1) emit(i);
        else fs.cand[atomicAdd(&sh.taken, 1u)] = make_uint2(rel, (unsigned int)i);
    });
    __syncthreads();
    if (need == n_cand) {
        for (unsigned int c = tid; c < n_cand; c += SEL_THREADS) emit((int)fs.cand[c].y);
        return true;
    }
    // ---- the remaining `shift` (<= 20) bits of the candidates, 10 at a time
    uint32_t prefix = 0, known = 0;
#pragma unroll 1
    for (int sh10 = 10; sh10 >= 0; sh10 -= 10) {
        for (int b = tid; b < 1024; b += SEL_THREADS) fs.hist[b] = 0;
        __syncthreads();
        for (unsigned int c = tid; c < n_cand; c += SEL_THREADS) {
            const uint32_t u = fs.cand[c].x;
            if ((u & known) == prefix) atomicAdd(&fs.hist[(u >> sh10) & 1023u], 1u);
        }
        __syncthreads();
        find_bin_desc<1024>(fs.hist, need, sh);
        prefix |= sh.prefix << sh10;
        known |= 1023u << sh10;
        need = sh.need;
        const unsigned int n_eq = sh.n_equal;
        __syncthreads();
        if (sh10 == 0) {
            if (need != n_eq && n_eq > (unsigned int)FS_TIES) return false;
            for (unsigned int c = tid; c < n_cand; c += SEL_THREADS) {
                const uint2 e = fs.cand[c];
                if (e.x > prefix) {
                    emit((int)e.y);
                } else if (e.x == prefix) {
                    bool take = need == n_eq;
                    if (!take) {  // more rows at the threshold value than needed: lowest row indices first
                        unsigned int before = 0;
                        for (unsigned int d = 0; d < n_cand; ++d) before += (fs.cand[d].x == prefix && fs.cand[d].y < e.y);
                        take = before < need;
                    }
                    if (take) emit((int)e.y);
                }
            }
        }
    }
    return true;
}

// Expected size of the candidate set the sampled path aims for, and whether a column qualifies for it.
#ifndef MOC_SEL_TARGET_MUL
#define MOC_SEL_TARGET_MUL 3
#define MOC_SEL_TARGET_ADD 400
#endif
__device__ __forceinline__ int sampled_target(int j) { return MOC_SEL_TARGET_MUL * j + MOC_SEL_TARGET_ADD; }
__device__ __forceinline__ bool sampled_applies(int n, int j) {
    return n >= 8192 && sampled_target(j) <= 2560 && 4 * sampled_target(j) <= n;
}

// f(i, u) for every key of a contiguous unmasked column; 16-byte loads, FOUR in flight per thread (the column's first
// touch comes from DRAM: with two, a CTA in its scan has 16 KB in flight, and the three CTAs of an SM are in a scan only
// about half of the time).
template <bool SMALLEST, typename Col, typename F>
__device__ __forceinline__ void scan_column4(const Col& col, int n, F f) {
    constexpr int DEPTH = Col::kScanDepth;      // 16-byte pieces in flight per thread: four loads either way
    const int tid = threadIdx.x;
    auto visit = [&](int i, float x) { f(i, sel_key<SMALLEST>(x)); };
    if (!col.vec()) {
        for (int i = tid; i < n; i += SEL_THREADS) visit(i, col.at(i));
        return;
    }
    int head = col.head();
    if (head > n) head = n;
    if (tid < head) visit(tid, col.at(tid));
    const int n4 = (n - head) >> 2;
    int k = tid;
    for (; k + (DEPTH - 1) * SEL_THREADS < n4; k += DEPTH * SEL_THREADS) {
        float4 a[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) a[d] = col.at4(head, k + d * SEL_THREADS);
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const int i = head + 4 * (k + d * SEL_THREADS);
            visit(i, a[d].x); visit(i + 1, a[d].y); visit(i + 2, a[d].z); visit(i + 3, a[d].w);
        }
    }
    for (; k < n4; k += SEL_THREADS) {
        const float4 a = col.at4(head, k);
        const int ia = head + 4 * k;
        visit(ia, a.x); visit(ia + 1, a.y); visit(ia + 2, a.z); visit(ia + 3, a.w);
    }
    const int t0 = head + 4 * n4;
    if (t0 + tid < n) visit(t0 + tid, col.at(t0 + tid));
}

constexpr int FS_SMALL = 64;    // keys of the threshold bin ranked by brute force

// ONE scan instead of three, for long unmasked columns and j << n (the evaluation pass: j = 400 of 20 000 - 100 000
// patches, 2C + 2 selections per slide).  A CTA of the three-scan path spends its time in DRAM / L2 latency - three
// dependent scans - and in a dozen block-wide histogram rounds; this path has one scan and two rounds:
//   1. sample one 32-byte sector (8 keys) out of every `stride` keys, at most 4096 keys, into shared memory; one
//      4096-bin histogram of the sample's own range gives the bin that holds the sample's key of rank r, r chosen so
//      that about 3j + 400 keys of the whole column lie above that bin's lower edge t0 (independent draws would put
//      the real count within +-10 % of it; sectors of neighbouring patches are correlated, hence the wide margins:
//      anything in [j, 4096] is accepted);
//   2. the scan parks every key >= t0 (with its row) in shared memory;
//   3. if between j and FS_CAND keys were parked, the column's top j are the top j of those: one 4096-bin histogram
//      of [t0, max] finds the threshold bin, keys above it are taken, and its (usually one to three) keys are ranked
//      by brute force - ties at equal value to the lowest row indices, exactly like the other paths.
// Returns false (block-uniform) with nothing emitted when the parked count falls outside [j, FS_CAND] or the threshold
// bin is crowded with equal keys; the caller then runs the three-scan path.
template <bool SMALLEST, typename Col, typename Emit>
__device__ bool select_rows_sampled(const Col& v, int n, int j, SelShared& sh, FastShared& fs, Emit emit) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- 1. sample, its range, one histogram round
    int stride = 128;
    if (n > 65536) stride = ((n + 511) / 512 + 7) & ~7;
    const unsigned int m = (unsigned int)(n / stride) * 8u;             // <= 4096
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (unsigned int c = tid; c < m; c += SEL_THREADS) {
        const uint32_t u = sel_key<SMALLEST>(v.at((int)((c >> 3) * stride + (c & 7u))));
        fs.cand[c].x = u;
        lo = min(lo, u);
        hi = max(hi, u);
    }
    for (int b = tid; b < FS_BINS; b += SEL_THREADS) fs.hist[b] = 0;
    lo = __reduce_min_sync(FULL, lo);
    hi = __reduce_max_sync(FULL, hi);
    if (lane == 0) { sh.hist[warp] = lo; sh.hist[SEL_WARPS + warp] = hi; }
    if (tid == 0) { sh.taken = 0; sh.n_kept = 0; }
    __syncthreads();
    for (int w = 0; w < SEL_WARPS; ++w) { lo = min(lo, sh.hist[w]); hi = max(hi, sh.hist[SEL_WARPS + w]); }
    int bits = 32 - __clz(hi - lo);
    int shift = bits > FS_BITS ? bits - FS_BITS : 0;
    for (unsigned int c = tid; c < m; c += SEL_THREADS) atomicAdd(&fs.hist[(fs.cand[c].x - lo) >> shift], 1u);
    __syncthreads();
    unsigned int r = (unsigned int)(((int64_t)sampled_target(j) * m + n - 1) / n);
    r = r < 1u ? 1u : (r > m ? m : r);
    find_bin_desc<FS_BINS>(fs.hist, r, sh);
    const uint32_t t0 = lo + (sh.prefix << shift);
    __syncthreads();
    for (int b = tid; b < FS_BINS; b += SEL_THREADS) fs.hist[b] = 0;    // for round 3; nobody reads it before the next barrier
    // ---- 2. the scan
    uint32_t mx = 0u;
    scan_column4<SMALLEST>(v, n, [&](int i, uint32_t u) {
        if (u >= t0) {
            mx = max(mx, u);
            const unsigned int slot = atomicAdd(&sh.taken, 1u);
            if (slot < (unsigned int)FS_CAND) fs.cand[slot] = make_uint2(u, (unsigned int)i);
        }
    });
    mx = __reduce_max_sync(FULL, mx);
    if (lane == 0 && mx) atomicMax(&sh.n_kept, mx);
    __syncthreads();
    const unsigned int n_cand = sh.taken;
    const uint32_t top = sh.n_kept;
    if (n_cand < (unsigned int)j || n_cand > (unsigned int)FS_CAND) {
        __syncthreads();
        return false;
    }
    // ---- 3. threshold bin of [t0, top]
    bits = 32 - __clz(top - t0);
    shift = bits > FS_BITS ? bits - FS_BITS : 0;
    for (unsigned int c = tid; c < n_cand; c += SEL_THREADS) atomicAdd(&fs.hist[(fs.cand[c].x - t0) >> shift], 1u);
    __syncthreads();
    find_bin_desc<FS_BINS>(fs.hist, (unsigned int)j, sh);
    const uint32_t bin_lo = t0 + (sh.prefix << shift);
    const uint32_t bin_hi = bin_lo + (shift == 0 ? 0u : ((1u << shift) - 1u));      // inclusive
    const unsigned int need = sh.need, n_bin = sh.n_equal;
    __syncthreads();
    if (need == n_bin) {            // the whole threshold bin belongs to the top j
        for (unsigned int c = tid; c < n_cand; c += SEL_THREADS)
            if (fs.cand[c].x >= bin_lo) emit((int)fs.cand[c].y);
        return true;
    }
    if (n_bin > (unsigned int)FS_SMALL) return false;
    uint2* small = reinterpret_cast<uint2*>(fs.hist);                               // the histogram is done with
    if (tid == 0) sh.taken = 0;
    __syncthreads();
    for (unsigned int c = tid; c < n_cand; c += SEL_THREADS) {
        const uint2 e = fs.cand[c];
        if (e.x > bin_hi) emit((int)e.y);
        else if (e.x >= bin_lo) small[atomicAdd(&sh.taken, 1u)] = e;
    }
    __syncthreads();
    if ((unsigned int)tid < n_bin) {
        const uint2 e = small[tid];
        unsigned int before = 0;
        for (unsigned int d = 0; d < n_bin; ++d)
            before += (small[d].x > e.x) || (small[d].x == e.x && small[d].y < e.y);
        if (before < need) emit((int)e.y);
    }
    return true;
}

template <bool HAS_MASK>
__device__ int count_kept(const uint8_t* __restrict__ mk, int n, SelShared& sh) {
    if (!HAS_MASK) return n;
    if (threadIdx.x == 0) sh.n_kept = 0;
    __syncthreads();
    unsigned int c = 0;
    for (int i = threadIdx.x; i < n; i += SEL_THREADS) c += mk[i] != 0;
    c = (unsigned int)__reduce_add_sync(FULL, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sh.n_kept, c);
    __syncthreads();
    return (int)sh.n_kept;
}

// One selection of one slide, whatever its column type: the one-scan path for long unmasked columns, else the
// three-scan path, else (degenerate columns) the generic four-pass radix select.
template <bool HAS_MASK, typename Col, typename Emit>
__device__ __forceinline__ void select_one(const Col& col, const uint8_t* __restrict__ mk, int n, int n_kept, int j,
                                           bool smallest, bool sampled, SelShared& sh, FastShared& fs, Emit mark) {
    if (j > 0 && j < n_kept) {
        if (!HAS_MASK && sampled && sampled_applies(n, j)) {
            const bool done = smallest ? select_rows_sampled<true>(col, n, j, sh, fs, mark)
                                       : select_rows_sampled<false>(col, n, j, sh, fs, mark);
            if (done) return;
            __syncthreads();
        }
        const bool done = smallest ? select_rows_fast<true, HAS_MASK>(col, mk, n, j, sh, fs, mark)
                                   : select_rows_fast<false, HAS_MASK>(col, mk, n, j, sh, fs, mark);
        if (done) return;
        __syncthreads();
    }
    if (smallest) select_rows<true, HAS_MASK>(col, mk, n, n_kept, j, sh, mark);
    else select_rows<false, HAS_MASK>(col, mk, n, n_kept, j, sh, mark);
}

// grid (2C+2, n_slides): one selection of one slide per CTA; marks the chosen rows in the global bitmap.  The CTAs of a
// slide are neighbours in the grid, so they run at about the same time: what several of them read - the row mask, and
// in the compact key layout (COMPACT; C >= MOC_KEYS_COMPACT_MIN_CLASSES) the lse plane every softmax selection needs
// and the L_c plane the top-J and the softmax selection of class c share - comes from DRAM once.
#ifndef MOC_SEL_COMPACT_CTAS
#define MOC_SEL_COMPACT_CTAS 3
#endif
template <bool HAS_MASK, bool COMPACT>
__global__ void __launch_bounds__(SEL_THREADS, COMPACT ? MOC_SEL_COMPACT_CTAS : 3)   // 40 registers: three CTAs per SM (two: -20 %; four: worse at C = 30)
select_mark_kernel(const float* __restrict__ keys, int64_t key_stride, const int64_t* __restrict__ offsets, int C,
                   int topj, unsigned discard_mask, const uint8_t* __restrict__ row_mask,
                   unsigned int* __restrict__ bitmap, int sampled) {
    extern __shared__ __align__(16) unsigned char sel_dyn_smem[];
    FastShared& fs = *reinterpret_cast<FastShared*>(sel_dyn_smem);
    __shared__ SelShared sh;
    // q: 2c = top-J of class c, 2c + 1 = softmax selection of class c (neighbours: they share L_c), 2C = |top1 - top2|,
    // 2C + 1 = bottom-J of the background sum
    const int q = blockIdx.x, slide = blockIdx.y;
    const KeyLayout kl = key_layout(C);
    unsigned cls;
    int plane;
    bool smallest = false, softmax_col = false;
    if (q < 2 * C) {
        const int c = q >> 1;
        if ((q & 1) == 0) { cls = MOC_CLS_TOPK; plane = c; }
        else { cls = MOC_CLS_DELTA_SOFTMAX; plane = COMPACT ? c : kl.softmax0 + c; softmax_col = COMPACT; }
    } else if (q == 2 * C) { cls = MOC_CLS_DELTA_DIFF; plane = kl.diff; }
    else { cls = MOC_CLS_BOTTOMK; plane = kl.bg_sum; smallest = true; }
    if (discard_mask & cls) return;
    const int64_t row0 = offsets[slide];
    const int n = (int)(offsets[slide + 1] - row0);
    if (n <= 0) return;
    const uint8_t* mk = HAS_MASK ? row_mask + row0 : nullptr;
    const int n_kept = count_kept<HAS_MASK>(mk, n, sh);
    const int j = topj < n_kept ? topj : n_kept;
    auto mark = [&](int i) {
        const int64_t r = row0 + i;
        atomicOr(&bitmap[r >> 5], 1u << (r & 31));
    };
    const float* v = keys + (int64_t)plane * key_stride + row0;
    if (COMPACT && softmax_col) {
        const ColLogSoftmax col = {v, keys + (int64_t)kl.lse * key_stride + row0};
        select_one<HAS_MASK>(col, mk, n, n_kept, j, smallest, sampled != 0, sh, fs, mark);
    } else {
        const ColPlain col = {v, 1};
        select_one<HAS_MASK>(col, mk, n, n_kept, j, smallest, sampled != 0, sh, fs, mark);
    }
}

// grid n_slides: bitmap -> ascending row list (+ index inside the masked bag), count per slide.
constexpr int CMP_THREADS = 256;
template <bool HAS_MASK>
__global__ void __launch_bounds__(CMP_THREADS)
compact_kernel(const unsigned int* __restrict__ bitmap, const int64_t* __restrict__ offsets,
               const uint8_t* __restrict__ row_mask, const int64_t* __restrict__ sel_base,
               int32_t* __restrict__ sel_rows, int32_t* __restrict__ sel_local, int32_t* __restrict__ sel_count) {
    __shared__ unsigned long long warp_tot[CMP_THREADS / 32];
    __shared__ unsigned long long running;
    const int slide = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = offsets[slide], row1 = offsets[slide + 1];
    if (tid == 0) running = 0;
    __syncthreads();
    if (row1 > row0) {
        const int64_t w0 = row0 >> 5, w1 = (row1 - 1) >> 5;
        const int64_t out0 = sel_base[slide], out_end = sel_base[slide + 1];
        for (int64_t wb = w0; wb <= w1; wb += CMP_THREADS) {
            const int64_t w = wb + tid;
            unsigned int bits = 0, kept = 0;
            if (w <= w1) {
                unsigned int valid = 0xffffffffu;
                const int64_t first = w << 5;
                if (first < row0) valid &= 0xffffffffu << (row0 - first);
                if (first + 32 > row1) valid &= 0xffffffffu >> (first + 32 - row1);
                bits = bitmap[w] & valid;
                if (HAS_MASK) {
                    for (int b = 0; b < 32; ++b)
                        if ((valid >> b) & 1u) kept |= (row_mask[first + b] != 0 ? 1u : 0u) << b;
                } else {
                    kept = valid;
                }
            }
            // one scan for both counts: high word = kept rows, low word = selected rows
            const unsigned long long mine = ((unsigned long long)__popc(kept) << 32) | (unsigned long long)__popc(bits);
            unsigned long long incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            unsigned long long before = running, total = 0;
            for (int k = 0; k < CMP_THREADS / 32; ++k) {
                const unsigned long long t = warp_tot[k];
                if (k < warp) before += t;
                total += t;
            }
            const unsigned long long excl = before + incl - mine;
            int64_t pos = out0 + (int64_t)(excl & 0xffffffffull);
            const int kept_before = (int)(excl >> 32);
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                if (pos < out_end) {   // a region smaller than the selection (a caller's stale layout) must not overrun
                    sel_rows[pos] = (int32_t)((w << 5) + b);
                    sel_local[pos] = kept_before + __popc(kept & ((1u << b) - 1u));
                }
                ++pos;
            }
            __syncthreads();
            if (tid == 0) running += total;
            __syncthreads();
        }
    }
    __syncthreads();
    int64_t count = (int64_t)(running & 0xffffffffull);
    if (count > sel_base[slide + 1] - sel_base[slide]) count = sel_base[slide + 1] - sel_base[slide];
    if (tid == 0) sel_count[slide] = (int32_t)count;
    // unused tail of this slide's region: -1, so consumers can walk regions without the counts
    for (int64_t p = sel_base[slide] + count + tid; p < sel_base[slide + 1]; p += CMP_THREADS) sel_rows[p] = -1;
}

// ---- stand-alone sorted top-J ------------------------------------------------------------------
// grid n_cols; dynamic smem: P uint64 sort keys.
__global__ void __launch_bounds__(SEL_THREADS)
topj_sorted_kernel(const float* __restrict__ values, int n, int64_t ld, int64_t col_stride, int j, int largest,
                   int sort_pow2, int64_t* __restrict__ idx_out, int64_t out_ld, float* __restrict__ val_out) {
    extern __shared__ unsigned long long skeys[];
    __shared__ SelShared sh;
    __shared__ unsigned int n_out;
    const int col = blockIdx.x, tid = threadIdx.x;
    const float* v = values + (int64_t)col * col_stride;
    if (tid == 0) n_out = 0;
    for (int i = tid; i < sort_pow2; i += SEL_THREADS) skeys[i] = 0ull;
    __syncthreads();
    const bool small = !largest;
    auto push = [&](int i) {
        const uint32_t u = small ? ~f2ord(v[(int64_t)i * ld]) : f2ord(v[(int64_t)i * ld]);
        const unsigned int slot = atomicAdd(&n_out, 1u);
        // descending sort of (key, ~index): larger value first, lower index first among equals
        skeys[slot] = (unsigned long long)u << 32 | (unsigned long long)(0xffffffffu - (uint32_t)i);
    };
    const ColPlain column = {v, ld};
    if (small) select_rows<true, false>(column, nullptr, n, n, j, sh, push);
    else select_rows<false, false>(column, nullptr, n, n, j, sh, push);
    __syncthreads();
    // bitonic sort, descending
    for (int k = 2; k <= sort_pow2; k <<= 1) {
        for (int s = k >> 1; s > 0; s >>= 1) {
            for (int i = tid; i < sort_pow2; i += SEL_THREADS) {
                const int p = i ^ s;
                if (p > i) {
                    const unsigned long long a = skeys[i], b = skeys[p];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { skeys[i] = b; skeys[p] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int r = tid; r < j; r += SEL_THREADS) {
        const unsigned long long kk = skeys[r];
        const uint32_t i = 0xffffffffu - (uint32_t)(kk & 0xffffffffull);
        idx_out[(int64_t)r * out_ld + col] = (int64_t)i;
        if (val_out) val_out[(int64_t)r * out_ld + col] = v[(int64_t)i * ld];
    }
}

// ---- pooled top-K over whole bags (zero-shot path) ---------------------------------------------
// grid (n_slides, C): rows chosen by one key plane, the mean taken over another.
constexpr int POOL_MAX_K = 64;
__global__ void __launch_bounds__(SEL_THREADS)
pool_topk_kernel(const float* __restrict__ keys, int64_t key_stride, const int64_t* __restrict__ offsets, int C,
                 int topk, int sel_plane0, int sel_step, int sel_smallest, int val_plane0, int val_step,
                 float* __restrict__ bag_logits) {
    __shared__ SelShared sh;
    __shared__ unsigned int n_out;
    __shared__ unsigned long long win[POOL_MAX_K];
    const int c = blockIdx.y, slide = blockIdx.x, tid = threadIdx.x;
    const int64_t row0 = offsets[slide];
    const int n = (int)(offsets[slide + 1] - row0);
    const float* sv = keys + (int64_t)(sel_plane0 + c * sel_step) * key_stride + row0;
    const float* vv = keys + (int64_t)(val_plane0 + c * val_step) * key_stride + row0;
    if (tid == 0) n_out = 0;
    __syncthreads();
    const int k = topk < n ? topk : n;
    const bool small = sel_smallest != 0;
    auto push = [&](int i) {
        const uint32_t u = small ? ~f2ord(sv[i]) : f2ord(sv[i]);
        const unsigned int slot = atomicAdd(&n_out, 1u);
        win[slot] = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
    };
    const ColPlain col = {sv, 1};
    if (small) select_rows<true, false>(col, nullptr, n, n, k, sh, push);
    else select_rows<false, false>(col, nullptr, n, n, k, sh, push);
    __syncthreads();
    if (tid == 0) {
        // order the k winners (descending key, ascending row) so the fp32 sum has a fixed order
        for (int a = 1; a < k; ++a) {
            const unsigned long long x = win[a];
            int b = a - 1;
            while (b >= 0 && win[b] < x) { win[b + 1] = win[b]; --b; }
            win[b + 1] = x;
        }
        float s = 0.f;
        for (int a = 0; a < k; ++a) s += vv[0xffffffffu - (uint32_t)(win[a] & 0xffffffffull)];
        bag_logits[(int64_t)slide * C + c] = k > 0 ? s / (float)k : 0.f;
    }
}


// ---- helpers for the stand-alone selector / pooling API (inputs are logits [N,Ct] row-major) -------------
// key planes from already-computed logits: columns < n_fg are classes, the rest background.
__global__ void row_keys_kernel(const float* __restrict__ logits, int64_t n, int64_t ld, int n_fg, int n_total,
                                float* __restrict__ keys, int64_t key_stride) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* x = logits + i * ld;
    float m1 = -INFINITY, m2 = -INFINITY;
    for (int c = 0; c < n_fg; ++c) {
        const float v = x[c];
        if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) { m2 = v; }
    }
    float es = 0.f;
    for (int c = 0; c < n_fg; ++c) es += expf(x[c] - m1);
    const float inv = 1.0f / es;
    float* kp = keys + i;
    for (int c = 0; c < n_fg; ++c) {
        kp[(int64_t)c * key_stride] = x[c];
        kp[(int64_t)(n_fg + c) * key_stride] = expf(x[c] - m1) * inv;
    }
    float bs = 0.f, bm = -INFINITY;
    for (int c = n_fg; c < n_total; ++c) { bs += x[c]; bm = fmaxf(bm, x[c]); }
    kp[(int64_t)(2 * n_fg) * key_stride] = fabsf(m1 - m2);
    kp[(int64_t)(2 * n_fg + 1) * key_stride] = bs;
    kp[(int64_t)(2 * n_fg + 2) * key_stride] = bm;
}

// dst[r][c] = src[idx[r]][c], c < n_cols
__global__ void take_rows_kernel(const float* __restrict__ src, int64_t ld, const int64_t* __restrict__ idx,
                                 int64_t n_idx, int n_cols, float* __restrict__ dst) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_idx * n_cols) return;
    const int64_t r = t / n_cols;
    const int c = (int)(t % n_cols);
    dst[t] = src[idx[r] * ld + c];
}

// out[c] = mean(vals[0..j-1][c]) summed in row order (fixed order => reproducible)
__global__ void col_prefix_mean_kernel(const float* __restrict__ vals, int64_t ld, int n_cols, int j,
                                       float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    float s = 0.f;
    for (int r = 0; r < j; ++r) s += vals[(int64_t)r * ld + c];
    out[c] = s / (float)j;
}

}  // namespace moc

using namespace moc;

extern "C" int64_t moc_select_capacity(int64_t n_rows_of_slide, int n_classes, int topj) {
    const int64_t bound = (int64_t)topj * (2 * n_classes + 2);
    return n_rows_of_slide < bound ? n_rows_of_slide : bound;
}

extern "C" size_t moc_select_workspace_bytes(int64_t total_rows, int n_slides) {
    (void)n_slides;
    return (size_t)((total_rows + 31) / 32 + 1) * sizeof(unsigned int);
}

extern "C" int moc_select_union(const float* keys, int64_t key_stride, const int64_t* offsets, int n_slides,
                                int64_t total_rows, int n_classes, int topj, unsigned discard_mask,
                                const uint8_t* row_mask, const int64_t* sel_base, int32_t* sel_rows,
                                int32_t* sel_local, int32_t* sel_count, void* workspace, size_t workspace_bytes,
                                void* stream) {
    MOC_CHECK_ARG(keys && offsets && sel_base && sel_rows && sel_local && sel_count && workspace,
                  "moc_select_union: null pointer");
    MOC_CHECK_ARG(n_slides >= 0 && total_rows >= 0 && topj >= 0, "moc_select_union: negative size");
    MOC_CHECK_SHAPE(n_classes >= 2 && n_classes < MOC_MAX_COLS, "moc_select_union: bad class count %d", n_classes);
    MOC_CHECK_SHAPE(total_rows < (1ll << 31), "moc_select_union: more than 2^31 rows in one store");
    const size_t need = moc_select_workspace_bytes(total_rows, n_slides);
    if (workspace_bytes < need) {
        set_error("moc_select_union: workspace %zu B < required %zu B", workspace_bytes, need);
        return MOC_E_WORKSPACE;
    }
    if (n_slides == 0) return MOC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* bitmap = reinterpret_cast<unsigned int*>(workspace);
    MOC_CUDA(cudaMemsetAsync(bitmap, 0, need, st));
    constexpr int MAX_GRID_Y = 65535;       // slides per launch of the marking kernel (grid.y)
    static int sampled = -1;        // MOC_SELECT_SAMPLED=0: three-scan selection only (developer A/B switch)
    if (sampled < 0) {
        const char* e = getenv("MOC_SELECT_SAMPLED");
        sampled = (e && e[0] == '0') ? 0 : 1;
    }
    const bool compact = key_layout(n_classes).compact;
#define MOC_SEL_LAUNCH(MASKED, COMPACT_)                                                                                  \
    do {                                                                                                                  \
        MOC_CUDA(cudaFuncSetAttribute(select_mark_kernel<MASKED, COMPACT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                      (int)sizeof(FastShared)));                                                          \
        for (int s0 = 0; s0 < n_slides; s0 += MAX_GRID_Y) {                                                               \
            const dim3 grid(2 * n_classes + 2, n_slides - s0 < MAX_GRID_Y ? n_slides - s0 : MAX_GRID_Y);                  \
            select_mark_kernel<MASKED, COMPACT_><<<grid, SEL_THREADS, sizeof(FastShared), st>>>(                          \
                keys, key_stride, offsets + s0, n_classes, topj, discard_mask, row_mask, bitmap, sampled);                \
        }                                                                                                                 \
    } while (0)
    if (row_mask) {
        if (compact) MOC_SEL_LAUNCH(true, true); else MOC_SEL_LAUNCH(true, false);
        MOC_LAUNCH_CHECK("select_mark_kernel");
        compact_kernel<true><<<n_slides, CMP_THREADS, 0, st>>>(bitmap, offsets, row_mask, sel_base, sel_rows, sel_local,
                                                             sel_count);
    } else {
        if (compact) MOC_SEL_LAUNCH(false, true); else MOC_SEL_LAUNCH(false, false);
        MOC_LAUNCH_CHECK("select_mark_kernel");
        compact_kernel<false><<<n_slides, CMP_THREADS, 0, st>>>(bitmap, offsets, nullptr, sel_base, sel_rows, sel_local,
                                                              sel_count);
    }
#undef MOC_SEL_LAUNCH
    MOC_LAUNCH_CHECK("compact_kernel");
    return MOC_OK;
}

extern "C" int moc_topj_sorted(const float* values, int64_t n, int64_t ld, int n_cols, int64_t col_stride, int j,
                               int largest, int64_t* idx_out, int64_t out_ld, float* val_out, void* stream) {
    MOC_CHECK_ARG(values && idx_out, "moc_topj_sorted: null pointer");
    MOC_CHECK_ARG(n >= 0 && n < (1ll << 31) && ld >= 1 && n_cols >= 0 && j >= 0 && out_ld >= n_cols,
                  "moc_topj_sorted: bad sizes");
    if (j > n) j = (int)n;
    if (n_cols == 0 || j == 0) return MOC_OK;
    int p2 = 1;
    while (p2 < j) p2 <<= 1;
    MOC_CHECK_SHAPE(p2 <= 16384, "moc_topj_sorted: topj %d exceeds the 16384 rows this build sorts on chip", j);
    const size_t smem = (size_t)p2 * sizeof(unsigned long long);
    MOC_CUDA(cudaFuncSetAttribute(topj_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    topj_sorted_kernel<<<n_cols, SEL_THREADS, smem, (cudaStream_t)stream>>>(values, (int)n, ld, col_stride, j, largest,
                                                                           p2, idx_out, out_ld, val_out);
    MOC_LAUNCH_CHECK("topj_sorted_kernel");
    return MOC_OK;
}

extern "C" int moc_pool_topk(const float* keys, int64_t key_stride, const int64_t* offsets, int n_slides,
                             int n_classes, int topk, int sel_plane0, int sel_step, int sel_smallest, int val_plane0,
                             int val_step, float* bag_logits, void* stream) {
    MOC_CHECK_ARG(keys && offsets && bag_logits, "moc_pool_topk: null pointer");
    MOC_CHECK_ARG(n_slides >= 0 && n_classes >= 1, "moc_pool_topk: bad sizes");
    MOC_CHECK_SHAPE(topk >= 1 && topk <= POOL_MAX_K, "moc_pool_topk: topk must be in [1,%d], got %d", POOL_MAX_K, topk);
    if (n_slides == 0) return MOC_OK;
    pool_topk_kernel<<<dim3(n_slides, n_classes), SEL_THREADS, 0, (cudaStream_t)stream>>>(
        keys, key_stride, offsets, n_classes, topk, sel_plane0, sel_step, sel_smallest, val_plane0, val_step,
        bag_logits);
    MOC_LAUNCH_CHECK("pool_topk_kernel");
    return MOC_OK;
}

extern "C" int moc_row_keys(const float* logits, int64_t n, int64_t ld, int n_fg, int n_total, float* keys,
                            int64_t key_stride, void* stream) {
    MOC_CHECK_ARG(logits && keys && n >= 0 && ld >= n_total && key_stride >= n, "moc_row_keys: bad arguments");
    MOC_CHECK_SHAPE(n_fg >= 1 && n_total >= n_fg, "moc_row_keys: bad column split %d/%d", n_fg, n_total);
    if (n == 0) return MOC_OK;
    row_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(logits, n, ld, n_fg, n_total, keys,
                                                                                  key_stride);
    MOC_LAUNCH_CHECK("row_keys_kernel");
    return MOC_OK;
}

extern "C" int moc_take_rows(const float* src, int64_t ld, const int64_t* idx, int64_t n_idx, int n_cols, float* dst,
                             void* stream) {
    MOC_CHECK_ARG(src && idx && dst && n_idx >= 0 && n_cols >= 0 && ld >= n_cols, "moc_take_rows: bad arguments");
    if (n_idx * n_cols == 0) return MOC_OK;
    take_rows_kernel<<<(unsigned)((n_idx * n_cols + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, ld, idx, n_idx,
                                                                                                n_cols, dst);
    MOC_LAUNCH_CHECK("take_rows_kernel");
    return MOC_OK;
}

extern "C" int moc_col_prefix_mean(const float* vals, int64_t ld, int n_cols, int j, float* out, void* stream) {
    MOC_CHECK_ARG(vals && out && n_cols >= 1 && j >= 1 && ld >= n_cols, "moc_col_prefix_mean: bad arguments");
    col_prefix_mean_kernel<<<(n_cols + 63) / 64, 64, 0, (cudaStream_t)stream>>>(vals, ld, n_cols, j, out);
    MOC_LAUNCH_CHECK("col_prefix_mean_kernel");
    return MOC_OK;
}
