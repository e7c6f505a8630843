// pbsc_walk.cuh — one FM-index walk between two seeds, executed by one warp.
//
// Restates LongReadSelfCorrectByOverlap (PacBio/LongReadCorrectByOverlap.cpp:17-878) as a
// level-synchronous traversal: the warp holds the leaf frontier (<= 32 live leaves, <= 128 new
// leaves per level) in a per-warp scratch area, lanes own leaves / (leaf, base, strand) probes /
// (leaf, strand) re-searches, and survivors are compacted in list order with ballots.  All double
// arithmetic uses the explicit round-to-nearest intrinsics in the reference's source order (no FMA).
//
// Interval-tree queries (PacBio/IntervalTree.cpp:66-86) are replaced by what they compute on this
// path: every stored interval is the SA interval of one idmer/5-mer of the query, and a leaf interval
// lies inside a stored one iff the leaf string ends with that k-mer.  So a query is "positions of the
// query whose k-mer equals the leaf's last k bases", returned in the order the reference's
// std::sort(greater-by-start) leaves equal keys in (stl_sort_emul.cuh).
#ifndef PBSC_WALK_CUH
#define PBSC_WALK_CUH

#include "fm_table.cuh"
#include "stl_sort_emul.cuh"
#include "pbsc_status.h"

namespace pbsc {

constexpr int OLD_CAP = 32;      // -l / --max-leaves upper bound of this build
constexpr int NEW_CAP = 128;     // 4 children per live leaf
constexpr int RING_SLOTS = 160;
constexpr int RING_LEN = 100;    // m_localSimilarlykmerSize
constexpr int RES_CAP = 256;
constexpr int TERM_CAP = 128;
constexpr unsigned FULL = 0xffffffffu;

struct __align__(16) Leaf
{
    uint64_t f_lo, f_hi, r_lo, r_hi;   // fwdInterval (RBWT) / rvcInterval (BWT), half-open
    double redeem;                     // numRedeemSeed
    double local_err;                  // LocalErrorRateRecord.back()
    double global_err;                 // GlobalErrorRateRecord.back()
    uint64_t rt_hi, rt_lo;             // last 64 bases, newest base in the top 2 bits of rt_hi
    uint32_t lastSeedIdx, lastOverlapLen, totalSeeds, tailCount;
    int32_t seedOff;                   // lastSeedIdxOffset
    int32_t res_first, res_second;     // resultindex
    int32_t kmerFreq;                  // leafInfo::kmerFrequency
    uint32_t node;                     // id in the label tree
    uint16_t ring;                     // slot of the 100-deep GlobalErrorRateRecord window
    uint8_t tailLetter, alive;
    uint32_t aux;                      // thread engine: accepted bases of a parent / terminal index + 1 of a child (pbsc_walk_thread.cuh)
};
static_assert(sizeof(Leaf) == 128, "Leaf must be 128 bytes");

struct ProbeIv { uint64_t f_lo, f_hi, r_lo, r_hi; };
struct WalkResult { double err; uint32_t node; int32_t i; uint32_t depth; uint32_t pad; };

struct ExtParamsDev
{
    int32_t max_leaves, seed_size, min_overlap, pb_coverage, high_freq_thr;
    double redeem_a;          // (m_seedSize-1)*m_PacBioErrorRate
    double redeem_b;          // 1-m_PacBioErrorRate
    double walk_error_rate;   // m_errorRate = 0.25
    int32_t freq_int[101];    // (int)freqsOfKmerSize[k]
    // per-warp scratch capacities
    uint32_t q_cap, node_cap, merged_cap;
};

// per-warp scratch (global memory) carved from one allocation
struct WarpScratch
{
    Leaf* oldL; Leaf* newL;
    ProbeIv* probes;
    double* rings;
    uint32_t* nodes;
    WalkResult* res;
    Interval* termF; Interval* termR;
    uint8_t* q;
    uint64_t* sF; uint64_t* sR;
    uint16_t* c5;
    uint8_t* merged;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline size_t warp_scratch_bytes(uint32_t q_cap, uint32_t node_cap, uint32_t merged_cap)
{
    size_t b = 0;
    b += sizeof(Leaf) * (OLD_CAP + NEW_CAP);
    b += sizeof(ProbeIv) * NEW_CAP;
    b += sizeof(double) * RING_SLOTS * RING_LEN;
    b += align_up(sizeof(uint32_t) * (size_t)node_cap, 16);
    b += sizeof(WalkResult) * RES_CAP;
    b += sizeof(Interval) * TERM_CAP * 2;
    b += align_up(q_cap, 16);
    b += sizeof(uint64_t) * (size_t)q_cap * 2;
    b += align_up(sizeof(uint16_t) * (size_t)q_cap, 16);
    b += align_up(merged_cap, 16);
    return align_up(b, 128);
}

__device__ inline void carve_scratch(uint8_t* base, const ExtParamsDev& P, WarpScratch& w)
{
    uint8_t* p = base;
    w.oldL = (Leaf*)p; p += sizeof(Leaf) * OLD_CAP;
    w.newL = (Leaf*)p; p += sizeof(Leaf) * NEW_CAP;
    w.probes = (ProbeIv*)p; p += sizeof(ProbeIv) * NEW_CAP;
    w.rings = (double*)p; p += sizeof(double) * RING_SLOTS * RING_LEN;
    w.nodes = (uint32_t*)p; p += align_up(sizeof(uint32_t) * (size_t)P.node_cap, 16);
    w.res = (WalkResult*)p; p += sizeof(WalkResult) * RES_CAP;
    w.termF = (Interval*)p; p += sizeof(Interval) * TERM_CAP;
    w.termR = (Interval*)p; p += sizeof(Interval) * TERM_CAP;
    w.q = p; p += align_up(P.q_cap, 16);
    w.sF = (uint64_t*)p; p += sizeof(uint64_t) * (size_t)P.q_cap;
    w.sR = (uint64_t*)p; p += sizeof(uint64_t) * (size_t)P.q_cap;
    w.c5 = (uint16_t*)p; p += align_up(sizeof(uint16_t) * (size_t)P.q_cap, 16);
    w.merged = p;
}

// per-warp shared memory
struct WarpShared
{
    uint16_t win5[1024];      // occurrences of each 5-mer of the query inside the current +-maxIndel window
    int32_t pfreq[NEW_CAP];   // kmerFrequency of probe (leaf, base)
    uint32_t ptotal[OLD_CAP]; // totalcount of a leaf's four probes
    int32_t pmax[OLD_CAP];    // maxfreqsofleave
    uint8_t pmatch[OLD_CAP];  // bit b: ismatchedbykmer(probe b)
    uint32_t ringFree[RING_SLOTS / 32];
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// base at distance d from the end of the leaf string (d = 0 newest), d < 64
__device__ __forceinline__ int tail_base(uint64_t rt_hi, uint64_t rt_lo, int d)
{
    return d < 32 ? (int)((rt_hi >> (62 - 2 * d)) & 3) : (int)((rt_lo >> (62 - 2 * (d - 32))) & 3);
}
__device__ __forceinline__ void tail_push(uint64_t& rt_hi, uint64_t& rt_lo, int c)
{
    rt_lo = (rt_lo >> 2) | (rt_hi << 62);
    rt_hi = (rt_hi >> 2) | ((uint64_t)c << 62);
}

// single out-of-line copy of updateInterval for the walk code: the fully inlined kernel was 313 KB of SASS and spent
// 56% of its stall samples waiting for instructions (profiles/, round 1)
static __device__ __noinline__ Interval update_interval_ool(const FmTable& t, Interval iv, int c) { return update_interval(t, iv, c); }

// findInterval of the last K bases of a leaf on one strand (refineSAInterval, LongReadCorrectByOverlap.cpp:355-369):
//   strand 0: findInterval(RBWT, reverse(w))      processes w[0], w[1], ...
//   strand 1: findInterval(BWT,  revcomp(w))      processes comp(w[0]), comp(w[1]), ...
// with w[j] = base at distance K-1-j.  Stops at the first empty interval (BWTAlgorithms.cpp:25-29).
static __device__ __noinline__ Interval suffix_interval(const FmIndexDev& idx, uint64_t rt_hi, uint64_t rt_lo, int K, int strand)
{
    const FmTable& t = idx.t[strand == 0 ? PBSC_RBWT : PBSC_BWT];
    Interval iv;
    int j;
    if (idx.prefix != nullptr && K >= idx.k0)
    {
        // key = sum_j w[j]*4^j for j < k0: the k0 oldest bases of the window, oldest in the low bits
        const int sh = 128 - 2 * K;   // bit position of w[0] in the 128-bit tail
        uint64_t bits = sh >= 64 ? (rt_hi >> (sh - 64)) : (sh == 0 ? rt_lo : ((rt_lo >> sh) | (rt_hi << (64 - sh))));
        const uint64_t key = bits & ((1ull << (2 * idx.k0)) - 1ull);
        Interval f, r;
        prefix_lookup(idx, key, f, r);
        iv = strand == 0 ? f : r;
        j = idx.k0;
        if (!iv.valid()) return iv;
    }
    else
    {
        const int c = tail_base(rt_hi, rt_lo, K - 1);
        iv = init_interval(t, strand == 0 ? c : 3 - c);
        j = 1;
    }
    #pragma unroll 1
    for (; j < K; j++)
    {
        const int c = tail_base(rt_hi, rt_lo, K - 1 - j);
        iv = update_interval_ool(t, iv, strand == 0 ? c : 3 - c);
        if (!iv.valid()) break;
    }
    return iv;
}

// findInterval of a k-mer of the query on one strand, same conventions as suffix_interval
static __device__ __noinline__ Interval query_interval(const FmIndexDev& idx, const uint8_t* w, int K, int strand)
{
    const FmTable& t = idx.t[strand == 0 ? PBSC_RBWT : PBSC_BWT];
    Interval iv;
    int j;
    if (idx.prefix != nullptr && K >= idx.k0)
    {
        uint64_t key = 0;
        #pragma unroll 1
        for (int i = 0; i < idx.k0; i++) key |= (uint64_t)w[i] << (2 * i);
        Interval f, r;
        prefix_lookup(idx, key, f, r);
        iv = strand == 0 ? f : r;
        j = idx.k0;
        if (!iv.valid()) return iv;
    }
    else
    {
        iv = init_interval(t, strand == 0 ? w[0] : 3 - w[0]);
        j = 1;
    }
    #pragma unroll 1
    for (; j < K; j++)
    {
        iv = update_interval_ool(t, iv, strand == 0 ? w[j] : 3 - w[j]);
        if (!iv.valid()) break;
    }
    return iv;
}

struct KeyGreater { __host__ __device__ bool operator()(uint64_t a, uint64_t b) const { return (a >> 32) > (b >> 32); } };

static __device__ __noinline__ void sort_desc_ool(uint64_t* a, long n) { stlsort::sort(a, n, KeyGreater()); }

// first index of the run of `key` in an array sorted by key descending; n if absent
__device__ __forceinline__ uint32_t group_find(const uint64_t* a, uint32_t n, uint32_t key, uint32_t& len)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if ((uint32_t)(a[m] >> 32) > key) lo = m + 1; else hi = m; }
    uint32_t e = lo;
    while (e < n && (uint32_t)(a[e] >> 32) == key) e++;
    len = e - lo;
    return lo;
}

struct WalkState
{
    uint32_t n;          // live leaves (m_leaves.size())
    uint64_t curLen, curK, maxLength, minLength, maxIndel, minSA;
    uint32_t qlen, k, maxOverlap, trgLen, nTerm, n9F, n9R, n5;
    uint32_t nNodes, nRes, level;   // level = GlobalErrorRateRecord.size() of every live leaf
    int status;
};

__device__ __forceinline__ void ring_free(WarpShared& sh, uint32_t slot) { atomicOr(&sh.ringFree[slot >> 5], 1u << (slot & 31)); }
// uniform: every lane calls it and gets the same slot
__device__ __forceinline__ int ring_alloc_uniform(WarpShared& sh)
{
    int slot = -1;
    for (int w = 0; w < RING_SLOTS / 32; w++)
    {
        uint32_t m = sh.ringFree[w];
        if (m) { slot = w * 32 + (__ffs(m) - 1); break; }
    }
    __syncwarp();
    if (slot >= 0 && lane_id() == 0) sh.ringFree[slot >> 5] &= ~(1u << (slot & 31));
    __syncwarp();
    return slot;
}

// refineSAInterval (LongReadCorrectByOverlap.cpp:355-369) over `cnt` leaves of a bank
static __device__ __noinline__ void refine_bank(const FmIndexDev& idx, Leaf* bank, uint32_t cnt, int K)
{
    const int lane = lane_id();
    for (uint32_t it = lane; it < 2 * cnt; it += 32)
    {
        Leaf& L = bank[it >> 1];
        const int strand = it & 1;
        Interval iv = suffix_interval(idx, L.rt_hi, L.rt_lo, K, strand);
        if (strand == 0) { L.f_lo = iv.lo; L.f_hi = iv.hi; } else { L.r_lo = iv.lo; L.r_hi = iv.hi; }
    }
    __syncwarp();
}

// SelectFreqsOfrange (LongReadCorrectByOverlap.cpp:281-331)
static __device__ __noinline__ uint64_t select_freqs(const FmIndexDev& idx, const ExtParamsDev& P, Leaf* bank, uint32_t cnt,
                                                 uint64_t LB, uint64_t UB)
{
    const int lane = lane_id();
    const int extra = (int)(UB - LB);     // 0..2 further left extensions
    int mx[3] = {0, 0, 0};
    #pragma unroll 1
    for (uint32_t base = 0; base < 2 * cnt; base += 32)
    {
        const uint32_t it = base + lane;
        int64_t sz[3] = {0, 0, 0};
        if (it < 2 * cnt)
        {
            const Leaf& L = bank[it >> 1];
            const int strand = it & 1;
            // strand 0: Fwdinterval = findInterval(BWT, startkmer), processes the newest base first
            // strand 1: Rvcinterval = findInterval(RBWT, complement(startkmer)), same order, complemented
            const FmTable& t = idx.t[strand == 0 ? PBSC_BWT : PBSC_RBWT];
            Interval iv;
            int d;
            if (idx.prefix != nullptr && (int)LB >= idx.k0)
            {
                // table key x[j] = comp(base at distance j): entry.rvc is the BWT interval after processing
                // comp(x[j]) = the bases themselves, entry.fwd the RBWT interval after processing x[j] = their complements
                uint64_t key = 0;
                #pragma unroll 1
                for (int j = 0; j < idx.k0; j++) key |= (uint64_t)(3 - tail_base(L.rt_hi, L.rt_lo, j)) << (2 * j);
                Interval f, r;
                prefix_lookup(idx, key, f, r);
                iv = strand == 0 ? r : f;
                d = idx.k0;
            }
            else
            {
                const int c = tail_base(L.rt_hi, L.rt_lo, 0);
                iv = init_interval(t, strand == 0 ? c : 3 - c);
                d = 1;
            }
            #pragma unroll 1
            for (; d < (int)LB && iv.valid(); d++)
            {
                const int c = tail_base(L.rt_hi, L.rt_lo, d);
                iv = update_interval_ool(t, iv, strand == 0 ? c : 3 - c);
            }
            sz[0] = (int64_t)iv.size();
            #pragma unroll 1
            for (int e = 1; e <= extra; e++)
            {
                const int c = tail_base(L.rt_hi, L.rt_lo, (int)LB - 1 + e);
                if (iv.valid()) iv = update_interval_ool(t, iv, strand == 0 ? c : 3 - c);
                sz[e] = (int64_t)iv.size();
            }
        }
        #pragma unroll 1
        for (int e = 0; e <= extra; e++)
        {
            int64_t other = __shfl_xor_sync(FULL, sz[e], 1);
            int f = (int)(sz[e] + other);   // FMidx::kmerFrequency (int)
            for (int o = 16; o > 0; o >>= 1) f = max(f, __shfl_xor_sync(FULL, f, o));
            mx[e] = max(mx[e], f);
        }
    }
    if (mx[0] - P.freq_int[LB] < 5) return LB;
    for (int e = 1; e <= extra; e++) if (mx[e] - P.freq_int[LB + e] < 5) return LB + e;
    return UB;
}

// isInsufficientFreqs (LongReadCorrectByOverlap.cpp:334-352)
static __device__ __noinline__ bool insufficient_freqs(const ExtParamsDev& P, const Leaf* bank, uint32_t cnt)
{
    const int lane = lane_id();
    uint32_t high = 0;
    for (uint32_t base = 0; base < cnt; base += 32)
    {
        const uint32_t j = base + lane;
        bool h = j < cnt && bank[j].kmerFreq > P.high_freq_thr;
        high += __popc(__ballot_sync(FULL, h));
    }
    if (high == 0) return true;
    if (high <= 2 && cnt >= 5) return true;
    if (high <= 1 && cnt >= 3) return true;
    return false;
}

// getFMIndexExtensions' acceptance rule for one leaf (LongReadCorrectByOverlap.cpp:725-781); returns a 4-bit base mask
static __device__ __noinline__ uint32_t eval_extensions(const WarpShared& sh, uint32_t leaf, uint32_t tailCount, uint64_t cutoffSA)
{
    const int maxfreq = sh.pmax[leaf];
    const uint64_t totalcount = sh.ptotal[leaf];
    uint32_t mask = 0;
    #pragma unroll 1
    for (int b = 0; b < 4; b++)
    {
        const uint64_t kmerFreq = (uint64_t)(int64_t)sh.pfreq[leaf * 4 + b];
        const double kmerRatio = __ddiv_rn((double)kmerFreq, (double)maxfreq);
        const bool isHomopolymer = tailCount >= 3;
        const bool isMatchedBy5mer = (sh.pmatch[leaf] >> b) & 1;
        const bool isFreqPass = kmerFreq >= cutoffSA;
        const bool isLowCoverage = totalcount >= cutoffSA + 2;
        const bool isRepeat = maxfreq > 100, isHighlyRepeat = maxfreq > 150, isLowlyRepeat = maxfreq > 50;
        double cutoff;
        if (isMatchedBy5mer && isHighlyRepeat) cutoff = 0.125;
        else if (isMatchedBy5mer && isLowlyRepeat) cutoff = 0.2;
        else if (isFreqPass) cutoff = 0.25;
        else if (isLowCoverage) cutoff = 0.6;
        else cutoff = 2;
        if (isHomopolymer && isRepeat) cutoff = fmax(cutoff, 0.3);
        else if (isHomopolymer) cutoff = fmax(cutoff, 0.6);
        if (kmerRatio >= cutoff) mask |= 1u << b;
    }
    return mask;
}

// attempToExtend (LongReadCorrectByOverlap.cpp:373-465) + updateLeaves (:468-488).  Returns the number of new leaves.
static __device__ __noinline__ uint32_t attempt_extend(const FmIndexDev& idx, const ExtParamsDev& P, WarpScratch& ws, WarpShared& sh,
                                                   WalkState& S, uint64_t thr)
{
    const int lane = lane_id();
    // ---- drop leaves whose local error rate is far above the best one ----
    double e = lane < (int)S.n ? ws.oldL[lane].local_err : 1.0;
    double minErr = e;
    for (int o = 16; o > 0; o >>= 1) minErr = fmin(minErr, __shfl_xor_sync(FULL, minErr, o));
    minErr = fmin(minErr, 1.0);
    {
        const double diff = __dsub_rn(e, minErr);
        const bool drop = lane < (int)S.n && ((diff > 0.05 && S.curLen > (uint64_t)(RING_LEN / 2)) || (diff > 0.1 && S.curLen > 15));
        const unsigned keep = __ballot_sync(FULL, lane < (int)S.n && !drop);
        const unsigned dropm = __ballot_sync(FULL, drop);
        if (dropm)
        {
            Leaf tmp;
            if (lane < (int)S.n) tmp = ws.oldL[lane];
            if (drop) ring_free(sh, tmp.ring);
            __syncwarp();
            if (lane < (int)S.n && !drop) ws.oldL[__popc(keep & ((1u << lane) - 1u))] = tmp;
            S.n = __popc(keep);
            __syncwarp();
        }
    }
    const uint32_t n = S.n;
    // ---- probes: 8 lanes per leaf = (base, strand) ----
    #pragma unroll 1
    for (uint32_t base = 0; base < n; base += 4)
    {
        const uint32_t li = base + (lane >> 3);
        const int b = (lane >> 1) & 3, strand = lane & 1;
        int64_t sz = 0;
        Interval iv; iv.lo = iv.hi = 0;
        uint64_t rt_hi = 0;
        if (li < n)
        {
            const Leaf& L = ws.oldL[li];
            rt_hi = L.rt_hi;
            if (strand == 0) { iv.lo = L.f_lo; iv.hi = L.f_hi; if (iv.valid()) iv = update_interval_ool(idx.t[PBSC_RBWT], iv, b); }
            else { iv.lo = L.r_lo; iv.hi = L.r_hi; if (iv.valid()) iv = update_interval_ool(idx.t[PBSC_BWT], iv, 3 - b); }
            sz = (int64_t)iv.size();
            ProbeIv& pr = ws.probes[li * 4 + b];
            if (strand == 0) { pr.f_lo = iv.lo; pr.f_hi = iv.hi; } else { pr.r_lo = iv.lo; pr.r_hi = iv.hi; }
        }
        const int64_t other = __shfl_xor_sync(FULL, sz, 1);
        const int freq = (int)(sz + other);
        uint32_t tot = (uint32_t)freq;
        tot += __shfl_xor_sync(FULL, tot, 2);
        tot += __shfl_xor_sync(FULL, tot, 4);
        int mx = freq;
        mx = max(mx, __shfl_xor_sync(FULL, mx, 2));
        mx = max(mx, __shfl_xor_sync(FULL, mx, 4));
        // ismatchedbykmer (LongReadCorrectByOverlap.cpp:787-821): any query 5-mer equal to the probe's last five bases
        // within curLen +- maxIndel; needs one valid strand
        const bool anyValid = (sz > 0) || (other > 0);
        const uint32_t code5 = ((uint32_t)b << 8) | (uint32_t)(rt_hi >> 56);
        const bool m5 = anyValid && sh.win5[code5] != 0;
        unsigned mball = __ballot_sync(FULL, m5 && strand == 0);
        if (li < n && strand == 0)
        {
            sh.pfreq[li * 4 + b] = freq;
            if (b == 0)
            {
                sh.ptotal[li] = tot; sh.pmax[li] = mx;
                const unsigned grp = (mball >> (lane & ~7)) & 0xffu;   // bits 0,2,4,6 = bases A,C,G,T
                sh.pmatch[li] = (uint8_t)((grp & 1) | ((grp >> 1) & 2) | ((grp >> 2) & 4) | ((grp >> 3) & 8));
            }
        }
    }
    __syncwarp();
    // ---- decide per leaf (lane = leaf), with the one retry at threshold-1 for minimum-error leaves ----
    uint32_t mask = 0;
    Leaf parent;
    if (lane < (int)n)
    {
        parent = ws.oldL[lane];
        mask = eval_extensions(sh, lane, parent.tailCount, thr);
        if (!mask && parent.local_err == minErr && n > 1) mask = eval_extensions(sh, lane, parent.tailCount, thr - 1);
    }
    const uint32_t cnt = __popc(mask);
    uint32_t pos = cnt;
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(FULL, pos, o); if (lane >= o) pos += v; }
    const uint32_t total = __shfl_sync(FULL, pos, 31);
    pos -= cnt;
    if (total == 0) return 0;
    if (S.nNodes + total > P.node_cap) { S.status = PBSC_WALK_OVERFLOW; return 0; }
    // ---- create children in list order; first child keeps the parent's history ring ----
    if (lane < (int)n)
    {
        if (cnt == 0) ring_free(sh, parent.ring);
        uint32_t j = 0;
        #pragma unroll 1
        for (int b = 0; b < 4; b++)
        {
            if (!((mask >> b) & 1)) continue;
            Leaf c = parent;
            const ProbeIv pr = ws.probes[lane * 4 + b];
            c.f_lo = pr.f_lo; c.f_hi = pr.f_hi; c.r_lo = pr.r_lo; c.r_hi = pr.r_hi;
            c.kmerFreq = sh.pfreq[lane * 4 + b];
            tail_push(c.rt_hi, c.rt_lo, b);
            if (parent.tailLetter == b) c.tailCount = parent.tailCount + 1; else { c.tailLetter = (uint8_t)b; c.tailCount = 1; }
            c.node = S.nNodes + pos + j;
            ws.nodes[c.node] = (parent.node << 2) | (uint32_t)b;
            c.alive = (j == 0) ? 1 : 2;   // 2: needs its own ring (copied from the parent's) below
            ws.newL[pos + j] = c;
            j++;
        }
    }
    S.nNodes += total;
    __syncwarp();
    // ---- extra children: allocate a ring and copy the 100-deep history (createChild copies both records) ----
    unsigned need = __ballot_sync(FULL, cnt > 1);
    while (need)
    {
        const int src_lane = __ffs(need) - 1;
        need &= need - 1;
        const uint32_t p0 = __shfl_sync(FULL, pos, src_lane), c0 = __shfl_sync(FULL, cnt, src_lane);
        const uint32_t src_ring = ws.newL[p0].ring;
        for (uint32_t j = 1; j < c0; j++)
        {
            const int slot = ring_alloc_uniform(sh);
            if (slot < 0) { S.status = PBSC_WALK_OVERFLOW; return 0; }
            for (int x = lane; x < RING_LEN; x += 32) ws.rings[slot * RING_LEN + x] = ws.rings[src_ring * RING_LEN + x];
            if (lane == 0) { ws.newL[p0 + j].ring = (uint16_t)slot; ws.newL[p0 + j].alive = 1; }
        }
    }
    __syncwarp();
    return total;
}

// PrunedBySeedSupport + isSupportedByNewSeed + computeErrorRate (LongReadCorrectByOverlap.cpp:491-664)
static __device__ __noinline__ void prune_by_seed_support(const ExtParamsDev& P, WarpScratch& ws, WarpShared& sh, WalkState& S, uint32_t m)
{
    const int lane = lane_id();
    const uint64_t seedSize = (uint64_t)P.seed_size;
    const uint64_t curLen = S.curLen;
    const uint64_t currSeedIdx = curLen - seedSize;
    const uint64_t indelOffset = seedSize + S.maxIndel;
    const uint64_t smallSeedIdx = currSeedIdx <= indelOffset ? 0 : currSeedIdx - indelOffset;
    const uint64_t largeSeedIdx = (currSeedIdx + indelOffset) >= ((uint64_t)S.qlen - seedSize) ? ((uint64_t)S.qlen - seedSize) : currSeedIdx + indelOffset;
    const uint32_t keyMask = (1u << (2 * P.seed_size)) - 1u;
    #pragma unroll 1
    for (uint32_t base = 0; base < m; base += 32)
    {
        const uint32_t j = base + lane;
        if (j < m)
        {
            Leaf L = ws.newL[j];
            bool found = false;
            const uint64_t d = curLen - (uint64_t)L.lastOverlapLen;
            if (d > seedSize || d <= 1)
            {
                const uint64_t preSeedIdx = L.lastSeedIdx;
                // ---- isSupportedByNewSeed ----
                const uint64_t seedIdxOffset = (uint64_t)L.lastOverlapLen < curLen - seedSize ? seedSize : curLen - (uint64_t)L.lastOverlapLen;
                const uint64_t startSeedIdx = max(smallSeedIdx, (uint64_t)L.lastSeedIdx + seedIdxOffset);
                const bool fV = L.f_hi > L.f_lo, rV = L.r_hi > L.r_lo;
                const uint32_t keyF = (uint32_t)(L.rt_hi >> (64 - 2 * P.seed_size));
                uint32_t nf = 0, nr = 0, gf = 0, gr = 0;
                if (fV) gf = group_find(ws.sF, S.n9F, keyF, nf);
                if (rV) gr = group_find(ws.sR, S.n9R, keyMask - keyF, nr);
                int minIdxDiff = 10000;
                const uint32_t lim = max(nf, nr);
                #pragma unroll 1
                for (uint32_t i = 0; i < lim; i++)
                {
                    uint64_t v;
                    bool hit = false;
                    if (i < nf) { v = (uint32_t)ws.sF[gf + i]; hit = v >= startSeedIdx && v <= largeSeedIdx; }
                    if (!hit && i < nr) { v = (uint32_t)ws.sR[gr + i]; hit = v >= startSeedIdx && v <= largeSeedIdx; }
                    if (hit)
                    {
                        const int diff = abs((int)v - (int)currSeedIdx);
                        if (diff < minIdxDiff) { L.lastSeedIdx = (uint32_t)v; minIdxDiff = diff; }
                        L.lastOverlapLen = (uint32_t)curLen;
                        found = true;
                    }
                }
                if (found) L.totalSeeds++;
                // ---- back in PrunedBySeedSupport ----
                if (found)
                {
                    if (currSeedIdx + (uint64_t)(int64_t)L.seedOff - preSeedIdx > seedSize) L.redeem = __dadd_rn(L.redeem, P.redeem_a);
                    L.seedOff = (int)L.lastSeedIdx - (int)currSeedIdx;
                }
                else
                {
                    const uint64_t v = currSeedIdx + (uint64_t)(int64_t)L.seedOff - (uint64_t)L.lastSeedIdx;
                    if (v % seedSize == 1) { /* numOfErrors++ : never read */ }
                    else if (v > seedSize - 1) L.redeem = __dadd_rn(L.redeem, P.redeem_b);
                }
            }
            else L.redeem = __dadd_rn(L.redeem, P.redeem_b);
            // ---- computeErrorRate ----
            double matchedLen = __dsub_rn(__dadd_rn((double)L.totalSeeds, (double)seedSize), 1.0);
            matchedLen = __dadd_rn(matchedLen, L.redeem);
            const double totalLen = (double)curLen;   // currOverlapLen == m_currentLength for every live leaf
            const double unmatchedLen = __dsub_rn(totalLen, matchedLen);
            double err = __ddiv_rn(unmatchedLen, totalLen);
            double* ring = ws.rings + (size_t)L.ring * RING_LEN;
            ring[S.level % RING_LEN] = err;
            L.global_err = err;
            if (S.level + 1 >= (uint32_t)RING_LEN)
            {
                const double old = ring[(S.level + 1) % RING_LEN];
                err = __ddiv_rn(__dsub_rn(__dmul_rn(err, totalLen), __dmul_rn(old, __dsub_rn(totalLen, (double)RING_LEN))), (double)RING_LEN);
            }
            L.local_err = err;
            if (err > P.walk_error_rate) { L.alive = 0; ring_free(sh, L.ring); }
            ws.newL[j] = L;
        }
    }
    __syncwarp();
}

// isTerminated (LongReadCorrectByOverlap.cpp:825-878) over the alive new leaves, in list order
static __device__ __noinline__ void check_terminated(const ExtParamsDev& P, WarpScratch& ws, WalkState& S, uint32_t m)
{
    const int lane = lane_id();
    #pragma unroll 1
    for (uint32_t base = 0; base < m; base += 32)
    {
        const uint32_t j = base + lane;
        int ilast = -1, first = -1;
        if (j < m && ws.newL[j].alive)
        {
            const Leaf& L = ws.newL[j];
            const bool fV = L.f_hi > L.f_lo, rV = L.r_hi > L.r_lo;
            first = L.res_first;
            #pragma unroll 1
            for (int i = max(L.res_second, 0); i < (int)S.nTerm; i++)
            {
                const Interval tf = ws.termF[i], tr = ws.termR[i];
                // lower >= T.lower && upper <= T.upper on inclusive bounds == lo >= T.lo && hi <= T.hi on half-open ones;
                // an empty terminal interval can never contain a valid one
                const bool ft = fV && tf.valid() && L.f_lo >= tf.lo && L.f_hi <= tf.hi;
                const bool rt = rV && tr.valid() && L.r_lo >= tr.lo && L.r_hi <= tr.hi;
                if (ft || rt) ilast = i;
            }
        }
        const unsigned hit = __ballot_sync(FULL, ilast >= 0);
        const unsigned fresh = __ballot_sync(FULL, ilast >= 0 && first == -1);
        if (hit)
        {
            if (S.nRes + __popc(fresh) > RES_CAP) { S.status = PBSC_WALK_OVERFLOW; return; }
            int slot = first;   // 1-based
            if (ilast >= 0 && first == -1) slot = (int)S.nRes + __popc(fresh & ((1u << lane) - 1u)) + 1;
            // later leaves overwrite earlier ones that share a slot (siblings inherit resultindex)
            unsigned pend = hit;
            while (pend)
            {
                const int l = __ffs(pend) - 1;
                pend &= pend - 1;
                if (lane == l)
                {
                    Leaf& L = ws.newL[j];
                    WalkResult r; r.err = L.global_err; r.node = L.node; r.i = ilast; r.depth = (uint32_t)S.curLen; r.pad = 0;
                    ws.res[slot - 1] = r;
                    L.res_first = slot; L.res_second = ilast;
                }
                __syncwarp();
            }
            S.nRes += __popc(fresh);
        }
    }
    __syncwarp();
}

// One complete walk.  ws.q[0..qlen) = beginningkmer(k) + strBetweenSrcTarget(dis bases) + targetSeed(trgLen), as 2-bit codes.
// On success returns 1 and leaves the merged sequence (codes) in ws.merged[0..*mergedLen).
static __device__ __noinline__ int walk_pair(const FmIndexDev& idx, const ExtParamsDev& P, WarpScratch& ws, WarpShared& sh,
                                uint32_t qlen, uint32_t k, int32_t dis, uint32_t trgLen, uint64_t minSA, uint32_t* mergedLen)
{
    const int lane = lane_id();
    WalkState S;
    S.status = 0;
    S.qlen = qlen; S.k = k; S.maxOverlap = k + 2; S.trgLen = trgLen; S.minSA = minSA;
    if (trgLen < (uint32_t)P.min_overlap || k < (uint32_t)P.seed_size || k + 3 > 64 || qlen > P.q_cap || trgLen - P.min_overlap + 1 > TERM_CAP || qlen != k + (uint32_t)dis + trgLen)
        return PBSC_WALK_UNSUPPORTED;
    // LongReadCorrectByOverlap.cpp:54-58,77-79
    S.maxIndel = dis > 100 ? (uint64_t)__dmul_rn((double)dis, 0.2) : 20;
    S.maxLength = (uint64_t)__dadd_rn(__dmul_rn(1.2, (double)(dis + 10)), (double)(2 * (uint64_t)k));
    S.minLength = (uint64_t)__dadd_rn(__dmul_rn(0.8, (double)(dis - 20)), (double)(2 * (uint64_t)k));
    S.curLen = S.curK = k;
    S.nTerm = trgLen - P.min_overlap + 1;
    S.nNodes = 1; S.nRes = 0; S.level = 1;
    if (S.maxLength + trgLen + 8 > P.merged_cap) return PBSC_WALK_OVERFLOW;
    const uint8_t* q = ws.q;

    // ---- ring allocator + 5-mer window ----
    for (int x = lane; x < 1024; x += 32) sh.win5[x] = 0;
    if (lane < RING_SLOTS / 32) sh.ringFree[lane] = 0xffffffffu;
    __syncwarp();
    // ---- terminal intervals: every min_overlap-mer of the target on both strands (:82-88) ----
    const uint8_t* trg = q + k + dis;
    for (uint32_t it = lane; it < 2 * S.nTerm; it += 32)
    {
        const Interval iv = query_interval(idx, trg + (it >> 1), P.min_overlap, it & 1);
        if (it & 1) ws.termR[it >> 1] = iv; else ws.termF[it >> 1] = iv;
    }
    // ---- query idmers (buildOverlapbyFMindex, :127-152): keep the valid ones, in position order ----
    const int s9 = P.seed_size;
    const uint32_t n9 = qlen - s9 + 1;
    S.n9F = 0; S.n9R = 0;
    for (uint32_t base = 0; base < n9; base += 32)
    {
        const uint32_t p = base + lane;
        bool vf = false, vr = false;
        uint32_t keyF = 0;
        if (p < n9)
        {
            for (int j = 0; j < s9; j++) keyF |= (uint32_t)q[p + j] << (2 * j);
            vf = query_interval(idx, q + p, s9, 0).valid();
            vr = query_interval(idx, q + p, s9, 1).valid();
        }
        const unsigned bf = __ballot_sync(FULL, vf), br = __ballot_sync(FULL, vr);
        if (vf) ws.sF[S.n9F + __popc(bf & ((1u << lane) - 1u))] = ((uint64_t)keyF << 32) | p;
        if (vr) ws.sR[S.n9R + __popc(br & ((1u << lane) - 1u))] = ((uint64_t)(((1u << (2 * s9)) - 1u) - keyF) << 32) | p;
        S.n9F += __popc(bf); S.n9R += __popc(br);
    }
    // ---- query 5-mers: code with the newest base most significant ----
    S.n5 = qlen >= 5 ? qlen - 4 : 0;
    for (uint32_t p = lane; p < S.n5; p += 32)
        ws.c5[p] = (uint16_t)(q[p] | (q[p + 1] << 2) | (q[p + 2] << 4) | (q[p + 3] << 6) | (q[p + 4] << 8));
    __syncwarp();
    // the reference sorts each idmer list with std::sort(greater-by-start); start order == key order
    if (lane < 2) sort_desc_ool(lane == 0 ? ws.sF : ws.sR, (long)(lane == 0 ? S.n9F : S.n9R));
    // window of curLen = k: positions [max(k - maxIndel, 0), k + maxIndel]
    {
        const int64_t lo = max((int64_t)k - (int64_t)S.maxIndel, (int64_t)0);
        const int64_t hi = min((int64_t)k + (int64_t)S.maxIndel, (int64_t)S.n5 - 1);
        if (lane == 0) for (int64_t p = lo; p <= hi; p++) sh.win5[ws.c5[p]]++;
    }
    // ---- root leaf (initialRootNode, :106-124; leafInfo ctor, LongReadCorrectByOverlap.h:160-178) ----
    if (lane < 2)
    {
        const Interval iv = query_interval(idx, q, k, lane);
        if (lane == 0) { ws.oldL[0].f_lo = iv.lo; ws.oldL[0].f_hi = iv.hi; } else { ws.oldL[0].r_lo = iv.lo; ws.oldL[0].r_hi = iv.hi; }
    }
    __syncwarp();
    if (lane == 0)
    {
        Leaf& R = ws.oldL[0];
        R.redeem = 0; R.local_err = 0; R.global_err = 0;
        R.rt_hi = R.rt_lo = 0;
        for (uint32_t j = 0; j < k; j++) tail_push(R.rt_hi, R.rt_lo, q[j]);
        R.lastOverlapLen = k; R.lastSeedIdx = k - s9; R.totalSeeds = k - s9 + 1; R.seedOff = 0;
        R.res_first = -1; R.res_second = -1;
        R.kmerFreq = (int)((int64_t)(R.f_hi - R.f_lo) + (int64_t)(R.r_hi - R.r_lo));
        R.tailLetter = q[k - 1];
        uint32_t tc = 0;
        for (int j = (int)k - 1; j >= 0 && q[j] == R.tailLetter; j--) tc++;
        R.tailCount = tc;
        R.node = 0; R.ring = 0; R.alive = 1;
        ws.nodes[0] = 0;
        ws.rings[0] = 0.0;   // GlobalErrorRateRecord = {0}
        sh.ringFree[0] &= ~1u;
    }
    S.n = 1;
    __syncwarp();

    // ---- extendOverlap (:155-211) ----
    while (S.n > 0 && S.n <= (uint32_t)P.max_leaves && S.curLen <= S.maxLength)
    {
        // extendLeaves (:239-278)
        if (S.curK > S.maxOverlap) { refine_bank(idx, ws.oldL, S.n, (int)S.maxOverlap); S.curK = S.maxOverlap; }
        uint32_t m = attempt_extend(idx, P, ws, sh, S, S.minSA);
        if (S.status) return S.status;
        if (m == 0 && S.n > 0)
        {
            const uint64_t LB = max(S.curK - 2, (uint64_t)P.min_overlap);
            const uint64_t R = select_freqs(idx, P, ws.oldL, S.n, LB, S.curK);
            refine_bank(idx, ws.oldL, S.n, (int)R);
            S.curK = R;
            m = attempt_extend(idx, P, ws, sh, S, S.minSA);
            if (S.status) return S.status;
            if (m == 0) { m = attempt_extend(idx, P, ws, sh, S, S.minSA - 1); if (S.status) return S.status; }
        }
        if (m > 0)
        {
            S.curLen++;
            S.curK++;
            if (insufficient_freqs(P, ws.newL, m))
            {
                const uint64_t LB = max(S.curK - 2, (uint64_t)P.min_overlap);
                const uint64_t R = select_freqs(idx, P, ws.newL, m, LB, S.curK);
                refine_bank(idx, ws.newL, m, (int)R);
                S.curK = R;
            }
            // slide the 5-mer window to the new curLen
            if (lane == 0)
            {
                const int64_t add = (int64_t)S.curLen + (int64_t)S.maxIndel;
                const int64_t rem = (int64_t)S.curLen - 1 - (int64_t)S.maxIndel;
                if (add < (int64_t)S.n5) sh.win5[ws.c5[add]]++;
                if (rem >= 0 && rem < (int64_t)S.n5) sh.win5[ws.c5[rem]]--;
            }
            prune_by_seed_support(P, ws, sh, S, m);
            S.level++;
        }
        // m_leaves = newLeaves; isTerminated
        if (m > 0 && S.curLen >= S.minLength) { check_terminated(P, ws, S, m); if (S.status) return S.status; }
        // compact the survivors into the old bank
        uint32_t nn = 0;
        for (uint32_t base = 0; base < m; base += 32)
        {
            const uint32_t j = base + lane;
            const bool a = j < m && ws.newL[j].alive;
            const unsigned bal = __ballot_sync(FULL, a);
            const uint32_t dst = nn + __popc(bal & ((1u << lane) - 1u));
            if (a && dst < OLD_CAP) ws.oldL[dst] = ws.newL[j];
            nn += __popc(bal);
        }
        S.n = nn;
        __syncwarp();
    }

    // ---- findTheBestPath (:214-236) ----
    if (S.nRes > 0)
    {
        double best = 1.0;
        int bi = -1;
        for (uint32_t i = 0; i < S.nRes; i++) { const double e = ws.res[i].err; if (e < best) { best = e; bi = (int)i; } }
        if (bi < 0) return PBSC_WALK_NO_PATH;
        const WalkResult r = ws.res[bi];
        const uint32_t chain = r.depth - k;
        const uint32_t tailFrom = (uint32_t)r.i + P.min_overlap;
        const uint32_t tailLen = trgLen > (uint32_t)P.min_overlap ? trgLen - tailFrom : 0;
        const uint32_t len = r.depth + tailLen;
        if (len > P.merged_cap) return PBSC_WALK_OVERFLOW;
        for (uint32_t x = lane; x < k; x += 32) ws.merged[x] = q[x];
        for (uint32_t x = lane; x < tailLen; x += 32) ws.merged[r.depth + x] = trg[tailFrom + x];
        if (lane == 0)
        {
            uint32_t node = r.node;
            for (uint32_t x = 0; x < chain; x++) { const uint32_t v = ws.nodes[node]; ws.merged[r.depth - 1 - x] = (uint8_t)(v & 3); node = v >> 2; }
        }
        __syncwarp();
        *mergedLen = len;
        return 1;
    }
    if (S.n == 0) return -1;
    if (S.curLen > S.maxLength) return -2;
    if (S.n > (uint32_t)P.max_leaves) return -3;
    return -4;
}

}  // namespace pbsc
#endif
