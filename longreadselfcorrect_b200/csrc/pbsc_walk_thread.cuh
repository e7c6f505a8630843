// pbsc_walk_thread.cuh — one FM-index walk per THREAD (scalar restatement of
// LongReadSelfCorrectByOverlap, PacBio/LongReadCorrectByOverlap.cpp:17-878).
//
// Why a second engine: the warp-per-walk engine (pbsc_walk.cuh) spends its time fetching instructions
// (ncu, profiles/r1: 43-56 % of stall samples are `no_inst`, 246 k warp instructions per walk with 8 of
// 32 lanes active) because a typical walk has one or two live leaves and ~40 levels: there is no
// parallelism inside a walk to give to 32 lanes, but there are millions of independent walks.  Here
// every lane runs its own walk; the loop structure (one level per iteration) is shared by all lanes,
// and each lane keeps several independent rank sectors in flight.
//
// Differences from the warp engine that do not change results:
//   * the four one-base probes of a leaf read the four counts of ONE sector per interval bound
//     (occ4): 4 sectors per leaf and level instead of the reference's 16 getOcc calls;
//   * kmerRatio >= cutoff (LongReadCorrectByOverlap.cpp:740-781) is evaluated as an exact integer
//     cross-multiplication: for integers 0 <= a <= b < 2^31 and cutoff p/q in {1/8, 1/5, 1/4, 3/10, 3/5, 2}
//     |a/b - p/q| is either 0 or >= 1/(10 b) >> 2^-53, so fl(a/b) >= fl(p/q) <=> a q >= p b;
//   * query idmers are looked up in a per-walk hash table; only if the query contains the same idmer
//     twice (where the reference's std::sort order of equal keys matters) are the validity-filtered
//     lists built and sorted with the libstdc++ permutation (stl_sort_emul.cuh).
#ifndef PBSC_WALK_THREAD_CUH
#define PBSC_WALK_THREAD_CUH

#include "pbsc_walk.cuh"

namespace pbsc {
namespace tw {

// A pass of the thread engine carries walks of at most `old_cap` live leaves and 4 * old_cap children per level.  A walk
// that outgrows its pass (repeats, ExceedLeaves/ExceedDepth cases: a few percent of the walks, but each one costs 10-100x a
// light walk) stops with PBSC_WALK_HEAVY and is re-walked from scratch in the next pass, among walks of its own weight.
struct Caps { uint32_t old_cap, new_cap, rings, res; };   // per-lane capacities of one pass of the thread engine (new_cap = 4 * old_cap)
#define PBSC_WALK_HEAVY (-102)

// Leaves live in two banks of new_cap slots each.  The live leaves of a level are the slots named by a list (n entries); the
// children of the i-th listed leaf go to slots 4 i + b of the other bank (b = appended base), so any lane can create them
// without knowing what the other leaves do.  Nothing is ever moved: a level ends by writing the list of the surviving
// children and flipping the roles of the banks.
// Per-lane scratch (global memory), the same layout for every lane of a pass:
struct Layout
{
    uint32_t bank1, rings, nodes, res, ringStack, list0, list1;   // byte offsets from the lane's base (bank 0 is at 0)
    uint32_t node_cap;
    Caps cap;
};

__host__ __device__ inline uint32_t pow2_ceil(uint32_t x) { uint32_t p = 1; while (p < x) p <<= 1; return p; }

__host__ __device__ inline size_t thread_scratch_bytes(uint32_t node_cap, Caps c)
{
    size_t b = 0;
    b += sizeof(Leaf) * (size_t)(2 * c.new_cap);   // two banks of the same size: the level loop flips them instead of copying
    b += sizeof(double) * (size_t)c.rings * RING_LEN;
    b += align_up(sizeof(uint32_t) * (size_t)node_cap, 16);
    b += align_up(sizeof(WalkResult) * (size_t)c.res, 16);
    b += align_up(c.rings, 16);
    b += 2 * align_up(c.new_cap, 16);
    return align_up(b, 128);
}

__host__ __device__ inline Layout make_layout(uint32_t node_cap, Caps c)
{
    Layout y;
    size_t p = sizeof(Leaf) * (size_t)c.new_cap;
    y.bank1 = (uint32_t)p; p += sizeof(Leaf) * (size_t)c.new_cap;
    y.rings = (uint32_t)p; p += sizeof(double) * (size_t)c.rings * RING_LEN;
    y.nodes = (uint32_t)p; p += align_up(sizeof(uint32_t) * (size_t)node_cap, 16);
    y.res = (uint32_t)p; p += align_up(sizeof(WalkResult) * (size_t)c.res, 16);
    y.ringStack = (uint32_t)p; p += align_up(c.rings, 16);
    y.list0 = (uint32_t)p; p += align_up(c.new_cap, 16);
    y.list1 = (uint32_t)p;
    y.node_cap = node_cap; y.cap = c;
    return y;
}

// Everything a walk carries from level to level besides its leaves: ONE record per lane in shared memory.  It used to be a
// 576-byte struct on the thread's stack (local memory): with 768 walks per SM that is 440 KB against ~90 KB of L1, so nine in
// ten accesses to it went to L2 and, L2 being thrashed by the leaf banks, often to DRAM (ncu, profiles/r2_walk_levels_cfg2_*:
// 12.9 G of the kernel's 60 G L2 sector transactions were local memory, L1 hit rate 10 %).  In shared memory it is also what
// the OTHER lanes of the warp read when the pooled stages give them one of this walk's leaves.
struct __align__(16) Lane
{
    uint8_t* base;                                   // this lane's scratch (Layout)
    const uint16_t* start4; const uint16_t* pos4; const uint8_t* q;   // the walk's setup record
    const uint64_t* hash; const uint64_t* sF; const uint64_t* sR;
    const Interval* termF; const Interval* termR;
    double minErr;
    uint32_t curLen, curK, maxLength, minLength, maxIndel, minSA, thr;
    uint32_t qlen, k, maxOverlap, trgLen, nTerm, n9F, n9R, n5, nNodes, nRes, level, hashMask, nFree, nFresh, phase, n;
    // SelectFreqsOfrange / refineSAInterval behind it (pooled): the k-mer range, the per-size maxima, which bank
    uint32_t selLB, selExtra, selK;
    int selMx[3];
    int status;
    uint8_t dup, flip, selNew, pad;
};
static_assert(sizeof(Lane) == 208, "Lane: 208 bytes of shared memory per thread");

__device__ __forceinline__ Leaf* old_bank(const Lane& s, const Layout& y) { return (Leaf*)(s.base + (s.flip ? y.bank1 : 0u)); }
__device__ __forceinline__ Leaf* new_bank(const Lane& s, const Layout& y) { return (Leaf*)(s.base + (s.flip ? 0u : y.bank1)); }
__device__ __forceinline__ uint8_t* old_list(const Lane& s, const Layout& y) { return s.base + (s.flip ? y.list1 : y.list0); }
__device__ __forceinline__ uint8_t* new_list(const Lane& s, const Layout& y) { return s.base + (s.flip ? y.list0 : y.list1); }
__device__ __forceinline__ double* rings_of(const Lane& s, const Layout& y) { return (double*)(s.base + y.rings); }
__device__ __forceinline__ uint32_t* nodes_of(const Lane& s, const Layout& y) { return (uint32_t*)(s.base + y.nodes); }
__device__ __forceinline__ WalkResult* res_of(const Lane& s, const Layout& y) { return (WalkResult*)(s.base + y.res); }
__device__ __forceinline__ uint8_t* ring_stack(const Lane& s, const Layout& y) { return s.base + y.ringStack; }

static __device__ __noinline__ Interval update1(const FmTable& t, Interval iv, int c) { return update_interval(t, iv, c); }
// one copy of occ4 for the four probes of a leaf: the level loop is bound by instruction fetch (ncu: the GPC instruction cache
// runs at 86 % of its request peak with 84 KB of SASS), so hot code is kept small rather than inlined
static __device__ __noinline__ void occ4_ool(const FmTable& t, uint64_t p, uint64_t r[4]) { occ4(t, p, r); }
// copy of one 128-byte leaf record
__device__ __forceinline__ void leaf_copy(Leaf* dst, const Leaf* src)
{
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    #pragma unroll
    for (int x = 0; x < (int)(sizeof(Leaf) / 16); x++) d4[x] = s4[x];
}

// both strands of the k-mer w[0..K): fwd = findInterval(RBWT, reverse(w)), rvc = findInterval(BWT, revcomp(w));
// `get(j)` returns w[j].  The two chains are advanced together so that their sectors are in flight at the same time.
template <class Get>
__device__ __forceinline__ void both_strands(const FmIndexDev& idx, Get get, int K, Interval& f, Interval& r)
{
    int j;
    if (idx.prefix != nullptr && K >= idx.k0)
    {
        uint64_t key = 0;
        #pragma unroll 1
        for (int i = 0; i < idx.k0; i++) key |= (uint64_t)get(i) << (2 * i);
        prefix_lookup(idx, key, f, r);
        j = idx.k0;
    }
    else
    {
        const int c = get(0);
        f = init_interval(idx.t[PBSC_RBWT], c);
        r = init_interval(idx.t[PBSC_BWT], 3 - c);
        j = 1;
    }
    #pragma unroll 1
    for (; j < K && (f.valid() || r.valid()); j++)
    {
        const int c = get(j);
        if (f.valid()) f = update1(idx.t[PBSC_RBWT], f, c);
        if (r.valid()) r = update1(idx.t[PBSC_BWT], r, 3 - c);
    }
}

// Deal `cnt` items of every lane out over the warp: f(owner lane, item index within the owner) runs once per item, 32 items
// per round.  Must be called by all 32 lanes converged (cnt may be 0).
template <class F>
__device__ __forceinline__ void pool_run(uint32_t cnt, F f)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
    uint32_t incl = cnt;
    #pragma unroll
    for (int off = 1; off < 32; off <<= 1)
    {
        const uint32_t o = __shfl_up_sync(FULL, incl, off);
        if (lane >= off) incl += o;
    }
    const uint32_t excl = incl - cnt;
    const uint32_t T = __shfl_sync(FULL, incl, 31);
    #pragma unroll 1
    for (uint32_t base = 0; base < T; base += 32)
    {
        const bool has = base + lane < T;
        const uint32_t x = has ? base + lane : T - 1;
        // owner = last lane whose first item is <= x
        int lo = 0, hi = 31;
        #pragma unroll
        for (int step = 0; step < 5; step++)
        {
            const int mid = (lo + hi + 1) >> 1;
            const uint32_t e = __shfl_sync(FULL, excl, mid);
            if (e <= x) lo = mid; else hi = mid - 1;
        }
        const uint32_t first = __shfl_sync(FULL, excl, lo);
        if (has) f(lo, x - first);
    }
    __syncwarp();
}

__device__ __forceinline__ void ring_release(Lane& S, const Layout& Y, uint32_t slot) { ring_stack(S, Y)[S.nFree++] = (uint8_t)slot; }
// released slots first, then slots never used by this walk (slot 0 is the root's)
__device__ __forceinline__ int ring_take(Lane& S, const Layout& Y)
{
    if (S.nFree) return (int)ring_stack(S, Y)[--S.nFree];
    return S.nFresh < Y.cap.rings ? (int)S.nFresh++ : -1;
}

// refineSAInterval(maxOverlap) for a whole warp at once: every lane passes the number of leaves of its own walk that start
// this level by cutting their k-mer back (0 if none)
__device__ __forceinline__ void refine_pool(const FmIndexDev& idx, const Layout& Y, const Lane* ctx, uint32_t cnt)
{
    pool_run(cnt, [&](int owner, uint32_t i) {
        const Lane& c = ctx[owner];
        Leaf* L = old_bank(c, Y) + old_list(c, Y)[i];
        const uint64_t rhi = L->rt_hi, rlo = L->rt_lo;
        const int K = (int)c.maxOverlap;
        Interval f, r;
        both_strands(idx, [&](int j) { return tail_base(rhi, rlo, K - 1 - j); }, K, f, r);
        L->f_lo = f.lo; L->f_hi = f.hi; L->r_lo = r.lo; L->r_hi = r.hi;
    });
}

// isInsufficientFreqs (LongReadCorrectByOverlap.cpp:334-352)
__device__ __forceinline__ bool insufficient(const ExtParamsDev& P, const Leaf* bank, const uint8_t* list, uint32_t cnt)
{
    uint32_t high = 0;
    #pragma unroll 1
    for (uint32_t j = 0; j < cnt; j++) high += bank[list[j]].kmerFreq > P.high_freq_thr;
    if (high == 0) return true;
    if (high <= 2 && cnt >= 5) return true;
    if (high <= 1 && cnt >= 3) return true;
    return false;
}

// acceptance rule of getFMIndexExtensions (LongReadCorrectByOverlap.cpp:725-781) with exact integer ratio tests
static __device__ __noinline__ uint32_t eval4(const int freq[4], uint64_t totalcount, int maxfreq, uint32_t match5, uint32_t tailCount, uint64_t cutoffSA)
{
    uint32_t mask = 0;
    if (maxfreq <= 0) return 0;   // 0/0 is NaN in the reference: never >= cutoff
    const bool isHomopolymer = tailCount >= 3;
    const bool isRepeat = maxfreq > 100, isHighlyRepeat = maxfreq > 150, isLowlyRepeat = maxfreq > 50;
    const bool isLowCoverage = totalcount >= cutoffSA + 2;
    #pragma unroll 1
    for (int b = 0; b < 4; b++)
    {
        const uint64_t kmerFreq = (uint64_t)(int64_t)freq[b];
        const bool isMatchedBy5mer = (match5 >> b) & 1;
        const bool isFreqPass = kmerFreq >= cutoffSA;
        // cutoff as p/q
        int p, q;
        if (isMatchedBy5mer && isHighlyRepeat) { p = 1; q = 8; }
        else if (isMatchedBy5mer && isLowlyRepeat) { p = 1; q = 5; }
        else if (isFreqPass) { p = 1; q = 4; }
        else if (isLowCoverage) { p = 3; q = 5; }
        else { p = 2; q = 1; }
        // std::max(cutoff, 0.3) / std::max(cutoff, 0.6)
        if (isHomopolymer && isRepeat) { if (p * 10 < 3 * q) { p = 3; q = 10; } }
        else if (isHomopolymer) { if (p * 5 < 3 * q) { p = 3; q = 5; } }
        if ((int64_t)freq[b] * q >= (int64_t)p * maxfreq) mask |= 1u << b;
    }
    return mask;
}

// ismatchedbykmer (LongReadCorrectByOverlap.cpp:787-821) for the four probes of one leaf at once: bit b is set when the
// query holds, starting within curLen +- maxIndel, the leaf's last four bases followed by base b
__device__ __forceinline__ uint32_t match5_mask(const Lane& c, uint32_t tail4)
{
    const int64_t lo = max((int64_t)c.curLen - (int64_t)c.maxIndel, (int64_t)0);
    const int64_t hi = min((int64_t)c.curLen + (int64_t)c.maxIndel, (int64_t)c.n5 - 1);
    uint32_t mask = 0;
    const uint32_t e1 = c.start4[tail4 + 1];
    #pragma unroll 1
    for (uint32_t e = c.start4[tail4]; e < e1; e++)
    {
        const int64_t p = c.pos4[e];
        if (p >= lo && p <= hi) mask |= 1u << c.q[p + 4];
    }
    return mask;
}

// attempToExtend, first part (LongReadCorrectByOverlap.cpp:373-398), by the walk's own lane: drop the leaves whose local error
// rate is far above the best one.  Only the list changes.
static __device__ __noinline__ void filter_leaves(Lane& S, const Layout& Y)
{
    Leaf* oldL = old_bank(S, Y);
    uint8_t* list = old_list(S, Y);
    double minErr = 1.0;
    #pragma unroll 1
    for (uint32_t i = 0; i < S.n; i++) minErr = fmin(minErr, oldL[list[i]].local_err);
    uint32_t w = 0;
    #pragma unroll 1
    for (uint32_t i = 0; i < S.n; i++)
    {
        const Leaf& L = oldL[list[i]];
        const double diff = __dsub_rn(L.local_err, minErr);
        const bool drop = (diff > 0.05 && S.curLen > (uint32_t)(RING_LEN / 2)) || (diff > 0.1 && S.curLen > 15);
        if (drop) { ring_release(S, Y, L.ring); continue; }
        if (w != i) list[w] = list[i];
        w++;
    }
    S.n = w;
    S.minErr = minErr;
}

// getFMIndexExtensions + the leaf part of updateLeaves (LongReadCorrectByOverlap.cpp:468-488, 667-784) for ONE leaf, by
// whichever lane the pool gave it to: the eight one-base probes from four sectors, the acceptance rule, and the accepted
// children written to slots 4 i + b of the other bank.  The accepted bases are left in the parent's `aux` for its owner,
// who numbers the children and gives them their history rings (adopt_children).
__device__ __forceinline__ void probe_leaf(const FmIndexDev& idx, const Layout& Y, const Lane& c, uint32_t i)
{
    Leaf* parent = old_bank(c, Y) + old_list(c, Y)[i];
    uint64_t fl[4], fh[4], rl[4], rh[4];
    const uint64_t pf_lo = parent->f_lo, pf_hi = parent->f_hi, pr_lo = parent->r_lo, pr_hi = parent->r_hi;
    const bool fV = pf_hi > pf_lo, rV = pr_hi > pr_lo;
    if (fV) { occ4_ool(idx.t[PBSC_RBWT], pf_lo, fl); occ4_ool(idx.t[PBSC_RBWT], pf_hi, fh); }
    if (rV) { occ4_ool(idx.t[PBSC_BWT], pr_lo, rl); occ4_ool(idx.t[PBSC_BWT], pr_hi, rh); }
    int freq[4];
    Interval pf[4], pr[4];
    uint64_t total = 0;
    int mx = 0;
    uint32_t match5 = 0;
    const uint64_t p_rt_hi = parent->rt_hi;
    const uint32_t near5 = match5_mask(c, (uint32_t)(p_rt_hi >> 56));
    #pragma unroll 1
    for (int b = 0; b < 4; b++)
    {
        if (fV) { pf[b].lo = idx.t[PBSC_RBWT].C[b] + fl[b]; pf[b].hi = idx.t[PBSC_RBWT].C[b] + fh[b]; }
        else { pf[b].lo = pf_lo; pf[b].hi = pf_hi; }
        if (rV) { pr[b].lo = idx.t[PBSC_BWT].C[3 - b] + rl[3 - b]; pr[b].hi = idx.t[PBSC_BWT].C[3 - b] + rh[3 - b]; }
        else { pr[b].lo = pr_lo; pr[b].hi = pr_hi; }
        freq[b] = (int)((int64_t)pf[b].size() + (int64_t)pr[b].size());
        total += (uint64_t)(int64_t)freq[b];
        mx = max(mx, freq[b]);
        if ((pf[b].valid() || pr[b].valid()) && ((near5 >> b) & 1)) match5 |= 1u << b;
    }
    const uint32_t p_tailCount = parent->tailCount;
    uint32_t mask = eval4(freq, total, mx, match5, p_tailCount, (uint64_t)c.thr);
    if (!mask && parent->local_err == c.minErr && c.n > 1) mask = eval4(freq, total, mx, match5, p_tailCount, (uint64_t)c.thr - 1);
    parent->aux = mask;
    if (!mask) return;
    const uint32_t p_tailLetter = parent->tailLetter;
    const uint64_t p_rt_lo = parent->rt_lo;
    Leaf* newL = new_bank(c, Y);
    #pragma unroll 1
    for (int b = 0; b < 4; b++)
    {
        if (!((mask >> b) & 1)) continue;
        Leaf* ch = newL + 4 * i + b;
        leaf_copy(ch, parent);
        ch->f_lo = pf[b].lo; ch->f_hi = pf[b].hi; ch->r_lo = pr[b].lo; ch->r_hi = pr[b].hi;
        ch->kmerFreq = freq[b];
        uint64_t th = p_rt_hi, tl = p_rt_lo;
        tail_push(th, tl, b);
        ch->rt_hi = th; ch->rt_lo = tl;
        if (p_tailLetter == (uint32_t)b) ch->tailCount = p_tailCount + 1; else { ch->tailLetter = (uint8_t)b; ch->tailCount = 1; }
        ch->alive = 1;
        ch->aux = 0;
    }
}

// The sequential rest of updateLeaves (LongReadCorrectByOverlap.cpp:468-488; SAIOverlapNode3::createChild,
// FMIndexWalk/SAINode.cpp:166-189), by the walk's own lane: children are numbered in the reference's order (leaves in list
// order, bases A..T), the first child of a leaf inherits its history ring, the others get a copy.  Returns the number of
// children and writes their slots to the child list.
static __device__ __noinline__ uint32_t adopt_children(Lane& S, const Layout& Y)
{
    Leaf* oldL = old_bank(S, Y);
    Leaf* newL = new_bank(S, Y);
    const uint8_t* list = old_list(S, Y);
    uint8_t* cl = new_list(S, Y);
    uint32_t* nodes = nodes_of(S, Y);
    const uint32_t n = S.n;
    uint32_t total = 0;
    #pragma unroll 1
    for (uint32_t i = 0; i < n; i++) total += __popc(oldL[list[i]].aux);
    if (total == 0) return 0;
    if (S.nNodes + total > Y.node_cap) { S.status = PBSC_WALK_HEAVY; return 0; }
    uint32_t m = 0;
    #pragma unroll 1
    for (uint32_t i = 0; i < n; i++)
    {
        const Leaf& parent = oldL[list[i]];
        const uint32_t mask = parent.aux;
        const uint32_t p_ring = parent.ring;
        if (!mask) { ring_release(S, Y, p_ring); continue; }   // not extended: nobody inherits its ring
        const uint32_t p_node = parent.node;
        uint32_t j = 0;
        #pragma unroll 1
        for (int b = 0; b < 4; b++)
        {
            if (!((mask >> b) & 1)) continue;
            Leaf* c = newL + 4 * i + b;
            const uint32_t node = S.nNodes++;
            c->node = node;
            nodes[node] = (p_node << 2) | (uint32_t)b;
            if (j > 0)
            {
                // createChild copies both error-rate records
                // (the copy itself is left to whichever lane prunes this child: aux = parent's ring + 1)
                const int slot = ring_take(S, Y);
                if (slot < 0) { S.status = PBSC_WALK_HEAVY; return 0; }
                c->ring = (uint16_t)slot;
                c->aux = p_ring + 1;
            }
            cl[m++] = (uint8_t)(4 * i + b);
            j++;
        }
    }
    return m;
}

// first position of the query whose idmer has `key`, from the hash (no duplicate idmers in this query)
__device__ __forceinline__ int hash_find(const Lane& c, uint32_t key)
{
    uint32_t h = (key * 2654435761u) & c.hashMask;
    #pragma unroll 1
    for (;;)
    {
        const uint64_t e = c.hash[h];
        if (e == ~0ull) return -1;
        if ((uint32_t)(e >> 32) == key) return (int)(uint32_t)e;
        h = (h + 1) & c.hashMask;
    }
}

// out-of-line IEEE division and 64-bit remainder: each expands to a few hundred bytes of SASS
static __device__ __noinline__ double ddiv_ool(double a, double b) { return __ddiv_rn(a, b); }
static __device__ __noinline__ uint64_t urem_ool(uint64_t a, uint64_t b) { return a % b; }

// PrunedBySeedSupport + isSupportedByNewSeed + computeErrorRate (LongReadCorrectByOverlap.cpp:491-664) for ONE new leaf, by
// whichever lane the pool gave it to (c.curLen and c.level are the walk's values after curLen++ and before level++).  A leaf
// whose local error rate is too high is only marked dead; its owner releases the ring (finish_level).
__device__ __forceinline__ void prune_leaf(const ExtParamsDev& P, const Layout& Y, const Lane& c, uint32_t j)
{
    const uint64_t seedSize = (uint64_t)P.seed_size;
    const uint64_t curLen = c.curLen;
    const uint64_t currSeedIdx = curLen - seedSize;
    const uint64_t indelOffset = seedSize + c.maxIndel;
    const uint64_t smallSeedIdx = currSeedIdx <= indelOffset ? 0 : currSeedIdx - indelOffset;
    const uint64_t largeSeedIdx = (currSeedIdx + indelOffset) >= ((uint64_t)c.qlen - seedSize) ? ((uint64_t)c.qlen - seedSize) : currSeedIdx + indelOffset;
    const uint32_t keyMask = (1u << (2 * P.seed_size)) - 1u;
    Leaf& L = new_bank(c, Y)[new_list(c, Y)[j]];
    bool found = false;
    const uint64_t d = curLen - (uint64_t)L.lastOverlapLen;
    if (d > seedSize || d <= 1)
    {
        const uint64_t preSeedIdx = L.lastSeedIdx;
        const uint64_t seedIdxOffset = (uint64_t)L.lastOverlapLen < curLen - seedSize ? seedSize : curLen - (uint64_t)L.lastOverlapLen;
        const uint64_t startSeedIdx = max(smallSeedIdx, (uint64_t)L.lastSeedIdx + seedIdxOffset);
        const bool fV = L.f_hi > L.f_lo, rV = L.r_hi > L.r_lo;
        const uint32_t keyF = (uint32_t)(L.rt_hi >> (64 - 2 * P.seed_size));
        if (!c.dup)
        {
            // every idmer of the query is unique: both result lists hold the same single position
            if (fV || rV)
            {
                const int p = hash_find(c, keyF);
                if (p >= 0 && (uint64_t)p >= startSeedIdx && (uint64_t)p <= largeSeedIdx)
                {
                    if (abs(p - (int)currSeedIdx) < 10000) L.lastSeedIdx = (uint32_t)p;   // minIdxDiff starts at 10000 (:580)
                    L.lastOverlapLen = (uint32_t)curLen;
                    found = true;
                }
            }
        }
        else
        {
            uint32_t nf = 0, nr = 0, gf = 0, gr = 0;
            if (fV) gf = group_find(c.sF, c.n9F, keyF, nf);
            if (rV) gr = group_find(c.sR, c.n9R, keyMask - keyF, nr);
            int minIdxDiff = 10000;
            const uint32_t lim = max(nf, nr);
            #pragma unroll 1
            for (uint32_t i = 0; i < lim; i++)
            {
                uint64_t v = 0;
                bool hit = false;
                if (i < nf) { v = (uint32_t)c.sF[gf + i]; hit = v >= startSeedIdx && v <= largeSeedIdx; }
                if (!hit && i < nr) { v = (uint32_t)c.sR[gr + i]; hit = v >= startSeedIdx && v <= largeSeedIdx; }
                if (hit)
                {
                    const int diff = abs((int)v - (int)currSeedIdx);
                    if (diff < minIdxDiff) { L.lastSeedIdx = (uint32_t)v; minIdxDiff = diff; }
                    L.lastOverlapLen = (uint32_t)curLen;
                    found = true;
                }
            }
        }
        if (found)
        {
            L.totalSeeds++;
            if (currSeedIdx + (uint64_t)(int64_t)L.seedOff - preSeedIdx > seedSize) L.redeem = __dadd_rn(L.redeem, P.redeem_a);
            L.seedOff = (int)L.lastSeedIdx - (int)currSeedIdx;
        }
        else
        {
            const uint64_t v = currSeedIdx + (uint64_t)(int64_t)L.seedOff - (uint64_t)L.lastSeedIdx;
            if (urem_ool(v, seedSize) == 1) { /* numOfErrors++ : never read */ }
            else if (v > seedSize - 1) L.redeem = __dadd_rn(L.redeem, P.redeem_b);
        }
    }
    else L.redeem = __dadd_rn(L.redeem, P.redeem_b);
    // computeErrorRate
    double matchedLen = __dsub_rn(__dadd_rn((double)L.totalSeeds, (double)seedSize), 1.0);
    matchedLen = __dadd_rn(matchedLen, L.redeem);
    const double totalLen = (double)curLen;
    double err = ddiv_ool(__dsub_rn(totalLen, matchedLen), totalLen);
    double* ring = rings_of(c, Y) + (size_t)L.ring * RING_LEN;
    if (L.aux)
    {
        // second or later child of its parent: createChild's copy of both error-rate records (FMIndexWalk/SAINode.cpp:166-189).
        // The first child shares the parent's ring and may already have written this level's entry into it; the same slot of
        // the copy is overwritten just below, every other slot is older.
        const double* src = rings_of(c, Y) + (size_t)(L.aux - 1) * RING_LEN;
        const int have = min((int)c.level, RING_LEN);   // GlobalErrorRateRecord holds `level` entries so far
        #pragma unroll 4
        for (int x = 0; x < have; x++) ring[x] = src[x];
        L.aux = 0;
    }
    ring[c.level % RING_LEN] = err;
    L.global_err = err;
    if (c.level + 1 >= (uint32_t)RING_LEN)
    {
        const double old = ring[(c.level + 1) % RING_LEN];
        err = ddiv_ool(__dsub_rn(__dmul_rn(err, totalLen), __dmul_rn(old, __dsub_rn(totalLen, (double)RING_LEN))), (double)RING_LEN);
    }
    L.local_err = err;
    if (err > P.walk_error_rate) L.alive = 0;
}

// isTerminated, the search (LongReadCorrectByOverlap.cpp:825-860) for ONE new leaf: the last terminal interval that holds
// the leaf's interval on either strand, at or after the one it matched before; left in `aux` as index + 1 (0 = none)
__device__ __forceinline__ void term_leaf(const Layout& Y, const Lane& c, uint32_t j)
{
    Leaf& L = new_bank(c, Y)[new_list(c, Y)[j]];
    if (!L.alive) { L.aux = 0; return; }
    const bool fV = L.f_hi > L.f_lo, rV = L.r_hi > L.r_lo;
    const uint64_t f_lo = L.f_lo, f_hi = L.f_hi, r_lo = L.r_lo, r_hi = L.r_hi;
    int ilast = -1;
    #pragma unroll 1
    for (int i = max(L.res_second, 0); i < (int)c.nTerm; i++)
    {
        const Interval tf = c.termF[i], tr = c.termR[i];
        const bool ft = fV && tf.valid() && f_lo >= tf.lo && f_hi <= tf.hi;
        const bool rt = rV && tr.valid() && r_lo >= tr.lo && r_hi <= tr.hi;
        if (ft || rt) ilast = i;
    }
    L.aux = (uint32_t)(ilast + 1);
}

// End of a level, by the walk's own lane: isTerminated's bookkeeping (one result slot per leaf, overwritten while the leaf
// keeps terminating, LongReadCorrectByOverlap.cpp:861-875), then the surviving children become the next level's leaves: the
// banks swap roles (no copy; the walk kernel's DRAM traffic is mostly this scratch), the child list loses the dead entries.
static __device__ __noinline__ void finish_level(Lane& S, const Layout& Y, const ExtParamsDev& P, uint32_t m, bool check_term)
{
    Leaf* newL = new_bank(S, Y);
    uint8_t* cl = new_list(S, Y);
    WalkResult* res = res_of(S, Y);
    if (check_term)
    {
        #pragma unroll 1
        for (uint32_t j = 0; j < m; j++)
        {
            Leaf& L = newL[cl[j]];
            if (!L.alive || L.aux == 0) continue;
            const int ilast = (int)L.aux - 1;
            int slot = L.res_first;
            if (slot == -1)
            {
                if (S.nRes >= Y.cap.res) { S.status = PBSC_WALK_HEAVY; return; }
                slot = (int)++S.nRes;
            }
            WalkResult r; r.err = L.global_err; r.node = L.node; r.i = ilast; r.depth = (uint32_t)S.curLen; r.pad = 0;
            res[slot - 1] = r;
            L.res_first = slot; L.res_second = ilast;
        }
    }
    uint32_t nn = 0;
    #pragma unroll 1
    for (uint32_t j = 0; j < m; j++)
    {
        const uint8_t slot = cl[j];
        const Leaf& L = newL[slot];
        if (!L.alive) { ring_release(S, Y, L.ring); continue; }
        if (nn != j && nn < Y.cap.old_cap) cl[nn] = slot;
        nn++;
    }
    S.flip ^= 1;   // the banks and the lists change roles
    S.n = nn;
    // more live leaves than this pass carries, and the reference's loop would go on: hand the walk over
    if (nn > Y.cap.old_cap && nn <= (uint32_t)P.max_leaves && S.curLen <= S.maxLength) S.status = PBSC_WALK_HEAVY;
}

// ------------------------------------------------------------------------------------------------------------
// Per-task setup record, produced once by setup_tasks_kernel and read by the level loop.  Everything in
// LongReadSelfCorrectByOverlap's constructor (LongReadCorrectByOverlap.cpp:17-95,106-152) lands here.
// ------------------------------------------------------------------------------------------------------------
struct __align__(16) SetupHdr
{
    uint64_t rf_lo, rf_hi, rr_lo, rr_hi;   // root fwd / rvc intervals
    uint64_t maxLength, minLength, maxIndel;
    uint32_t qlen, k, trgLen, nTerm, n5, hashMask, n9F, n9R;
    int32_t dis;
    uint32_t dup;
    int32_t status0;    // 0, PBSC_WALK_UNSUPPORTED or PBSC_WALK_OVERFLOW decided before any walking
    // result of a successful light walk, materialised later by materialize_kernel
    uint32_t res_node, res_depth, n_nodes;
    int32_t res_i;
    uint32_t pad2[3];
    uint64_t node_off;   // where this walk's label tree was copied in the node pool
    uint64_t pad3;
};
static_assert(sizeof(SetupHdr) % 16 == 0, "SetupHdr must keep 16-byte alignment");

struct SetupView
{
    SetupHdr* hdr;
    uint8_t* q;
    uint16_t* start4; uint16_t* cur4; uint16_t* pos4;
    Interval* termF; Interval* termR;
    uint64_t* hash;
    uint64_t* sF; uint64_t* sR;
};

__host__ __device__ inline size_t setup_record_bytes(uint32_t qlen, uint32_t trgLen, int min_overlap, int s9)
{
    const uint32_t nTerm = trgLen >= (uint32_t)min_overlap ? trgLen - min_overlap + 1 : 0;
    const uint32_t n9 = qlen >= (uint32_t)s9 ? qlen - s9 + 1 : 0;
    size_t b = sizeof(SetupHdr);
    b += align_up(qlen, 16);
    b += 2 * align_up(sizeof(uint16_t) * 258, 16) + align_up(sizeof(uint16_t) * (size_t)qlen, 16);
    b += sizeof(Interval) * (size_t)nTerm * 2;
    b += sizeof(uint64_t) * (size_t)pow2_ceil(2 * (n9 ? n9 : 1));
    b += sizeof(uint64_t) * (size_t)n9 * 2;
    return align_up(b, 128);
}

__device__ inline void setup_view(uint8_t* base, uint32_t qlen, uint32_t trgLen, int min_overlap, int s9, SetupView& v)
{
    const uint32_t nTerm = trgLen >= (uint32_t)min_overlap ? trgLen - min_overlap + 1 : 0;
    const uint32_t n9 = qlen >= (uint32_t)s9 ? qlen - s9 + 1 : 0;
    uint8_t* p = base;
    v.hdr = (SetupHdr*)p; p += sizeof(SetupHdr);
    v.q = p; p += align_up(qlen, 16);
    v.start4 = (uint16_t*)p; p += align_up(sizeof(uint16_t) * 258, 16);
    v.cur4 = (uint16_t*)p; p += align_up(sizeof(uint16_t) * 258, 16);
    v.pos4 = (uint16_t*)p; p += align_up(sizeof(uint16_t) * (size_t)qlen, 16);
    v.termF = (Interval*)p; p += sizeof(Interval) * (size_t)nTerm;
    v.termR = (Interval*)p; p += sizeof(Interval) * (size_t)nTerm;
    v.hash = (uint64_t*)p; p += sizeof(uint64_t) * (size_t)pow2_ceil(2 * (n9 ? n9 : 1));
    v.sF = (uint64_t*)p; p += sizeof(uint64_t) * (size_t)n9;
    v.sR = (uint64_t*)p;
}

// v.q[0..qlen) already holds the query.  Fills the rest of the record.
static __device__ __noinline__ void setup_task(const FmIndexDev& idx, const ExtParamsDev& P, SetupView& v, uint32_t qlen, uint32_t k, int32_t dis,
                                               uint32_t trgLen, uint32_t outCap)
{
    SetupHdr H;
    memset(&H, 0, sizeof H);
    H.qlen = qlen; H.k = k; H.trgLen = trgLen; H.dis = dis;
    if (trgLen < (uint32_t)P.min_overlap || k < (uint32_t)P.seed_size || k + 3 > 64 || trgLen - P.min_overlap + 1 > TERM_CAP || qlen != k + (uint32_t)dis + trgLen)
    { H.status0 = PBSC_WALK_UNSUPPORTED; *v.hdr = H; return; }
    // LongReadCorrectByOverlap.cpp:54-58,77-79
    H.maxIndel = dis > 100 ? (uint64_t)__dmul_rn((double)dis, 0.2) : 20;
    H.maxLength = (uint64_t)__dadd_rn(__dmul_rn(1.2, (double)(dis + 10)), (double)(2 * (uint64_t)k));
    H.minLength = (uint64_t)__dadd_rn(__dmul_rn(0.8, (double)(dis - 20)), (double)(2 * (uint64_t)k));
    H.nTerm = trgLen - P.min_overlap + 1;
    if (H.maxLength + trgLen + 8 > outCap) { H.status0 = PBSC_WALK_OVERFLOW; *v.hdr = H; return; }
    const uint8_t* q = v.q;
    const uint8_t* trg = q + k + dis;
    const int s9 = P.seed_size;
    const uint32_t n9 = qlen - s9 + 1;
    // terminal intervals (:82-88)
    #pragma unroll 1
    for (uint32_t i = 0; i < H.nTerm; i++)
    {
        Interval f, r;
        const uint8_t* w = trg + i;
        both_strands(idx, [&](int j) { return (int)w[j]; }, P.min_overlap, f, r);
        v.termF[i] = f; v.termR[i] = r;
    }
    // query idmers into the hash; a repeated idmer switches this walk to the exact sorted lists
    H.hashMask = pow2_ceil(2 * n9) - 1;
    {
        ulonglong2* hz = reinterpret_cast<ulonglong2*>(v.hash);
        #pragma unroll 1
        for (uint32_t h = 0; h <= H.hashMask / 2; h++) hz[h] = make_ulonglong2(~0ull, ~0ull);
    }
    bool dup = false;
    {
        uint32_t key = 0;
        const uint32_t keyMask = (1u << (2 * s9)) - 1u;
        #pragma unroll 1
        for (int j = 0; j < s9 - 1; j++) key |= (uint32_t)q[j] << (2 * (j + 1));
        #pragma unroll 1
        for (uint32_t p = 0; p < n9 && !dup; p++)
        {
            key = ((key >> 2) | ((uint32_t)q[p + s9 - 1] << (2 * (s9 - 1)))) & keyMask;
            uint32_t h = (key * 2654435761u) & H.hashMask;
            #pragma unroll 1
            for (;;)
            {
                const uint64_t e = v.hash[h];
                if (e == ~0ull) { v.hash[h] = ((uint64_t)key << 32) | p; break; }
                if ((uint32_t)(e >> 32) == key) { dup = true; break; }
                h = (h + 1) & H.hashMask;
            }
        }
    }
    H.dup = dup ? 1 : 0;
    if (dup)
    {
        // buildOverlapbyFMindex (:127-152): only idmers with a valid interval on a strand enter that strand's list
        #pragma unroll 1
        for (uint32_t p = 0; p < n9; p++)
        {
            uint32_t keyF = 0;
            #pragma unroll 1
            for (int j = 0; j < s9; j++) keyF |= (uint32_t)q[p + j] << (2 * j);
            bool fv, rv;
            if (idx.idmer_valid != nullptr && idx.idmer_len == s9) { const uint8_t b = __ldg(idx.idmer_valid + keyF); fv = b & 1; rv = (b & 2) != 0; }
            else
            {
                Interval f, r;
                const uint8_t* w = q + p;
                both_strands(idx, [&](int j) { return (int)w[j]; }, s9, f, r);
                fv = f.valid(); rv = r.valid();
            }
            if (fv) v.sF[H.n9F++] = ((uint64_t)keyF << 32) | p;
            if (rv) v.sR[H.n9R++] = ((uint64_t)(((1u << (2 * s9)) - 1u) - keyF) << 32) | p;
        }
        sort_desc_ool(v.sF, (long)H.n9F);
        sort_desc_ool(v.sR, (long)H.n9R);
    }
    // query 5-mers, newest base most significant
    H.n5 = qlen >= 5 ? qlen - 4 : 0;
    // positions grouped by their 4-mer (counting sort): ismatchedbykmer asks, for a leaf ending in some 4-mer, which next
    // bases follow an occurrence of that 4-mer near the current length
    {
        uint4* z = reinterpret_cast<uint4*>(v.start4);
        #pragma unroll 1
        for (int x = 0; x < 33; x++) z[x] = make_uint4(0, 0, 0, 0);
        auto c4 = [&](uint32_t p) { return (uint32_t)(q[p] | (q[p + 1] << 2) | (q[p + 2] << 4) | (q[p + 3] << 6)); };
        #pragma unroll 1
        for (uint32_t p = 0; p < H.n5; p++) v.start4[c4(p) + 1]++;
        #pragma unroll 1
        for (int c = 1; c <= 256; c++) v.start4[c] = (uint16_t)(v.start4[c] + v.start4[c - 1]);
        #pragma unroll 1
        for (int c = 0; c < 256; c++) v.cur4[c] = v.start4[c];
        #pragma unroll 1
        for (uint32_t p = 0; p < H.n5; p++) v.pos4[v.cur4[c4(p)]++] = (uint16_t)p;
    }
    // root intervals (:106-124)
    {
        Interval f, r;
        both_strands(idx, [&](int j) { return (int)q[j]; }, (int)k, f, r);
        H.rf_lo = f.lo; H.rf_hi = f.hi; H.rr_lo = r.lo; H.rr_hi = r.hi;
    }
    *v.hdr = H;
}

// start a walk from its setup record
__device__ __forceinline__ void begin_walk(Lane& S, const Layout& Y, const ExtParamsDev& P, uint8_t* lane_scratch, const SetupView& v, uint64_t minSA)
{
    const SetupHdr H = *v.hdr;
    S.base = lane_scratch;
    S.q = v.q; S.start4 = v.start4; S.pos4 = v.pos4; S.termF = v.termF; S.termR = v.termR; S.hash = v.hash; S.sF = v.sF; S.sR = v.sR;
    S.status = H.status0;
    S.qlen = H.qlen; S.k = H.k; S.maxOverlap = H.k + 2; S.trgLen = H.trgLen; S.minSA = (uint32_t)minSA; S.thr = (uint32_t)minSA;
    S.maxIndel = (uint32_t)H.maxIndel; S.maxLength = (uint32_t)H.maxLength; S.minLength = (uint32_t)H.minLength;
    S.curLen = S.curK = H.k;
    S.nTerm = H.nTerm; S.n5 = H.n5; S.hashMask = H.hashMask; S.n9F = H.n9F; S.n9R = H.n9R; S.dup = H.dup != 0;
    S.nNodes = 1; S.nRes = 0; S.level = 1; S.phase = 0; S.flip = 0; S.selNew = 0; S.minErr = 0.0;
    S.n = 0;
    if (S.status) return;
    S.nFree = 0;
    S.nFresh = 1;
    const uint8_t* q = S.q;
    const uint32_t k = H.k;
    const int s9 = P.seed_size;
    Leaf R;
    R.f_lo = H.rf_lo; R.f_hi = H.rf_hi; R.r_lo = H.rr_lo; R.r_hi = H.rr_hi;
    R.redeem = 0; R.local_err = 0; R.global_err = 0;
    R.rt_hi = R.rt_lo = 0;
    #pragma unroll 1
    for (uint32_t j = 0; j < k; j++) tail_push(R.rt_hi, R.rt_lo, q[j]);
    R.lastOverlapLen = k; R.lastSeedIdx = k - s9; R.totalSeeds = k - s9 + 1; R.seedOff = 0;
    R.res_first = -1; R.res_second = -1;
    R.kmerFreq = (int)((int64_t)(R.f_hi - R.f_lo) + (int64_t)(R.r_hi - R.r_lo));
    R.tailLetter = q[k - 1];
    uint32_t tc = 0;
    #pragma unroll 1
    for (int j = (int)k - 1; j >= 0 && q[j] == R.tailLetter; j--) tc++;
    R.tailCount = tc;
    R.node = 0; R.ring = 0; R.alive = 1; R.aux = 0;
    old_bank(S, Y)[0] = R;
    old_list(S, Y)[0] = 0;
    nodes_of(S, Y)[0] = 0;
    rings_of(S, Y)[0] = 0.0;
    S.n = 1;
}

// does extendLeaves start by cutting the k-mer back to maxOverlap (LongReadCorrectByOverlap.cpp:242-243)?
__device__ __forceinline__ bool needs_refine(const Lane& S) { return S.phase == 0 && S.curK > S.maxOverlap; }

// does extendOverlap's loop (LongReadCorrectByOverlap.cpp:161) run another level?
__device__ __forceinline__ bool walk_continues(const Lane& S, const ExtParamsDev& P)
{
    return S.status == 0 && S.n > 0 && S.n <= (uint32_t)P.max_leaves && S.curLen <= S.maxLength;
}

// SelectFreqsOfrange (LongReadCorrectByOverlap.cpp:281-331) for ONE leaf, by whichever lane the pool gave it to: the
// frequency of the leaf's last LB .. LB + extra bases, newest base first on BWT and complemented on RBWT; the walk-wide maxima
// are collected in the owner's record (shared-memory atomics).
__device__ __forceinline__ void select_leaf(const FmIndexDev& idx, const Layout& Y, Lane& c, uint32_t i)
{
    const Leaf& L = (c.selNew ? new_bank(c, Y) : old_bank(c, Y))[(c.selNew ? new_list(c, Y) : old_list(c, Y))[i]];
    const uint64_t hi = L.rt_hi, lo = L.rt_lo;
    const int LB = (int)c.selLB, extra = (int)c.selExtra;
    Interval a, b;   // a on BWT, b on RBWT
    int d;
    if (idx.prefix != nullptr && LB >= idx.k0)
    {
        uint64_t key = 0;
        #pragma unroll 1
        for (int j = 0; j < idx.k0; j++) key |= (uint64_t)(3 - tail_base(hi, lo, j)) << (2 * j);
        Interval f, r;
        prefix_lookup(idx, key, f, r);
        a = r; b = f;
        d = idx.k0;
    }
    else
    {
        const int ch = tail_base(hi, lo, 0);
        a = init_interval(idx.t[PBSC_BWT], ch);
        b = init_interval(idx.t[PBSC_RBWT], 3 - ch);
        d = 1;
    }
    #pragma unroll 1
    for (; d < LB && (a.valid() || b.valid()); d++)
    {
        const int ch = tail_base(hi, lo, d);
        if (a.valid()) a = update1(idx.t[PBSC_BWT], a, ch);
        if (b.valid()) b = update1(idx.t[PBSC_RBWT], b, 3 - ch);
    }
    atomicMax(&c.selMx[0], (int)((int64_t)a.size() + (int64_t)b.size()));
    #pragma unroll 1
    for (int e = 1; e <= extra; e++)
    {
        const int ch = tail_base(hi, lo, LB - 1 + e);
        if (a.valid()) a = update1(idx.t[PBSC_BWT], a, ch);
        if (b.valid()) b = update1(idx.t[PBSC_RBWT], b, 3 - ch);
        atomicMax(&c.selMx[e], (int)((int64_t)a.size() + (int64_t)b.size()));
    }
}

// refineSAInterval(selK) of one leaf of the selected bank (the refinement that follows SelectFreqsOfrange)
__device__ __forceinline__ void reselect_leaf(const FmIndexDev& idx, const Layout& Y, const Lane& c, uint32_t i)
{
    Leaf& L = (c.selNew ? new_bank(c, Y) : old_bank(c, Y))[(c.selNew ? new_list(c, Y) : old_list(c, Y))[i]];
    const uint64_t hi = L.rt_hi, lo = L.rt_lo;
    const int K = (int)c.selK;
    Interval f, r;
    both_strands(idx, [&](int j) { return tail_base(hi, lo, K - 1 - j); }, K, f, r);
    L.f_lo = f.lo; L.f_hi = f.hi; L.r_lo = r.lo; L.r_hi = r.hi;
}

// One step of extendOverlap's loop (:161-197) is spread over stages that the level loop of walk_levels_kernel runs for all 32
// lanes of a warp together (pooled = the leaves of all lanes dealt out over the warp; own = each lane for its walk):
//   refine_pool      pooled   refineSAInterval(maxOverlap) that opens extendLeaves (:242-243)
//   filter_leaves    own      attempToExtend's error-rate filter
//   probe_leaf       pooled   getFMIndexExtensions + the children's leaf records
//   adopt_children   own      node numbers, history rings
//   level_middle     own      extendLeaves' bookkeeping and fallbacks (below)
//   select_leaf      pooled   SelectFreqsOfrange's searches (only walks whose frequencies are insufficient)
//   level_select     own      the reduced k-mer size
//   reselect_leaf    pooled   refineSAInterval to it
//   prune_leaf       pooled   PrunedBySeedSupport
//   term_leaf        pooled   isTerminated's interval search
//   finish_level     own      results, next level's leaf list
// extendLeaves' fallbacks (reduce the k-mer, then lower the SA threshold, :249-262) are rare per walk but long; run inline
// they would stall the other lanes of the warp, so a walk that needs them spends extra iterations in phases 1 and 2 while its
// neighbours keep extending:
//   phase 0: normal level            probes with thr
//   phase 1: after the k-mer was reduced (select + refine on the old leaves happened at the end of the previous iteration)
//   phase 2: last resort             probes with thr - 1
// Returns true when the level produced children that go on to prune / terminal search / finish_level.
// `sel` = how many leaves go through SelectFreqsOfrange + refineSAInterval before the level goes on (selNew says which bank);
// level_select then turns the pooled maxima into the reduced k-mer size.
static __device__ __noinline__ bool level_middle(Lane& S, const Layout& Y, const ExtParamsDev& P, uint32_t m, uint32_t& sel)
{
    uint32_t cnt = 0;
    sel = 0;
    if (m > 0)
    {
        S.curLen++;
        S.curK++;
        S.phase = 0;
        if (insufficient(P, new_bank(S, Y), new_list(S, Y), m)) { S.selNew = 1; cnt = m; }
    }
    else if (S.phase == 0) { S.selNew = 0; cnt = S.n; S.phase = 1; }   // level 1: reduce the k-mer size, then retry
    else if (S.phase == 1) { S.phase = 2; return false; }               // level 2: retry with the lower threshold
    else { S.n = 0; return false; }                                     // newLeaves stays empty: the walk ends
    if (cnt)
    {
        const uint32_t LB = max(S.curK - 2, (uint32_t)P.min_overlap);
        S.selLB = LB; S.selExtra = S.curK - LB;
        S.selMx[0] = S.selMx[1] = S.selMx[2] = 0;
        sel = cnt;
    }
    return m > 0;
}

// the decision of SelectFreqsOfrange (:318-330) from the maxima the pool collected
__device__ __forceinline__ void level_select(Lane& S, const ExtParamsDev& P)
{
    const uint32_t LB = S.selLB, UB = LB + S.selExtra;
    uint32_t R = UB;
    if (S.selMx[0] - P.freq_int[LB] < 5) R = LB;
    else
    {
        #pragma unroll 1
        for (uint32_t e = 1; e <= S.selExtra; e++) if (S.selMx[e] - P.freq_int[LB + e] < 5) { R = LB + e; break; }
    }
    S.selK = R;
    S.curK = R;
}

#define PBSC_TASK_MATERIALIZE 2   // walk succeeded; the merged sequence is still to be written from the saved label tree

// extendOverlap's return value and findTheBestPath (:199-236).  A successful walk does not chase its label chain here
// (a serial pointer chase by one lane while 31 wait): it copies its label tree to the node pool and leaves the rest to
// materialize_kernel.
static __device__ __noinline__ int finish_walk(Lane& S, const Layout& Y, const ExtParamsDev& P, SetupHdr* hdr, uint32_t* nodepool, unsigned long long* pool_used,
                                               uint64_t pool_cap, bool direct, uint8_t* out, uint32_t outCap, uint32_t* outLen)
{
    if (S.status) return S.status;
    const WalkResult* res = res_of(S, Y);
    const uint32_t* nodes = nodes_of(S, Y);
    if (S.nRes > 0)
    {
        double best = 1.0;
        int bi = -1;
        #pragma unroll 1
        for (uint32_t i = 0; i < S.nRes; i++) { const double e = res[i].err; if (e < best) { best = e; bi = (int)i; } }
        if (bi < 0) return PBSC_WALK_NO_PATH;
        const WalkResult r = res[bi];
        if (direct)
        {
            // a pass of few, long walks (most lanes of the warp only help): write the merged sequence here, from the lane's own
            // label tree, instead of parking a tree of up to node_cap nodes in the shared pool
            const uint8_t* q = S.q;
            const uint8_t* trg = q + S.qlen - S.trgLen;
            const uint32_t chain = r.depth - S.k;
            const uint32_t tailFrom = (uint32_t)r.i + (uint32_t)P.min_overlap;
            const uint32_t tailLen = S.trgLen > (uint32_t)P.min_overlap ? S.trgLen - tailFrom : 0;
            const uint32_t len = r.depth + tailLen;
            if (len > outCap) return PBSC_WALK_OVERFLOW;
            #pragma unroll 1
            for (uint32_t x = 0; x < S.k; x++) out[x] = q[x];
            #pragma unroll 1
            for (uint32_t x = 0; x < tailLen; x++) out[r.depth + x] = trg[tailFrom + x];
            uint32_t node = r.node;
            #pragma unroll 1
            for (uint32_t x = 0; x < chain; x++) { const uint32_t w = nodes[node]; out[r.depth - 1 - x] = (uint8_t)(w & 3); node = w >> 2; }
            *outLen = len;
            return 1;
        }
        const uint32_t nn = (S.nNodes + 3u) & ~3u;
        const uint64_t off = atomicAdd(pool_used, (unsigned long long)nn);
        if (off + nn > pool_cap) return PBSC_OVF_POOL;
        const uint4* src = reinterpret_cast<const uint4*>(nodes);
        uint4* dst = reinterpret_cast<uint4*>(nodepool + off);
        #pragma unroll 4
        for (uint32_t x = 0; x < nn / 4; x++) dst[x] = src[x];
        hdr->res_node = r.node; hdr->res_depth = r.depth; hdr->res_i = r.i; hdr->n_nodes = S.nNodes; hdr->node_off = off;
        return PBSC_TASK_MATERIALIZE;
    }
    if (S.n == 0) return -1;
    if (S.curLen > S.maxLength) return -2;
    if (S.n > (uint32_t)P.max_leaves) return -3;
    return -4;
}

// merged sequence of a successful light walk: beginningkmer + labels along the best leaf's chain + rest of the target
__device__ __forceinline__ int materialize(const SetupView& v, int min_overlap, const uint32_t* nodepool, uint8_t* out, uint32_t outCap, uint32_t* outLen)
{
    const SetupHdr H = *v.hdr;
    const uint8_t* q = v.q;
    const uint8_t* trg = q + H.qlen - H.trgLen;
    const uint32_t k = H.k;
    const uint32_t chain = H.res_depth - k;
    const uint32_t tailFrom = (uint32_t)H.res_i + min_overlap;
    const uint32_t tailLen = H.trgLen > (uint32_t)min_overlap ? H.trgLen - tailFrom : 0;
    const uint32_t len = H.res_depth + tailLen;
    if (len > outCap) return PBSC_WALK_OVERFLOW;
    #pragma unroll 1
    for (uint32_t x = 0; x < k; x++) out[x] = q[x];
    #pragma unroll 1
    for (uint32_t x = 0; x < tailLen; x++) out[H.res_depth + x] = trg[tailFrom + x];
    const uint32_t* nodes = nodepool + H.node_off;
    uint32_t node = H.res_node;
    #pragma unroll 1
    for (uint32_t x = 0; x < chain; x++) { const uint32_t w = nodes[node]; out[H.res_depth - 1 - x] = (uint8_t)(w & 3); node = w >> 2; }
    *outLen = len;
    return 1;
}

}  // namespace tw
}  // namespace pbsc
#endif
