// pbsc_extend_thread.cu — FM-extend phase, thread-per-walk engine with speculative pair scheduling.
//
// PacBioSelfCorrectionProcess::initCorrect (PacBio/PacBioSelfCorrectionProcess.cpp:56-157) walks the seed pairs of a
// read in order because the source of walk i+1 is the piece corrected by walk i.  But the only things walk i+1
// takes from that piece are its last k bases and its length, and they almost always equal what the raw read
// says (the walk result ends with the target seed itself; SURVEY.md H1 measured 99.9 %).  So:
//   1. make_spec_tasks_kernel  one task per consecutive seed pair, source taken from the raw seed (speculation)
//   2. walk_tasks_kernel       every task is an independent walk: one thread each (pbsc_walk_thread.cuh)
//   3. stitch_kernel           thread per read replays initCorrect in order with the ACTUAL source state; a task whose
//                              speculated inputs (k, strand swap, source k-mer) match is consumed, otherwise the read
//                              files one exact request and stalls
//   4. walk the requests, stitch again, until no read is stalled (also how -n/--next-target > 1 look-aheads run)
// Results are identical to the sequential chain by construction: a walk is a pure function of its inputs.
#include <cub/cub.cuh>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include "pbsc_batch.cuh"
#include "pbsc_task.cuh"
#include "pbsc_dp.cuh"

namespace pbsc {

constexpr int TW_BLOCK = 128;
struct ReadState
{
    int32_t t, next, started, done, rstatus, firstType;
    int32_t srcEnd, srcEndBest, srcRepeat;
    uint32_t nPieces;
    int64_t srcLen;
    uint64_t plen;
    int32_t pending_trg;   // seed index the pending request targets
    int32_t srcFreq;       // source.maxFixedMerFreq
    int32_t srcStart;      // source.seedStartPos: SeedFeature::append takes it from the target (SeedFeature.h:22-33)
    uint32_t nLog;         // --debugseed: failed walks logged so far
    // DP fallback result of the current target's first (next == 0) walk, kept while the look-ahead walks run
    int32_t dpStatus0;
    uint32_t dpLen0;
    uint64_t dpOff0;
    pbsc_read_stats st;
};

__device__ __forceinline__ void pack_src(const uint8_t* bases, int k, uint64_t& hi, uint64_t& lo)
{
    hi = lo = 0;
    for (int j = 0; j < k; j++) tail_push(hi, lo, bases[j]);
}

// extendKmerSize / isFromRtoU of correctByFMExtension (PacBioSelfCorrectionProcess.cpp:163-175)
__device__ __forceinline__ void pair_inputs(int srcEndBest, int srcRepeat, int64_t srcLen, const pbsc_seed& tg, int start_kmer, int& k, bool& rtou)
{
    k = min(srcEndBest, tg.start_best_k) - 2;
    if (srcRepeat || tg.is_repeat)
    {
        k = (int)min(srcLen, (int64_t)tg.len);
        k = min(k, start_kmer + 2);
    }
    rtou = srcRepeat && !tg.is_repeat;
}

__device__ __forceinline__ uint32_t task_out_cap(int dis, int k, int trgLenWalk)
{
    return (uint32_t)((uint64_t)(1.2 * (double)(dis + 10)) + 2 * (uint64_t)k + (uint64_t)trgLenWalk + 24);
}

// one speculative task per seed t >= 1 of every read: source = raw seed t-1
__global__ void make_spec_tasks_kernel(uint64_t n_reads, const uint8_t* __restrict__ codes, const uint64_t* __restrict__ offsets,
                                       const pbsc_seed* __restrict__ seeds, const uint64_t* __restrict__ region,
                                       const uint32_t* __restrict__ seed_count, const uint64_t* __restrict__ task_base,
                                       WalkTask* __restrict__ tasks, uint64_t* __restrict__ caps, uint64_t* __restrict__ rec_caps,
                                       int start_kmer, int min_overlap, int s9, int no_dp)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const pbsc_seed* sv = seeds + region[r];
    const uint32_t ns = seed_count[r];
    const uint8_t* read = codes + offsets[r];
    for (uint32_t t = 0; t < ns; t++)
    {
        WalkTask tk;
        memset(&tk, 0, sizeof tk);
        tk.read = (uint32_t)r;
        if (t >= 1)
        {
            const pbsc_seed s = sv[t - 1], tg = sv[t];
            int k; bool rtou;
            // source.seedLen is the length of the whole piece corrected so far (SeedFeature::append): only for the first pair is
            // it the raw seed's; later it is long, and where the walk k-mer reaches back beyond the seed (repeat seeds shorter
            // than startKmerLen + 2) the raw read is the best guess of the corrected bases in front of it
            pair_inputs(s.end_best_k, s.is_repeat, t == 1 ? (int64_t)s.len : ((int64_t)1 << 40), tg, start_kmer, k, rtou);
            tk.src_end = s.start + s.len - 1;
            tk.trg_start = tg.start; tk.trg_len = tg.len;
            tk.k = k; tk.rtou = rtou ? 1 : 0;
            tk.status = PBSC_TASK_PENDING;
            tk.freq_sum = s.max_fixed_freq + tg.max_fixed_freq;
            tk.dp_wanted = no_dp ? 0 : 1;
            if (k > 0 && k <= s.start + s.len && k <= 64)
            {
                pack_src(read + s.start + s.len - k, k, tk.src_hi, tk.src_lo);
                tk.valid = 1;
                tk.out_cap = task_out_cap(tg.start - tk.src_end - 1, k, rtou ? k : tg.len);
            }
        }
        caps[task_base[r] + t] = tk.out_cap;
        uint64_t rc = 0;
        if (tk.valid)
        {
            const int interval = tk.trg_start - tk.src_end - 1;
            const uint32_t trgLen = tk.rtou ? (uint32_t)tk.k : (uint32_t)tk.trg_len;
            rc = tw::setup_record_bytes((uint32_t)(tk.k + interval) + trgLen, trgLen, min_overlap, s9);
        }
        rec_caps[task_base[r] + t] = rc;
        tasks[task_base[r] + t] = tk;
    }
}

__global__ void set_out_offsets_kernel(uint64_t n_tasks, WalkTask* tasks, const uint64_t* off)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n_tasks) tasks[i].out_off = off[i];
}

// which idmers occur on which strand (replaces the 2 x s findInterval steps per query position that buildOverlapbyFMindex,
// LongReadCorrectByOverlap.cpp:127-152, spends on deciding whether an idmer enters a strand's list)
__global__ void idmer_valid_kernel(FmIndexDev idx, int s9, uint64_t n_keys, uint8_t* __restrict__ valid)
{
    const uint64_t key = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (key >= n_keys) return;
    Interval f, r;
    tw::both_strands(idx, [&](int j) { return (int)((key >> (2 * j)) & 3); }, s9, f, r);
    valid[key] = (uint8_t)((f.valid() ? 1 : 0) | (r.valid() ? 2 : 0));
}

// Second-guessing the speculation.  A pair whose walk failed and whose DP fallback produced a consensus hands its successor a
// source that ends with that consensus instead of the raw target seed; waiting for stitch_kernel to discover this costs one
// round per such pair along a read (14 rounds on config 2).  After every round of walks this kernel looks, for ALL pairs at
// once, at what each pair's current result would hand to its successor and, where that differs from what the successor was
// walked with, files an alternative task for the successor.  Alternatives are only ever hints: stitch_kernel still checks the
// exact inputs before it consumes any task.
__global__ void make_alt_tasks_kernel(uint64_t n_reads, const pbsc_seed* __restrict__ seeds, const uint64_t* __restrict__ region,
                                      const uint32_t* __restrict__ seed_count, const uint64_t* __restrict__ task_base, const WalkTask* __restrict__ spec,
                                      WalkTask* alt, const uint8_t* __restrict__ outpool, uint64_t alt_pool_base, int start_kmer, uint32_t* alt_list,
                                      unsigned int* n_alt)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const pbsc_seed* sv = seeds + region[r];
    const uint32_t ns = seed_count[r];
    const uint64_t b = task_base[r];
    for (uint32_t t = 1; t + 1 < ns; t++)
    {
        const WalkTask* e = (alt[b + t].valid && alt[b + t].status != PBSC_TASK_PENDING) ? &alt[b + t] : &spec[b + t];
        if (!e->valid || e->status == PBSC_TASK_PENDING) continue;
        // what the piece ends with after this pair: the DP consensus, the walk's merged sequence (plain orientation), or raw read
        const bool fromDp = e->status < 0 && e->dp_status == PBSC_DP_OK;
        const bool fromWalk = e->status == 1 && !e->rtou;
        if (!fromDp && !fromWalk) continue;
        const WalkTask& sp = spec[b + t + 1];
        if (!sp.valid) continue;
        int k2; bool rtou2;
        pair_inputs(sv[t].end_best_k, sv[t].is_repeat, (int64_t)1 << 40, sv[t + 1], start_kmer, k2, rtou2);
        if (k2 != sp.k || (rtou2 ? 1 : 0) != sp.rtou) continue;
        if ((int64_t)e->out_len - e->k < k2) continue;
        uint64_t hi, lo;
        pack_src(outpool + e->out_off + e->out_len - k2, k2, hi, lo);
        if (hi == sp.src_hi && lo == sp.src_lo) continue;
        WalkTask& al = alt[b + t + 1];
        if (al.valid && al.src_hi == hi && al.src_lo == lo) continue;
        WalkTask nt = sp;
        nt.src_hi = hi; nt.src_lo = lo; nt.status = PBSC_TASK_PENDING; nt.dp_status = PBSC_DP_NONE; nt.out_len = 0;
        nt.dp_wanted = sp.dp_wanted ? 1 : 0;   // (2 marks a task the DP stage has already collected)
        nt.out_off = sp.out_off + alt_pool_base;
        al = nt;
        alt_list[atomicAdd(n_alt, 1u)] = (uint32_t)(b + t + 1);
    }
}

// query of a task: beginningkmer + strBetweenSrcTarget + targetSeed, or its reverse complement when the walk runs from the
// target towards the source (isFromRtoU, PacBioSelfCorrectionProcess.cpp:176-184)
__device__ __forceinline__ void task_shape(const WalkTask& tk, int& interval, uint32_t& trgLen, uint32_t& qlen)
{
    interval = tk.trg_start - tk.src_end - 1;
    trgLen = tk.rtou ? (uint32_t)tk.k : (uint32_t)tk.trg_len;
    qlen = (uint32_t)(tk.k + interval) + trgLen;
}

// setup record of task `ti`: speculative tasks have scanned offsets, a read's pending request has a fixed-size slot
// (pend_cap != 0 selects the latter)
__device__ __forceinline__ uint8_t* task_record(uint64_t ti, uint8_t* recpool, const uint64_t* rec_off, uint64_t pend_base, uint64_t pend_cap)
{
    return pend_cap ? recpool + pend_base + ti * pend_cap : recpool + rec_off[ti];
}

// thread per task: everything the constructor of LongReadSelfCorrectByOverlap computes (terminal intervals, query idmer
// table, query 5-mers, root intervals) goes into the task's setup record
__global__ void __launch_bounds__(128)
setup_tasks_kernel(const __grid_constant__ FmIndexDev idx, const __grid_constant__ ExtParamsDev P, uint64_t n_items, const uint32_t* __restrict__ list,
                   WalkTask* tasks, const uint8_t* __restrict__ codes, const uint64_t* __restrict__ offsets, uint8_t* recpool,
                   const uint64_t* __restrict__ rec_off, uint64_t pend_base, uint64_t pend_cap)
{
    const uint64_t it = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (it >= n_items) return;
    const uint64_t ti = list ? list[it] : it;
    WalkTask& tk = tasks[ti];
    if (!tk.valid) return;
    int interval; uint32_t trgLen, qlen;
    task_shape(tk, interval, trgLen, qlen);
    if (interval < 0 || tk.k <= 0) { tk.valid = 0; tk.status = PBSC_WALK_UNSUPPORTED; return; }
    tw::SetupView v;
    tw::setup_view(task_record(ti, recpool, rec_off, pend_base, pend_cap), qlen, trgLen, P.min_overlap, P.seed_size, v);
    const uint8_t* read = codes + offsets[tk.read];
    const uint8_t* pth = read + tk.src_end + 1;
    const uint8_t* trgS = read + tk.trg_start;
    const int k = tk.k;
    if (!tk.rtou)
    {
        for (int x = 0; x < k; x++) v.q[x] = (uint8_t)tail_base(tk.src_hi, tk.src_lo, k - 1 - x);
        for (int x = 0; x < interval; x++) v.q[k + x] = pth[x];
        for (uint32_t x = 0; x < trgLen; x++) v.q[k + interval + x] = trgS[x];
    }
    else
    {
        for (uint32_t x = 0; x < qlen; x++)
        {
            const uint32_t y = qlen - 1 - x;
            const uint8_t c = y < (uint32_t)k ? (uint8_t)tail_base(tk.src_hi, tk.src_lo, k - 1 - y)
                                              : (y < (uint32_t)(k + interval) ? pth[y - k] : trgS[y - k - interval]);
            v.q[x] = 3 - c;
        }
    }
    tw::setup_task(idx, P, v, qlen, (uint32_t)k, interval, trgLen, tk.out_cap);
}

// -DPBSC_STAGE_CLOCKS: a measurement build (tools/stage_clocks.py) that adds up, per warp, the clock64() time of every stage of the
// level loop, the iterations, the walking lanes and the leaves dealt out by the pooled stages.  Not for timing runs.
#ifdef PBSC_STAGE_CLOCKS
#define PBSC_N_STAGE 24
__device__ unsigned long long g_stage_clk[PBSC_N_STAGE];
#define STAGE_BEGIN() long long stg_t = clock64(); unsigned long long stg_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define STAGE_MARK(i) do { const long long stg_n = clock64(); stg_acc[i] += (unsigned long long)(stg_n - stg_t); stg_t = stg_n; } while (0)
#define STAGE_COUNT(i, v) do { const unsigned stg_v = __reduce_add_sync(FULL, (unsigned)(v)); stg_acc[i] += stg_v; } while (0)
#define STAGE_END() do { if ((threadIdx.x & 31) == 0) for (int stg_i = 0; stg_i < 12; stg_i++) atomicAdd(&g_stage_clk[stg_i], stg_acc[stg_i]); } while (0)
#else
#define STAGE_BEGIN()
#define STAGE_MARK(i)
#define STAGE_COUNT(i, v)
#define STAGE_END()
#endif

// Level loop.  Every lane owns one walk at a time; all lanes of a warp run the same loop (one level of extendOverlap per
// iteration), and a lane whose walk ended picks the next task at the top of the next iteration, so the warp stays converged
// at the granularity of a level.
__device__ __forceinline__ void
walk_levels_body(const FmIndexDev& idx, const ExtParamsDev& P, uint8_t* scratch, size_t stride,
                   unsigned long long* counter, uint64_t n_items, const uint32_t* __restrict__ list, WalkTask* tasks, uint8_t* recpool,
                   const uint64_t* __restrict__ rec_off, uint64_t pend_base, uint64_t pend_cap, uint8_t* outpool, uint64_t minSA,
                   unsigned long long* walk_counter, uint32_t* heavy_list, unsigned int* n_heavy, uint32_t* nodepool,
                   unsigned long long* pool_used, uint64_t pool_cap, tw::Caps caps, const unsigned int* n_items_dev, int last_pass, int owners)
{
    // lanes 0 .. owners-1 of every warp take walks (and own scratch); the others only work in the pooled stages
    const bool owner = (int)(threadIdx.x & 31) < owners;
    const size_t tid = (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (size_t)owners + (threadIdx.x & 31);
    __shared__ tw::Lane lanes_all[TW_BLOCK];
    tw::Lane* ctx = lanes_all + (threadIdx.x & ~31u);    // this warp's 32 records
    tw::Lane& S = ctx[threadIdx.x & 31];
    const tw::Layout Y = tw::make_layout(P.node_cap, caps);
    uint8_t* const my_scratch = scratch + (owner ? tid : 0) * stride;
    if (n_items_dev) n_items = *n_items_dev;      // a later pass: its item count was produced on the device
    S.status = 0; S.n = 0;
    WalkTask* tk = nullptr;
    tw::SetupHdr* hdr = nullptr;
    // lane states: a walk in progress (active), a walk that ended and still has to be filed (ended), nothing (idle).
    // Filing a result and starting the next walk happen at the top of the iteration for every lane that is waiting.  Batching
    // them (waiting until 8 lanes are idle) was measured and lost: 783 -> 832 ms on config 2, idle lanes cost more than the
    // divergent refill path.
    constexpr int REFILL_BATCH = 1;
    bool active = false, ended = false, exhausted = !owner;
    unsigned long long done = 0;
    STAGE_BEGIN();
    for (;;)
    {
        const unsigned waiting = __ballot_sync(FULL, !active && !exhausted);
        const unsigned walking = __ballot_sync(FULL, active);
        if (waiting && (__popc(waiting) >= REFILL_BATCH || !walking))
        {
            if (ended)
            {
                uint32_t mlen = 0;
                int st = tw::finish_walk(S, Y, P, hdr, nodepool, pool_used, pool_cap, last_pass != 0, outpool + tk->out_off, tk->out_cap, &mlen);
                // the last pass carries everything the reference's loop can hold (-l leaves, 4 children each); only the
                // label tree and the result list are bounded, and running out of those is reported, not hidden
                if (st == PBSC_WALK_HEAVY && last_pass) st = PBSC_OVF_TREE;
                tk->out_len = st == 1 ? mlen : 0;
                tk->status = st;
                if (st == PBSC_WALK_HEAVY) heavy_list[atomicAdd(n_heavy, 1u)] = (uint32_t)(tk - tasks);
                ended = false;
            }
            if (!active && !exhausted)
            {
                for (;;)
                {
                    const unsigned long long it = atomicAdd(counter, 1ull);
                    if (it >= n_items) { exhausted = true; break; }
                    const uint64_t ti = list ? list[it] : it;
                    tk = &tasks[ti];
                    if (!tk->valid) continue;
                    int interval; uint32_t trgLen, qlen;
                    task_shape(*tk, interval, trgLen, qlen);
                    tw::SetupView v;
                    tw::setup_view(task_record(ti, recpool, rec_off, pend_base, pend_cap), qlen, trgLen, P.min_overlap, P.seed_size, v);
                    tw::begin_walk(S, Y, P, my_scratch, v, minSA);
                    hdr = v.hdr;
                    active = true;
                    done++;
                    break;
                }
            }
        }
        if (__all_sync(FULL, !active)) break;   // only reached with nobody waiting either: every lane is exhausted
        // ---- one level of extendOverlap's loop for every walking lane; the stages that touch the index or the per-walk
        //      tables are pooled over the warp (pbsc_walk_thread.cuh) ----
        STAGE_MARK(0);                                    // filing results, next task, begin_walk
        const bool lv = active && tw::walk_continues(S, P);
        {
            const bool need = lv && tw::needs_refine(S);
            STAGE_COUNT(9, 1u);                           // [9] lane-iterations (32 per warp iteration)
            STAGE_COUNT(10, lv ? 1u : 0u);                // [10] walking lanes
            STAGE_COUNT(11, lv ? S.n : 0u);               // [11] leaves probed
            tw::refine_pool(idx, Y, ctx, need ? S.n : 0u);
            if (need) S.curK = S.maxOverlap;
        }
        STAGE_MARK(1);
        if (lv)
        {
            tw::filter_leaves(S, Y);
            S.thr = S.phase == 2 ? S.minSA - 1 : S.minSA;
        }
        STAGE_MARK(2);
        tw::pool_run(lv ? S.n : 0u, [&](int owner_lane, uint32_t i) { tw::probe_leaf(idx, Y, ctx[owner_lane], i); });
        STAGE_MARK(3);
        uint32_t m = 0, sel = 0;
        bool go = false;
        if (lv)
        {
            m = tw::adopt_children(S, Y);
            if (S.status == 0) go = tw::level_middle(S, Y, P, m, sel);
        }
        STAGE_MARK(4);
        if (__any_sync(FULL, sel != 0))
        {
            tw::pool_run(sel, [&](int owner_lane, uint32_t i) { tw::select_leaf(idx, Y, ctx[owner_lane], i); });
            if (sel) tw::level_select(S, P);
            tw::pool_run(sel, [&](int owner_lane, uint32_t i) { tw::reselect_leaf(idx, Y, ctx[owner_lane], i); });
        }
        STAGE_MARK(5);
        // (prune_leaf reads the walk's curLen after curLen++ and its level before level++)
        tw::pool_run(go ? m : 0u, [&](int owner_lane, uint32_t j) { tw::prune_leaf(P, Y, ctx[owner_lane], j); });
        STAGE_MARK(6);
        bool check_term = false;
        if (go) { S.level++; check_term = S.curLen >= S.minLength; }
        tw::pool_run(check_term ? m : 0u, [&](int owner_lane, uint32_t j) { tw::term_leaf(Y, ctx[owner_lane], j); });
        STAGE_MARK(7);
        if (go) tw::finish_level(S, Y, P, m, check_term);
        if (active && !tw::walk_continues(S, P)) { active = false; ended = true; }
        STAGE_MARK(8);
    }
    STAGE_END();
    if (done) atomicAdd(walk_counter, done);
}

// The level loop is latency-bound (ncu: 59 % of stall samples wait on a rank sector), so the register budget decides how many
// walks an SM keeps in flight: MINB resident blocks of 128 lanes (3 -> 168 registers, 4 -> 128, 5 -> 96 with spills to L1).
template <int MINB>
__global__ void __launch_bounds__(TW_BLOCK, MINB)
walk_levels_kernel(const __grid_constant__ FmIndexDev idx, const __grid_constant__ ExtParamsDev P, uint8_t* scratch, size_t stride,
                   unsigned long long* counter, uint64_t n_items, const uint32_t* __restrict__ list, WalkTask* tasks, uint8_t* recpool,
                   const uint64_t* __restrict__ rec_off, uint64_t pend_base, uint64_t pend_cap, uint8_t* outpool, uint64_t minSA,
                   unsigned long long* walk_counter, uint32_t* heavy_list, unsigned int* n_heavy, uint32_t* nodepool,
                   unsigned long long* pool_used, uint64_t pool_cap, tw::Caps caps, const unsigned int* n_items_dev, int last_pass, int owners)
{
    walk_levels_body(idx, P, scratch, stride, counter, n_items, list, tasks, recpool, rec_off, pend_base, pend_cap, outpool, minSA, walk_counter,
                     heavy_list, n_heavy, nodepool, pool_used, pool_cap, caps, n_items_dev, last_pass, owners);
}
typedef void (*WalkKernel)(FmIndexDev, ExtParamsDev, uint8_t*, size_t, unsigned long long*, uint64_t, const uint32_t*, WalkTask*, uint8_t*,
                           const uint64_t*, uint64_t, uint64_t, uint8_t*, uint64_t, unsigned long long*, uint32_t*, unsigned int*, uint32_t*,
                           unsigned long long*, uint64_t, tw::Caps, const unsigned int*, int, int);
static WalkKernel walk_kernel_for(int minb)
{
    switch (minb) { case 3: return walk_levels_kernel<3>; case 4: return walk_levels_kernel<4>; case 5: return walk_levels_kernel<5>; case 7: return walk_levels_kernel<7>; case 8: return walk_levels_kernel<8>; default: return walk_levels_kernel<6>; }
}
static int walk_min_blocks()
{
    const char* e = getenv("PBSC_TW_MINB");
    const int v = e ? atoi(e) : 6;   // measured on config 2: 1002 ms at 3 blocks/SM, 867 ms at 6
    return (v >= 3 && v <= 8) ? v : 6;
}

// thread per task: write the merged sequence of every light walk that succeeded
__global__ void __launch_bounds__(128)
materialize_kernel(uint64_t n_items, const uint32_t* __restrict__ list, WalkTask* tasks, uint8_t* recpool, const uint64_t* __restrict__ rec_off,
                   uint64_t pend_base, uint64_t pend_cap, const uint32_t* __restrict__ nodepool, uint8_t* outpool, int min_overlap, int s9)
{
    const uint64_t it = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (it >= n_items) return;
    const uint64_t ti = list ? list[it] : it;
    WalkTask& tk = tasks[ti];
    if (!tk.valid || tk.status != PBSC_TASK_MATERIALIZE) return;
    int interval; uint32_t trgLen, qlen;
    task_shape(tk, interval, trgLen, qlen);
    tw::SetupView v;
    tw::setup_view(task_record(ti, recpool, rec_off, pend_base, pend_cap), qlen, trgLen, min_overlap, s9, v);
    uint32_t mlen = 0;
    const int st = tw::materialize(v, min_overlap, nodepool, outpool + tk.out_off, tk.out_cap, &mlen);
    tk.out_len = st == 1 ? mlen : 0;
    tk.status = st;
}

struct StitchParams { int32_t start_kmer, next_target, split, no_dp; };

// thread per read: initCorrect (PacBioSelfCorrectionProcess.cpp:56-157) over finished walk tasks
__global__ void __launch_bounds__(128)
stitch_kernel(StitchParams C, uint64_t n_reads, const uint8_t* __restrict__ codes, const uint64_t* __restrict__ offsets,
              const pbsc_seed* __restrict__ seeds, const uint64_t* __restrict__ region, const uint32_t* __restrict__ seed_count,
              const uint64_t* __restrict__ task_base, WalkTask* spec, const WalkTask* alt, WalkTask* pending, const uint8_t* __restrict__ outpool,
              uint64_t pending_pool_off, uint32_t pending_cap, ReadState* states, uint8_t* __restrict__ pieces,
              const uint64_t* __restrict__ piece_region, uint32_t* __restrict__ piece_bounds, const uint64_t* __restrict__ bounds_region,
              pbsc_read_stats* __restrict__ stats, int32_t* __restrict__ read_status, uint32_t* stalled_list, unsigned int* n_stalled,
              pbsc_walk_log* __restrict__ dbg_log, uint32_t* __restrict__ dbg_log_n)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    ReadState S = states[r];
    if (S.done) return;
    const uint8_t* read = codes + offsets[r];
    const int64_t L = (int64_t)(offsets[r + 1] - offsets[r]);
    const pbsc_seed* sv = seeds + region[r];
    const uint32_t ns = seed_count[r];
    uint8_t* piece = pieces + piece_region[r];
    const uint64_t pieceCap = piece_region[r + 1] - piece_region[r];
    uint32_t* bounds = piece_bounds + bounds_region[r];
    const uint64_t boundsCap = bounds_region[r + 1] - bounds_region[r];
    if (!S.started)
    {
        memset(&S, 0, sizeof S);
        S.started = 1;
        S.st.total_reads_len = L;
        S.st.total_seed_num = ns;
        S.pending_trg = -1;
        if (ns >= 2)
        {
            const pbsc_seed s0 = sv[0];
            if ((uint64_t)s0.len > pieceCap || boundsCap < 2) S.rstatus = PBSC_OVF_PIECES;
            else
            {
                for (int x = 0; x < s0.len; x++) piece[x] = read[s0.start + x];
                S.plen = s0.len;
                bounds[0] = 0;
            }
            S.srcLen = s0.len; S.srcEnd = s0.start + s0.len - 1; S.srcEndBest = s0.end_best_k; S.srcRepeat = s0.is_repeat;
            S.srcFreq = s0.max_fixed_freq;
            S.srcStart = s0.start;
            S.nPieces = 1;
            S.t = 1; S.next = 0;
        }
    }
    bool stalled = false;
    if (ns >= 2 && S.rstatus == 0)
    {
        while ((uint32_t)S.t < ns)
        {
            bool success = false;
            // S.next > 0 only when resuming inside the look-ahead loop; firstType was saved with it
            for (; S.next < C.next_target && (uint32_t)(S.t + S.next) < ns; S.next++)
            {
                const int ti = S.t + S.next;
                const pbsc_seed tg = sv[ti];
                const int interval = tg.start - S.srcEnd - 1;
                int k; bool rtou;
                pair_inputs(S.srcEndBest, S.srcRepeat, S.srcLen, tg, C.start_kmer, k, rtou);
                if (k <= 0 || k > S.srcLen || k > 64 || interval < 0) { S.rstatus = PBSC_WALK_UNSUPPORTED; break; }
                uint64_t hi, lo;
                pack_src(piece + S.plen - k, k, hi, lo);
                // find a finished task with exactly these inputs
                const WalkTask* use = nullptr;
                if (S.next == 0)
                {
                    const WalkTask* sp = spec + task_base[r] + ti;
                    if (sp->valid && sp->k == k && sp->rtou == (rtou ? 1 : 0) && sp->src_hi == hi && sp->src_lo == lo && sp->status != PBSC_TASK_PENDING) use = sp;
                    if (!use && alt)
                    {
                        const WalkTask* al = alt + task_base[r] + ti;
                        if (al->valid && al->k == k && al->rtou == (rtou ? 1 : 0) && al->src_hi == hi && al->src_lo == lo && al->status != PBSC_TASK_PENDING) use = al;
                    }
                }
                if (!use)
                {
                    WalkTask* pd = pending + r;
                    if (pd->valid && S.pending_trg == ti && pd->k == k && pd->rtou == (rtou ? 1 : 0) && pd->src_hi == hi && pd->src_lo == lo && pd->src_end == S.srcEnd)
                    {
                        if (pd->status != PBSC_TASK_PENDING) use = pd;
                    }
                    if (!use)
                    {
                        // file the exact request and stall
                        WalkTask tk;
                        memset(&tk, 0, sizeof tk);
                        tk.src_hi = hi; tk.src_lo = lo; tk.read = (uint32_t)r; tk.src_end = S.srcEnd; tk.trg_start = tg.start; tk.trg_len = tg.len;
                        tk.k = k; tk.rtou = rtou ? 1 : 0; tk.status = PBSC_TASK_PENDING; tk.valid = 1;
                        tk.freq_sum = S.srcFreq + tg.max_fixed_freq;
                        tk.dp_wanted = (S.next == 0 && !C.no_dp) ? 1 : 0;   // correctByMSAlignment only runs on the first target
                        tk.out_off = pending_pool_off + r * (uint64_t)pending_cap;
                        tk.out_cap = pending_cap;
                        *pd = tk;
                        S.pending_trg = ti;
                        stalled = true;
                        break;
                    }
                }
                const int stw = use->status;
                if (is_overflow(stw) || stw == PBSC_WALK_UNSUPPORTED) { S.rstatus = stw; break; }
                if (S.next == 0)
                {
                    S.firstType = stw;
                    S.dpStatus0 = stw > 0 ? PBSC_DP_NONE : use->dp_status; S.dpLen0 = use->out_len; S.dpOff0 = use->out_off;
                }
                if (stw > 0)
                {
                    const uint8_t* merged = outpool + use->out_off;
                    const uint32_t mlen = use->out_len;
                    uint64_t outLen;
                    if (!rtou)
                    {
                        outLen = mlen - k;
                        if (S.plen + outLen > pieceCap) { S.rstatus = PBSC_OVF_PIECES; break; }
                        for (uint64_t x = 0; x < outLen; x++) piece[S.plen + x] = merged[k + x];
                    }
                    else
                    {
                        const uint64_t tailLen = tg.len - k;
                        outLen = (mlen - k) + tailLen;
                        if (S.plen + outLen > pieceCap) { S.rstatus = PBSC_OVF_PIECES; break; }
                        for (uint64_t x = 0; x < mlen - k; x++) piece[S.plen + x] = 3 - merged[mlen - 1 - (k + x)];
                        for (uint64_t x = 0; x < tailLen; x++) piece[S.plen + (mlen - k) + x] = read[tg.start + k + x];
                    }
                    S.plen += outLen;
                    S.st.corrected_len += outLen;
                    S.st.seed_dis += interval;
                    S.st.fm_num++;
                    S.st.total_walk_num++;
                    S.srcLen += outLen;
                    S.srcEnd = tg.start + tg.len - 1; S.srcStart = tg.start; S.srcEndBest = tg.end_best_k; S.srcRepeat = tg.is_repeat; S.srcFreq = tg.max_fixed_freq;
                    S.t += S.next;
                    success = true;
                    break;
                }
            }
            if (stalled || S.rstatus) break;
            if (!success)
            {
                const pbsc_seed tg = sv[S.t];
                if (S.firstType == -1) S.st.high_error_num++;
                else if (S.firstType == -2) S.st.exceed_depth_num++;
                else if (S.firstType == -3) S.st.exceed_leave_num++;
                else { S.rstatus = PBSC_WALK_NO_PATH; break; }
                S.st.total_walk_num++;
                // correctByMSAlignment (PacBioSelfCorrectionProcess.cpp:134-136, 208-245)
                if (is_overflow(S.dpStatus0) || S.dpStatus0 == PBSC_WALK_UNSUPPORTED) { S.rstatus = S.dpStatus0; break; }
                if (dbg_log)
                {
                    // extend/<id>.ext and .dp (PacBioSelfCorrectionProcess.cpp:130-131, 139-140); one slot per seed of the read
                    pbsc_walk_log e;
                    e.src_start = S.srcStart; e.trg_start = tg.start; e.code = S.firstType + 4; e.dp_failed = S.dpStatus0 == PBSC_DP_OK ? 0 : 1;
                    dbg_log[region[r] + S.nLog++] = e;
                }
                if (S.dpStatus0 == PBSC_DP_OK)
                {
                    int k; bool rtou;
                    pair_inputs(S.srcEndBest, S.srcRepeat, S.srcLen, tg, C.start_kmer, k, rtou);
                    const uint8_t* cons = outpool + S.dpOff0;
                    const uint64_t outLen = S.dpLen0 - (uint32_t)k;
                    if (S.plen + outLen > pieceCap) { S.rstatus = PBSC_OVF_PIECES; break; }
                    for (uint64_t x = 0; x < outLen; x++) piece[S.plen + x] = cons[k + x];
                    S.plen += outLen;
                    S.srcLen += outLen;
                    S.st.corrected_len += outLen;
                    S.st.seed_dis += tg.start - S.srcEnd - 1;
                    S.st.dp_num++;
                }
                else
                {
                    if (C.split)
                    {
                        if (S.plen + tg.len > pieceCap || S.nPieces + 1 >= boundsCap) { S.rstatus = PBSC_OVF_PIECES; break; }
                        bounds[S.nPieces] = (uint32_t)S.plen;
                        S.nPieces++;
                        for (int x = 0; x < tg.len; x++) piece[S.plen + x] = read[tg.start + x];
                        S.plen += tg.len;
                        S.srcLen = tg.len;
                    }
                    else
                    {
                        const int tgEnd = tg.start + tg.len - 1;
                        const uint64_t n = (uint64_t)(tgEnd - S.srcEnd);
                        if (S.plen + n > pieceCap) { S.rstatus = PBSC_OVF_PIECES; break; }
                        for (uint64_t x = 0; x < n; x++) piece[S.plen + x] = read[S.srcEnd + 1 + x];
                        S.plen += n;
                        S.srcLen += n;
                    }
                    S.st.corrected_len += tg.len;
                }
                S.srcEnd = tg.start + tg.len - 1; S.srcStart = tg.start; S.srcEndBest = tg.end_best_k; S.srcRepeat = tg.is_repeat; S.srcFreq = tg.max_fixed_freq;
            }
            S.t++;
            S.next = 0;
            S.firstType = 0;
            S.dpStatus0 = PBSC_DP_NONE;
        }
    }
    if (stalled)
    {
        states[r] = S;
        stalled_list[atomicAdd(n_stalled, 1u)] = (uint32_t)r;
        return;
    }
    S.done = 1;
    S.st.merge = (ns >= 2) ? 1 : 0;
    S.st.n_pieces = (int32_t)S.nPieces;
    if (S.nPieces && S.nPieces < boundsCap) bounds[S.nPieces] = (uint32_t)S.plen;
    stats[r] = S.st;
    read_status[r] = S.rstatus;
    if (dbg_log_n) dbg_log_n[r] = S.nLog;
    states[r] = S;
}

void make_ext_params(const pbsc_params* p, ExtParamsDev& d, uint32_t q_cap, uint32_t node_cap, uint32_t merged_cap);
__global__ void chain_bounds_kernel(uint64_t n_reads, const pbsc_seed* __restrict__ seeds, const uint64_t* __restrict__ region,
                                    const uint32_t* __restrict__ seed_count, int next_target, unsigned int* max_gap, unsigned int* max_trg);

// buffers of the thread engine; they live in the index's grow-only arena
template <class T>
struct ArenaPtr
{
    T* p = nullptr;
    cudaError_t get(pbsc_index* idx, const char* name, size_t count) { return arena_get(idx, name, (count ? count : 1) * sizeof(T), (void**)&p); }
};
struct ThreadEngine
{
    ArenaPtr<WalkTask> spec, pending, alt;
    ArenaPtr<uint32_t> alt_list;
    ArenaPtr<unsigned int> n_alt;
    uint64_t alt_pool_base = 0, alt_rec_base = 0;
    ArenaPtr<uint64_t> task_base, caps, cap_off, rec_caps, rec_off;
    ArenaPtr<uint8_t> outpool, recpool, scratch, wscratch, hscratch;
    ArenaPtr<uint32_t> heavy_list, nodepool;
    uint64_t heavy_cap = 0;
    tw::Caps light, heavy;
    size_t stride = 0, hstride = 0;
    int blocks = 0, hblocks = 0;
    int heavy_owners = 16;  // most lanes of a warp that own a walk in a full-capacity pass (PBSC_TW_HEAVY_OWNERS)
    bool no_dp = true;
    const pbsc_params* params = nullptr;
    WalkKernel kernel = nullptr;
    // CUDA events around every walk_levels_kernel launch (the dominant kernel): its own duration, for the roofline
    std::vector<cudaEvent_t> ev;
    ~ThreadEngine() { for (auto e : ev) cudaEventDestroy(e); }
    void mark(cudaStream_t st) { cudaEvent_t e; if (cudaEventCreate(&e) == cudaSuccess) { cudaEventRecord(e, st); ev.push_back(e); } }
    std::vector<std::pair<const char*, uint64_t>> ev_what;   // PBSC_ROUND_TRACE: pass name and item count of every walk launch
    ArenaPtr<unsigned long long> pool_used;
    uint64_t pool_cap = 0;
    ArenaPtr<unsigned int> n_heavy;
    ExtParamsDev Pw;
    size_t wstride = 0;
    int wblocks = 0;
    DevBuf<uint8_t> cubtmp;
    ArenaPtr<ReadState> states;
    ArenaPtr<uint32_t> stalled;
    ArenaPtr<unsigned int> n_stalled;
    uint64_t pend_rec_base = 0, pend_rec_cap = 0;
};

// Light pass, then the walks that outgrew it among walks of their own weight, then materialisation.  `is_pending` selects
// where the setup records live (a read's pending request has a fixed-size slot).
static int launch_walk(pbsc_index* idx, const ExtParamsDev& P, ThreadEngine& E, Workspace& w, DeviceBatch& b, uint64_t n_items, const uint32_t* list,
                       WalkTask* tasks, bool is_pending, uint64_t minSA, uint64_t* launches, uint8_t* recpool = nullptr)
{
    cudaStream_t st = idx->stream;
    if (!recpool) recpool = E.recpool.p;
    const uint64_t pcap = is_pending ? E.pend_rec_cap : 0;
    PBSC_CUDA(cudaMemsetAsync(w.counters.p, 0, 8, st));
    setup_tasks_kernel<<<(unsigned)((n_items + 127) / 128), 128, 0, st>>>(idx->dev, P, n_items, list, tasks, b.codes.p, b.offsets.p, recpool, E.rec_off.p,
                                                                         E.pend_rec_base, pcap);
    PBSC_OCC_TAKE(1, st);
    const uint64_t threads = (uint64_t)E.blocks * TW_BLOCK;
    int nb = E.blocks;
    if (n_items < threads) nb = (int)((n_items + TW_BLOCK - 1) / TW_BLOCK);
    if (nb < 1) nb = 1;
    PBSC_CUDA(cudaMemsetAsync(E.n_heavy.p, 0, 8, st));
    PBSC_CUDA(cudaMemsetAsync(E.pool_used.p, 0, 8, st));
    // A pass of few, long walks (the full-capacity pass, the re-walk rounds) lasts as long as its longest walk, and a walk
    // advances one level per round of the warp's pooled stages: the fewer walks share a warp, the fewer rounds a level takes.
    // Such passes therefore give walks to only `owners` lanes of every warp (the other lanes only help), as few as the number
    // of resident warps allows.  The light pass of a large round is bound by throughput and uses all 32.
    const uint64_t warps_max = (uint64_t)E.blocks * (TW_BLOCK / 32);
    auto geometry = [&](uint64_t walks, int& owners, int& blocks) {
        owners = (int)std::min<uint64_t>((uint64_t)E.heavy_owners, std::max<uint64_t>(1, (walks + warps_max - 1) / warps_max));
        blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)E.blocks, (walks + (uint64_t)owners * (TW_BLOCK / 32) - 1) / ((uint64_t)owners * (TW_BLOCK / 32))));
    };
    // A small round (re-walks of mispredicted pairs): one pass at full capacities instead of a light and a heavy one.
    const bool single_pass = n_items <= warps_max * (uint64_t)E.heavy_owners / 2;
    if (single_pass)
    {
        int owners, hb;
        geometry(n_items, owners, hb);
        E.ev_what.push_back({"single", n_items});
        E.mark(st);
        E.kernel<<<hb, TW_BLOCK, 0, st>>>(idx->dev, E.Pw, E.hscratch.p, E.hstride, w.counters.p, n_items, list, tasks, recpool, E.rec_off.p,
                                          E.pend_rec_base, pcap, E.outpool.p, minSA, w.counters.p + 1, E.heavy_list.p, E.n_heavy.p,
                                          E.nodepool.p, E.pool_used.p, E.pool_cap, E.heavy, nullptr, 1, owners);
        E.mark(st);
    }
    else
    {
        E.ev_what.push_back({"light", n_items});
        E.mark(st);
        E.kernel<<<nb, TW_BLOCK, 0, st>>>(idx->dev, P, E.scratch.p, E.stride, w.counters.p, n_items, list, tasks, recpool, E.rec_off.p,
                                          E.pend_rec_base, pcap, E.outpool.p, minSA, w.counters.p + 1, E.heavy_list.p, E.n_heavy.p,
                                          E.nodepool.p, E.pool_used.p, E.pool_cap, E.light, nullptr, 0, 32);
        E.mark(st);
        PBSC_CUDA(cudaMemsetAsync(w.counters.p, 0, 8, st));
        // the walks that outgrew the light pass: how many there are decides the geometry (4-byte round trip)
        unsigned int nh = 0;
        PBSC_CUDA(cudaMemcpyAsync(&nh, E.n_heavy.p, 4, cudaMemcpyDeviceToHost, st));
        PBSC_CUDA(cudaStreamSynchronize(st));
        if (nh)
        {
            int owners, hb;
            geometry(nh, owners, hb);
            E.ev_what.push_back({"heavy", nh});
            E.mark(st);
            E.kernel<<<hb, TW_BLOCK, 0, st>>>(idx->dev, E.Pw, E.hscratch.p, E.hstride, w.counters.p, nh, E.heavy_list.p, tasks, recpool, E.rec_off.p,
                                              E.pend_rec_base, pcap, E.outpool.p, minSA, w.counters.p + 1, E.heavy_list.p + E.heavy_cap,
                                              E.n_heavy.p + 1, E.nodepool.p, E.pool_used.p, E.pool_cap, E.heavy, nullptr, 1, owners);
            E.mark(st);
        }
        PBSC_CUDA(cudaGetLastError());
    }
    PBSC_OCC_TAKE(2, st);
    materialize_kernel<<<(unsigned)((n_items + 127) / 128), 128, 0, st>>>(n_items, list, tasks, recpool, E.rec_off.p, E.pend_rec_base, pcap,
                                                                         E.nodepool.p, E.outpool.p, P.min_overlap, P.seed_size);
    PBSC_CUDA(cudaGetLastError());
    if (launches) *launches += 4;
    if (!E.no_dp)
    {
        const int rc = run_dp_fallback(idx, E.params, b, tasks, n_items, list, E.outpool.p, w.q_cap, launches);
        if (rc != PBSC_OK) return rc;
    }
    return PBSC_OK;
}

static int thread_geometry(int device, int* blocks)
{
    int sms = 0;
    PBSC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    int per_sm = 0;
    PBSC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, walk_kernel_for(walk_min_blocks()), TW_BLOCK, 0));
    if (per_sm < 1) per_sm = 1;
    const char* e = getenv("PBSC_TW_BLOCKS_PER_SM");
    if (e && atoi(e) > 0) per_sm = std::min(per_sm, atoi(e));
    *blocks = sms * per_sm;
    return PBSC_OK;
}

static int scan_u64(DevBuf<uint8_t>& tmp, const uint64_t* in, uint64_t* out, uint64_t n, cudaStream_t st)
{
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (int)n, st);
    if (tb > tmp.n) PBSC_CUDA(tmp.alloc(tb));
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, (int)n, st);
    return PBSC_OK;
}

// which idmers occur on which strand, for every possible idmer (4^-i entries): built once per index and -i value
int ensure_idmer_table(pbsc_index* idx, int idmer_len)
{
    if (idx->dev.idmer_len == idmer_len && idx->dev.idmer_valid) return PBSC_OK;
    PBSC_CUDA(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    const uint64_t n_keys = 1ull << (2 * idmer_len);
    if (idx->d_idmer_valid)
    {
        if (!idx->blob || (uint8_t*)idx->d_idmer_valid < (uint8_t*)idx->blob || (uint8_t*)idx->d_idmer_valid >= (uint8_t*)idx->blob + idx->blob_bytes) cudaFree(idx->d_idmer_valid);
        idx->device_bytes -= 1ull << (2 * idx->dev.idmer_len);
        idx->d_idmer_valid = nullptr;
    }
    idx->dev.idmer_valid = nullptr; idx->dev.idmer_len = 0;
    PBSC_CUDA(cudaMalloc((void**)&idx->d_idmer_valid, n_keys));
    idmer_valid_kernel<<<(unsigned)((n_keys + 255) / 256), 256, 0, st>>>(idx->dev, idmer_len, n_keys, idx->d_idmer_valid);
    PBSC_CUDA(cudaGetLastError());
    PBSC_CUDA(cudaStreamSynchronize(st));
    idx->dev.idmer_valid = idx->d_idmer_valid;
    idx->dev.idmer_len = idmer_len;
    idx->device_bytes += n_keys;
    PBSC_OCC_TAKE(4, st);
    return PBSC_OK;
}

int run_extend_threads(pbsc_index* idx, const pbsc_params* p, DeviceBatch& b, SeedBuffers& s, Workspace& w, uint64_t* launches)
{
    cudaStream_t st = idx->stream;
    const uint64_t n = b.n_reads;
    if (n == 0) return PBSC_OK;
    ThreadEngine E;
    E.no_dp = p->no_dp != 0;
    E.params = p;
    // PBSC_ROUND_TRACE=1: host wall clock of every stage of the round structure (each ends with a stream synchronisation)
    const bool rtrace = getenv("PBSC_ROUND_TRACE") != nullptr;
    auto rt0 = std::chrono::steady_clock::now();
    auto trace = [&](const char* what, uint64_t items) {
        if (!rtrace) return;
        cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[pbsc round trace] %-28s %10llu items %9.2f ms (DP so far %.1f ms)\n", what, (unsigned long long)items,
                std::chrono::duration<double, std::milli>(now - rt0).count(), last_dp_stats().ms);
        rt0 = now;
    };
    if (idx->dev.idmer_len != p->idmer_len)
    {
        // lane_acquire() builds the table before any lane runs; this is the direct-call path of a single-lane index
        if (idx->primary) { set_error("idmer table of the lane does not match -i %d", p->idmer_len); return PBSC_ERR_INTERNAL; }
        const int rc = ensure_idmer_table(idx, p->idmer_len);
        if (rc != PBSC_OK) return rc;
    }
    // ---- task index space: one slot per surviving seed ----
    PBSC_CUDA(E.task_base.get(idx, "tw.task_base", n + 1));
    {
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, s.count.p, E.task_base.p, (int)n, st);
        PBSC_CUDA(E.cubtmp.alloc(tb));
        cub::DeviceScan::ExclusiveSum(E.cubtmp.p, tb, s.count.p, E.task_base.p, (int)n, st);
    }
    PBSC_CUDA(cudaMemsetAsync(w.maxima.p, 0, 8, st));
    chain_bounds_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, s.seeds.p, s.region.p, s.count.p, p->next_target, w.maxima.p, w.maxima.p + 1);
    uint64_t last_base = 0; uint32_t last_cnt = 0; unsigned int hmax[2] = {0, 0};
    PBSC_CUDA(cudaMemcpyAsync(&last_base, E.task_base.p + n - 1, 8, cudaMemcpyDeviceToHost, st));
    PBSC_CUDA(cudaMemcpyAsync(&last_cnt, s.count.p + n - 1, 4, cudaMemcpyDeviceToHost, st));
    PBSC_CUDA(cudaMemcpyAsync(hmax, w.maxima.p, 8, cudaMemcpyDeviceToHost, st));
    PBSC_CUDA(cudaStreamSynchronize(st));
    const uint64_t n_tasks = last_base + last_cnt;
    uint64_t nl = 2;
    // ---- capacities ----
    const uint32_t need_q = (uint32_t)align_up((size_t)hmax[0] + hmax[1] + 64 + 16, 16);
    if (need_q > w.q_cap) w.q_cap = need_q;
    const uint32_t pending_cap = (uint32_t)align_up((size_t)(1.2 * (hmax[0] + 10)) + 2 * 64 + hmax[1] + 64, 16);
    E.pend_rec_cap = tw::setup_record_bytes(need_q, std::max<uint32_t>(hmax[1], 64), p->min_kmer, p->idmer_len);
    ExtParamsDev P;
    uint32_t t_node_cap = 4096;                       // light walks; a walk that needs more goes to the heavy pass
    tw::Caps light_caps{8, 32, 40, 32};
    if (const char* e = getenv("PBSC_TW_LIGHT"))
    {
        // "live leaves,children per level,history rings,results,label-tree nodes" of the light pass (experiments)
        unsigned a = 0, b2 = 0, c = 0, d = 0, n2 = 0;
        if (sscanf(e, "%u,%u,%u,%u,%u", &a, &b2, &c, &d, &n2) == 5 && a >= 1 && a <= (unsigned)OLD_CAP && c >= a + 1 && c <= 255 && d >= 1 && n2 >= 64)
        { (void)b2; light_caps = tw::Caps{a, 4 * a, c, d}; t_node_cap = n2; }   // children per level is always 4 x live leaves (slot = 4 i + base)
    }
    make_ext_params(p, P, w.q_cap, t_node_cap, pending_cap);
    make_ext_params(p, E.Pw, w.q_cap, std::max<uint32_t>(w.node_cap, 1u << 15), pending_cap);
    const uint64_t minSA = p->pb_coverage > 60 ? (uint64_t)((p->pb_coverage / 60) * 3) : 3;
    E.kernel = walk_kernel_for(walk_min_blocks());
    int rc = thread_geometry(idx->device, &E.blocks);
    if (rc != PBSC_OK) return rc;
    E.light = light_caps;
    // everything extendOverlap's loop can hold: -l live leaves, four children each (LongReadCorrectByOverlap.cpp:161)
    E.heavy = tw::Caps{(uint32_t)OLD_CAP, (uint32_t)NEW_CAP, (uint32_t)RING_SLOTS, (uint32_t)RES_CAP};
    E.stride = tw::thread_scratch_bytes(t_node_cap, E.light);
    PBSC_CUDA(E.scratch.get(idx, "tw.scratch", E.stride * (size_t)E.blocks * TW_BLOCK));
    {
        E.hstride = tw::thread_scratch_bytes(E.Pw.node_cap, E.heavy);
        if (const char* e = getenv("PBSC_TW_HEAVY_OWNERS")) { if (atoi(e) >= 1 && atoi(e) <= 32) E.heavy_owners = atoi(e); }
        // scratch of the full-capacity passes: hstride bytes (280 KB with the default label-tree size) per OWNER lane.  Never
        // ask for more than the device has left (a grown node_cap multiplies hstride): fewer owners per warp first, then a
        // clean PBSC_ERR_LIMIT instead of a CUDA out-of-memory error.  cudaMemGetInfo costs tens of milliseconds, so it is
        // only asked when the arena has to grow.
        {
            const size_t have = idx->arena.count("tw.hscratch") ? idx->arena["tw.hscratch"].cap : 0;
            auto need = [&]() { return E.hstride * (size_t)E.blocks * (TW_BLOCK / 32) * (size_t)E.heavy_owners; };
            if (need() > have)
            {
                size_t free_b = 0, total_b = 0;
                PBSC_CUDA(cudaMemGetInfo(&free_b, &total_b));
                const size_t avail = (size_t)((double)(free_b + have) * 0.8);
                while (need() > avail && E.heavy_owners > 1) E.heavy_owners /= 2;
                if (need() > avail)
                {
                    set_error("walk scratch of %zu bytes per lane (label tree of %u nodes) does not fit the %zu MB left on the device", E.hstride, E.Pw.node_cap, free_b >> 20);
                    return PBSC_ERR_LIMIT;
                }
            }
        }
        E.hblocks = E.blocks;
        PBSC_CUDA(E.hscratch.get(idx, "tw.hscratch", E.hstride * (size_t)E.blocks * (TW_BLOCK / 32) * (size_t)E.heavy_owners));
    }
    PBSC_CUDA(E.spec.get(idx, "tw.spec", n_tasks)); PBSC_CUDA(E.pending.get(idx, "tw.pending", n)); PBSC_CUDA(E.caps.get(idx, "tw.caps", n_tasks + 1)); PBSC_CUDA(E.cap_off.get(idx, "tw.cap_off", n_tasks + 1));
    PBSC_CUDA(E.rec_caps.get(idx, "tw.rec_caps", n_tasks + 1)); PBSC_CUDA(E.rec_off.get(idx, "tw.rec_off", n_tasks + 1));
    E.heavy_cap = std::max<uint64_t>(n_tasks, n) + 1;
    PBSC_CUDA(E.heavy_list.get(idx, "tw.heavy_list", 2 * E.heavy_cap)); PBSC_CUDA(E.n_heavy.get(idx, "tw.n_heavy", 2));
    E.pool_cap = std::max<uint64_t>(n_tasks, n) * (uint64_t)w.pool_nodes + 65536;
    PBSC_CUDA(E.pool_used.get(idx, "tw.pool_used", 1));
    PBSC_CUDA(E.states.get(idx, "tw.states", n)); PBSC_CUDA(E.stalled.get(idx, "tw.stalled", n)); PBSC_CUDA(E.n_stalled.get(idx, "tw.n_stalled", 1));
    PBSC_CUDA(cudaMemsetAsync(E.states.p, 0, n * sizeof(ReadState), st));
    PBSC_CUDA(cudaMemsetAsync(E.pending.p, 0, n * sizeof(WalkTask), st));
    PBSC_CUDA(cudaMemsetAsync(w.counters.p, 0, 16, st));
    uint64_t pool_spec = 0, rec_spec = 0;
    if (n_tasks)
    {
        make_spec_tasks_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, b.codes.p, b.offsets.p, s.seeds.p, s.region.p, s.count.p, E.task_base.p,
                                                                            E.spec.p, E.caps.p, E.rec_caps.p, p->start_kmer, p->min_kmer, p->idmer_len, p->no_dp);
        rc = scan_u64(E.cubtmp, E.caps.p, E.cap_off.p, n_tasks, st);
        if (rc != PBSC_OK) return rc;
        rc = scan_u64(E.cubtmp, E.rec_caps.p, E.rec_off.p, n_tasks, st);
        if (rc != PBSC_OK) return rc;
        set_out_offsets_kernel<<<(unsigned)((n_tasks + 255) / 256), 256, 0, st>>>(n_tasks, E.spec.p, E.cap_off.p);
        uint64_t tail[4] = {0, 0, 0, 0};
        PBSC_CUDA(cudaMemcpyAsync(&tail[0], E.cap_off.p + n_tasks - 1, 8, cudaMemcpyDeviceToHost, st));
        PBSC_CUDA(cudaMemcpyAsync(&tail[1], E.caps.p + n_tasks - 1, 8, cudaMemcpyDeviceToHost, st));
        PBSC_CUDA(cudaMemcpyAsync(&tail[2], E.rec_off.p + n_tasks - 1, 8, cudaMemcpyDeviceToHost, st));
        PBSC_CUDA(cudaMemcpyAsync(&tail[3], E.rec_caps.p + n_tasks - 1, 8, cudaMemcpyDeviceToHost, st));
        PBSC_CUDA(cudaStreamSynchronize(st));
        pool_spec = tail[0] + tail[1];
        rec_spec = tail[2] + tail[3];
        nl += 6;
    }
    // label-tree pool of the light passes: a successful light walk parks one node per child it created, about 1.3 per base
    // of its output on average; the output slots' total (just scanned) bounds that with room to spare
    E.pool_cap = std::max<uint64_t>(E.pool_cap, 2 * pool_spec + 65536);
    PBSC_CUDA(E.nodepool.get(idx, "tw.nodepool", E.pool_cap));
    // pools: [speculative tasks | their alternatives (same layout) | one pending request per read]
    int alt_rounds = 4;
    if (const char* e = getenv("PBSC_ALT_ROUNDS")) alt_rounds = std::max(0, atoi(e));
    const bool use_alt = alt_rounds > 0 && n_tasks > 0;
    E.alt_pool_base = align_up(pool_spec, 16);
    const uint64_t pending_pool_off = use_alt ? 2 * E.alt_pool_base : E.alt_pool_base;
    PBSC_CUDA(E.outpool.get(idx, "tw.outpool", pending_pool_off + n * (uint64_t)pending_cap + 16));
    E.alt_rec_base = align_up(rec_spec, 128);
    E.pend_rec_base = use_alt ? 2 * E.alt_rec_base : E.alt_rec_base;
    PBSC_CUDA(E.recpool.get(idx, "tw.recpool", E.pend_rec_base + n * E.pend_rec_cap + 128));
    if (use_alt)
    {
        PBSC_CUDA(E.alt.get(idx, "tw.alt", n_tasks)); PBSC_CUDA(E.alt_list.get(idx, "tw.alt_list", n_tasks)); PBSC_CUDA(E.n_alt.get(idx, "tw.n_alt", 1));
        PBSC_CUDA(cudaMemsetAsync(E.alt.p, 0, n_tasks * sizeof(WalkTask), st));
    }
    else E.alt.p = nullptr;
    trace("task setup (scans, pools)", n_tasks);
    // ---- round 1: all speculative walks ----
    if (n_tasks)
    {
        rc = launch_walk(idx, P, E, w, b, n_tasks, nullptr, E.spec.p, false, minSA, &nl);
        if (rc != PBSC_OK) return rc;
    }
    trace("round 1 (speculative walks)", n_tasks);
    // ---- alternatives for the successors of pairs the DP fallback corrected (see make_alt_tasks_kernel) ----
    for (int ar = 0; use_alt && ar < alt_rounds; ar++)
    {
        PBSC_CUDA(cudaMemsetAsync(E.n_alt.p, 0, 4, st));
        make_alt_tasks_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, s.seeds.p, s.region.p, s.count.p, E.task_base.p, E.spec.p, E.alt.p, E.outpool.p,
                                                                           E.alt_pool_base, p->start_kmer, E.alt_list.p, E.n_alt.p);
        unsigned int na = 0;
        PBSC_CUDA(cudaMemcpyAsync(&na, E.n_alt.p, 4, cudaMemcpyDeviceToHost, st));
        PBSC_CUDA(cudaStreamSynchronize(st));
        nl++;
        if (na == 0) break;
        rc = launch_walk(idx, P, E, w, b, na, E.alt_list.p, E.alt.p, false, minSA, &nl, E.recpool.p + E.alt_rec_base);
        if (rc != PBSC_OK) return rc;
        trace("alternative round", na);
    }
    StitchParams C;
    C.start_kmer = p->start_kmer; C.next_target = p->next_target; C.split = p->split; C.no_dp = p->no_dp;
    for (int round = 0;; round++)
    {
        PBSC_CUDA(cudaMemsetAsync(E.n_stalled.p, 0, 4, st));
        stitch_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(C, n, b.codes.p, b.offsets.p, s.seeds.p, s.region.p, s.count.p, E.task_base.p, E.spec.p,
                                                                   E.alt.p, E.pending.p, E.outpool.p, pending_pool_off, pending_cap, E.states.p, w.pieces.p,
                                                                   w.piece_region.p, w.bounds.p, w.bounds_region.p, w.stats.p, w.status.p, E.stalled.p,
                                                                   E.n_stalled.p, w.dbg_log.p, w.dbg_log_n.p);
        nl++;
        unsigned int ns = 0;
        PBSC_CUDA(cudaMemcpyAsync(&ns, E.n_stalled.p, 4, cudaMemcpyDeviceToHost, st));
        PBSC_CUDA(cudaStreamSynchronize(st));
        if (ns == 0) break;
        if (round > 100000) { set_error("stitch did not converge"); return PBSC_ERR_INTERNAL; }
        // the stalled reads' requests sit in pending[read]
        rc = launch_walk(idx, P, E, w, b, ns, E.stalled.p, E.pending.p, true, minSA, &nl);
        if (rc != PBSC_OK) return rc;
        trace("stitch + re-walk round", ns);
    }
    trace("last stitch", 0);
    if (launches) *launches += nl;
    {
        // every launch was followed by a stream synchronisation (the stitch loop reads its counter), so the events are complete
        Timing& T = last_timing();
        T.walk_ms = 0; T.walk_launches = 0;
        for (size_t i = 0; i + 1 < E.ev.size(); i += 2)
        {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, E.ev[i], E.ev[i + 1]) == cudaSuccess) { T.walk_ms += ms; T.walk_launches++; }
            if (getenv("PBSC_ROUND_TRACE") && i / 2 < E.ev_what.size())
                fprintf(stderr, "[pbsc round trace] walk launch %zu: %s pass, %llu items, %.2f ms\n", i / 2, E.ev_what[i / 2].first, (unsigned long long)E.ev_what[i / 2].second, ms);
        }
    }
#ifdef PBSC_STAGE_CLOCKS
    {
        unsigned long long h[PBSC_N_STAGE], z[PBSC_N_STAGE] = {0};
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_stage_clk, sizeof h);
        cudaMemcpyToSymbol(g_stage_clk, z, sizeof z);
        static const char* names[9] = {"refill", "refine", "filter", "probe", "adopt+middle", "select", "prune", "term", "finish"};
        unsigned long long tot = 0;
        for (int i = 0; i < 9; i++) tot += h[i];
        fprintf(stderr, "[pbsc stage clocks] warp-cycles %llu; lane-iterations %llu, walking %.1f %%, leaves per walking lane %.2f\n", tot, h[9],
                h[9] ? 100.0 * (double)h[10] / (double)h[9] : 0.0, h[10] ? (double)h[11] / (double)h[10] : 0.0);
        for (int i = 0; i < 9; i++) fprintf(stderr, "[pbsc stage clocks]   %-14s %5.1f %%\n", names[i], tot ? 100.0 * (double)h[i] / (double)tot : 0.0);
    }
#endif
    return PBSC_OK;
}

}  // namespace pbsc
