// pbsc_dp_msa.cuh — multiple alignment of a failed walk's pile-up on per-column counts, and its consensus
// (MultipleAlignment::_addSequence + calculateBaseConsensus, Thirdparty/multiple_alignment.cpp:240-393, 517-594).
//
// Column model.  MultipleAlignment::_addSequence places every incoming row against row 0 (the query); a new gap column is
// only ever inserted immediately before a BASE column of row 0, i.e. appended to the run of gap columns in front of that
// base, so the alignment is: for each query position p a list run(p) of inserted columns, then the base column p.  The
// consensus only needs, per column, how many rows show A/C/G/T/'-' there.  A new column inserted before base column p gets a
// '-' from every row that already spans it: rows covering base column p minus rows whose first column IS base column p
// (MultipleAlignmentElement::insertGapBeforeColumn, multiple_alignment.cpp:112-134).
//
// Plain scalar C++ over raw pointers: dp_msa_kernel (pbsc_dp.cu) runs it one job per thread, and tests/cpp/test_dp_msa.cpp
// compiles the same function with g++ and checks it against the pile-ups answered by the reference's own MultipleAlignment
// (tests/golden/dp_units*.txt, written by oracle/_ref/dp_dump).
#ifndef PBSC_DP_MSA_CUH
#define PBSC_DP_MSA_CUH

#include <stdint.h>

#if defined(__CUDACC__)
#define PBSC_MSA_HD __host__ __device__ __forceinline__
#define PBSC_MSA_ALIGN(n) __align__(n)
#else
#define PBSC_MSA_HD inline
#define PBSC_MSA_ALIGN(n) alignas(n)
#endif

namespace pbsc {

constexpr int OP_M = 0, OP_I = 1, OP_D = 2;

struct PBSC_MSA_ALIGN(16) DpRow
{
    uint32_t job, local;       // job index, row index inside the job
    uint32_t len, seq_start;   // the retrieved read is buf[seq_start, seq_start + len)
    uint32_t nops;             // alignment columns (ops are stored last column first)
    int32_t start0, start1;    // match[0].start, match[1].start
    uint32_t pass;             // 0 dropped, 1 enters the multiple alignment, 2 retrieved and waiting for alignment
};

struct PBSC_MSA_ALIGN(8) GapCol { uint16_t cnt[5]; uint16_t pad; uint32_t next; };   // one inserted column: A,C,G,T,'-' counts

// scratch of one job: the query, one slot per row (retrieved read, then its alignment columns), the column counts
struct JobView
{
    uint8_t* q; uint8_t* rows;
    uint16_t* baseCnt; uint16_t* startAt; uint32_t* head; uint32_t* tail; GapCol* pool;
    uint64_t rowBytes, seqBytes;
};

PBSC_MSA_HD uint32_t dp_gap_cap(uint32_t qlen) { return 3 * qlen + 128; }

namespace msa {

// R[0, nr): the job's rows.  ctx.min_call() = min_call_coverage, ctx.out() / ctx.cap() = where the consensus goes (asked
// for only when they are needed).  Returns 0: n_out consensus bases written; 1: fewer than three rows passed the filters
// (maquery.getNumRows() <= 3, PacBioSelfCorrectionProcess.cpp:238); 2: outside this build's limits (more inserted columns
// than dp_gap_cap, consensus longer than its slot or shorter than k).
template <class Ctx>
PBSC_MSA_HD int consensus(const JobView& v, const uint32_t qlen, const uint32_t k, const DpRow* R, const uint32_t nr, const Ctx& ctx, uint32_t& n_out)
{
    uint32_t passing = 0;
    for (uint32_t r = 0; r < nr; r++) passing += R[r].pass == 1;
    if (passing < 3) return 1;
    const uint8_t* q = v.q;
    for (uint32_t p = 0; p <= qlen; p++)
    {
#if defined(__CUDA_ARCH__)
        #pragma unroll
#endif
        for (int c = 0; c < 5; c++) v.baseCnt[p * 5 + c] = 0;
        v.startAt[p] = 0; v.head[p] = 0; v.tail[p] = 0;
        if (p < qlen) v.baseCnt[p * 5 + q[p]] = 1;
    }
    v.startAt[0] = 1;   // row 0 starts at base column 0
    const uint32_t gapCap = dp_gap_cap(qlen);
    uint32_t nGap = 0;
    bool bad = false;
    for (uint32_t r = 0; r < nr && !bad; r++)
    {
        if (R[r].pass != 1) continue;
        const uint8_t* buf = v.rows + (uint64_t)R[r].local * v.rowBytes;
        const uint8_t* s2 = buf + R[r].seq_start;
        const uint8_t* ops = buf + v.seqBytes;
        uint32_t p = (uint32_t)R[r].start0, inc = (uint32_t)R[r].start1;
        uint32_t cur = 0;   // 0: at base column p; otherwise gap column cur-1 of run(p)
        bool firstOp = true;
        int c = (int)R[r].nops - 1;
        while (c >= 0)
        {
            const int op = ops[c];
            if (cur)
            {
                GapCol& g = v.pool[cur - 1];
                if (op == OP_I) { g.cnt[s2[inc]]++; inc++; c--; firstOp = false; }
                else g.cnt[4]++;
                cur = g.next;
            }
            else if (op == OP_I)
            {
                if (nGap >= gapCap) { bad = true; break; }
                uint32_t cover = 0;
        #if defined(__CUDA_ARCH__)
        #pragma unroll
#endif
                for (int s = 0; s < 5; s++) cover += v.baseCnt[p * 5 + s];
                GapCol g;
                g.cnt[0] = g.cnt[1] = g.cnt[2] = g.cnt[3] = 0; g.pad = 0; g.next = 0;
                g.cnt[4] = (uint16_t)(cover - v.startAt[p]);
                g.cnt[s2[inc]] = 1;
                v.pool[nGap] = g;
                nGap++;
                if (v.tail[p]) v.pool[v.tail[p] - 1].next = nGap; else v.head[p] = nGap;
                v.tail[p] = nGap;
                inc++; c--; firstOp = false;
            }
            else
            {
                if (p >= qlen) { bad = true; break; }
                v.baseCnt[p * 5 + (op == OP_M ? s2[inc] : 4)]++;
                if (op == OP_M) inc++;
                if (firstOp) v.startAt[p]++;
                firstOp = false;
                p++; c--;
                cur = v.head[p];
            }
        }
    }
    if (bad) return 2;
    // calculateBaseConsensus(min_call_coverage, -1) (multiple_alignment.cpp:517-594)
    const int minCall = ctx.min_call();
    uint8_t* out = ctx.out();
    uint32_t n = 0;
    const uint32_t cap = ctx.cap();
    auto call = [&](const uint16_t* cnt, int baseSym) -> int
    {
        int maxSym = -1, maxCount = -1;
#if defined(__CUDA_ARCH__)
        #pragma unroll
#endif
        for (int s = 0; s < 5; s++) if ((int)cnt[s] > maxCount) { maxSym = s; maxCount = cnt[s]; }   // order A,C,G,T,(N),'-'
        const int baseCount = cnt[baseSym];
        return (maxCount >= baseCount && baseCount < minCall) ? maxSym : baseSym;
    };
    bool over = false;
    for (uint32_t p = 0; p < qlen && !over; p++)
    {
        if (p >= 1)
            for (uint32_t g = v.head[p]; g; g = v.pool[g - 1].next)
            {
                const int s = call(v.pool[g - 1].cnt, 4);
                if (s != 4) { if (n >= cap) { over = true; break; } out[n++] = (uint8_t)s; }
            }
        if (over) break;
        const int s = call(v.baseCnt + p * 5, (int)q[p]);
        if (s != 4) { if (n >= cap) { over = true; break; } out[n++] = (uint8_t)s; }
    }
    // out.erase(0, extendKmerSize) needs at least k bases
    if (over || n < k) return 2;
    n_out = n;
    return 0;
}

#if defined(__CUDACC__)
// The same multiple alignment by ONE WARP per job, for the jobs whose pile-up is large (hundreds of rows over a query of a
// thousand bases: one thread would walk ~10^5-10^6 alignment columns while the rest of the launch waits for it).  Rows are
// still added one after the other -- the column structure depends on their order -- but the alignment columns of a row are cut
// into 32 consecutive segments, one per lane.  A segment may only start right after a base column (never inside a run of
// insertions), so that the run of inserted columns in front of base column p and base column p itself are always touched by the
// same lane: within one row no two lanes write the same counter, list head or tail, and every lane's state at the start of its
// segment (query position, read position) is a prefix count over the segments before it.  New gap columns come from a shared
// counter (their numbering differs from the one-thread version, the linked lists and counts do not).
// All 32 lanes must call it converged.  `gap_counter`: one word of shared memory of this warp.
template <class Ctx>
__device__ __forceinline__ int consensus_warp(const JobView& v, const uint32_t qlen, const uint32_t k, const DpRow* R, const uint32_t nr, const Ctx& ctx,
                                              uint32_t& n_out, unsigned int* gap_counter)
{
    const unsigned FULLW = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    uint32_t passing = 0;
    for (uint32_t r = lane; r < nr; r += 32) passing += R[r].pass == 1;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) passing += __shfl_xor_sync(FULLW, passing, o);
    if (passing < 3) return 1;
    const uint8_t* q = v.q;
    for (uint32_t p = lane; p <= qlen; p += 32)
    {
        #pragma unroll
        for (int c = 0; c < 5; c++) v.baseCnt[p * 5 + c] = 0;
        v.startAt[p] = p == 0 ? 1 : 0; v.head[p] = 0; v.tail[p] = 0;   // row 0 starts at base column 0
        if (p < qlen) v.baseCnt[p * 5 + q[p]] = 1;
    }
    if (lane == 0) *gap_counter = 0;
    __syncwarp();
    const uint32_t gapCap = dp_gap_cap(qlen);
    bool bad = false;
    for (uint32_t r = 0; r < nr; r++)
    {
        if (R[r].pass != 1) continue;   // uniform: every lane reads the same row record
        const uint8_t* buf = v.rows + (uint64_t)R[r].local * v.rowBytes;
        const uint8_t* s2 = buf + R[r].seq_start;
        const uint8_t* ops = buf + v.seqBytes;
        const int nops = (int)R[r].nops;
        // forward index t = 0 .. nops-1 is column ops[nops-1-t] (columns are stored last first)
        const int chunk = (nops + 31) / 32;
        int t0 = min(lane * chunk, nops);
        while (t0 > 0 && t0 < nops && ops[nops - t0] == OP_I) t0++;   // op(t0 - 1) is an insertion: not a segment start
        int t1 = __shfl_down_sync(FULLW, t0, 1);
        if (lane == 31) t1 = nops;
        // query / read positions consumed by this segment, then the prefix over the segments before it
        uint32_t md = 0, mi = 0;
        for (int t = t0; t < t1; t++) { const int op = ops[nops - 1 - t]; md += op != OP_I; mi += op != OP_D; }
        uint32_t pmd = md, pmi = mi;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t a = __shfl_up_sync(FULLW, pmd, o), b = __shfl_up_sync(FULLW, pmi, o);
            if (lane >= o) { pmd += a; pmi += b; }
        }
        uint32_t p = (uint32_t)R[r].start0 + (pmd - md), inc = (uint32_t)R[r].start1 + (pmi - mi);
        bool firstOp = t0 == 0;
        uint32_t cur = (t0 == 0 || t0 >= t1) ? 0 : v.head[p];   // 0: at base column p; otherwise gap column cur-1 of run(p)
        int t = t0;
        while (t < t1)
        {
            const int op = ops[nops - 1 - t];
            if (cur)
            {
                GapCol& g = v.pool[cur - 1];
                if (op == OP_I) { g.cnt[s2[inc]]++; inc++; t++; firstOp = false; }
                else g.cnt[4]++;
                cur = g.next;
            }
            else if (op == OP_I)
            {
                const uint32_t slot = atomicAdd(gap_counter, 1u);
                if (slot >= gapCap) { bad = true; break; }
                uint32_t cover = 0;
                #pragma unroll
                for (int s = 0; s < 5; s++) cover += v.baseCnt[p * 5 + s];
                GapCol g;
                g.cnt[0] = g.cnt[1] = g.cnt[2] = g.cnt[3] = 0; g.pad = 0; g.next = 0;
                g.cnt[4] = (uint16_t)(cover - v.startAt[p]);
                g.cnt[s2[inc]] = 1;
                v.pool[slot] = g;
                if (v.tail[p]) v.pool[v.tail[p] - 1].next = slot + 1; else v.head[p] = slot + 1;
                v.tail[p] = slot + 1;
                inc++; t++; firstOp = false;
            }
            else
            {
                if (p >= qlen) { bad = true; break; }
                v.baseCnt[p * 5 + (op == OP_M ? s2[inc] : 4)]++;
                if (op == OP_M) inc++;
                if (firstOp) v.startAt[p]++;
                firstOp = false;
                p++; t++;
                cur = v.head[p];
            }
        }
        __syncwarp();
        if (__any_sync(FULLW, bad)) return 2;
    }
    // calculateBaseConsensus(min_call_coverage, -1) (multiple_alignment.cpp:517-594): lane l calls a block of consecutive query
    // positions; the number of bases each block emits gives every lane its place in the output
    const int minCall = ctx.min_call();
    uint8_t* out = ctx.out();
    const uint32_t cap = ctx.cap();
    auto call = [&](const uint16_t* cnt, int baseSym) -> int
    {
        int maxSym = -1, maxCount = -1;
        #pragma unroll
        for (int s = 0; s < 5; s++) if ((int)cnt[s] > maxCount) { maxSym = s; maxCount = cnt[s]; }   // order A,C,G,T,(N),'-'
        const int baseCount = cnt[baseSym];
        return (maxCount >= baseCount && baseCount < minCall) ? maxSym : baseSym;
    };
    const uint32_t blk = (qlen + 31) / 32;
    const uint32_t pa = min((uint32_t)lane * blk, qlen), pb = min(pa + blk, qlen);
    uint32_t n = 0;
    #pragma unroll 1
    for (int pass = 0; pass < 2; pass++)
    {
        uint32_t w = 0;
        if (pass == 1)
        {
            uint32_t incl = n;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t a = __shfl_up_sync(FULLW, incl, o); if (lane >= o) incl += a; }
            const uint32_t total = __shfl_sync(FULLW, incl, 31);
            if (total > cap || total < k) return 2;   // out.erase(0, extendKmerSize) needs at least k bases
            n_out = total;
            w = incl - n;
        }
        for (uint32_t p = pa; p < pb; p++)
        {
            if (p >= 1)
                for (uint32_t g = v.head[p]; g; g = v.pool[g - 1].next)
                {
                    const int s = call(v.pool[g - 1].cnt, 4);
                    if (s != 4) { if (pass == 0) n++; else out[w++] = (uint8_t)s; }
                }
            const int s = call(v.baseCnt + p * 5, (int)q[p]);
            if (s != 4) { if (pass == 0) n++; else out[w++] = (uint8_t)s; }
        }
    }
    return 0;
}
#endif

}  // namespace msa
}  // namespace pbsc

#endif
