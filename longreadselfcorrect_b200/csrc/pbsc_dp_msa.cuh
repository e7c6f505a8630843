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

}  // namespace msa
}  // namespace pbsc

#endif
