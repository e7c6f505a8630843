// pbsc_dp_thread.cuh — Overlapper::extendMatch (Thirdparty/overlapper.cpp:421-701) as ONE ALIGNMENT PER THREAD.
//
// The warp-per-row kernel (dp_align_kernel) spends its instructions on lanes the matrix does not use: a query of the DP
// fallback is ~140 bases, the read retrieved for it ~1.1x that, so the 201-cell band is wider than the whole matrix, a band
// column holds ~150 computed cells of 224 lane slots, and every cell pays for the band-clipping predicates and the column
// max-scan.  Here a thread walks its own matrix column by column exactly like the reference's loop (no idle lanes, no
// predicates, no scan: the "up" dependency is a register), a warp holds 32 rows of (nearly) equal query length, and the
// previous column lives in shared memory as a 256-entry circular array of 16-bit scores indexed by the matrix row.
//
// The code below is plain scalar C++ over four small accessor types, so that tests/cpp/test_dp_thread.cpp can compile the
// very same fill/traceback with g++ and compare it with the oracle; the kernel instantiates it with shared-memory and
// interleaved-global accessors (pbsc_dp.cu).
//
// Rules reproduced (all relative to the reference's band table, band row r = j - (origin + i), 0 <= r <= 200):
//   * the table is zero-initialised and column 0 / row 0 are never written: they read as 0;
//   * first computed cell of a column: "up" is not consulted; "left" only if it is inside the band (r < 200);
//   * last computed cell (when it is not also the first): "left" is ignored, even if the band was clipped by the matrix
//     and the cell exists;
//   * the traceback compares the cell with ALL in-band neighbours (also the ones the fill ignored), a neighbour outside the
//     band never compares equal; its three tie-break orders are the reference's (:620-680).
// A computed cell only ever reads cells computed in the previous column, row 0, or (first computed column) cells nothing has
// written yet, so an in-place array indexed by the matrix row reproduces the zero-initialised table.  Queries longer than
// QMAX stay with the warp-per-row kernel.
#ifndef PBSC_DP_THREAD_CUH
#define PBSC_DP_THREAD_CUH

#include <stdint.h>

#if defined(__CUDACC__)
#define PBSC_DPT_HD __host__ __device__ __forceinline__
#else
#define PBSC_DPT_HD inline
#endif

namespace pbsc { namespace dpt {

constexpr int HALF = 100;              // bandwidth 200 (LongReadOverlap.cpp:629)
constexpr int BW = 2 * HALF + 1;       // cells per band column
constexpr int CPW = 10;                // cells per 32-bit word of traceback flags (3 bits each)
constexpr int WMAX = (BW + CPW - 1) / CPW;
constexpr int NEVER = -(1 << 28);      // a neighbour that is not there: never the maximum, never equal
constexpr int OP_M = 0, OP_I = 1, OP_D = 2;
constexpr int HSLOTS = 256;            // circular previous-column array: the live rows are [jb - 1, jb + 200]

// append "v != x" to a word of flags, for v >= x: x - v is negative exactly when they differ, and its sign bit is shifted in
// by one funnel shift (the caller inverts the finished word)
PBSC_DPT_HD uint32_t push_ne(uint32_t w, int v, int x)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l((uint32_t)(x - v), w, 1);
#else
    return (w << 1) | ((uint32_t)(x - v) >> 31);
#endif
}
PBSC_DPT_HD int imax(int a, int b) { return a > b ? a : b; }
PBSC_DPT_HD int imin(int a, int b) { return a < b ? a : b; }

// longest read retrieved for a query of qmax bases: query.length()*1.1+20 (LongReadOverlap.cpp:611), rounded up
PBSC_DPT_HD constexpr int max_len_bound(int qmax) { return qmax + qmax / 10 + 21; }

// the scores fit 16 bits (|score| <= 8 * qlen) and the read fits its shared-memory slot.  Columns whose band misses the
// matrix (the reference's `continue`, overlapper.cpp:470-471) can only be the first ones (band above the matrix) or the
// last ones (band below it); the fill skips them like the reference: the first computed column then reads a column array
// that is still all zero, like the reference's zero-initialised table, and nothing reads the trailing ones (the traceback
// starts in a computed cell and leaves the matrix before it could enter a skipped column).
PBSC_DPT_HD bool eligible(int qlen, int mlen, int /*origin*/, int qmax) { return qlen >= 1 && qlen <= qmax && mlen >= 1 && mlen <= max_len_bound(qmax); }
// flag words per column: ordinals count from the even row at or below the first computed one, the widest column has
// min(201, mlen) cells
PBSC_DPT_HD int words_per_col(int mlen) { return (imin(BW, mlen) + 1 + CPW - 1) / CPW; }

// Fill.  H: previous/current column, zero on entry: get(j) / set(j, v) one score; pairs_ok(j) / get2(j, u) / set2(j, u, w)
// the packed scores of rows j + 2u, j + 2u + 1 (j even, u = 0..4) when those ten rows do not wrap around the circular array.
// S: the read, S.base(x) = code of s2[x], S.bits(x) = 2-bit codes of s2[x .. x + 9] in bits 0..19.  F: flag words
// (put(n, w)).  q(x) = code of s1[x].
// Flags of cell (i, j): ordinal k = j - (first & ~1), word (i - 1) * W + k / 10, bits 3 * (9 - k % 10) ..+2 =
// {equals diagonal, equals up - 1, equals left - 1}.  Ordinals start at an even row so that whole flag words are five
// aligned pairs of rows: one 32-bit shared-memory load and store per two cells.
template <class HS, class SS, class FS, class QF>
PBSC_DPT_HD void fill(int qlen, int mlen, int origin, HS& H, const SS& S, FS& F, const QF& q, int& bi, int& bj)
{
    const int nRows = mlen + 1;
    const int W = words_per_col(mlen);
    int bestRowVal = 0, bestRowI = 0;
    bool anyRow = false;
    int wbase = 0;
#if defined(__CUDA_ARCH__)
    #pragma unroll 1
#endif
    for (int i = 1; i <= qlen; i++, wbase += W)
    {
        const int jb = origin + i;
        const int first = imax(jb, 1);
        const int last = imin(jb + BW, nRows) - 1;
        if (last < first) continue;   // the band misses the matrix
        const int c1 = q(i - 1);
        const uint32_t c1rep = (uint32_t)c1 * 0x55555u;
        int j = first;
        int diagOld = H.get(j - 1);
        int widx = wbase, kin = (first & 1) + 1;
        int up, v;
        uint32_t w;
        {
            // first cell of the column (overlapper.cpp:497-509)
            const int leftOld = H.get(j);
            const int d = diagOld + (S.base(j - 1) == c1 ? 1 : -8);
            const int l1 = (j == jb + BW - 1) ? NEVER : leftOld - 1;
            const int u1 = first > jb ? -1 : NEVER;   // row 0 of the matrix reads 0; above band row 0 there is nothing
            v = imax(d, l1);
            w = (v == d ? 1u : 0u) | (v == u1 ? 2u : 0u) | (v == l1 ? 4u : 0u);
            H.set(j, v);
            up = v; diagOld = leftOld;
            j++;
        }
        // middle cells [first + 1, last): all three neighbours
        auto rolled = [&](int jend)
        {
#if defined(__CUDA_ARCH__)
            #pragma unroll 1
#endif
            for (; j < jend; j++)
            {
                const int leftOld = H.get(j);
                const int d = diagOld + (S.base(j - 1) == c1 ? 1 : -8);
                const int l1 = leftOld - 1, u1 = up - 1;
                const int vv = imax(imax(d, l1), u1);
                w = (w << 3) | (vv == d ? 1u : 0u) | (vv == u1 ? 2u : 0u) | (vv == l1 ? 4u : 0u);
                H.set(j, vv);
                up = vv; diagOld = leftOld;
                if (++kin == CPW) { F.put(widx++, w); w = 0; kin = 0; }
            }
        };
        rolled(imin(last, (first & ~1) + CPW));   // up to the end of the first flag word
#if defined(__CUDA_ARCH__)
        #pragma unroll 1
#endif
        while (j + CPW <= last)                   // whole flag words: kin == 0 and j is even here
        {
            if (!H.pairs_ok(j)) { rolled(j + CPW); continue; }
            const uint32_t x = S.bits(j - 1) ^ c1rep;
            uint32_t ww = 0;
#if defined(__CUDA_ARCH__)
            #pragma unroll
#endif
            for (int u = 0; u < CPW / 2; u++)
            {
                const uint32_t old2 = H.get2(j, u);
                const int leftA = (int)(int16_t)(old2 & 0xFFFFu), leftB = (int)old2 >> 16;
                // max(d, left - 1, up - 1) = max(d + 1, left, up) - 1: the three candidates without their own decrements
                const int eA = diagOld + ((x & (3u << (4 * u))) == 0u ? 2 : -7);
                const int xA = imax(imax(eA, leftA), up);
                const int vA = xA - 1;
                const int eB = leftA + ((x & (12u << (4 * u))) == 0u ? 2 : -7);
                const int xB = imax(imax(eB, leftB), vA);
                const int vB = xB - 1;
                ww = push_ne(push_ne(push_ne(ww, xA, leftA), xA, up), xA, eA);   // bit 2 left, bit 1 up, bit 0 diagonal
                ww = push_ne(push_ne(push_ne(ww, xB, leftB), xB, vA), xB, eB);
                H.set2(j, u, ((uint32_t)vA & 0xFFFFu) | ((uint32_t)vB << 16));
                up = vB; diagOld = leftB;
            }
            F.put(widx++, ww ^ 0x3FFFFFFFu);   // "differs" bits -> "equals" bits
            j += CPW;
        }
        rolled(last);
        if (last > first)
        {
            // last cell of the column (:510-516): no left in the fill; the traceback still sees it when it is in the band
            const int d = diagOld + (S.base(last - 1) == c1 ? 1 : -8);
            const int u1 = up - 1;
            const int l1 = (last == jb + BW - 1) ? NEVER : H.get(last) - 1;
            v = imax(d, u1);
            w = (w << 3) | (v == d ? 1u : 0u) | (v == u1 ? 2u : 0u) | (v == l1 ? 4u : 0u);
            H.set(last, v);
            if (++kin == CPW) { F.put(widx++, w); w = 0; kin = 0; }
        }
        if (kin) F.put(widx, w << (3 * (CPW - kin)));
        // best cell of the last row: first column with the strictly largest score (:553-561)
        if (last == nRows - 1 && (!anyRow || v > bestRowVal)) { bestRowVal = v; bestRowI = i; anyRow = true; }
    }
    // best cell of the last column: first row with the strictly largest score (:564-570); H holds that column now
    int bestColVal = 0, bestColJ = 0;
    bool anyCol = false;
    {
        const int jb = origin + qlen;
        const int first = imax(jb, 1), last = imin(jb + BW, nRows) - 1;
        for (int j = first; j <= last; j++)
        {
            const int v = H.get(j);
            if (!anyCol || v > bestColVal) { bestColVal = v; bestColJ = j; anyCol = true; }
        }
    }
    // start of the traceback (:577-586)
    if (anyCol && (!anyRow || bestColVal > bestRowVal)) { bi = qlen; bj = bestColJ; }
    else { bi = anyRow ? bestRowI : 0; bj = nRows - 1; }
    if (!(anyRow || anyCol) || bi <= 0 || bj <= 0) bi = 0;
}

// Traceback from (bi, bj), bi > 0.  ops.put(n, op) receives the alignment columns last column first.
template <class SS, class FS, class QF, class OS>
PBSC_DPT_HD void traceback(int qlen, int mlen, int origin, const SS& S, const FS& F, const QF& q, int bi, int bj, OS& ops,
                           int& n_out, int& ed_out, int& i_out, int& j_out)
{
    const int W = words_per_col(mlen);
    int i = bi, j = bj, n = 0, ed = 0;
    int ckey = -1;
    uint32_t cword = 0;
#if defined(__CUDA_ARCH__)
    #pragma unroll 1
#endif
    while (i > 0 && j > 0)
    {
        const int k = j - (imax(origin + i, 1) & ~1);
        const int kw = k / CPW;
        const int key = (i - 1) * W + kw;
        if (key != ckey) { cword = F.get(key); ckey = key; }
        const uint32_t f = (cword >> (3 * (CPW - 1 - (k - kw * CPW)))) & 7u;
        const int s2p = S.base(j - 1), s2n = j < mlen ? S.base(j) : -1;
        const int s1p = q(i - 1), s1n = i < qlen ? q(i) : -2;
        const bool eqD = (f & 1u) != 0, eqU = (f & 2u) != 0, eqL = (f & 4u) != 0;
        int op;
        if (s2p == s2n) op = eqU ? OP_I : (eqL ? OP_D : OP_M);        // s2 homopolymer: prefer consuming s2 (:620-640)
        else if (s1p == s1n) op = eqL ? OP_D : (eqU ? OP_I : OP_M);   // s1 homopolymer: prefer consuming s1 (:642-661)
        else op = eqD ? OP_M : (eqL ? OP_D : OP_I);                   // (:663-680)
        if (op == OP_M) { if (s1p != s2p) ed++; i--; j--; }
        else if (op == OP_I) { ed++; j--; }
        else { ed++; i--; }
        ops.put(n, op);
        n++;
    }
    n_out = n; ed_out = ed; i_out = i; j_out = j;
}

}}  // namespace pbsc::dpt

#endif
