// fm_table.cuh — GPU-resident flat occurrence table and the backward-search primitives.
//
// Replaces RLBWT::getOcc / getPC (SuffixTools/RLBWT.h:118-140; run-length units + two marker
// levels, ~3 dependent cache lines per query) with one aligned 32-byte sector per occ(c,i):
//
//   struct FmBlock { u32 cnt[4]; u32 bases[4]; }      64 BWT symbols per block
//     cnt[c]   = occurrences of base c (A,C,G,T = 0..3) in bwt[0 .. 64*blk)
//     bases    = the 64 symbols, 2 bits each, symbol j in word j>>4 at bits 2*(j&15)
//     '$' is stored as code 0; bit 31 of cnt[0] flags a block that contains a '$', in which
//     case occ(A,.) is corrected from a side bit-vector (one 64-bit word per block, read only for flagged
//     blocks: one '$' per read, under 1 % of the blocks).
//
// Intervals are kept half-open [lo, hi) on the device; the reference's inclusive
// (lower, upper) = (lo, hi-1).  Invalid intervals always have hi == lo (BWTAlgorithms.h:66-72
// keeps upper == lower-1 once an interval is empty), so size 0 <=> !isValid().
#ifndef PBSC_FM_TABLE_CUH
#define PBSC_FM_TABLE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace pbsc {

struct __align__(32) FmBlock
{
    uint32_t cnt[4];
    uint32_t bases[4];
};

// 32-byte prefix-table entry: both strands' intervals of one k0-mer w (key = sum_j w[j]*4^j, w[0] in the low bits):
//   fwd = interval of reverse(w) in RBWT, rvc = interval of revcomp(w) in BWT.
struct __align__(32) PrefixEntry
{
    uint64_t fwd_lo, rvc_lo;
    uint32_t fwd_size, rvc_size;
    uint32_t pad[2];
};

struct FmTable
{
    const FmBlock* blocks;      // n/64 + 1 blocks
    const uint32_t* dollar_pos; // sorted positions of '$' in the BWT
    const uint64_t* dollar_mask;// one bit per BWT position, word per block: which symbols of a flagged block are '$'
    uint64_t n;                 // BWT length (symbols incl. '$')
    uint64_t C[4];              // C[c] = #symbols lexicographically smaller than base c (RLBWT.cpp:243-247)
    uint64_t total[4];          // occurrences of each base in the whole BWT
    uint32_t n_dollar;
};

struct FmIndexDev
{
    FmTable t[2];               // [PBSC_BWT], [PBSC_RBWT]
    const PrefixEntry* prefix;  // 4^k0 entries or nullptr
    int k0;
    // per idmer (key = sum_j w[j]*4^j): bit 0 = reverse(w) occurs in RBWT, bit 1 = revcomp(w) occurs in BWT
    const uint8_t* idmer_valid;
    int idmer_len;
};

struct Interval   // half-open
{
    uint64_t lo, hi;
    __host__ __device__ __forceinline__ bool valid() const { return hi > lo; }
    __host__ __device__ __forceinline__ uint64_t size() const { return hi - lo; }
};

// The primitives below also compile for the host (tests/cpp/test_fm_occ.cpp checks them against naive counting).
#if defined(__CUDA_ARCH__)
#define PBSC_LDG(p) __ldg(p)
#define PBSC_POPCLL(x) __popcll(x)
#else
#define PBSC_LDG(p) (*(p))
#define PBSC_POPCLL(x) __builtin_popcountll(x)
#endif

// -DPBSC_COUNT_OCC: a measurement build (libpbsc_count.so) that counts the DISTINCT 32-byte sectors of the rank tables and
// of the prefix table every lookup asks for ("issued sectors"): one per occ / occ4 / lf_step / prefix_lookup, one
// (not two) for an updateInterval whose two bounds fall into the same block.  Host code reads and clears the counter at the
// phase boundaries (occ_take below; every translation unit has its own copy of the symbol).
#ifdef PBSC_COUNT_OCC
static __device__ unsigned long long g_occ_counter;
#if defined(__CUDA_ARCH__)
#define PBSC_OCC_TICK(n) atomicAdd(&g_occ_counter, (unsigned long long)(long long)(n))
#else
#define PBSC_OCC_TICK(n)
#endif
#else
#define PBSC_OCC_TICK(n)
#endif

__host__ __device__ __forceinline__ uint32_t count_dollars(const FmTable& t, uint64_t from, uint64_t to)
{
    // number of '$' positions in [from, to): two lower_bounds on the sorted list
    uint32_t lo = 0, hi = t.n_dollar;
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if ((uint64_t)PBSC_LDG(t.dollar_pos + m) < from) lo = m + 1; else hi = m; }
    uint32_t a = lo;
    hi = t.n_dollar;
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if ((uint64_t)PBSC_LDG(t.dollar_pos + m) < to) lo = m + 1; else hi = m; }
    return lo - a;
}

// occ(c, p): occurrences of base c in bwt[0, p)  ==  RLBWT::getOcc(c, p-1)  (RLBWT.h:121-140)
__host__ __device__ __forceinline__ uint64_t occ(const FmTable& t, int c, uint64_t p)
{
    PBSC_OCC_TICK(1);
    const uint64_t blk = p >> 6;
    const uint32_t off = (uint32_t)p & 63u;
    const uint4* bp = reinterpret_cast<const uint4*>(t.blocks + blk);
    const uint4 cn = PBSC_LDG(bp);
    const uint4 bs = PBSC_LDG(bp + 1);
    uint32_t base = c == 0 ? cn.x : c == 1 ? cn.y : c == 2 ? cn.z : cn.w;
    const bool has_dollar = (cn.x >> 31) != 0;
    if (c == 0) base &= 0x7fffffffu;
    const uint64_t w0 = (uint64_t)bs.x | ((uint64_t)bs.y << 32);
    const uint64_t w1 = (uint64_t)bs.z | ((uint64_t)bs.w << 32);
    const uint64_t pat = 0x5555555555555555ull * (uint64_t)c;
    uint64_t x0 = w0 ^ pat, x1 = w1 ^ pat;
    uint64_t m0 = ~(x0 | (x0 >> 1)) & 0x5555555555555555ull;
    uint64_t m1 = ~(x1 | (x1 >> 1)) & 0x5555555555555555ull;
    // keep the first `off` symbols
    if (off < 32) { m0 &= (1ull << (2 * off)) - 1ull; m1 = 0; }
    else { m1 &= (1ull << (2 * (off - 32))) - 1ull; }
    uint64_t r = (uint64_t)base + PBSC_POPCLL(m0) + PBSC_POPCLL(m1);
    if (c == 0 && has_dollar && off) r -= PBSC_POPCLL(PBSC_LDG(t.dollar_mask + blk) & ((1ull << off) - 1ull));
    return r;
}

// occurrences of all four bases in bwt[0, p) from one 32-byte sector
__host__ __device__ __forceinline__ void occ4(const FmTable& t, uint64_t p, uint64_t r[4])
{
    PBSC_OCC_TICK(1);
    const uint64_t blk = p >> 6;
    const uint32_t off = (uint32_t)p & 63u;
    const uint4* bp = reinterpret_cast<const uint4*>(t.blocks + blk);
    const uint4 cn = PBSC_LDG(bp);
    const uint4 bs = PBSC_LDG(bp + 1);
    const uint64_t w0 = (uint64_t)bs.x | ((uint64_t)bs.y << 32);
    const uint64_t w1 = (uint64_t)bs.z | ((uint64_t)bs.w << 32);
    const uint64_t M = 0x5555555555555555ull;
    uint64_t k0, k1;   // position masks: one bit per symbol kept
    if (off < 32) { k0 = ((1ull << (2 * off)) - 1ull) & M; k1 = 0; }
    else { k0 = M; k1 = ((1ull << (2 * (off - 32))) - 1ull) & M; }
    const uint64_t l0 = w0 & M, h0 = (w0 >> 1) & M, l1 = w1 & M, h1 = (w1 >> 1) & M;
    const uint32_t nT = PBSC_POPCLL(h0 & l0 & k0) + PBSC_POPCLL(h1 & l1 & k1);
    const uint32_t nG = PBSC_POPCLL(h0 & ~l0 & k0) + PBSC_POPCLL(h1 & ~l1 & k1);
    const uint32_t nC = PBSC_POPCLL(~h0 & l0 & k0) + PBSC_POPCLL(~h1 & l1 & k1);
    uint32_t nA = off - nT - nG - nC;
    if ((cn.x >> 31) && off) nA -= PBSC_POPCLL(PBSC_LDG(t.dollar_mask + blk) & ((1ull << off) - 1ull));
    r[0] = (uint64_t)(cn.x & 0x7fffffffu) + nA;
    r[1] = (uint64_t)cn.y + nC;
    r[2] = (uint64_t)cn.z + nG;
    r[3] = (uint64_t)cn.w + nT;
}

// occ(c, lo) and occ(c, hi) for lo <= hi: one sector load and one decode when both fall into the same 64-symbol block (the
// usual case once an interval is down to ~coverage rows).  EXPERIMENT: used by update_interval only when the library is built
// with -DPBSC_FUSED_UPDATE (not measured yet); checked against occ() on the host.
__host__ __device__ __forceinline__ void occ_pair(const FmTable& t, int c, uint64_t lo, uint64_t hi, uint64_t& a, uint64_t& b)
{
    if ((lo >> 6) != (hi >> 6)) { a = occ(t, c, lo); b = occ(t, c, hi); return; }
    PBSC_OCC_TICK(1);
    const uint64_t blk = lo >> 6;
    const uint32_t off_a = (uint32_t)lo & 63u, off_b = (uint32_t)hi & 63u;
    const uint4* bp = reinterpret_cast<const uint4*>(t.blocks + blk);
    const uint4 cn = PBSC_LDG(bp);
    const uint4 bs = PBSC_LDG(bp + 1);
    uint32_t base = c == 0 ? cn.x : c == 1 ? cn.y : c == 2 ? cn.z : cn.w;
    const bool has_dollar = (cn.x >> 31) != 0;
    if (c == 0) base &= 0x7fffffffu;
    const uint64_t w0 = (uint64_t)bs.x | ((uint64_t)bs.y << 32);
    const uint64_t w1 = (uint64_t)bs.z | ((uint64_t)bs.w << 32);
    const uint64_t pat = 0x5555555555555555ull * (uint64_t)c;
    const uint64_t x0 = w0 ^ pat, x1 = w1 ^ pat;
    const uint64_t m0 = ~(x0 | (x0 >> 1)) & 0x5555555555555555ull;
    const uint64_t m1 = ~(x1 | (x1 >> 1)) & 0x5555555555555555ull;
    // symbols [0, off) of the block: all of word 0 and off - 32 of word 1 when off >= 32
    const uint64_t ka0 = off_a < 32 ? (1ull << (2 * off_a)) - 1ull : ~0ull, ka1 = off_a < 32 ? 0ull : (1ull << (2 * (off_a - 32))) - 1ull;
    const uint64_t kb0 = off_b < 32 ? (1ull << (2 * off_b)) - 1ull : ~0ull, kb1 = off_b < 32 ? 0ull : (1ull << (2 * (off_b - 32))) - 1ull;
    a = (uint64_t)base + PBSC_POPCLL(m0 & ka0) + PBSC_POPCLL(m1 & ka1);
    b = (uint64_t)base + PBSC_POPCLL(m0 & kb0) + PBSC_POPCLL(m1 & kb1);
    if (c == 0 && has_dollar && off_b)
    {
        const uint64_t dm = PBSC_LDG(t.dollar_mask + blk);
        if (off_a) a -= PBSC_POPCLL(dm & ((1ull << off_a) - 1ull));
        b -= PBSC_POPCLL(dm & ((1ull << off_b) - 1ull));
    }
}

// BWTAlgorithms::updateInterval (SuffixTools/BWTAlgorithms.h:66-72) on a half-open interval
__host__ __device__ __forceinline__ Interval update_interval(const FmTable& t, Interval iv, int c)
{
    Interval r;
#ifdef PBSC_FUSED_UPDATE
    uint64_t a, b;
    occ_pair(t, c, iv.lo, iv.hi, a, b);
#else
    const uint64_t a = occ(t, c, iv.lo);
    const uint64_t b = (iv.hi == iv.lo) ? a : occ(t, c, iv.hi);
    if (iv.hi != iv.lo && (iv.hi >> 6) == (iv.lo >> 6)) PBSC_OCC_TICK(-1);   // the second load hit the sector the first one fetched
#endif
    r.lo = t.C[c] + a;
    r.hi = t.C[c] + b;
    return r;
}

// BWTAlgorithms::initInterval (BWTAlgorithms.h:136-140)
__host__ __device__ __forceinline__ Interval init_interval(const FmTable& t, int c)
{
    Interval r;
    r.lo = t.C[c];
    r.hi = t.C[c] + t.total[c];
    return r;
}

// BWT symbol at idx and the LF step from it: returns the symbol (0..3) or -1 for '$'; idx becomes C[b] + occ(b, idx - 1)
// (RLBWT::getChar + getPC + getOcc, LongReadOverlap.cpp:713-718) from one 32-byte sector
__host__ __device__ __forceinline__ int lf_step(const FmTable& t, uint64_t& idx)
{
    PBSC_OCC_TICK(1);
    const uint64_t blk = idx >> 6;
    const uint32_t off = (uint32_t)idx & 63u;
    const uint4* bp = reinterpret_cast<const uint4*>(t.blocks + blk);
    const uint4 cn = PBSC_LDG(bp);
    const uint4 bs = PBSC_LDG(bp + 1);
    const uint64_t w0 = (uint64_t)bs.x | ((uint64_t)bs.y << 32);
    const uint64_t w1 = (uint64_t)bs.z | ((uint64_t)bs.w << 32);
    const int c = (int)(((off < 32 ? w0 : w1) >> (2 * (off & 31))) & 3);
    const bool has_dollar = (cn.x >> 31) != 0;
    uint64_t dmask = 0;
    if (has_dollar)
    {
        dmask = PBSC_LDG(t.dollar_mask + blk);
        if ((dmask >> off) & 1) return -1;
    }
    uint32_t base = c == 0 ? (cn.x & 0x7fffffffu) : c == 1 ? cn.y : c == 2 ? cn.z : cn.w;
    const uint64_t pat = 0x5555555555555555ull * (uint64_t)c;
    const uint64_t x0 = w0 ^ pat, x1 = w1 ^ pat;
    uint64_t m0 = ~(x0 | (x0 >> 1)) & 0x5555555555555555ull;
    uint64_t m1 = ~(x1 | (x1 >> 1)) & 0x5555555555555555ull;
    if (off < 32) { m0 &= (1ull << (2 * off)) - 1ull; m1 = 0; }
    else { m1 &= (1ull << (2 * (off - 32))) - 1ull; }
    uint64_t r = (uint64_t)base + PBSC_POPCLL(m0) + PBSC_POPCLL(m1);
    if (c == 0 && has_dollar && off) r -= PBSC_POPCLL(dmask & ((1ull << off) - 1ull));
    idx = t.C[c] + r;
    return c;
}

// read one prefix-table entry (one 32-byte sector) and return the strand the caller wants
__host__ __device__ __forceinline__ void prefix_lookup(const FmIndexDev& idx, uint64_t key, Interval& fwd, Interval& rvc)
{
    PBSC_OCC_TICK(1);
    const uint4* ep = reinterpret_cast<const uint4*>(idx.prefix + key);
    const uint4 a = PBSC_LDG(ep), b = PBSC_LDG(ep + 1);
    fwd.lo = (uint64_t)a.x | ((uint64_t)a.y << 32); fwd.hi = fwd.lo + b.x;
    rvc.lo = (uint64_t)a.z | ((uint64_t)a.w << 32); rvc.hi = rvc.lo + b.y;
}

#ifdef PBSC_COUNT_OCC
// sectors counted by this translation unit's kernels since the last call (synchronises the stream)
static inline unsigned long long occ_take(cudaStream_t st)
{
    unsigned long long v = 0, z = 0;
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(&v, g_occ_counter, sizeof v);
    cudaMemcpyToSymbol(g_occ_counter, &z, sizeof z);
    return v;
}
#define PBSC_OCC_TAKE(slot, st) (pbsc::occ_counts()[slot] += pbsc::occ_take(st))
#else
#define PBSC_OCC_TAKE(slot, st)
#endif
// issued sectors per kernel family of the measurement build: [0] seed phase, [1] walk setup (setup_tasks_kernel),
// [2] level loop (walk_levels_kernel), [3] DP fallback (collect + retrieve), [4] everything else
unsigned long long* occ_counts();

}  // namespace pbsc
#endif
