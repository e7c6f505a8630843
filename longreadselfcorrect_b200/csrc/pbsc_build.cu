// pbsc_build.cu — `stride index` (StriDe/index.cpp:86-214) on the GPU: the BWT of the read collection and of the reversed
// reads as the run-length files the reference reads back (PREFIX.bwt / PREFIX.rbwt), and the lexicographic read order
// (PREFIX.sai / PREFIX.rsai).
//
// What the reference computes.  BWTCA::runRopebwt2 (SuffixTools/BWTCARopebwt.cpp:160-247) inserts the reads into ropebwt2 in
// input order (MR_SO_IO): the result is the BWT of the collection with one sentinel after every read, the sentinel of read i
// smaller than the sentinel of read j for i < j and every sentinel smaller than A < C < G < T.  BWTWriterBinary
// (SuffixTools/BWTWriterBinary.cpp:28-94) writes it as run-length units of at most 31 symbols (RLUnit.h:13-16) behind a 30-byte
// header; SampledSuffixArray::buildLexicoIndex / writeLexicoIndex (SuffixTools/SampledSuffixArray.cpp:158-190,248-258) writes,
// for the r-th '$' of the BWT, the index of the read that row is the first base of.  The files written here are byte-identical
// to the reference's (tests/golden/tiny.*, index_edge.*: duplicates, one-base reads, reads that are prefixes of others).
//
// How it is computed here.  The collection is one text in HBM (a byte per symbol, 0 = sentinel).  Suffixes are sorted bucket
// by bucket (first two symbols; bounded scratch whatever the input size), each bucket in rounds that compare 21 symbols at a
// time: a round builds one 63-bit key per unresolved suffix (3 bits per symbol, zero after the read's sentinel), sorts by
// (group, key) with two stable radix sorts, writes the suffixes that are alone in their group to the suffix array and keeps
// the rest.  A suffix whose window held its sentinel is compared by read index in the next round, which is the order of the
// sentinels.  Reads with 13 % error share long substrings only with the reads of the same locus: after 21 symbols most
// suffixes still have company, after 42 a quarter, after 84 almost none, so three to five rounds of shrinking size sort a
// bucket.  The BWT is then one gather, the run-length units one flag/scan/scatter.
//
// oracle/index_model.py is the same algorithm in numpy (test infrastructure), checked against the reference's files on the CPU.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <string>
#include <vector>
#include "../../include/pbsc.h"
#include "pbsc_internal.h"

namespace pbsc { namespace build {

constexpr int SYM_PER_KEY = 21;
constexpr int N_BUCKETS = 40;   // (first symbol << 3) | second symbol, the second reading 0 after a sentinel

struct Buf
{
    void* p = nullptr; size_t bytes = 0;
    ~Buf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t b) { if (p) cudaFree(p); p = nullptr; bytes = b ? b : 1; return cudaMalloc(&p, bytes); }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return (T*)p; }
};

// ---- the text ---------------------------------------------------------------------------------------------------------------
// one block per read (grid-stride): text[offsets[r] + r + x] = base of the read (the reversed read for the .rbwt) + 1, a zero
// after it; dollar[r] = position of that zero.  bad[0] counts letters that are not ACGT.
__global__ void __launch_bounds__(256)
text_kernel(const char* __restrict__ reads, const uint64_t* __restrict__ offsets, uint64_t n_reads, int reverse, uint8_t* text, uint32_t* dollar,
            unsigned int* bad)
{
    for (uint64_t r = blockIdx.x; r < n_reads; r += gridDim.x)
    {
        const uint64_t a = offsets[r], e = offsets[r + 1], len = e - a;
        uint8_t* out = text + a + r;
        unsigned int nbad = 0;
        for (uint64_t x = threadIdx.x; x < len; x += blockDim.x)
        {
            const char ch = reads[reverse ? e - 1 - x : a + x];
            int c;
            switch (ch) { case 'A': case 'a': c = 1; break; case 'C': case 'c': c = 2; break; case 'G': case 'g': c = 3; break;
                          case 'T': case 't': c = 4; break; default: c = 1; nbad++; }
            out[x] = (uint8_t)c;
        }
        if (nbad) atomicAdd(bad, nbad);
        if (threadIdx.x == 0) { out[len] = 0; dollar[r] = (uint32_t)(e + r); }
    }
}

__device__ __forceinline__ int bucket_of(const uint8_t* __restrict__ text, uint64_t p)
{
    const int s = text[p];
    return s == 0 ? 0 : ((s << 3) | text[p + 1]);   // a base is never the last symbol of the text
}

__global__ void __launch_bounds__(256)
bucket_hist_kernel(const uint8_t* __restrict__ text, uint64_t N, unsigned long long* hist)
{
    __shared__ unsigned int h[N_BUCKETS];
    if (threadIdx.x < N_BUCKETS) h[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < N; p += (uint64_t)gridDim.x * blockDim.x) atomicAdd(&h[bucket_of(text, p)], 1u);
    __syncthreads();
    if (threadIdx.x < N_BUCKETS && h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

struct InBucket
{
    const uint8_t* text; uint64_t base; int b;
    __device__ __forceinline__ bool operator()(uint32_t x) const { return bucket_of(text, base + x) == b; }
};
__global__ void add_base_kernel(uint32_t* P, uint64_t m, uint32_t base)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < m) P[i] += base;
}
__global__ void fill_kernel(uint32_t* R, uint8_t* T, uint64_t m, uint32_t row)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < m) { R[i] = row; T[i] = 0; }
}

// ---- one round --------------------------------------------------------------------------------------------------------------
// number of sentinels before position p = the read p belongs to
__device__ __forceinline__ uint32_t read_of(const uint32_t* __restrict__ dollar, uint32_t n, uint32_t p)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (dollar[mid] < p) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(256)
keys_kernel(uint64_t m, const uint32_t* __restrict__ P, const uint8_t* __restrict__ T, uint32_t off, const uint8_t* __restrict__ text, uint64_t N,
            const uint32_t* __restrict__ dollar, uint32_t n, uint64_t* K, uint32_t* I)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t p = P[i];
    uint64_t key;
    if (T[i]) key = read_of(dollar, n, p);
    else
    {
        key = 0;
        bool alive = true;
        const uint64_t q0 = (uint64_t)p + off;
        #pragma unroll 1
        for (int j = 0; j < SYM_PER_KEY; j++)
        {
            const uint64_t q = q0 + j;
            const uint64_t s = (alive && q < N) ? text[q] : 0;
            alive = alive && s != 0;
            key = (key << 3) | s;
        }
    }
    K[i] = key;
    I[i] = (uint32_t)i;
}

__global__ void gather_u32_kernel(uint64_t m, const uint32_t* __restrict__ src, const uint32_t* __restrict__ I, uint32_t* dst)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < m) dst[i] = src[I[i]];
}

// everything in the new order; oh[i] = i where an old group starts, else 0 (for the max-scan)
__global__ void __launch_bounds__(256)
reorder_kernel(uint64_t m, const uint32_t* __restrict__ I, const uint32_t* __restrict__ P, const uint32_t* __restrict__ R, const uint64_t* __restrict__ K,
               const uint8_t* __restrict__ T, uint32_t* Pn, uint32_t* Rn, uint64_t* Kn, uint8_t* Tn)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t s = I[i];
    Pn[i] = P[s]; Rn[i] = R[s]; Kn[i] = K[s]; Tn[i] = T[s];
}
__global__ void __launch_bounds__(256)
old_heads_kernel(uint64_t m, const uint32_t* __restrict__ Rn, uint32_t* oh)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    oh[i] = (i > 0 && Rn[i] != Rn[i - 1]) ? (uint32_t)i : 0u;
}
// row[i] = SA row of element i; nh[i] = i where a new group starts (old boundary or another key), else 0
__global__ void __launch_bounds__(256)
rows_kernel(uint64_t m, const uint32_t* __restrict__ Rn, const uint64_t* __restrict__ Kn, const uint32_t* __restrict__ ogs, uint32_t* row, uint32_t* nh)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    row[i] = Rn[i] + ((uint32_t)i - ogs[i]);
    nh[i] = (i > 0 && (Rn[i] != Rn[i - 1] || Kn[i] != Kn[i - 1])) ? (uint32_t)i : 0u;
}
// a suffix alone in its group is sorted: write it; the others stay for the next round (keep = 1)
__global__ void __launch_bounds__(256)
resolve_kernel(uint64_t m, const uint32_t* __restrict__ Pn, const uint32_t* __restrict__ ngs, const uint32_t* __restrict__ row, uint32_t* sa, uint32_t* keep,
               unsigned int* dup)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const bool head = ngs[i] == (uint32_t)i;
    const bool next_head = (i + 1 == m) || ngs[i + 1] == (uint32_t)(i + 1);
    const bool single = head && next_head;
    if (single) sa[row[i]] = Pn[i];
    keep[i] = single ? 0u : 1u;
    (void)dup;
}
__global__ void __launch_bounds__(256)
compact_kernel(uint64_t m, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ Pn, const uint32_t* __restrict__ ngs,
               const uint32_t* __restrict__ row, const uint64_t* __restrict__ Kn, const uint8_t* __restrict__ Tn, uint32_t* P, uint32_t* R, uint8_t* T,
               unsigned int* dup)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= m || !keep[i]) return;
    const uint32_t o = pos[i];
    P[o] = Pn[i];
    R[o] = row[ngs[i]];
    if (Tn[i]) atomicAdd(dup, 1u);   // compared by read index and still not alone: cannot happen
    T[o] = (uint8_t)(Tn[i] | ((Kn[i] & 7ull) == 0ull));
}

// ---- BWT, run-length units, lexicographic read order -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bwt_kernel(uint64_t N, const uint32_t* __restrict__ sa, const uint8_t* __restrict__ text, uint8_t* bwt)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t p = sa[i];
    bwt[i] = p == 0 ? (uint8_t)0 : text[p - 1];
}
struct RunHead
{
    const uint8_t* bwt; uint64_t base;
    __device__ __forceinline__ bool operator()(uint32_t x) const { const uint64_t i = base + x; return i == 0 || bwt[i] != bwt[i - 1]; }
};
struct IsDollar
{
    const uint8_t* bwt; uint64_t base;
    __device__ __forceinline__ bool operator()(uint32_t x) const { return bwt[base + x] == 0; }
};
// (n_runs runs of a chunk; `end` = N when the chunk holds the last run of the BWT, 0 when starts[n_runs] is the next chunk's first run)
__global__ void __launch_bounds__(256)
run_units_kernel(uint64_t n_runs, const uint32_t* __restrict__ starts, uint64_t end, uint32_t* units)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const uint64_t len = ((r + 1 < n_runs || end == 0) ? (uint64_t)starts[r + 1] : end) - starts[r];
    units[r] = (uint32_t)((len + 30) / 31);
}
__global__ void __launch_bounds__(256)
run_bytes_kernel(uint64_t n_runs, const uint32_t* __restrict__ starts, const uint32_t* __restrict__ first, uint64_t end, const uint8_t* __restrict__ bwt, uint8_t* out)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    uint64_t len = ((r + 1 < n_runs || end == 0) ? (uint64_t)starts[r + 1] : end) - starts[r];
    const uint8_t sym = (uint8_t)(bwt[starts[r]] << 5);
    uint8_t* o = out + first[r];
    while (len > 31) { *o++ = (uint8_t)(sym | 31u); len -= 31; }
    *o = (uint8_t)(sym | (uint8_t)len);
}
__global__ void __launch_bounds__(256)
lex_kernel(uint64_t n, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ sa, const uint32_t* __restrict__ dollar, uint32_t* lex)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) lex[i] = read_of(dollar, (uint32_t)n, sa[rows[i]]);
}

// positions x in [0, N) with pred(x), in order, appended to out (32-bit positions; the counting iterator restarts every 2^30)
template <class Pred>
static cudaError_t select_positions(uint64_t N, Pred pred, uint32_t* out, uint64_t cap, uint64_t* n_out, Buf& tmp, unsigned long long* d_count, cudaStream_t st)
{
    uint64_t got = 0;
    const uint64_t CH = 1ull << 30;
    for (uint64_t base = 0; base < N; base += CH)
    {
        const int cnt = (int)std::min<uint64_t>(CH, N - base);
        pred.base = base;
        thrust::counting_iterator<uint32_t> it(0);
        size_t tb = 0;
        cudaError_t e = cub::DeviceSelect::If(nullptr, tb, it, out + got, d_count, cnt, pred, st);
        if (e != cudaSuccess) return e;
        if (tmp.bytes < tb) { e = tmp.alloc(tb); if (e != cudaSuccess) return e; }
        e = cub::DeviceSelect::If(tmp.p, tb, it, out + got, d_count, cnt, pred, st);
        if (e != cudaSuccess) return e;
        unsigned long long c = 0;
        e = cudaMemcpyAsync(&c, d_count, 8, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return e;
        if (base)
        {
            if (c) add_base_kernel<<<(unsigned)((c + 255) / 256), 256, 0, st>>>(out + got, c, (uint32_t)base);
        }
        got += c;
        if (got > cap) return cudaErrorInvalidValue;
    }
    *n_out = got;
    return cudaSuccess;
}

struct Result
{
    uint8_t* runs = nullptr;       // RLUnit bytes (malloc: handed to the caller of pbsc_build_bwt as it is)
    uint64_t n_runs = 0;
    std::vector<uint32_t> lex;     // read index of the r-th '$' of the BWT
    uint64_t n_symbols = 0;
    int rounds_max = 0;
    double ms_text = 0, ms_sort = 0, ms_bwt = 0;
    ~Result() { free(runs); }
    uint8_t* take_runs() { uint8_t* r = runs; runs = nullptr; return r; }
};

// one device allocation carved into the buffers of a phase (cudaMalloc / cudaFree of twenty buffers per strand cost more than the sort)
struct Arena
{
    Buf mem; size_t used = 0;
    template <class T> T* take(size_t count)
    {
        const size_t b = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
        if (used + b > mem.bytes) return nullptr;
        T* p = (T*)((char*)mem.p + used);
        used += b;
        return p;
    }
    void reset() { used = 0; }
};

#define BUILD_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return pbsc::cuda_fail(_e, #call, __FILE__, __LINE__); } while (0)
static inline unsigned grid_for(uint64_t m) { return (unsigned)((m + 255) / 256); }

static int build_strand(const char* d_reads, const uint64_t* d_offsets, const uint64_t* h_offsets, uint64_t n, int reverse, Result& out)
{
    const bool trace = getenv("PBSC_TRACE") != nullptr;
    const uint64_t total = h_offsets[n];
    const uint64_t N = total + n;
    if (N >= 0xffffffffull) { set_error("pbsc_build: %llu symbols; this build supports < 2^32-1", (unsigned long long)N); return PBSC_ERR_LIMIT; }
    if (n == 0) { set_error("pbsc_build: no reads"); return PBSC_ERR_ARG; }
    out.n_symbols = N;
    cudaStream_t st = 0;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&]() { cudaStreamSynchronize(st); const auto t = std::chrono::steady_clock::now(); const double ms = std::chrono::duration<double, std::milli>(t - t0).count(); t0 = t; return ms; };
    Buf text, dollar, misc, sa;
    BUILD_CUDA(text.alloc(N + 32));
    BUILD_CUDA(dollar.alloc(n * 4));
    BUILD_CUDA(misc.alloc(64 * 8));
    BUILD_CUDA(sa.alloc(N * 4));
    BUILD_CUDA(cudaMemsetAsync(misc.p, 0, 64 * 8, st));
    BUILD_CUDA(cudaMemsetAsync(text.as<uint8_t>() + N, 0, 32, st));
    unsigned long long* d_hist = misc.as<unsigned long long>();            // [0, 40)
    unsigned long long* d_count = misc.as<unsigned long long>() + 48;       // select counter
    unsigned int* d_flags = (unsigned int*)(misc.as<unsigned long long>() + 56);   // [0] letters that are not ACGT, [1] unresolved read-index ties
    text_kernel<<<(unsigned)std::min<uint64_t>(n, 148 * 16), 256, 0, st>>>(d_reads, d_offsets, n, reverse, text.as<uint8_t>(), dollar.as<uint32_t>(), d_flags);
    bucket_hist_kernel<<<148 * 8, 256, 0, st>>>(text.as<uint8_t>(), N, d_hist);
    unsigned long long hist[N_BUCKETS];
    unsigned int flags[2];
    BUILD_CUDA(cudaMemcpyAsync(hist, d_hist, sizeof hist, cudaMemcpyDeviceToHost, st));
    BUILD_CUDA(cudaMemcpyAsync(flags, d_flags, sizeof flags, cudaMemcpyDeviceToHost, st));
    BUILD_CUDA(cudaStreamSynchronize(st));
    if (flags[0]) { set_error("pbsc_build: %u letters other than ACGT in the reads (the reference's index holds $ACGT only)", flags[0]); return PBSC_ERR_ARG; }
    out.ms_text = lap();
    uint64_t M = 0;
    for (int b = 0; b < N_BUCKETS; b++) M = std::max<uint64_t>(M, hist[b]);
    if (M >= 0x7fffffffull) { set_error("pbsc_build: a bucket of %llu suffixes is outside this build's range", (unsigned long long)M); return PBSC_ERR_LIMIT; }
    // scratch: one arena, sized for the larger of the two phases (a bucket's rounds; the BWT and its run-length units)
    size_t sort_tmp = 0;
    {
        size_t a = 0, b2 = 0, c = 0, d = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, a, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)M, 0, 63, st);
        cub::DeviceRadixSort::SortPairs(nullptr, b2, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)M, 0, 32, st);
        cub::DeviceScan::InclusiveScan(nullptr, c, (uint32_t*)nullptr, (uint32_t*)nullptr, cub::Max(), (int)M, st);
        cub::DeviceScan::ExclusiveSum(nullptr, d, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)M, st);
        sort_tmp = std::max(std::max(a, b2), std::max(c, d)) + 256;
    }
    Arena A;
    Buf tmp;
    BUILD_CUDA(tmp.alloc(sort_tmp));
    BUILD_CUDA(A.mem.alloc(std::max<size_t>((size_t)M * 82 + 20 * 256, (size_t)N * 9 + (size_t)n * 8 + 8 * 256)));
    uint32_t *P = A.take<uint32_t>(M), *R = A.take<uint32_t>(M), *I = A.take<uint32_t>(M), *I2 = A.take<uint32_t>(M), *I3 = A.take<uint32_t>(M),
             *RK = A.take<uint32_t>(M), *RK2 = A.take<uint32_t>(M), *Pn = A.take<uint32_t>(M), *Rn = A.take<uint32_t>(M), *s1 = A.take<uint32_t>(M),
             *s2 = A.take<uint32_t>(M), *s3 = A.take<uint32_t>(M), *s4 = A.take<uint32_t>(M);
    uint64_t *K = A.take<uint64_t>(M), *K2 = A.take<uint64_t>(M), *Kn = A.take<uint64_t>(M);
    uint8_t *T = A.take<uint8_t>(M), *Tn = A.take<uint8_t>(M);
    if (!Tn) { set_error("pbsc_build: scratch arena too small"); return PBSC_ERR_INTERNAL; }
    uint64_t row0 = 0;
    for (int b = 0; b < N_BUCKETS; b++)
    {
        uint64_t m = hist[b];
        if (m == 0) continue;
        uint64_t got = 0;
        InBucket pred{text.as<uint8_t>(), 0, b};
        BUILD_CUDA(select_positions(N, pred, P, M, &got, tmp, d_count, st));
        if (got != m) { set_error("pbsc_build: bucket %d holds %llu suffixes, expected %llu", b, (unsigned long long)got, (unsigned long long)m); return PBSC_ERR_INTERNAL; }
        fill_kernel<<<grid_for(m), 256, 0, st>>>(R, T, m, (uint32_t)row0);
        row0 += m;
        uint32_t off = 0;
        int round = 0;
        while (m)
        {
            keys_kernel<<<grid_for(m), 256, 0, st>>>(m, P, T, off, text.as<uint8_t>(), N, dollar.as<uint32_t>(), (uint32_t)n,
                                                     K, I);
            size_t tb = tmp.bytes;
            BUILD_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, K, K2, I, I2, (int)m, 0, 63, st));
            const uint32_t* order = I2;
            if (round > 0)
            {
                gather_u32_kernel<<<grid_for(m), 256, 0, st>>>(m, R, I2, RK);
                tb = tmp.bytes;
                BUILD_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, RK, RK2, I2, I3, (int)m, 0, 32, st));
                order = I3;
            }
            reorder_kernel<<<grid_for(m), 256, 0, st>>>(m, order, P, R, K, T, Pn,
                                                        Rn, Kn, Tn);
            old_heads_kernel<<<grid_for(m), 256, 0, st>>>(m, Rn, s1);
            tb = tmp.bytes;
            BUILD_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tb, s1, s2, cub::Max(), (int)m, st));      // s2 = start of the old group
            rows_kernel<<<grid_for(m), 256, 0, st>>>(m, Rn, Kn, s2, s3, s1);   // s3 = row, s1 = new heads
            tb = tmp.bytes;
            BUILD_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tb, s1, s2, cub::Max(), (int)m, st));      // s2 = start of the new group
            resolve_kernel<<<grid_for(m), 256, 0, st>>>(m, Pn, s2, s3, sa.as<uint32_t>(), s1, d_flags + 1);   // s1 = keep
            tb = tmp.bytes;
            BUILD_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, s1, s4, (int)m, st));                    // s4 = slot in the next round
            compact_kernel<<<grid_for(m), 256, 0, st>>>(m, s1, s4, Pn, s2, s3,
                                                        Kn, Tn, P, R, T, d_flags + 1);
            uint32_t last[2] = {0, 0};
            BUILD_CUDA(cudaMemcpyAsync(&last[0], s1 + (m - 1), 4, cudaMemcpyDeviceToHost, st));
            BUILD_CUDA(cudaMemcpyAsync(&last[1], s4 + (m - 1), 4, cudaMemcpyDeviceToHost, st));
            BUILD_CUDA(cudaStreamSynchronize(st));
            const uint64_t left = (uint64_t)last[0] + last[1];
            if (trace) fprintf(stderr, "[pbsc build] %s bucket %2d round %d: %llu suffixes, %llu left\n", reverse ? "rbwt" : "bwt", b, round, (unsigned long long)m, (unsigned long long)left);
            m = left;
            off += SYM_PER_KEY;
            round++;
            if (round > 100000) { set_error("pbsc_build: suffix sorting does not converge"); return PBSC_ERR_INTERNAL; }
        }
        out.rounds_max = std::max(out.rounds_max, round);
    }
    BUILD_CUDA(cudaMemcpyAsync(flags, d_flags, sizeof flags, cudaMemcpyDeviceToHost, st));
    BUILD_CUDA(cudaStreamSynchronize(st));
    if (flags[1]) { set_error("pbsc_build: %u suffixes tied after the comparison by read index", flags[1]); return PBSC_ERR_INTERNAL; }
    if (row0 != N) { set_error("pbsc_build: buckets hold %llu of %llu suffixes", (unsigned long long)row0, (unsigned long long)N); return PBSC_ERR_INTERNAL; }
    out.ms_sort = lap();
    // ---- BWT, lexicographic read order (the read of the r-th '$' row), run-length units ----
    double ms_part[6] = {0, 0, 0, 0, 0, 0};
    A.reset();
    uint8_t* bwt = A.take<uint8_t>(N);
    uint32_t* rows = A.take<uint32_t>(n);
    uint32_t* lex = A.take<uint32_t>(n);
    bwt_kernel<<<grid_for(N), 256, 0, st>>>(N, sa.as<uint32_t>(), text.as<uint8_t>(), bwt);
    uint64_t n_dollar = 0;
    IsDollar isd{bwt, 0};
    BUILD_CUDA(select_positions(N, isd, rows, n, &n_dollar, tmp, d_count, st));
    if (n_dollar != n) { set_error("pbsc_build: %llu '$' in the BWT of %llu reads", (unsigned long long)n_dollar, (unsigned long long)n); return PBSC_ERR_INTERNAL; }
    lex_kernel<<<grid_for(n), 256, 0, st>>>(n, rows, sa.as<uint32_t>(), dollar.as<uint32_t>(), lex);
    out.lex.resize(n);
    BUILD_CUDA(cudaMemcpyAsync(out.lex.data(), lex, n * 4, cudaMemcpyDeviceToHost, st));
    if (trace) ms_part[0] = lap();
    // the suffix array and the text are not needed any more: their memory holds the run starts and the unit bytes
    uint32_t* starts = sa.as<uint32_t>();
    uint8_t* bytes = text.as<uint8_t>();
    uint64_t n_runs = 0;
    RunHead rh{bwt, 0};
    BUILD_CUDA(select_positions(N, rh, starts, N, &n_runs, tmp, d_count, st));
    if (trace) ms_part[1] = lap();
    // unit counts, their prefix sums and the unit bytes, 2^28 runs at a time (a 4 G-symbol BWT has 2.6 G runs)
    const char* chunk_env = getenv("PBSC_RUN_CHUNK");   // tests: small chunks
    const uint64_t CHR = std::min<uint64_t>(n_runs, chunk_env && atoll(chunk_env) > 0 ? (uint64_t)atoll(chunk_env) : 1ull << 28);
    uint32_t* units = A.take<uint32_t>(CHR + 1);
    uint32_t* first = A.take<uint32_t>(CHR + 1);
    if (!first) { set_error("pbsc_build: scratch arena too small"); return PBSC_ERR_INTERNAL; }
    {
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, units, first, (int)CHR, st);
        if (tmp.bytes < tb) BUILD_CUDA(tmp.alloc(tb));
    }
    uint64_t n_units = 0;
    for (uint64_t r0 = 0; r0 < n_runs; r0 += CHR)
    {
        const uint64_t m = std::min<uint64_t>(CHR, n_runs - r0);
        run_units_kernel<<<grid_for(m), 256, 0, st>>>(m, starts + r0, r0 + m < n_runs ? (uint64_t)0 : N, units);
        size_t tb = tmp.bytes;
        BUILD_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, units, first, (int)m, st));
        uint32_t lastu[2] = {0, 0};
        BUILD_CUDA(cudaMemcpyAsync(&lastu[0], units + (m - 1), 4, cudaMemcpyDeviceToHost, st));
        BUILD_CUDA(cudaMemcpyAsync(&lastu[1], first + (m - 1), 4, cudaMemcpyDeviceToHost, st));
        BUILD_CUDA(cudaStreamSynchronize(st));
        const uint64_t chunk_units = (uint64_t)lastu[0] + lastu[1];
        if (n_units + chunk_units > N) { set_error("pbsc_build: more run-length units than symbols (%llu)", (unsigned long long)N); return PBSC_ERR_INTERNAL; }
        run_bytes_kernel<<<grid_for(m), 256, 0, st>>>(m, starts + r0, first, r0 + m < n_runs ? (uint64_t)0 : N, bwt, bytes + n_units);
        n_units += chunk_units;
    }
    if (trace) ms_part[2] = lap();
    out.runs = (uint8_t*)malloc(n_units ? n_units : 1);
    if (!out.runs) { set_error("pbsc_build: out of host memory"); return PBSC_ERR_LIMIT; }
    out.n_runs = n_units;
    BUILD_CUDA(cudaMemcpyAsync(out.runs, bytes, n_units, cudaMemcpyDeviceToHost, st));
    BUILD_CUDA(cudaStreamSynchronize(st));
    BUILD_CUDA(cudaGetLastError());
    if (trace) { ms_part[3] = lap(); fprintf(stderr, "[pbsc build]   bwt + read order %.1f ms, run starts %.1f ms, units %.1f ms, copy to host %.1f ms\n", ms_part[0], ms_part[1], ms_part[2], ms_part[3]); }
    out.ms_bwt = ms_part[0] + ms_part[1] + ms_part[2] + ms_part[3];
    if (!trace) out.ms_bwt = lap();
    if (trace) fprintf(stderr, "[pbsc build] %s: %llu symbols, %llu units; text %.0f ms, suffix sort %.0f ms (deepest bucket %d rounds), bwt + units %.0f ms\n",
                       reverse ? "rbwt" : "bwt", (unsigned long long)N, (unsigned long long)n_units, out.ms_text, out.ms_sort, out.rounds_max, out.ms_bwt);
    return PBSC_OK;
}

// BWTWriterBinary::writeHeader (BWTWriterBinary.cpp:28-58): magic, strings, symbols, runs, flag (BWF_NOFMI = 0)
static int write_bwt_file(const std::string& path, const Result& r, uint64_t n_reads)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { set_error("pbsc_build: cannot write %s", path.c_str()); return PBSC_ERR_IO; }
    const uint16_t magic = 0xCACA;
    const uint64_t ns = n_reads, nsym = r.n_symbols, nr = r.n_runs;
    const uint32_t flag = 0;
    bool ok = fwrite(&magic, 2, 1, f) == 1 && fwrite(&ns, 8, 1, f) == 1 && fwrite(&nsym, 8, 1, f) == 1 && fwrite(&nr, 8, 1, f) == 1 && fwrite(&flag, 4, 1, f) == 1;
    ok = ok && (r.n_runs == 0 || fwrite(r.runs, 1, r.n_runs, f) == r.n_runs);
    ok = (fclose(f) == 0) && ok;
    if (!ok) { set_error("pbsc_build: short write to %s", path.c_str()); return PBSC_ERR_IO; }
    return PBSC_OK;
}
// SAWriter::writeHeader / writeElem (SuffixTools/SAWriter.cpp): text, "<read> 0" per line
static int write_sai_file(const std::string& path, const Result& r)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { set_error("pbsc_build: cannot write %s", path.c_str()); return PBSC_ERR_IO; }
    std::string buf;
    buf.reserve(r.lex.size() * 10 + 64);
    char tmp[64];
    snprintf(tmp, sizeof tmp, "51914\n%zu\n%zu\n", r.lex.size(), r.lex.size());
    buf += tmp;
    for (uint32_t v : r.lex) { snprintf(tmp, sizeof tmp, "%u 0\n", v); buf += tmp; }
    bool ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) { set_error("pbsc_build: short write to %s", path.c_str()); return PBSC_ERR_IO; }
    return PBSC_OK;
}

template <class F>
static int guarded(const char* who, F f)
{
    try { return f(); }
    catch (const std::bad_alloc&) { set_error("%s: out of host memory", who); return PBSC_ERR_LIMIT; }
    catch (const std::exception& e) { set_error("%s: %s", who, e.what()); return PBSC_ERR_INTERNAL; }
    catch (...) { set_error("%s: unknown exception", who); return PBSC_ERR_INTERNAL; }
}

struct Uploaded
{
    Buf reads, offsets;
    int upload(const char* reads_h, const uint64_t* offsets_h, uint64_t n)
    {
        BUILD_CUDA(reads.alloc(offsets_h[n]));
        BUILD_CUDA(offsets.alloc((n + 1) * 8));
        BUILD_CUDA(cudaMemcpy(reads.p, reads_h, offsets_h[n], cudaMemcpyHostToDevice));
        BUILD_CUDA(cudaMemcpy(offsets.p, offsets_h, (n + 1) * 8, cudaMemcpyHostToDevice));
        return PBSC_OK;
    }
};

}}  // namespace pbsc::build

extern "C" {

int pbsc_build_bwt(const char* reads, const uint64_t* offsets, uint64_t n_reads, int reverse, int device, uint8_t** runs, uint64_t* n_runs, uint64_t* n_symbols,
                   uint32_t** lex_order)
{
    using namespace pbsc;
    if (!reads || !offsets || !runs || !n_runs || !n_symbols) { set_error("pbsc_build_bwt: null argument"); return PBSC_ERR_ARG; }
    *runs = nullptr; *n_runs = 0; *n_symbols = 0;
    if (lex_order) *lex_order = nullptr;
    return build::guarded("pbsc_build_bwt", [&]() -> int {
        PBSC_CUDA(cudaSetDevice(device));
        build::Uploaded up;
        int rc = up.upload(reads, offsets, n_reads);
        if (rc != PBSC_OK) return rc;
        build::Result r;
        rc = build::build_strand(up.reads.as<char>(), up.offsets.as<uint64_t>(), offsets, n_reads, reverse, r);
        if (rc != PBSC_OK) return rc;
        uint32_t* l = lex_order ? (uint32_t*)malloc((r.lex.size() ? r.lex.size() : 1) * 4) : nullptr;
        if (lex_order && !l) { set_error("pbsc_build_bwt: out of host memory"); return PBSC_ERR_LIMIT; }
        if (l) memcpy(l, r.lex.data(), r.lex.size() * 4);
        *n_runs = r.n_runs; *n_symbols = r.n_symbols;
        *runs = r.take_runs();
        if (lex_order) *lex_order = l;
        return PBSC_OK;
    });
}

void pbsc_free(void* p) { free(p); }

int pbsc_build_index_files(const char* reads, const uint64_t* offsets, uint64_t n_reads, const char* prefix, int device, int flags)
{
    using namespace pbsc;
    if (!reads || !offsets || !prefix) { set_error("pbsc_build_index_files: null argument"); return PBSC_ERR_ARG; }
    return build::guarded("pbsc_build_index_files", [&]() -> int {
        PBSC_CUDA(cudaSetDevice(device));
        build::Uploaded up;
        int rc = up.upload(reads, offsets, n_reads);
        if (rc != PBSC_OK) return rc;
        const std::string pre(prefix);
        for (int rev = 0; rev < 2; rev++)
        {
            if (flags & (rev ? PBSC_BUILD_NO_REVERSE : PBSC_BUILD_NO_FORWARD)) continue;
            build::Result r;
            rc = build::build_strand(up.reads.as<char>(), up.offsets.as<uint64_t>(), offsets, n_reads, rev, r);
            if (rc != PBSC_OK) return rc;
            rc = build::write_bwt_file(pre + (rev ? ".rbwt" : ".bwt"), r, n_reads);
            if (rc != PBSC_OK) return rc;
            rc = build::write_sai_file(pre + (rev ? ".rsai" : ".sai"), r);
            if (rc != PBSC_OK) return rc;
        }
        return PBSC_OK;
    });
}

}  // extern "C"
