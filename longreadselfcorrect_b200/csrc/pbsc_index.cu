// pbsc_index.cu — index construction (flat occurrence table, '$' list, short-prefix table),
// batched backward search, and the parameter derivation of PacBioSelfCorrectionMain.
#include <math.h>
#include <string.h>
#include <sys/stat.h>
#include <algorithm>
#include <cub/cub.cuh>
#include <mutex>
#include <map>
#include <fstream>
#include <sstream>
#include "pbsc_internal.h"

namespace pbsc {

static thread_local std::string g_err;
void set_error(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e), file, line, what);
    return PBSC_ERR_CUDA;
}
// ---- per-device cache of device blocks (see DevBuf) ----
namespace {
struct DevCache
{
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks[16];      // per device: size -> block
    std::map<void*, std::pair<int, size_t>> live;      // block -> (device, size)
};
DevCache& dev_cache() { static DevCache c; return c; }
size_t cache_round(size_t bytes)
{
    // sizes are rounded up to 1/8 of their power of two (at least 1 MB granularity above 8 MB) so that batches of similar
    // size reuse each other's blocks
    if (bytes < 4096) return 4096;
    size_t p2 = 1; while (p2 * 2 <= bytes) p2 *= 2;
    const size_t step = p2 / 8;
    return (bytes + step - 1) / step * step;
}
}  // namespace

cudaError_t dev_cache_alloc(void** p, size_t bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t want = cache_round(bytes);
    DevCache& c = dev_cache();
    {
        std::lock_guard<std::mutex> g(c.mu);
        auto& fb = c.free_blocks[dev & 15];
        auto it = fb.lower_bound(want);
        if (it != fb.end() && it->first <= want + want / 4)
        {
            *p = it->second;
            c.live[*p] = std::make_pair(dev, it->first);
            fb.erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess)
    {
        // give the cache back and retry once
        cudaGetLastError();
        dev_cache_trim(dev);
        e = cudaMalloc(p, want);
        if (e != cudaSuccess) return e;
    }
    std::lock_guard<std::mutex> g(c.mu);
    c.live[*p] = std::make_pair(dev, want);
    return cudaSuccess;
}

void dev_cache_free(void* p)
{
    if (!p) return;
    DevCache& c = dev_cache();
    std::lock_guard<std::mutex> g(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) { cudaFree(p); return; }
    c.free_blocks[it->second.first & 15].insert(std::make_pair(it->second.second, p));
    c.live.erase(it);
}

void dev_cache_trim(int device)
{
    DevCache& c = dev_cache();
    std::vector<void*> blocks;
    {
        std::lock_guard<std::mutex> g(c.mu);
        for (auto& kv : c.free_blocks[device & 15]) blocks.push_back(kv.second);
        c.free_blocks[device & 15].clear();
    }
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    for (void* b : blocks) cudaFree(b);
    cudaSetDevice(cur);
}

unsigned long long* occ_counts()
{
    static unsigned long long c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    return c;
}

Timing& last_timing()
{
    static thread_local Timing t;
    return t;
}

// ------------------------------------------------------------------------------------------
// host decode of the on-disk run bytes into FmBlocks (BWTReaderBinary.cpp:79-85, RLUnit.h:118-143)
// ------------------------------------------------------------------------------------------
static int decode_runs(const uint8_t* runs, uint64_t n_runs, uint64_t n_symbols, std::vector<FmBlock>& blocks,
                       std::vector<uint32_t>& dollars, uint64_t total[5])
{
    if (n_symbols >= 0xffffffffull) { set_error("BWT has %llu symbols; this build supports < 2^32-1", (unsigned long long)n_symbols); return PBSC_ERR_LIMIT; }
    const uint64_t nb = n_symbols / 64 + 1;
    blocks.assign(nb, FmBlock{{0, 0, 0, 0}, {0, 0, 0, 0}});
    dollars.clear();
    uint64_t cnt[5] = {0, 0, 0, 0, 0};   // $ A C G T
    uint64_t pos = 0;
    for (uint64_t r = 0; r < n_runs; r++)
    {
        const uint8_t u = runs[r];
        const uint32_t sym = u >> 5, len = u & 0x1f;
        if (sym > 4 || len == 0) { set_error("malformed run byte 0x%02x at run %llu", u, (unsigned long long)r); return PBSC_ERR_FORMAT; }
        if (pos + len > n_symbols) { set_error("runs hold more symbols than the header's %llu", (unsigned long long)n_symbols); return PBSC_ERR_FORMAT; }
        for (uint32_t i = 0; i < len; i++, pos++)
        {
            const uint64_t b = pos >> 6;
            const uint32_t j = (uint32_t)pos & 63u;
            if (j == 0) { FmBlock& h = blocks[b]; h.cnt[0] = (uint32_t)cnt[1]; h.cnt[1] = (uint32_t)cnt[2]; h.cnt[2] = (uint32_t)cnt[3]; h.cnt[3] = (uint32_t)cnt[4]; }
            if (sym == 0) { dollars.push_back((uint32_t)pos); blocks[b].cnt[0] |= 0x80000000u; }
            else blocks[b].bases[j >> 4] |= (sym - 1) << (2 * (j & 15));
            cnt[sym]++;
        }
    }
    if (pos != n_symbols) { set_error("runs hold %llu symbols, header says %llu", (unsigned long long)pos, (unsigned long long)n_symbols); return PBSC_ERR_FORMAT; }
    if ((n_symbols & 63) == 0) { FmBlock& h = blocks[nb - 1]; h.cnt[0] = (uint32_t)cnt[1]; h.cnt[1] = (uint32_t)cnt[2]; h.cnt[2] = (uint32_t)cnt[3]; h.cnt[3] = (uint32_t)cnt[4]; }
    if (cnt[1] >= 0x80000000ull) { set_error("BWT has %llu 'A' symbols; this build supports < 2^31", (unsigned long long)cnt[1]); return PBSC_ERR_LIMIT; }
    for (int c = 0; c < 5; c++) total[c] = cnt[c];
    return PBSC_OK;
}

static void fill_table(FmTable& t, const FmBlock* d_blocks, const uint32_t* d_dollar, const uint64_t* d_dmask, uint64_t n, const uint64_t total[5])
{
    t.blocks = d_blocks;
    t.dollar_pos = d_dollar;
    t.dollar_mask = d_dmask;
    t.n = n;
    t.n_dollar = (uint32_t)total[0];
    t.C[0] = total[0];
    t.C[1] = t.C[0] + total[1];
    t.C[2] = t.C[1] + total[2];
    t.C[3] = t.C[2] + total[3];
    for (int c = 0; c < 4; c++) t.total[c] = total[c + 1];
}

// ------------------------------------------------------------------------------------------
// synthetic index (microbenchmark config 5): i.i.d. symbols generated on the device
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

__global__ void synth_dollar_kernel(uint32_t* pos, uint64_t n_strings, uint64_t n_symbols, uint64_t seed)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n_strings) pos[i] = (uint32_t)(splitmix64(seed ^ (0xD011A5ull + i * 0x100000001b3ull)) % n_symbols);
}

// one thread per block: 64 symbols from two 64-bit hashes; '$' positions override; per-block counts out
__global__ void synth_blocks_kernel(FmBlock* blocks, uint32_t* cntA, uint32_t* cntC, uint32_t* cntG, uint32_t* cntT,
                                    uint64_t n_blocks, uint64_t n_symbols, const uint32_t* dollar, uint32_t n_dollar, uint64_t seed)
{
    uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    uint64_t w0 = splitmix64(seed + 2 * b), w1 = splitmix64(seed + 2 * b + 1);
    const uint64_t start = b << 6;
    uint32_t valid = (start >= n_symbols) ? 0u : (uint32_t)((n_symbols - start) < 64 ? (n_symbols - start) : 64);
    if (valid < 32) { w0 &= valid ? ((1ull << (2 * valid)) - 1ull) : 0ull; w1 = 0; }
    else if (valid < 64) w1 &= (1ull << (2 * (valid - 32))) - 1ull;
    // '$' inside this block
    uint32_t lo = 0, hi = n_dollar;
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if ((uint64_t)dollar[m] < start) lo = m + 1; else hi = m; }
    bool has = false;
    uint32_t nd = 0;
    for (uint32_t k = lo; k < n_dollar && (uint64_t)dollar[k] < start + valid; k++)
    {
        if (k > lo && dollar[k] == dollar[k - 1]) continue;
        uint32_t j = dollar[k] - (uint32_t)start;
        if (j < 32) w0 &= ~(3ull << (2 * j)); else w1 &= ~(3ull << (2 * (j - 32)));
        has = true;
        nd++;
    }
    uint32_t c[4];
    for (int x = 0; x < 4; x++)
    {
        const uint64_t pat = 0x5555555555555555ull * (uint64_t)x;
        uint64_t x0 = w0 ^ pat, x1 = w1 ^ pat;
        uint64_t m0 = ~(x0 | (x0 >> 1)) & 0x5555555555555555ull, m1 = ~(x1 | (x1 >> 1)) & 0x5555555555555555ull;
        if (valid < 32) { m0 &= valid ? ((1ull << (2 * valid)) - 1ull) : 0ull; m1 = 0; }
        else if (valid < 64) m1 &= (1ull << (2 * (valid - 32))) - 1ull;
        c[x] = __popcll(m0) + __popcll(m1);
    }
    c[0] -= nd;
    FmBlock blk;
    blk.cnt[0] = has ? 0x80000000u : 0u; blk.cnt[1] = blk.cnt[2] = blk.cnt[3] = 0;
    blk.bases[0] = (uint32_t)w0; blk.bases[1] = (uint32_t)(w0 >> 32); blk.bases[2] = (uint32_t)w1; blk.bases[3] = (uint32_t)(w1 >> 32);
    blocks[b] = blk;
    cntA[b] = c[0]; cntC[b] = c[1]; cntG[b] = c[2]; cntT[b] = c[3];
}

__global__ void synth_headers_kernel(FmBlock* blocks, const uint32_t* cntA, const uint32_t* cntC, const uint32_t* cntG,
                                     const uint32_t* cntT, uint64_t n_blocks)
{
    uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    blocks[b].cnt[0] |= cntA[b];
    blocks[b].cnt[1] = cntC[b];
    blocks[b].cnt[2] = cntG[b];
    blocks[b].cnt[3] = cntT[b];
}

__global__ void dollar_mask_kernel(const uint32_t* __restrict__ dollar, uint32_t n, unsigned long long* __restrict__ mask)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) atomicOr(mask + (dollar[i] >> 6), 1ull << (dollar[i] & 63u));
}

__global__ void unique_count_kernel(const uint32_t* sorted, uint32_t n, uint32_t* out)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n && (i == 0 || sorted[i] != sorted[i - 1])) atomicAdd(out, 1u);
}

// ------------------------------------------------------------------------------------------
// device decode of the on-disk run bytes (BWTReaderBinary.cpp:79-85, RLUnit.h:118-143): the host only reads the file
// ------------------------------------------------------------------------------------------
// per run: its length (and the length again if it is a run of '$'); malformed bytes are counted
__global__ void run_lengths_kernel(const uint8_t* __restrict__ runs, uint64_t n_runs, uint64_t* __restrict__ len, uint64_t* __restrict__ dlen, unsigned int* bad)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const uint8_t u = runs[r];
    const uint32_t sym = u >> 5, l = u & 0x1f;
    if (sym > 4 || l == 0) atomicAdd(bad, 1u);
    len[r] = l;
    dlen[r] = sym == 0 ? l : 0;
}

// per run: its symbols into the 2-bit words of the blocks; '$' positions into the list and the per-block bit mask
// (the runs come in chunks: `start` / `dstart` are relative to the chunk, sym0 / dollar0 = symbols and '$' before it)
__global__ void run_scatter_kernel(const uint8_t* __restrict__ runs, uint64_t n_runs, const uint64_t* __restrict__ start, const uint64_t* __restrict__ dstart,
                                   uint64_t sym0, uint64_t dollar0, uint64_t n_symbols, uint64_t n_dollar, FmBlock* blocks, uint32_t* __restrict__ dollar,
                                   unsigned long long* dmask, unsigned int* bad)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const uint8_t u = runs[r];
    const uint32_t sym = u >> 5, l = u & 0x1f;
    uint64_t pos = sym0 + start[r];
    if (pos + l > n_symbols) { atomicAdd(bad + 2, 1u); return; }
    if (sym == 0)
    {
        uint64_t d = dollar0 + dstart[r];
        if (d + l > n_dollar) { atomicAdd(bad + 1, 1u); return; }
        for (uint32_t i = 0; i < l; i++, pos++, d++) { dollar[d] = (uint32_t)pos; atomicOr(dmask + (pos >> 6), 1ull << (pos & 63u)); }
        return;
    }
    if (sym == 1 || sym > 4) return;   // 'A' is code 0: nothing to set
    const uint32_t code = sym - 1;
    // at most three 32-bit words (16 symbols each) hold a run of up to 31 symbols
    uint32_t left = l;
    while (left)
    {
        const uint64_t b = pos >> 6;
        const uint32_t j = (uint32_t)pos & 63u, w = j >> 4, o = j & 15u;
        const uint32_t take = min(left, 16u - o);
        const uint32_t pat = (take == 16 ? 0xffffffffu : ((1u << (2 * take)) - 1u)) & (0x55555555u * code);
        atomicOr(&blocks[b].bases[w], pat << (2 * o));
        pos += take; left -= take;
    }
}

// per block: how many A, C, G, T it holds ('$' and the padding behind the last symbol are code 0 but not 'A')
__global__ void block_counts_kernel(const FmBlock* __restrict__ blocks, const uint64_t* __restrict__ dmask, uint64_t n_blocks, uint64_t n_symbols,
                                    uint32_t* cntA, uint32_t* cntC, uint32_t* cntG, uint32_t* cntT)
{
    const uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint64_t startp = b << 6;
    const uint32_t valid = startp >= n_symbols ? 0u : (uint32_t)((n_symbols - startp) < 64 ? (n_symbols - startp) : 64);
    const FmBlock blk = blocks[b];
    const uint64_t w0 = (uint64_t)blk.bases[0] | ((uint64_t)blk.bases[1] << 32), w1 = (uint64_t)blk.bases[2] | ((uint64_t)blk.bases[3] << 32);
    const uint64_t M = 0x5555555555555555ull;
    const uint64_t l0 = w0 & M, h0 = (w0 >> 1) & M, l1 = w1 & M, h1 = (w1 >> 1) & M;
    const uint32_t nT = __popcll(h0 & l0) + __popcll(h1 & l1), nG = __popcll(h0 & ~l0) + __popcll(h1 & ~l1), nC = __popcll(~h0 & l0) + __popcll(~h1 & l1);
    cntA[b] = valid - nT - nG - nC - (uint32_t)__popcll(dmask[b]);
    cntC[b] = nC; cntG[b] = nG; cntT[b] = nT;
}

__global__ void block_headers_kernel(FmBlock* blocks, const uint64_t* __restrict__ dmask, const uint32_t* cumA, const uint32_t* cumC, const uint32_t* cumG,
                                     const uint32_t* cumT, uint64_t n_blocks)
{
    const uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    blocks[b].cnt[0] = cumA[b] | (dmask[b] ? 0x80000000u : 0u);
    blocks[b].cnt[1] = cumC[b]; blocks[b].cnt[2] = cumG[b]; blocks[b].cnt[3] = cumT[b];
}

// ------------------------------------------------------------------------------------------
// short-prefix table: entry(w) for every k0-mer w, built level by level (one update per entry)
// ------------------------------------------------------------------------------------------
__global__ void prefix_level_kernel(FmIndexDev idx, const PrefixEntry* prev, PrefixEntry* cur, uint64_t n_cur, int level)
{
    // key convention: key = sum_j w[j] * 4^j  (w[0], the first processed symbol, in the low bits)
    uint64_t key = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (key >= n_cur) return;
    const int c = (int)(key >> (2 * (level - 1)));
    Interval f, r;
    if (level == 1) { f = init_interval(idx.t[PBSC_RBWT], c); r = init_interval(idx.t[PBSC_BWT], 3 - c); }
    else
    {
        const PrefixEntry p = prev[key & ((1ull << (2 * (level - 1))) - 1ull)];
        f.lo = p.fwd_lo; f.hi = p.fwd_lo + p.fwd_size;
        r.lo = p.rvc_lo; r.hi = p.rvc_lo + p.rvc_size;
        // findInterval stops at the first empty interval (BWTAlgorithms.cpp:25-29): empty stays empty
        if (f.valid()) f = update_interval(idx.t[PBSC_RBWT], f, c);
        if (r.valid()) r = update_interval(idx.t[PBSC_BWT], r, 3 - c);
    }
    PrefixEntry e;
    e.fwd_lo = f.lo; e.fwd_size = (uint32_t)f.size();
    e.rvc_lo = r.lo; e.rvc_size = (uint32_t)r.size();
    e.pad[0] = e.pad[1] = 0;
    cur[key] = e;
}

// ------------------------------------------------------------------------------------------
// batched backward search
// ------------------------------------------------------------------------------------------
// generic entry: byte strings, variable length (BWTAlgorithms::findInterval, BWTAlgorithms.cpp:14-31)
__global__ void findinterval_bytes_kernel(FmIndexDev idx, int which, const char* __restrict__ kmers,
                                          const uint64_t* __restrict__ offsets, uint64_t n, int64_t* __restrict__ lower,
                                          int64_t* __restrict__ upper, uint8_t* __restrict__ steps)
{
    uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    const FmTable& t = idx.t[which];
    const char* w = kmers + offsets[q];
    int len = (int)(offsets[q + 1] - offsets[q]);
    auto code = [](char b) { return b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3; };
    int j = len - 1;
    Interval iv = init_interval(t, code(w[j]));
    int st = 0;
    for (--j; j >= 0; --j)
    {
        iv = update_interval(t, iv, code(w[j]));
        st++;
        if (!iv.valid()) break;
    }
    lower[q] = (int64_t)iv.lo;
    upper[q] = (int64_t)iv.hi - 1;
    if (steps) steps[q] = (uint8_t)st;
}

// fixed-k, 2-bit packed queries resident on the device; USE_PREFIX consults the short-prefix table for the
// first k0 processed symbols (the last k0 bases of w) and issues sectors only for the remaining k-k0 steps
template <bool USE_PREFIX>
__global__ void __launch_bounds__(256)
findinterval_packed_kernel(FmIndexDev idx, int which, const uint64_t* __restrict__ kmers, int k, uint64_t n,
                           int64_t* __restrict__ lower, int64_t* __restrict__ upper, uint8_t* __restrict__ steps)
{
    uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    const FmTable& t = idx.t[which];
    const uint64_t w = kmers[q];
    int j = k - 1;
    Interval iv;
    int st = 0;
    if (USE_PREFIX && k >= idx.k0)
    {
        // processed order is w[k-1], w[k-2], ...; the table is keyed by x with x[i] = processed symbol i
        // (RBWT / fwd field) or its complement (BWT / rvc field)
        uint64_t key = 0;
        for (int i = 0; i < idx.k0; i++)
        {
            int c = (int)((w >> (2 * (k - 1 - i))) & 3);
            key |= (uint64_t)(which == PBSC_RBWT ? c : 3 - c) << (2 * i);
        }
        const uint4* ep = reinterpret_cast<const uint4*>(idx.prefix + key);
        const uint4 a = __ldg(ep), b = __ldg(ep + 1);
        if (which == PBSC_RBWT) { iv.lo = (uint64_t)a.x | ((uint64_t)a.y << 32); iv.hi = iv.lo + b.x; }
        else { iv.lo = (uint64_t)a.z | ((uint64_t)a.w << 32); iv.hi = iv.lo + b.y; }
        j = k - 1 - idx.k0;
        // steps the plain search would have executed inside the prefix are unknown once it is empty; report
        // the table-assisted count (k0-1) there, exact otherwise
        st = idx.k0 - 1;
        if (!iv.valid()) j = -1;
    }
    else
    {
        iv = init_interval(t, (int)((w >> (2 * j)) & 3));
        --j;
    }
    for (; j >= 0; --j)
    {
        iv = update_interval(t, iv, (int)((w >> (2 * j)) & 3));
        st++;
        if (!iv.valid()) break;
    }
    lower[q] = (int64_t)iv.lo;
    upper[q] = (int64_t)iv.hi - 1;
    if (steps) steps[q] = (uint8_t)st;
}

__global__ void get_symbols_kernel(FmTable t, uint64_t first, uint64_t count, char* out)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t p = first + i;
    const FmBlock& b = t.blocks[p >> 6];
    uint32_t j = (uint32_t)p & 63u;
    uint32_t c = (b.bases[j >> 4] >> (2 * (j & 15))) & 3u;
    char ch = "ACGT"[c];
    if (c == 0 && (b.cnt[0] >> 31))
    {
        uint32_t lo = 0, hi = t.n_dollar;
        while (lo < hi) { uint32_t m = (lo + hi) >> 1; if ((uint64_t)t.dollar_pos[m] < p) lo = m + 1; else hi = m; }
        if (lo < t.n_dollar && (uint64_t)t.dollar_pos[lo] == p) ch = '$';
    }
    out[i] = ch;
}

}  // namespace pbsc

using namespace pbsc;

extern "C" {

const char* pbsc_last_error(void) { return g_err.c_str(); }

int pbsc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static void apply_l2_policy(pbsc_index* idx, int device);

static int upload_strand(pbsc_index* idx, int which, const std::vector<FmBlock>& blocks, const std::vector<uint32_t>& dollars,
                         uint64_t n_symbols, uint64_t n_strings, const uint64_t total[5])
{
    PBSC_CUDA(cudaMalloc((void**)&idx->d_blocks[which], blocks.size() * sizeof(FmBlock)));
    PBSC_CUDA(cudaMemcpy(idx->d_blocks[which], blocks.data(), blocks.size() * sizeof(FmBlock), cudaMemcpyHostToDevice));
    PBSC_CUDA(cudaMalloc((void**)&idx->d_dollar[which], (dollars.size() + 1) * sizeof(uint32_t)));
    if (!dollars.empty()) PBSC_CUDA(cudaMemcpy(idx->d_dollar[which], dollars.data(), dollars.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    idx->n_symbols[which] = n_symbols;
    idx->n_strings[which] = n_strings;
    idx->n_blocks[which] = blocks.size();
    PBSC_CUDA(cudaMalloc((void**)&idx->d_dmask[which], blocks.size() * sizeof(uint64_t)));
    PBSC_CUDA(cudaMemset(idx->d_dmask[which], 0, blocks.size() * sizeof(uint64_t)));
    if (!dollars.empty()) dollar_mask_kernel<<<(unsigned)((dollars.size() + 255) / 256), 256>>>(idx->d_dollar[which], (uint32_t)dollars.size(), (unsigned long long*)idx->d_dmask[which]);
    PBSC_CUDA(cudaDeviceSynchronize());
    idx->device_bytes += blocks.size() * (sizeof(FmBlock) + sizeof(uint64_t)) + (dollars.size() + 1) * sizeof(uint32_t);
    fill_table(idx->dev.t[which], idx->d_blocks[which], idx->d_dollar[which], idx->d_dmask[which], n_symbols, total);
    return PBSC_OK;
}

// One strand from its run bytes, decoded on the device: H2D of the bytes, lengths, two scans, one scatter, per-block counts,
// four scans.  ~1 s for the 1.25 G symbols of config 3 (the host loop of decode_runs took 6 s per strand).
static int upload_strand_device(pbsc_index* idx, int which, const uint8_t* runs, uint64_t n_runs, uint64_t n_symbols, uint64_t n_strings)
{
    if (n_symbols >= 0xffffffffull) { set_error("BWT has %llu symbols; this build supports < 2^32-1", (unsigned long long)n_symbols); return PBSC_ERR_LIMIT; }
    if (n_runs == 0 || n_runs > n_symbols) { set_error("BWT with %llu runs is outside this build's range", (unsigned long long)n_runs); return PBSC_ERR_LIMIT; }
    const uint64_t nb = n_symbols / 64 + 1;
    // The runs are decoded in chunks (33 bytes of scratch per run: a 4 G-symbol BWT has 2.6 G of them).  One '$' per read: the
    // list of their positions has n_strings entries, and the decode checks that the runs hold exactly that many.
    const char* chunk_env = getenv("PBSC_RUN_CHUNK");   // tests: small chunks
    const uint64_t CH = std::min<uint64_t>(n_runs, chunk_env && atoll(chunk_env) > 0 ? (uint64_t)atoll(chunk_env) : 1ull << 28);
    const uint64_t nd = n_strings;
    DevBuf<uint8_t> d_runs, tmp; DevBuf<uint64_t> len, dlen, start, dstart; DevBuf<unsigned int> bad; DevBuf<uint32_t> cnt[4], cum[4];
    PBSC_CUDA(d_runs.alloc(CH)); PBSC_CUDA(len.alloc(CH + 1)); PBSC_CUDA(dlen.alloc(CH + 1)); PBSC_CUDA(start.alloc(CH + 1)); PBSC_CUDA(dstart.alloc(CH + 1));
    PBSC_CUDA(bad.alloc(4));   // [0] malformed bytes, [1] more '$' than strings, [2] more symbols than the header says
    PBSC_CUDA(cudaMemset(bad.p, 0, 16));
    PBSC_CUDA(cudaMalloc((void**)&idx->d_blocks[which], nb * sizeof(FmBlock)));
    PBSC_CUDA(cudaMemset(idx->d_blocks[which], 0, nb * sizeof(FmBlock)));
    PBSC_CUDA(cudaMalloc((void**)&idx->d_dollar[which], (nd + 1) * sizeof(uint32_t)));
    PBSC_CUDA(cudaMalloc((void**)&idx->d_dmask[which], nb * sizeof(uint64_t)));
    PBSC_CUDA(cudaMemset(idx->d_dmask[which], 0, nb * sizeof(uint64_t)));
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, len.p, start.p, (int)(CH + 1));
    PBSC_CUDA(tmp.alloc(tb));
    uint64_t tot[2] = {0, 0};
    for (uint64_t r0 = 0; r0 < n_runs; r0 += CH)
    {
        const uint64_t m = std::min<uint64_t>(CH, n_runs - r0);
        PBSC_CUDA(cudaMemcpy(d_runs.p, runs + r0, m, cudaMemcpyHostToDevice));
        PBSC_CUDA(cudaMemset(len.p + m, 0, 8)); PBSC_CUDA(cudaMemset(dlen.p + m, 0, 8));
        run_lengths_kernel<<<(unsigned)((m + 255) / 256), 256>>>(d_runs.p, m, len.p, dlen.p, bad.p);
        size_t tb1 = tb;
        cub::DeviceScan::ExclusiveSum(tmp.p, tb1, len.p, start.p, (int)(m + 1));
        tb1 = tb;
        cub::DeviceScan::ExclusiveSum(tmp.p, tb1, dlen.p, dstart.p, (int)(m + 1));
        uint64_t part[2] = {0, 0};
        PBSC_CUDA(cudaMemcpy(&part[0], start.p + m, 8, cudaMemcpyDeviceToHost));
        PBSC_CUDA(cudaMemcpy(&part[1], dstart.p + m, 8, cudaMemcpyDeviceToHost));
        run_scatter_kernel<<<(unsigned)((m + 255) / 256), 256>>>(d_runs.p, m, start.p, dstart.p, tot[0], tot[1], n_symbols, nd, idx->d_blocks[which],
                                                               idx->d_dollar[which], (unsigned long long*)idx->d_dmask[which], bad.p);
        tot[0] += part[0]; tot[1] += part[1];
    }
    unsigned int hbad2[4] = {0, 0, 0, 0};
    PBSC_CUDA(cudaMemcpy(hbad2, bad.p, 16, cudaMemcpyDeviceToHost));
    unsigned int hbad = hbad2[0];
    if (hbad) { set_error("malformed run bytes (%u) in the BWT", hbad); return PBSC_ERR_FORMAT; }
    if (tot[0] != n_symbols || hbad2[2]) { set_error("runs hold %llu symbols, header says %llu", (unsigned long long)tot[0], (unsigned long long)n_symbols); return PBSC_ERR_FORMAT; }
    if (tot[1] != nd || hbad2[1]) { set_error("runs hold %llu '$', header says %llu strings", (unsigned long long)tot[1], (unsigned long long)nd); return PBSC_ERR_FORMAT; }
    d_runs.release(); len.release(); dlen.release(); start.release(); dstart.release();
    for (int c = 0; c < 4; c++) { PBSC_CUDA(cnt[c].alloc(nb)); PBSC_CUDA(cum[c].alloc(nb)); }
    block_counts_kernel<<<(unsigned)((nb + 255) / 256), 256>>>(idx->d_blocks[which], idx->d_dmask[which], nb, n_symbols, cnt[0].p, cnt[1].p, cnt[2].p, cnt[3].p);
    uint64_t total[5] = {nd, 0, 0, 0, 0};
    for (int c = 0; c < 4; c++)
    {
        size_t tb2 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb2, cnt[c].p, cum[c].p, (int)nb);
        if (tb2 > tmp.n) PBSC_CUDA(tmp.alloc(tb2));
        cub::DeviceScan::ExclusiveSum(tmp.p, tb2, cnt[c].p, cum[c].p, (int)nb);
        uint32_t lastc = 0, lastv = 0;
        PBSC_CUDA(cudaMemcpy(&lastc, cum[c].p + nb - 1, 4, cudaMemcpyDeviceToHost));
        PBSC_CUDA(cudaMemcpy(&lastv, cnt[c].p + nb - 1, 4, cudaMemcpyDeviceToHost));
        total[c + 1] = (uint64_t)lastc + lastv;
    }
    if (total[1] >= 0x80000000ull) { set_error("BWT has %llu 'A' symbols; this build supports < 2^31", (unsigned long long)total[1]); return PBSC_ERR_LIMIT; }
    block_headers_kernel<<<(unsigned)((nb + 255) / 256), 256>>>(idx->d_blocks[which], idx->d_dmask[which], cum[0].p, cum[1].p, cum[2].p, cum[3].p, nb);
    PBSC_CUDA(cudaMemcpy(&hbad, bad.p, 4, cudaMemcpyDeviceToHost));
    PBSC_CUDA(cudaDeviceSynchronize());
    if (hbad) { set_error("runs hold more symbols than the header's %llu", (unsigned long long)n_symbols); return PBSC_ERR_FORMAT; }
    if (total[0] + total[1] + total[2] + total[3] + total[4] != n_symbols) { set_error("decoded symbol counts do not add up to %llu", (unsigned long long)n_symbols); return PBSC_ERR_INTERNAL; }
    idx->n_symbols[which] = n_symbols; idx->n_strings[which] = n_strings; idx->n_blocks[which] = nb;
    idx->device_bytes += nb * (sizeof(FmBlock) + sizeof(uint64_t)) + (nd + 1) * sizeof(uint32_t);
    fill_table(idx->dev.t[which], idx->d_blocks[which], idx->d_dollar[which], idx->d_dmask[which], n_symbols, total);
    return PBSC_OK;
}

int pbsc_index_create(const uint8_t* bwt_runs, uint64_t bwt_n_runs, uint64_t bwt_n_symbols, uint64_t bwt_n_strings,
                      const uint8_t* rbwt_runs, uint64_t rbwt_n_runs, uint64_t rbwt_n_symbols, uint64_t rbwt_n_strings,
                      int device, pbsc_index** out)
try
{
    if (!bwt_runs || !rbwt_runs || !out) { set_error("pbsc_index_create: null argument"); return PBSC_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    PBSC_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("pbsc_index_create: device %d not available (%d devices)", device, ndev); return PBSC_ERR_CUDA; }
    PBSC_CUDA(cudaSetDevice(device));
    pbsc_index* idx = new pbsc_index();
    idx->device = device;
    idx->dev.prefix = nullptr;
    idx->dev.k0 = 0;
    idx->dev.idmer_valid = nullptr;
    idx->dev.idmer_len = 0;
    int rc = PBSC_OK;
    const uint8_t* runs[2] = {bwt_runs, rbwt_runs};
    const uint64_t nr[2] = {bwt_n_runs, rbwt_n_runs}, ns[2] = {bwt_n_symbols, rbwt_n_symbols}, nstr[2] = {bwt_n_strings, rbwt_n_strings};
    // PBSC_HOST_DECODE=1 keeps the sequential host decoder (the first implementation; the tests compare the two)
    const bool host_decode = getenv("PBSC_HOST_DECODE") && atoi(getenv("PBSC_HOST_DECODE")) != 0;
    for (int w = 0; w < 2 && rc == PBSC_OK; w++)
    {
        if (!host_decode) { rc = upload_strand_device(idx, w, runs[w], nr[w], ns[w], nstr[w]); continue; }
        std::vector<FmBlock> blocks;
        std::vector<uint32_t> dollars;
        uint64_t total[5];
        rc = decode_runs(runs[w], nr[w], ns[w], blocks, dollars, total);
        if (rc == PBSC_OK) rc = upload_strand(idx, w, blocks, dollars, ns[w], nstr[w], total);
    }
    if (rc == PBSC_OK && cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "cudaStreamCreate", __FILE__, __LINE__); }
    if (rc == PBSC_OK) cudaDeviceGetAttribute(&idx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (rc != PBSC_OK) { pbsc_index_destroy(idx); return rc; }
    apply_l2_policy(idx, device);
    *out = idx;
    return PBSC_OK;
}
PBSC_CATCH_ALL("pbsc_index_create")

// EXPERIMENT, off unless PBSC_L2_PERSIST is set (not measured yet: prepared at the end of round 1 for round 2).
// walk_levels_kernel is bound by DRAM random access (2.1 TB/s of 32-byte sectors at a 33 % L2 hit rate) and the two rank
// tables are what its random reads hit (config 2: 2 x 119 MB against 126 MB of L2), while its per-lane scratch (~0.9 GB of
// leaf records per pass) streams through the same L2.  PBSC_L2_PERSIST=<percent> sets aside the largest persisting L2
// carve-out the device allows and puts an access-policy window on the launch stream over the span of the two block tables
// (when the span fits the device's window limit; otherwise over the first table), hit ratio = percent / 100 (default: the
// carve-out divided by the window), misses streaming.
static void apply_l2_policy(pbsc_index* idx, int device)
{
    const char* e = getenv("PBSC_L2_PERSIST");
    if (!e || !idx->stream || !idx->d_blocks[0] || !idx->d_blocks[1]) return;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
    if (max_persist <= 0 || max_window <= 0) { fprintf(stderr, "[pbsc] PBSC_L2_PERSIST: the device reports no persisting L2\n"); return; }
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) != cudaSuccess) { cudaGetLastError(); return; }
    const uintptr_t a0 = (uintptr_t)idx->d_blocks[0], a1 = (uintptr_t)idx->d_blocks[1];
    const size_t b0 = idx->n_blocks[0] * sizeof(FmBlock), b1 = idx->n_blocks[1] * sizeof(FmBlock);
    uintptr_t lo = std::min(a0, a1), hi = std::max(a0 + b0, a1 + b1);
    if (hi - lo > (uintptr_t)max_window) { lo = a0; hi = a0 + std::min(b0, (size_t)max_window); }
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = (void*)lo;
    attr.accessPolicyWindow.num_bytes = (size_t)(hi - lo);
    double ratio = atof(e) > 1.0 ? atof(e) / 100.0 : (double)max_persist / (double)(hi - lo);
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, std::max(0.0, ratio));
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    const cudaError_t rc = cudaStreamSetAttribute(idx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    if (rc != cudaSuccess) cudaGetLastError();
    fprintf(stderr, "[pbsc] PBSC_L2_PERSIST: carve-out %.1f MB, window %.1f MB (%s), hit ratio %.2f: %s\n", max_persist / 1048576.0, (hi - lo) / 1048576.0,
            (hi - lo) >= b0 + b1 ? "both rank tables" : "first rank table", attr.accessPolicyWindow.hitRatio, rc == cudaSuccess ? "set" : cudaGetErrorString(rc));
}

static int read_bwt_file(const std::string& path, std::vector<uint8_t>& runs, uint64_t& n_strings, uint64_t& n_symbols)
{
    std::ifstream in(path.c_str(), std::ios::binary);
    if (!in) { set_error("cannot open %s", path.c_str()); return PBSC_ERR_IO; }
    uint16_t magic = 0; uint64_t n_runs = 0; int32_t flag = 0;
    in.read((char*)&magic, 2);
    if (!in || magic != 0xCACA) { set_error("BWT file is not properly formatted, aborting"); return PBSC_ERR_FORMAT; }   // BWTReaderBinary.cpp:61-65
    in.read((char*)&n_strings, 8); in.read((char*)&n_symbols, 8); in.read((char*)&n_runs, 8); in.read((char*)&flag, 4);
    if (!in) { set_error("truncated header in %s", path.c_str()); return PBSC_ERR_FORMAT; }
    {
        // the header is untrusted: a run holds 1..31 symbols, and the runs must fit the file
        struct stat sb;
        if (stat(path.c_str(), &sb) != 0 || (uint64_t)sb.st_size < 30 + n_runs || n_runs > n_symbols || n_symbols > 31 * n_runs)
        { set_error("%s: header (%llu runs, %llu symbols) does not match the file size", path.c_str(), (unsigned long long)n_runs, (unsigned long long)n_symbols); return PBSC_ERR_FORMAT; }
    }
    runs.resize(n_runs);
    in.read((char*)runs.data(), (std::streamsize)n_runs);
    if ((uint64_t)in.gcount() != n_runs) { set_error("truncated run data in %s", path.c_str()); return PBSC_ERR_FORMAT; }
    return PBSC_OK;
}

int pbsc_index_load(const char* prefix, int device, int require_sai, pbsc_index** out)
try
{
    if (!prefix || !out) { set_error("pbsc_index_load: null argument"); return PBSC_ERR_ARG; }
    std::string p(prefix);
    std::vector<uint8_t> r0, r1;
    uint64_t s0, n0, s1, n1;
    int rc = read_bwt_file(p + ".bwt", r0, s0, n0);
    if (rc != PBSC_OK) return rc;
    rc = read_bwt_file(p + ".rbwt", r1, s1, n1);
    if (rc != PBSC_OK) return rc;
    if (require_sai)
    {
        struct stat st;
        if (stat((p + ".sai").c_str(), &st) != 0) { set_error("cannot open %s.sai", prefix); return PBSC_ERR_IO; }
    }
    rc = pbsc_index_create(r0.data(), r0.size(), n0, s0, r1.data(), r1.size(), n1, s1, device, out);
    if (rc == PBSC_OK) { (*out)->src_runs[0] = r0.size(); (*out)->src_runs[1] = r1.size(); }
    return rc;
}
PBSC_CATCH_ALL("pbsc_index_load")

int pbsc_index_create_synthetic(uint64_t n_symbols, uint64_t n_strings, uint64_t seed, int device, pbsc_index** out)
try
{
    if (!out || n_symbols == 0) { set_error("pbsc_index_create_synthetic: bad argument"); return PBSC_ERR_ARG; }
    if (n_symbols >= 0xffffffffull) { set_error("synthetic BWT of %llu symbols; this build supports < 2^32-1", (unsigned long long)n_symbols); return PBSC_ERR_LIMIT; }
    *out = nullptr;
    int ndev = 0;
    PBSC_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("device %d not available (%d devices)", device, ndev); return PBSC_ERR_CUDA; }
    PBSC_CUDA(cudaSetDevice(device));
    pbsc_index* idx = new pbsc_index();
    idx->device = device;
    idx->dev.prefix = nullptr; idx->dev.k0 = 0; idx->dev.idmer_valid = nullptr; idx->dev.idmer_len = 0;
    auto fail = [&](int rc) { pbsc_index_destroy(idx); return rc; };
    for (int w = 0; w < 2; w++)
    {
        const uint64_t nb = n_symbols / 64 + 1;
        const uint64_t sd = splitmix64(seed * 2 + w);
        DevBuf<uint32_t> dpos, dsorted, cnt[4], cum[4], uniq;
        DevBuf<uint8_t> tmp;
        if (cudaMalloc((void**)&idx->d_blocks[w], nb * sizeof(FmBlock)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc blocks", __FILE__, __LINE__));
        if (cudaMalloc((void**)&idx->d_dollar[w], (n_strings + 1) * sizeof(uint32_t)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc dollar", __FILE__, __LINE__));
        if (dpos.alloc(n_strings + 1) != cudaSuccess || uniq.alloc(1) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc", __FILE__, __LINE__));
        for (int c = 0; c < 4; c++) if (cnt[c].alloc(nb) != cudaSuccess || cum[c].alloc(nb) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc", __FILE__, __LINE__));
        if (n_strings)
        {
            synth_dollar_kernel<<<(unsigned)((n_strings + 255) / 256), 256>>>(dpos.p, n_strings, n_symbols, sd);
            size_t tb = 0;
            cub::DeviceRadixSort::SortKeys(nullptr, tb, dpos.p, idx->d_dollar[w], (int)n_strings);
            if (tmp.alloc(tb) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc", __FILE__, __LINE__));
            cub::DeviceRadixSort::SortKeys(tmp.p, tb, dpos.p, idx->d_dollar[w], (int)n_strings);
        }
        synth_blocks_kernel<<<(unsigned)((nb + 255) / 256), 256>>>(idx->d_blocks[w], cnt[0].p, cnt[1].p, cnt[2].p, cnt[3].p, nb, n_symbols,
                                                                    idx->d_dollar[w], (uint32_t)n_strings, sd);
        uint64_t total[5] = {0, 0, 0, 0, 0};
        for (int c = 0; c < 4; c++)
        {
            size_t tb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt[c].p, cum[c].p, (int)nb);
            DevBuf<uint8_t> t2;
            if (t2.alloc(tb) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc", __FILE__, __LINE__));
            cub::DeviceScan::ExclusiveSum(t2.p, tb, cnt[c].p, cum[c].p, (int)nb);
            uint32_t lastc = 0, lastv = 0;
            cudaMemcpy(&lastc, cum[c].p + nb - 1, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(&lastv, cnt[c].p + nb - 1, 4, cudaMemcpyDeviceToHost);
            total[c + 1] = (uint64_t)lastc + lastv;
        }
        synth_headers_kernel<<<(unsigned)((nb + 255) / 256), 256>>>(idx->d_blocks[w], cum[0].p, cum[1].p, cum[2].p, cum[3].p, nb);
        total[0] = n_symbols - total[1] - total[2] - total[3] - total[4];
        if (cudaDeviceSynchronize() != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "synthetic index kernels", __FILE__, __LINE__));
        idx->n_symbols[w] = n_symbols;
        idx->n_strings[w] = total[0];
        idx->n_blocks[w] = nb;
        if (cudaMalloc((void**)&idx->d_dmask[w], nb * sizeof(uint64_t)) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaMalloc dmask", __FILE__, __LINE__));
        cudaMemset(idx->d_dmask[w], 0, nb * sizeof(uint64_t));
        if (n_strings) dollar_mask_kernel<<<(unsigned)((n_strings + 255) / 256), 256>>>(idx->d_dollar[w], (uint32_t)n_strings, (unsigned long long*)idx->d_dmask[w]);
        if (cudaDeviceSynchronize() != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "dollar mask", __FILE__, __LINE__));
        idx->device_bytes += nb * (sizeof(FmBlock) + sizeof(uint64_t)) + (n_strings + 1) * sizeof(uint32_t);
        fill_table(idx->dev.t[w], idx->d_blocks[w], idx->d_dollar[w], idx->d_dmask[w], n_symbols, total);
        // duplicates in the '$' list are harmless for block generation (skipped) but the count used by
        // count_dollars() must see each position once: compact in place on the host side of the list
        if (n_strings)
        {
            std::vector<uint32_t> h(n_strings);
            cudaMemcpy(h.data(), idx->d_dollar[w], n_strings * 4, cudaMemcpyDeviceToHost);
            h.erase(std::unique(h.begin(), h.end()), h.end());
            cudaMemcpy(idx->d_dollar[w], h.data(), h.size() * 4, cudaMemcpyHostToDevice);
            idx->dev.t[w].n_dollar = (uint32_t)h.size();
        }
    }
    if (cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "cudaStreamCreate", __FILE__, __LINE__));
    cudaDeviceGetAttribute(&idx->sm_count, cudaDevAttrMultiProcessorCount, device);
    apply_l2_policy(idx, device);
    *out = idx;
    return PBSC_OK;
}
PBSC_CATCH_ALL("pbsc_index_create_synthetic")

int pbsc_index_build_prefix_table(pbsc_index* idx, int k0)
try
{
    if (!idx || k0 < 0 || k0 > 15) { set_error("pbsc_index_build_prefix_table: k0 must be in 0..15"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(idx->device));
    if (idx->primary) { set_error("pbsc_index_build_prefix_table: call it on the index, not on a lane"); return PBSC_ERR_ARG; }
    {
        // no batch may be running while the table is replaced
        std::unique_lock<std::mutex> lk(idx->run_mu);
        idx->lane_cv.wait(lk, [&] { return idx->lanes_active == 0; });
    }
    auto in_blob = [&](const void* p) { return idx->blob && (const uint8_t*)p >= (const uint8_t*)idx->blob && (const uint8_t*)p < (const uint8_t*)idx->blob + idx->blob_bytes; };
    if (idx->d_prefix) { idx->device_bytes -= (sizeof(PrefixEntry) << (2 * idx->dev.k0)); if (!in_blob(idx->d_prefix)) cudaFree(idx->d_prefix); idx->d_prefix = nullptr; }
    idx->dev.prefix = nullptr;
    idx->dev.k0 = 0;
    if (k0 == 0) return PBSC_OK;
    const uint64_t n_final = 1ull << (2 * k0);
    PrefixEntry *a = nullptr, *b = nullptr;
    PBSC_CUDA(cudaMalloc((void**)&a, n_final * sizeof(PrefixEntry)));
    if (k0 > 1 && cudaMalloc((void**)&b, (n_final >> 2) * sizeof(PrefixEntry)) != cudaSuccess) { cudaFree(a); return cuda_fail(cudaGetLastError(), "cudaMalloc prefix", __FILE__, __LINE__); }
    // levels 1..k0 ping-pong so that the last level lands in `a`
    PrefixEntry* bufs[2] = {a, b};
    int cur = (k0 & 1) ? 0 : 1;   // level 1 buffer such that level k0 is buffer 0
    const PrefixEntry* prev = nullptr;
    for (int level = 1; level <= k0; level++)
    {
        const uint64_t n_cur = 1ull << (2 * level);
        prefix_level_kernel<<<(unsigned)((n_cur + 255) / 256), 256, 0, idx->stream>>>(idx->dev, prev, bufs[cur], n_cur, level);
        prev = bufs[cur];
        cur ^= 1;
    }
    cudaError_t e = cudaStreamSynchronize(idx->stream);
    if (b) cudaFree(b);
    if (e != cudaSuccess) { cudaFree(a); return cuda_fail(e, "prefix table kernels", __FILE__, __LINE__); }
    idx->d_prefix = a;
    idx->dev.prefix = a;
    idx->dev.k0 = k0;
    idx->device_bytes += n_final * sizeof(PrefixEntry);
    return PBSC_OK;
}
PBSC_CATCH_ALL("pbsc_index_build_prefix_table")

void pbsc_index_destroy(pbsc_index* idx)
{
    if (!idx) return;
    cudaSetDevice(idx->device);
    for (pbsc_index* s : idx->shadows)
    {
        for (auto& kv : s->arena) if (kv.second.p) cudaFree(kv.second.p);
        if (s->stream) cudaStreamDestroy(s->stream);
        if (s->stream2) cudaStreamDestroy(s->stream2);
        if (s->ev_a) cudaEventDestroy(s->ev_a);
        if (s->ev_b) cudaEventDestroy(s->ev_b);
        delete s;
    }
    idx->shadows.clear();
    auto in_blob = [&](const void* p) { return idx->blob && (const uint8_t*)p >= (const uint8_t*)idx->blob && (const uint8_t*)p < (const uint8_t*)idx->blob + idx->blob_bytes; };
    for (int w = 0; w < 2; w++)
    {
        if (idx->d_blocks[w] && !in_blob(idx->d_blocks[w])) cudaFree(idx->d_blocks[w]);
        if (idx->d_dollar[w] && !in_blob(idx->d_dollar[w])) cudaFree(idx->d_dollar[w]);
        if (idx->d_dmask[w] && !in_blob(idx->d_dmask[w])) cudaFree(idx->d_dmask[w]);
    }
    if (idx->d_prefix && !in_blob(idx->d_prefix)) cudaFree(idx->d_prefix);
    if (idx->d_idmer_valid && !in_blob(idx->d_idmer_valid)) cudaFree(idx->d_idmer_valid);
    if (idx->blob) cudaFree(idx->blob);
    for (auto& kv : idx->arena) if (kv.second.p) cudaFree(kv.second.p);
    if (idx->stream) cudaStreamDestroy(idx->stream);
    if (idx->stream2) cudaStreamDestroy(idx->stream2);
    if (idx->ev_a) cudaEventDestroy(idx->ev_a);
    if (idx->ev_b) cudaEventDestroy(idx->ev_b);
    dev_cache_trim(idx->device);
    delete idx;
}

}  // extern "C"

// ---- random 32-byte-sector gather: the roofline denominator of the rank-query kernels ----
namespace pbsc {
template <int HALVES>
__global__ void random_sector_kernel(const uint4* __restrict__ buf, uint64_t sector_mask, uint64_t loads, unsigned long long* sink)
{
    // every thread issues `loads` independent reads of pseudo-random 32-byte sectors, sixteen 16-byte loads in flight:
    // HALVES = 2 reads both halves of 8 sectors (what a rank query does), HALVES = 1 one half of 16 sectors
    uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    unsigned long long acc = 0;
    constexpr int SECTORS = 16 / HALVES;
    for (uint64_t i = 0; i < loads; i += SECTORS)
    {
        uint4 a[16];
        #pragma unroll
        for (int u = 0; u < SECTORS; u++)
        {
            x = x * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t sct = (x >> 24) & sector_mask;
            #pragma unroll
            for (int h = 0; h < HALVES; h++) a[u * HALVES + h] = __ldg(buf + 2 * sct + h);
        }
        #pragma unroll
        for (int u = 0; u < 16; u++) acc += a[u].x ^ a[u].w;
    }
    if (acc == 0x12345) *sink = acc;
}
}  // namespace pbsc

extern "C" {

/* Measured peak of independent random 32-byte sector reads over a buffer of `bytes` bytes on `device` (GB/s): what a kernel
 * whose every lookup is one aligned sector of the rank table can reach at best. */
int pbsc_random_sector_bench(int device, uint64_t bytes, float* gbps)
{
    if (!gbps || bytes < (1u << 20)) { set_error("pbsc_random_sector_bench: bad argument"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(device));
    DevBuf<uint4> buf; DevBuf<unsigned long long> sink;
    uint64_t n_sectors = 1;
    while (n_sectors * 2 * 32 <= bytes) n_sectors *= 2;   // largest power of two that fits: the sector index is a mask
    PBSC_CUDA(buf.alloc(n_sectors * 2)); PBSC_CUDA(sink.alloc(1));
    PBSC_CUDA(cudaMemset(buf.p, 1, n_sectors * 32));
    int sms = 0;
    PBSC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sms * 8, threads = 256;
    const uint64_t loads = 8192;
    cudaEvent_t e0, e1;
    PBSC_CUDA(cudaEventCreate(&e0)); PBSC_CUDA(cudaEventCreate(&e1));
    float best = 0;
    for (int rep = 0; rep < 6; rep++)
    {
        cudaEventRecord(e0);
        if (rep & 1) pbsc::random_sector_kernel<1><<<blocks, threads>>>(buf.p, n_sectors - 1, loads, sink.p);
        else pbsc::random_sector_kernel<2><<<blocks, threads>>>(buf.p, n_sectors - 1, loads, sink.p);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const float g = (float)((double)blocks * threads * loads * 32.0 / (ms * 1e-3) / 1e9);
        if (rep >= 2 && g > best) best = g;   // the better of the two access shapes, after one warm-up of each
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    PBSC_CUDA(cudaGetLastError());
    *gbps = best;
    return PBSC_OK;
}

/* pinned host memory for the callers' read and result buffers: copies from/to it run at PCIe speed and asynchronously */
int pbsc_host_alloc(void** p, size_t bytes)
{
    if (!p) { set_error("pbsc_host_alloc: null argument"); return PBSC_ERR_ARG; }
    *p = nullptr;
    PBSC_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable));
    return PBSC_OK;
}
void pbsc_host_free(void* p) { if (p) cudaFreeHost(p); }
/* page-lock memory the caller already owns (e.g. a shared-memory mapping several processes write their results into) */
int pbsc_host_register(void* p, size_t bytes)
{
    if (!p || !bytes) { set_error("pbsc_host_register: null argument"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return PBSC_OK;
}
void pbsc_host_unregister(void* p) { if (p && cudaHostUnregister(p) != cudaSuccess) cudaGetLastError(); }
void pbsc_trim(int device) { dev_cache_trim(device); }

uint64_t pbsc_index_num_symbols(const pbsc_index* idx, int which) { return idx && (which == 0 || which == 1) ? idx->n_symbols[which] : 0; }
uint64_t pbsc_index_num_strings(const pbsc_index* idx, int which) { return idx && (which == 0 || which == 1) ? idx->n_strings[which] : 0; }
uint64_t pbsc_index_device_bytes(const pbsc_index* idx) { return idx ? idx->device_bytes : 0; }

int pbsc_index_get_symbols(const pbsc_index* idx, int which, uint64_t first, uint64_t count, char* out)
try
{
    if (!idx || !out || (which != 0 && which != 1) || first + count > idx->n_symbols[which]) { set_error("pbsc_index_get_symbols: bad argument"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(idx->device));
    DevBuf<char> d;
    PBSC_CUDA(d.alloc(count));
    if (count) get_symbols_kernel<<<(unsigned)((count + 255) / 256), 256, 0, idx->stream>>>(idx->dev.t[which], first, count, d.p);
    PBSC_CUDA(cudaStreamSynchronize(idx->stream));
    PBSC_CUDA(cudaMemcpy(out, d.p, count, cudaMemcpyDeviceToHost));
    return PBSC_OK;
}
PBSC_CATCH_ALL("pbsc_index_get_symbols")

int pbsc_findinterval_batch(pbsc_index* idx, int which, const char* kmers, const uint64_t* offsets, uint64_t n,
                            int64_t* lower, int64_t* upper, uint8_t* steps)
try
{
    if (!idx || !kmers || !offsets || !lower || !upper || (which != 0 && which != 1)) { set_error("pbsc_findinterval_batch: bad argument"); return PBSC_ERR_ARG; }
    if (n == 0) return PBSC_OK;
    for (uint64_t i = 0; i < n; i++) if (offsets[i + 1] <= offsets[i]) { set_error("pbsc_findinterval_batch: empty query %llu", (unsigned long long)i); return PBSC_ERR_ARG; }
    const uint64_t nbytes = offsets[n];
    for (uint64_t i = 0; i < nbytes; i++) if (base_code(kmers[i]) < 0) { set_error("pbsc_findinterval_batch: non-ACGT character in query"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(idx->device));
    DevBuf<char> dk; DevBuf<uint64_t> doff; DevBuf<int64_t> dl, du; DevBuf<uint8_t> ds;
    PBSC_CUDA(dk.alloc(nbytes)); PBSC_CUDA(doff.alloc(n + 1)); PBSC_CUDA(dl.alloc(n)); PBSC_CUDA(du.alloc(n)); PBSC_CUDA(ds.alloc(n));
    PBSC_CUDA(cudaMemcpyAsync(dk.p, kmers, nbytes, cudaMemcpyHostToDevice, idx->stream));
    PBSC_CUDA(cudaMemcpyAsync(doff.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, idx->stream));
    findinterval_bytes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, idx->stream>>>(idx->dev, which, dk.p, doff.p, n, dl.p, du.p, ds.p);
    PBSC_CUDA(cudaGetLastError());
    PBSC_CUDA(cudaMemcpyAsync(lower, dl.p, n * 8, cudaMemcpyDeviceToHost, idx->stream));
    PBSC_CUDA(cudaMemcpyAsync(upper, du.p, n * 8, cudaMemcpyDeviceToHost, idx->stream));
    if (steps) PBSC_CUDA(cudaMemcpyAsync(steps, ds.p, n, cudaMemcpyDeviceToHost, idx->stream));
    PBSC_CUDA(cudaStreamSynchronize(idx->stream));
    return PBSC_OK;
}
PBSC_CATCH_ALL("pbsc_findinterval_batch")

int pbsc_findinterval_device(pbsc_index* idx, int which, const uint64_t* d_kmers2bit, int k, uint64_t n,
                             int64_t* d_lower, int64_t* d_upper, uint8_t* d_steps, float* ms)
{
    if (!idx || !d_kmers2bit || !d_lower || !d_upper || k < 1 || k > 32 || (which != 0 && which != 1)) { set_error("pbsc_findinterval_device: bad argument"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(idx->device));
    cudaEvent_t e0, e1;
    PBSC_CUDA(cudaEventCreate(&e0)); PBSC_CUDA(cudaEventCreate(&e1));
    PBSC_CUDA(cudaEventRecord(e0, idx->stream));
    if (n)
    {
        unsigned grid = (unsigned)((n + 255) / 256);
        if (idx->dev.prefix && k >= idx->dev.k0) findinterval_packed_kernel<true><<<grid, 256, 0, idx->stream>>>(idx->dev, which, d_kmers2bit, k, n, d_lower, d_upper, d_steps);
        else findinterval_packed_kernel<false><<<grid, 256, 0, idx->stream>>>(idx->dev, which, d_kmers2bit, k, n, d_lower, d_upper, d_steps);
    }
    PBSC_CUDA(cudaEventRecord(e1, idx->stream));
    cudaError_t e = cudaStreamSynchronize(idx->stream);
    float t = 0;
    if (e == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess) return cuda_fail(e, "findinterval_packed_kernel", __FILE__, __LINE__);
    if (ms) *ms = t;
    return PBSC_OK;
}

// ------------------------------------------------------------------------------------------
// parameters — StriDe/PacBioSelfCorrection.cpp:71-101,195-231; PacBio/KmerThreshold.cpp:11-79
// ------------------------------------------------------------------------------------------
void pbsc_params_default(pbsc_params* p)
{
    memset(p, 0, sizeof *p);
    p->pb_coverage = 90; p->error_rate = 0.15; p->start_kmer = 19; p->next_target = 1; p->max_leaves = 32;
    p->idmer_len = 9; p->min_kmer = 13; p->genome = 10; p->mode = 1;
    p->scan_kmer = 19; p->kmer_up_bound = 50; p->radius = 100; p->hh_ratio = 0.6f;
}

// the float polynomial of KmerThreshold::calculate, evaluated in source order; volatile keeps each partial
// result rounded to float so that no compiler can contract or widen the chain
static float threshold_value(int mode, int x, int y)
{
    static const float formula[3][6] = {
        {0.0004799107143, -0.008037815126, 0.03673552754, 0.1850695903, -1.572552521, 18.0522088},
        {0.0003348214286, -0.009112394958, 0.04286714686, 0.240519958, -1.8793367350, 21.29319228},
        {0.01714285714, -0.6193907563, 2.266956783, 17.28450630, -100.6983493, 1103.571729}};
    const float* f = formula[mode];
    volatile float t0 = f[0] * x; t0 = t0 * x;
    volatile float t1 = f[1] * x; t1 = t1 * y;
    volatile float t2 = f[2] * y; t2 = t2 * y;
    volatile float t3 = f[3] * x;
    volatile float t4 = f[4] * y;
    volatile float v = t0 + t1; v = v + t2; v = v + t3; v = v + t4; v = v + f[5];
    return fmaxf(v, 2.0f);
}

int pbsc_params_derive(pbsc_params* p)
{
    if (!p) { set_error("pbsc_params_derive: null"); return PBSC_ERR_ARG; }
    // option validation — PacBioSelfCorrection.cpp:364-423
    if (p->pb_coverage <= 0) { set_error("invalid number of coverage: %d, must be greater than zero", p->pb_coverage); return PBSC_ERR_ARG; }
    if (p->error_rate < 0 || p->error_rate > 1) { set_error("invalid error rate: %g, must be 0 ~ 1", p->error_rate); return PBSC_ERR_ARG; }
    if (p->start_kmer <= 0) { set_error("invalid start kmer length: %d, must be greater than zero", p->start_kmer); return PBSC_ERR_ARG; }
    if (p->next_target <= 0) { set_error("invalid number of next target: %d, must be greater than zero", p->next_target); return PBSC_ERR_ARG; }
    if (p->max_leaves <= 0) { set_error("invalid number of max leaves:%d, must be greater than zero", p->max_leaves); return PBSC_ERR_ARG; }
    if (p->idmer_len <= 0) { set_error("invalid kmer length to identify similar reads%d, must be greater than zero", p->idmer_len); return PBSC_ERR_ARG; }
    if (p->min_kmer <= 0) { set_error("invalid min kmer length:%d, must be greater than zero", p->min_kmer); return PBSC_ERR_ARG; }
    if (p->genome != 5 && p->genome != 10 && p->genome != 100) { set_error("invalid genome size: %d, must be (5/10/100)[m]", p->genome); return PBSC_ERR_ARG; }
    if (p->mode < 0 || p->mode > 2) { set_error("invalid mode: %d, must be (0/1/2)", p->mode); return PBSC_ERR_ARG; }
    const int order = p->genome == 5 ? 0 : p->genome == 10 ? 1 : 2;
    static const int size[3] = {17, 19, 21};
    if (!p->adjust)
    {
        p->start_kmer = size[order];
        p->offset[1] = 2 * std::min(std::max((p->pb_coverage / 30 - 1), 0), (order + 1));
        p->offset[2] = -2 * (order + 1);
    }
    int pool[8] = {5, 9, 19, p->start_kmer + p->offset[0], p->start_kmer + p->offset[1], p->start_kmer + p->offset[2], 0, 0};
    std::sort(pool, pool + 6);
    p->n_pool = (int)(std::unique(pool, pool + 6) - pool);
    for (int i = 0; i < 8; i++) p->pool[i] = i < p->n_pool ? pool[i] : 0;
    p->scan_kmer = 19; p->kmer_up_bound = 50; p->radius = 100; p->hh_ratio = 0.6f;
    // limits of this build (documented in DESIGN.md)
    if (p->pool[0] < 1 || p->pool[p->n_pool - 1] > 50) { set_error("k-mer pool sizes must lie in 1..50 (got %d..%d)", p->pool[0], p->pool[p->n_pool - 1]); return PBSC_ERR_LIMIT; }
    if (p->start_kmer + 2 > 60) { set_error("start k-mer length %d too large for this build (max 58)", p->start_kmer); return PBSC_ERR_LIMIT; }
    if (p->idmer_len > 13 || p->idmer_len < 5) { set_error("idmer length %d outside this build's 5..13", p->idmer_len); return PBSC_ERR_LIMIT; }
    if (p->min_kmer < p->idmer_len) { set_error("min k-mer size %d below idmer length %d is not supported", p->min_kmer, p->idmer_len); return PBSC_ERR_LIMIT; }
    if (p->max_leaves > 32) { set_error("max leaves %d above this build's 32", p->max_leaves); return PBSC_ERR_LIMIT; }
    // KmerThreshold::initialize(-1, 50, cov, dir) — KmerThreshold.cpp:43-63
    for (int mode = 0; mode <= 2; mode++)
    {
        for (int k = 0; k < 52; k++) p->threshold[mode][k] = 0.0f;
        float cavity = 3.402823466e+38f;
        for (int k = 15; k <= 50; k++)
        {
            cavity = fminf(cavity, threshold_value(mode, p->pb_coverage, k));
            p->threshold[mode][k] = cavity;
        }
    }
    // LongReadCorrectByOverlap.cpp:68-70 (host libm pow, as in the reference)
    for (int i = 0; i <= 100; i++) p->freqs_of_kmer[i] = 0;
    for (int i = p->min_kmer; i <= 100; i++) p->freqs_of_kmer[i] = pow(1 - p->error_rate, i) * (size_t)p->pb_coverage;
    return PBSC_OK;
}

int pbsc_threshold_table_text(const pbsc_params* p, char* buf, size_t cap)
try
{
    if (!p || !buf) { set_error("pbsc_threshold_table_text: null"); return PBSC_ERR_ARG; }
    std::ostringstream out;
    out << "Coverage : " << p->pb_coverage << "\n" << "size\tlowcov\tunique\trepeat\n";
    for (int k = 15; k <= 50; k++) out << k << "\t" << p->threshold[0][k] << "\t" << p->threshold[1][k] << "\t" << p->threshold[2][k] << "\n";
    std::string s = out.str();
    if (s.size() + 1 > cap) { set_error("pbsc_threshold_table_text: buffer too small"); return PBSC_ERR_LIMIT; }
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}
PBSC_CATCH_ALL("pbsc_threshold_table_text")

/* measurement build only (-DPBSC_COUNT_OCC, libpbsc_count.so): distinct 32-byte index sectors asked for since the last reset,
 * per kernel family: out[0] seed phase, out[1] walk setup, out[2] level loop, out[3] DP fallback, out[4] other.
 * Returns 1 in the measurement build, 0 (and zeros) in the product build. */
int pbsc_occ_counts(uint64_t* out, int reset)
{
    if (!out) return PBSC_ERR_ARG;
    unsigned long long* c = occ_counts();
    for (int i = 0; i < 5; i++) { out[i] = c[i]; if (reset) c[i] = 0; }
#ifdef PBSC_COUNT_OCC
    return 1;
#else
    return 0;
#endif
}

int pbsc_last_timing(pbsc_timing* t)
{
    if (!t) return PBSC_ERR_ARG;
    const Timing& s = last_timing();
    t->h2d_ms = s.h2d_ms; t->seed_ms = s.seed_ms; t->extend_ms = s.extend_ms; t->d2h_ms = s.d2h_ms; t->total_ms = s.total_ms;
    t->kernel_launches = s.kernel_launches; t->seed_pairs = s.seed_pairs; t->rank_queries = s.rank_queries;
    t->dp_ms = s.dp_ms; t->dp_jobs = s.dp_jobs; t->dp_rows = s.dp_rows;
    t->walk_ms = s.walk_ms; t->walk_launches = s.walk_launches; t->dp_thread_rows = s.dp_thread_rows;
    return PBSC_OK;
}

}  // extern "C"
