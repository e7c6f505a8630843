// pbsc_seed.cu — seed phase: LongReadProbe::searchSeedsWithHybridKmers
// (PacBio/LongReadProbe.cpp:34-227) as four kernels over a device-resident batch of reads.
//
//   seed_features_kernel   thread per read base: the KmerFeature chain over the k-mer pool
//                          (KmerFeature.h:37-64, LongReadProbe.cpp:143-150); 2 strands x 2 bounds
//                          of independent rank sectors in flight per thread
//   seed_scan_kernel       thread per read: sliding repeat-ratio attribute (LongReadProbe.cpp:120-182)
//                          and the static/dynamic hybrid k-mer scan (:46-104)
//   seed_bestk_kernel      thread per (seed, pole): SeedFeature::modifyKmerSize (SeedFeature.cpp:48-78)
//   seed_hitchhike_kernel  thread per read: removeHitchhikingSeeds (LongReadProbe.cpp:187-227)
#include <algorithm>
#include "pbsc_batch.cuh"

namespace pbsc {

// 16 bases per thread; *bad is raised when anything but A/C/G/T shows up (SeqReader.cpp:118-125 exits on such reads)
__global__ void ascii_to_codes_kernel(const char* __restrict__ in, uint8_t* __restrict__ out, uint64_t n, unsigned int* bad)
{
    const uint64_t i0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 16;
    if (i0 >= n) return;
    unsigned char c[16];
    bool wrong = false;
    if (i0 + 16 <= n && ((uintptr_t)(in + i0) & 15) == 0)
    {
        const uint4 v = *reinterpret_cast<const uint4*>(in + i0);
        memcpy(c, &v, 16);
        #pragma unroll
        for (int x = 0; x < 16; x++)
        {
            const unsigned char b = c[x];
            wrong |= !(b == 'A' || b == 'C' || b == 'G' || b == 'T');
            c[x] = b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3;
        }
        uint4 o;
        memcpy(&o, c, 16);
        *reinterpret_cast<uint4*>(out + i0) = o;
    }
    else
    {
        for (uint64_t i = i0; i < n && i < i0 + 16; i++)
        {
            const unsigned char b = (unsigned char)in[i];
            wrong |= !(b == 'A' || b == 'C' || b == 'G' || b == 'T');
            out[i] = b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3;
        }
    }
    if (wrong) atomicOr(bad, 1u);
}

__device__ __forceinline__ uint64_t find_read(const uint64_t* __restrict__ offsets, uint64_t n_reads, uint64_t g)
{
    uint64_t lo = 0, hi = n_reads;   // largest r with offsets[r] <= g
    while (hi - lo > 1) { uint64_t m = (lo + hi) >> 1; if (__ldg(offsets + m) <= g) lo = m; else hi = m; }
    return lo;
}

__device__ __forceinline__ int64_t bi_freq(const Interval& f, const Interval& r) { return (int64_t)f.size() + (int64_t)r.size(); }

// KmerFeature::isLowComplexity (KmerFeature.h:116-126)
__device__ __forceinline__ bool low_complexity(const int* count, int size)
{
    int a = count[0], b = count[1], c = count[2], d = count[3];
    // top two of four
    int hi1 = max(max(a, b), max(c, d));
    int sum = a + b + c + d;
    int lo1 = min(min(a, b), min(c, d));
    // second largest = sum - max - min - (the other middle one); compute by sorting network
    int x0 = min(a, b), x1 = max(a, b), y0 = min(c, d), y1 = max(c, d);
    int second = max(min(x1, y1), max(x0, y0));
    (void)sum; (void)lo1;
    const float m = 0.7f, dthr = 0.9f;
    bool isMonmer = __fdiv_rn((float)hi1, (float)size) >= m;
    bool isDimer = __fdiv_rn((float)(second + hi1), (float)size) >= dthr;
    return isMonmer || isDimer;
}

// one KmerFeature chain per read position
__global__ void __launch_bounds__(256)
seed_features_kernel(FmIndexDev idx, SeedParamsDev P, const uint8_t* __restrict__ codes, const uint64_t* __restrict__ offsets,
                     uint64_t n_reads, uint64_t n_bases, StaticFeat* __restrict__ feats, uint8_t* __restrict__ cls)
{
    uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= n_bases) return;
    const uint64_t r = find_read(offsets, n_reads, g);
    const uint64_t rend = __ldg(offsets + r + 1);
    const int avail = (int)min((uint64_t)64, rend - g);   // bases available from this position (pool sizes <= 50)
    const uint8_t* s = codes + g;
    const FmTable& RB = idx.t[PBSC_RBWT];
    const FmTable& FB = idx.t[PBSC_BWT];
    float repeatValue = P.threshold[2][P.scan_kmer];

    int count[4] = {0, 0, 0, 0};
    int size;
    Interval f, rv;
    {
        // base k-mer: findBiInterval(word, count) — BWTAlgorithms.cpp:32-38 with the early break of :25-29 per strand
        const int K0 = P.pool[0];
        size = min(K0, avail);
        int c0 = s[0];
        count[c0]++;
        f = init_interval(RB, c0);
        for (int j = 1; j < size; j++) { int c = s[j]; count[c]++; f = update_interval(RB, f, c); if (!f.valid()) break; }
        rv = init_interval(FB, 3 - c0);
        for (int j = 1; j < size; j++) { rv = update_interval(FB, rv, 3 - s[j]); if (!rv.valid()) break; }
    }
    uint8_t cl = 0x22;   // both classifications "1" (unique) by default, encoded +1
    for (int pi = 0; pi < P.n_pool; pi++)
    {
        const int K = P.pool[pi];
        if (pi > 0)
        {
            // KmerFeature(indices, seq, pos, len, base) with base != nullptr — KmerFeature.h:54-60: expand() never breaks
            while (size < K && size < avail)
            {
                int c = s[size];
                size++;
                count[c]++;
                if (f.valid()) f = update_interval(RB, f, c);
                if (rv.valid()) rv = update_interval(FB, rv, 3 - c);
            }
        }
        const bool fake = (K != size);
        const int freq = (int)((f.valid() ? (int64_t)f.size() : 0) + (rv.valid() ? (int64_t)rv.size() : 0));
        if (K == P.scan_kmer)
        {
            // LongReadProbe.cpp:151-157 (entering) and :161-168 (leaving)
            int fq = low_complexity(count, size) ? -1 : (fake ? -1 : freq);
            int in = fq < 0 ? -1 : ((float)fq >= repeatValue ? 2 : 1);
            int out = fq <= 0 ? -1 : ((float)fq >= repeatValue ? 2 : 1);
            cl = (uint8_t)((in + 1) | ((out + 1) << 4));
        }
        for (int sl = 0; sl < P.n_static; sl++)
        {
            if (P.static_size[sl] == K)
            {
                StaticFeat o;
                o.fwd_lo = f.lo; o.fwd_size = (uint32_t)f.size();
                o.rvc_lo = rv.lo; o.rvc_size = (uint32_t)rv.size();
                o.freq = freq;
                o.count_fake = (uint32_t)count[0] | ((uint32_t)count[1] << 7) | ((uint32_t)count[2] << 14) | ((uint32_t)count[3] << 21) | (fake ? 0x80000000u : 0u);
                feats[(uint64_t)sl * n_bases + g] = o;
            }
        }
    }
    cls[g] = cl;
}

// ---- getSeqAttribute (LongReadProbe.cpp:120-182) in two data-parallel steps ----
// (1) warp per read: running counts of the four indicators the sliding box needs
//     .x = #entering k-mers classified -1 up to and including this position, .y = #entering classified 2,
//     .z = #leaving k-mers classified -1,                                     .w = #leaving classified 2
__global__ void __launch_bounds__(128)
seed_prefix_kernel(const uint64_t* __restrict__ offsets, uint64_t n_reads, const uint8_t* __restrict__ cls, uint4* __restrict__ pre)
{
    const uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    const uint64_t base = offsets[r];
    const int64_t L = (int64_t)(offsets[r + 1] - base);
    uint32_t run[4] = {0, 0, 0, 0};
    for (int64_t c = 0; c < L; c += 32)
    {
        const int64_t p = c + lane;
        uint32_t v = 0;
        if (p < L)
        {
            const uint8_t x = cls[base + p];
            const int in = (int)(x & 15) - 1, out = (int)(x >> 4) - 1;
            v = (uint32_t)(in == -1) | ((uint32_t)(in == 2) << 8) | ((uint32_t)(out == -1) << 16) | ((uint32_t)(out == 2) << 24);
        }
        // inclusive scan of four 8-bit counters packed in one word (a chunk holds at most 32 of each)
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (p < L) pre[base + p] = make_uint4(run[0] + (v & 255), run[1] + ((v >> 8) & 255), run[2] + ((v >> 16) & 255), run[3] + (v >> 24));
        const uint32_t tot = __shfl_sync(0xffffffffu, v, 31);
        run[0] += tot & 255; run[1] += (tot >> 8) & 255; run[2] += (tot >> 16) & 255; run[3] += tot >> 24;
    }
}

// (2) thread per position: box counts of the +-150 window from the running counts, then the 2 % repeat-ratio rule
__global__ void __launch_bounds__(256)
seed_attr_kernel(SeedParamsDev P, const uint64_t* __restrict__ offsets, uint64_t n_reads, uint64_t n_bases, const uint4* __restrict__ pre,
                 uint8_t* __restrict__ attr, float* __restrict__ ratio_out)
{
    const uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= n_bases) return;
    const uint64_t r = find_read(offsets, n_reads, g);
    const uint64_t base = __ldg(offsets + r);
    const int64_t L = (int64_t)(__ldg(offsets + r + 1) - base);
    const int pos = (int)(g - base);
    const int range = 300;
    int left = pos - (range >> 1), right = pos + (range >> 1);
    left = max(left, 0);
    right = min(right, (int)(L - 1));
    // entering k-mers: every position <= right has been added; leaving k-mers: every position < left has been removed
    const uint4 in = pre[base + right];
    uint32_t outN = 0, out2 = 0;
    if (left > 0) { const uint4 o = pre[base + left - 1]; outN = o.z; out2 = o.w; }
    const int boxN = (int)in.x - (int)outN, box2 = (int)in.y - (int)out2;
    const int size = (right - left + 1) - boxN;
    const float ratio = (float)((double)__fdiv_rn((float)box2, (float)size) + 0.0005);
    if (ratio_out) ratio_out[g] = ratio;   // --debugseed: extend/<id>.log (LongReadProbe.cpp:172-173)
    uint8_t a = ((double)ratio >= 0.02) ? 2 : 1;
    if (P.manual) a = (uint8_t)P.mode;
    attr[g] = a;
}

// ---- hybrid k-mer scan (LongReadProbe.cpp:46-104) ----
// One iteration of the reference's outer loop is a pure function of its starting position initPos: it returns the
// position the loop continues from and possibly one seed.  seed_candidates_kernel evaluates that function for EVERY
// position in parallel (positions inside a seed region redo part of the extension, a few percent of the feature work);
// seed_chain_kernel then follows the actual sequence initPos -> f(initPos) + 1 from 0, skipping the trivial positions
// (f(p) = p, no seed) 32 at a time.
struct __align__(8) SeedCand
{
    int32_t next;          // value of initPos when the outer iteration ends
    int32_t max_fixed_freq;
    uint8_t emit, len, is_repeat, static_k;
    uint32_t pad;
};

__global__ void __launch_bounds__(256)
seed_candidates_kernel(FmIndexDev idx, SeedParamsDev P, const uint8_t* __restrict__ codes, const uint64_t* __restrict__ offsets,
                       uint64_t n_reads, uint64_t n_bases, const StaticFeat* __restrict__ feats, const uint8_t* __restrict__ attr,
                       SeedCand* __restrict__ cand, uint8_t* __restrict__ ntriv)
{
    const uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= n_bases) return;
    const uint64_t r = find_read(offsets, n_reads, g);
    const uint64_t base = __ldg(offsets + r);
    const int64_t L = (int64_t)(__ldg(offsets + r + 1) - base);
    const int64_t p0 = (int64_t)(g - base);
    ntriv[g] = 0;
    if ((int)L < P.start_kmer) return;
    const uint8_t* s = codes + base;
    const uint8_t* at = attr + base;
    const FmTable& RB = idx.t[PBSC_RBWT];
    const FmTable& FB = idx.t[PBSC_BWT];
    const float inv_hh = __fdiv_rn(1.0f, P.hh_ratio);
    int64_t initPos = p0;
    const int dynamicMode = at[p0];
    const int staticSize = P.start_kmer + P.offset[dynamicMode];
    const StaticFeat* F = feats + (uint64_t)P.slot_of_mode[dynamicMode] * n_bases + base;
    const StaticFeat d0 = F[p0];
    if (d0.count_fake >> 31) return;                                   // static k-mer runs off the read: breaks at once
    // first iteration (currPos == p0) uses the static k-mer itself as dynamic k-mer: decide triviality without the full state
    {
        const float thr = P.threshold[dynamicMode][staticSize];
        const bool valid = d0.fwd_size > 0 && d0.rvc_size > 0;
        if ((float)d0.freq < thr || !valid || staticSize > P.kmer_up_bound) return;
    }
    int dcount[4] = {(int)(d0.count_fake & 127), (int)((d0.count_fake >> 7) & 127), (int)((d0.count_fake >> 14) & 127), (int)((d0.count_fake >> 21) & 127)};
    int dsize = staticSize;
    Interval df, dr;
    df.lo = d0.fwd_lo; df.hi = d0.fwd_lo + d0.fwd_size;
    dr.lo = d0.rvc_lo; dr.hi = d0.rvc_lo + d0.rvc_size;
    int dfreq = d0.freq;
    bool isSeed = false, isRepeat = false;
    int maxFixedMerFreq = dfreq;
    const int64_t seedPos = p0;
    for (int64_t currPos = p0; currPos < L; currPos++)
    {
        const int staticMode = at[currPos];
        const StaticFeat S = F[currPos];
        if (S.count_fake >> 31) break;
        if (isSeed)
        {
            const int b = s[currPos + staticSize - 1];
            dsize++;
            dcount[b]++;
            if (df.valid()) df = update_interval(RB, df, b);
            if (dr.valid()) dr = update_interval(FB, dr, 3 - b);
            dfreq = (int)((df.valid() ? (int64_t)df.size() : 0) + (dr.valid() ? (int64_t)dr.size() : 0));
        }
        const float dynamicThreshold = P.threshold[dynamicMode][dsize];
        const float staticThreshold = P.threshold[staticMode][staticSize];
        const float repeatThreshold = __fmul_rn((float)(5 - ((staticMode >> 1) << 2)), staticThreshold);
        if ((float)S.freq < staticThreshold || (float)dfreq < dynamicThreshold || !(df.valid() && dr.valid()) || dsize > P.kmer_up_bound)
        {
            if (isSeed) { dsize--; dcount[s[seedPos + dsize]]--; }
            break;
        }
        const float freqDiff = __fdiv_rn((float)S.freq, (float)maxFixedMerFreq);
        if (freqDiff < P.hh_ratio)
        {
            initPos++;
            dsize--; dcount[s[seedPos + dsize]]--;
            break;
        }
        else if (freqDiff > inv_hh)
        {
            initPos = currPos - 1;
            isSeed = false;
            break;
        }
        initPos = seedPos + dsize - 1;
        isSeed = true;
        isRepeat |= ((float)S.freq >= repeatThreshold);
        maxFixedMerFreq = max(maxFixedMerFreq, S.freq);
    }
    SeedCand c;
    c.next = (int32_t)initPos; c.max_fixed_freq = maxFixedMerFreq; c.emit = (isSeed && !low_complexity(dcount, dsize)) ? 1 : 0;
    c.len = (uint8_t)dsize; c.is_repeat = isRepeat ? 1 : 0; c.static_k = (uint8_t)staticSize; c.pad = 0;
    cand[g] = c;
    ntriv[g] = (c.emit || initPos != p0) ? 1 : 0;
}

// warp per read: follow initPos -> f(initPos) + 1, 32 trivial positions per step
__global__ void __launch_bounds__(128)
seed_chain_kernel(SeedParamsDev P, const uint64_t* __restrict__ offsets, uint64_t n_reads, const SeedCand* __restrict__ cand,
                  const uint8_t* __restrict__ ntriv, pbsc_seed* __restrict__ seeds, const uint64_t* __restrict__ region,
                  uint32_t* __restrict__ seed_count)
{
    const uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    const uint64_t base = offsets[r];
    const int64_t L = (int64_t)(offsets[r + 1] - base);
    pbsc_seed* out = seeds + region[r];
    uint32_t ns = 0;
    int64_t p = 0;
    if ((int)L < P.start_kmer) p = L;
    while (p < L)
    {
        const int64_t q = p + lane;
        const unsigned m = __ballot_sync(0xffffffffu, q < L && ntriv[base + q] != 0);
        if (!m) { p += 32; continue; }
        const int64_t e = p + (__ffs(m) - 1);
        const SeedCand c = cand[base + e];
        if (c.emit)
        {
            if (lane == 0)
            {
                pbsc_seed sd;
                sd.start = (int32_t)e; sd.len = c.len; sd.max_fixed_freq = c.max_fixed_freq; sd.is_repeat = c.is_repeat;
                sd.start_best_k = c.static_k; sd.end_best_k = c.static_k; sd.hitchhiked = 0; sd.static_k = c.static_k;
                out[ns] = sd;
            }
            ns++;
        }
        p = (int64_t)c.next + 1;
    }
    if (lane == 0) seed_count[r] = ns;
}

// countSequenceOccurrences(w, pSelBWT) (BWTAlgorithms.cpp:135-141) for the two poles of a seed.
// pole == 1 (start): pSelBWT = RBWT, w = reverse(P), P = first ks bases of the seed
//      findInterval(RBWT, w) processes P[0], P[1], ...;  findInterval(RBWT, revcomp(w)=comp(P)) processes comp(P[ks-1]), comp(P[ks-2]), ...
// pole == 0 (end):   pSelBWT = BWT,  w = S = last ks bases of the seed
//      findInterval(BWT, S) processes S[ks-1], S[ks-2], ...;  findInterval(BWT, revcomp(S)) processes comp(S[0]), comp(S[1]), ...
__device__ __forceinline__ int pole_freq(const FmTable& t, const uint8_t* seed, int seedLen, int ks, int pole)
{
    const uint8_t* p = pole ? seed : seed + (seedLen - ks);
    Interval a, b;
    if (pole)
    {
        a = init_interval(t, p[0]);
        for (int j = 1; j < ks; j++) { a = update_interval(t, a, p[j]); if (!a.valid()) break; }
        b = init_interval(t, 3 - p[ks - 1]);
        for (int j = ks - 2; j >= 0; j--) { b = update_interval(t, b, 3 - p[j]); if (!b.valid()) break; }
    }
    else
    {
        a = init_interval(t, p[ks - 1]);
        for (int j = ks - 2; j >= 0; j--) { a = update_interval(t, a, p[j]); if (!a.valid()) break; }
        b = init_interval(t, 3 - p[0]);
        for (int j = 1; j < ks; j++) { b = update_interval(t, b, 3 - p[j]); if (!b.valid()) break; }
    }
    return (int)((a.valid() ? (int64_t)a.size() : 0) + (b.valid() ? (int64_t)b.size() : 0));
}

// SeedFeature::modifyKmerSize (SeedFeature.cpp:48-78), one thread per (seed slot, pole)
__global__ void __launch_bounds__(128)
seed_bestk_kernel(FmIndexDev idx, SeedParamsDev P, const uint8_t* __restrict__ codes, const uint64_t* __restrict__ offsets,
                  uint64_t n_reads, pbsc_seed* __restrict__ seeds, const uint64_t* __restrict__ region,
                  const uint32_t* __restrict__ seed_count, uint64_t total_slots)
{
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= total_slots * 2) return;
    const uint64_t slot = t >> 1;
    const int pole = (int)(t & 1);
    const uint64_t r = find_read(region, n_reads, slot);
    if (slot - region[r] >= seed_count[r]) return;
    pbsc_seed& sd = seeds[slot];
    const uint8_t* seed = codes + offsets[r] + sd.start;
    const int seedLen = sd.len;
    const FmTable& tab = idx.t[pole ? PBSC_RBWT : PBSC_BWT];
    int kmerSize = sd.static_k;
    const int freqUpperBound = P.pb_coverage >> 1, freqLowerBound = P.pb_coverage >> 2;
    const int sizeUpperBound = seedLen, sizeLowerBound = sd.static_k;
    int kmerFreq = pole_freq(tab, seed, seedLen, kmerSize, pole);
    int bit = 0;
    if (kmerFreq > freqUpperBound) bit = 1;
    else if (kmerFreq < freqLowerBound) bit = -1;
    if (bit != 0)
    {
        const int freqBound = bit > 0 ? freqUpperBound : freqLowerBound;
        const int corsFreqBound = bit > 0 ? freqLowerBound : freqUpperBound;
        const int sizeBound = bit > 0 ? sizeUpperBound : sizeLowerBound;
        while ((bit ^ kmerFreq) > (bit ^ freqBound) && (bit ^ kmerSize) < (bit ^ sizeBound))
        {
            kmerSize += bit;
            kmerFreq = pole_freq(tab, seed, seedLen, kmerSize, pole);
        }
        if ((bit ^ kmerFreq) < (bit ^ corsFreqBound))
        {
            kmerSize -= bit;
            kmerFreq = pole_freq(tab, seed, seedLen, kmerSize, pole);
        }
    }
    if (pole) sd.start_best_k = kmerSize; else sd.end_best_k = kmerSize;
}

// removeHitchhikingSeeds (LongReadProbe.cpp:187-227); survivors are packed first, outcasts after them
__global__ void __launch_bounds__(128)
seed_hitchhike_kernel(SeedParamsDev P, uint64_t n_reads, pbsc_seed* __restrict__ in, pbsc_seed* __restrict__ out,
                      const uint64_t* __restrict__ region, uint32_t* __restrict__ seed_count, uint32_t* __restrict__ outcast_count)
{
    uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint32_t n = seed_count[r];
    pbsc_seed* a = in + region[r];
    pbsc_seed* o = out + region[r];
    outcast_count[r] = 0;
    if (n < 2) { for (uint32_t i = 0; i < n; i++) o[i] = a[i]; return; }
    const float inv_hh = __fdiv_rn(1.0f, P.hh_ratio);
    for (uint32_t q = 0; q + 1 < n; q++)
    {
        const int qEnd = a[q].start + a[q].len - 1;
        for (uint32_t sj = q + 1; sj < n; sj++)
        {
            if ((int)(a[sj].start - qEnd) > P.radius) break;
            const float freqDiff = __fdiv_rn((float)a[sj].max_fixed_freq, (float)a[q].max_fixed_freq);
            if (a[q].is_repeat && freqDiff < P.hh_ratio) a[sj].hitchhiked = 1;
            if (a[sj].is_repeat && freqDiff > inv_hh) a[q].hitchhiked = 1;
        }
    }
    uint32_t nsurv = 0;
    for (uint32_t i = 0; i < n; i++) if (!a[i].hitchhiked) nsurv++;
    uint32_t k = 0, m = nsurv;
    for (uint32_t i = 0; i < n; i++) { if (!a[i].hitchhiked) o[k++] = a[i]; else o[m++] = a[i]; }
    seed_count[r] = nsurv;
    outcast_count[r] = n - nsurv;
}

int upload_reads(pbsc_index* idx, const char* reads, const uint64_t* offsets, uint64_t n_reads, DeviceBatch& b, cudaStream_t st)
{
    const uint64_t n_bases = offsets[n_reads];
    for (uint64_t i = 0; i < n_reads; i++) if (offsets[i + 1] < offsets[i]) { set_error("read offsets must be non-decreasing"); return PBSC_ERR_ARG; }
    b.n_reads = n_reads;
    b.n_bases = n_bases;
    DevBuf<char> ascii;
    PBSC_CUDA(ascii.alloc(n_bases));
    PBSC_CUDA(b.codes.alloc(n_bases + 64));
    PBSC_CUDA(b.offsets.alloc(n_reads + 1));
    PBSC_CUDA(cudaMemcpyAsync(ascii.p, reads, n_bases, cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemcpyAsync(b.offsets.p, offsets, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaMemsetAsync(b.codes.p + n_bases, 0, 64, st));
    DevBuf<unsigned int> bad;
    PBSC_CUDA(bad.alloc(1));
    PBSC_CUDA(cudaMemsetAsync(bad.p, 0, 4, st));
    if (n_bases) ascii_to_codes_kernel<<<(unsigned)((n_bases / 16 + 256) / 256), 256, 0, st>>>(ascii.p, b.codes.p, n_bases, bad.p);
    PBSC_CUDA(cudaGetLastError());
    unsigned int hbad = 0;
    PBSC_CUDA(cudaMemcpyAsync(&hbad, bad.p, 4, cudaMemcpyDeviceToHost, st));
    PBSC_CUDA(cudaStreamSynchronize(st));
    if (hbad) { set_error("Error: read contains non-ACGT characters."); return PBSC_ERR_ARG; }   // SeqReader.cpp:118-125
    return PBSC_OK;
}

static void make_seed_params(const pbsc_params* p, SeedParamsDev& d)
{
    memset(&d, 0, sizeof d);
    for (int i = 0; i < 8; i++) d.pool[i] = p->pool[i];
    d.n_pool = p->n_pool;
    d.start_kmer = p->start_kmer; d.scan_kmer = p->scan_kmer; d.kmer_up_bound = p->kmer_up_bound; d.radius = p->radius;
    d.pb_coverage = p->pb_coverage; d.manual = p->manual; d.mode = p->mode; d.hh_ratio = p->hh_ratio;
    d.n_static = 0;
    for (int m = 0; m < 3; m++)
    {
        d.offset[m] = p->offset[m];
        int sz = p->start_kmer + p->offset[m];
        int sl = -1;
        for (int j = 0; j < d.n_static; j++) if (d.static_size[j] == sz) sl = j;
        if (sl < 0) { sl = d.n_static; d.static_size[d.n_static++] = sz; }
        d.slot_of_mode[m] = sl;
    }
    memcpy(d.threshold, p->threshold, sizeof d.threshold);
}

int alloc_seed_workspace(const pbsc_params* p, const std::vector<uint64_t>& off, DeviceBatch& b, SeedBuffers& s, Workspace& w, cudaStream_t st)
{
    SeedParamsDev P;
    make_seed_params(p, P);
    // per-read seed regions: seeds never overlap and are at least min(static size) long
    int min_static = P.static_size[0];
    for (int j = 1; j < P.n_static; j++) min_static = std::min(min_static, P.static_size[j]);
    if (min_static < 1) { set_error("static k-mer size must be positive"); return PBSC_ERR_ARG; }
    std::vector<uint64_t> region(b.n_reads + 1);
    region[0] = 0;
    for (uint64_t r = 0; r < b.n_reads; r++) region[r + 1] = region[r] + (off[r + 1] - off[r]) / (uint64_t)min_static + 2;
    s.total_slots = region[b.n_reads];
    PBSC_CUDA(w.feats.alloc((uint64_t)P.n_static * b.n_bases));
    PBSC_CUDA(w.cls.alloc(b.n_bases)); PBSC_CUDA(w.attr.alloc(b.n_bases));
    PBSC_CUDA(w.prefix.alloc(b.n_bases)); PBSC_CUDA(w.cand.alloc(b.n_bases * 2)); PBSC_CUDA(w.ntriv.alloc(b.n_bases));
    PBSC_CUDA(w.seed_tmp.alloc(s.total_slots)); PBSC_CUDA(s.seeds.alloc(s.total_slots));
    PBSC_CUDA(s.region.alloc(b.n_reads + 1)); PBSC_CUDA(s.count.alloc(b.n_reads)); PBSC_CUDA(s.outcast.alloc(b.n_reads));
    PBSC_CUDA(cudaMemcpyAsync(s.region.p, region.data(), (b.n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    PBSC_CUDA(cudaStreamSynchronize(st));
    return PBSC_OK;
}

int run_seed_phase(pbsc_index* idx, const pbsc_params* p, DeviceBatch& b, SeedBuffers& s, Workspace& w, uint64_t* launches)
{
    SeedParamsDev P;
    make_seed_params(p, P);
    cudaStream_t st = idx->stream;
    uint64_t nl = 0;
    if (b.n_bases)
    {
        seed_features_kernel<<<(unsigned)((b.n_bases + 255) / 256), 256, 0, st>>>(idx->dev, P, b.codes.p, b.offsets.p, b.n_reads, b.n_bases, w.feats.p, w.cls.p);
        nl++;
    }
    if (b.n_reads)
    {
        const unsigned warp_blocks = (unsigned)((b.n_reads * 32 + 127) / 128);
        const unsigned base_blocks = (unsigned)((b.n_bases + 255) / 256);
        seed_prefix_kernel<<<warp_blocks, 128, 0, st>>>(b.offsets.p, b.n_reads, w.cls.p, (uint4*)w.prefix.p);
        seed_attr_kernel<<<base_blocks, 256, 0, st>>>(P, b.offsets.p, b.n_reads, b.n_bases, (const uint4*)w.prefix.p, w.attr.p, w.dbg_ratio.p);
        seed_candidates_kernel<<<base_blocks, 256, 0, st>>>(idx->dev, P, b.codes.p, b.offsets.p, b.n_reads, b.n_bases, w.feats.p, w.attr.p,
                                                            (SeedCand*)w.cand.p, w.ntriv.p);
        seed_chain_kernel<<<warp_blocks, 128, 0, st>>>(P, b.offsets.p, b.n_reads, (const SeedCand*)w.cand.p, w.ntriv.p, w.seed_tmp.p, s.region.p, s.count.p);
        nl += 3;
        seed_bestk_kernel<<<(unsigned)((s.total_slots * 2 + 127) / 128), 128, 0, st>>>(idx->dev, P, b.codes.p, b.offsets.p, b.n_reads, w.seed_tmp.p, s.region.p, s.count.p, s.total_slots);
        seed_hitchhike_kernel<<<(unsigned)((b.n_reads + 127) / 128), 128, 0, st>>>(P, b.n_reads, w.seed_tmp.p, s.seeds.p, s.region.p, s.count.p, s.outcast.p);
        nl += 3;
    }
    PBSC_CUDA(cudaGetLastError());
    PBSC_OCC_TAKE(0, st);
    if (launches) *launches += nl;
    return PBSC_OK;
}

}  // namespace pbsc

using namespace pbsc;

extern "C" int pbsc_seed_batch(pbsc_index* idx, const pbsc_params* p, const char* reads, const uint64_t* offsets, uint64_t n_reads,
                               pbsc_seed* seeds_out, uint64_t seeds_cap, uint64_t* seed_offsets, uint64_t* seeds_needed, int keep_outcast)
try
{
    if (!idx || !p || !reads || !offsets || !seed_offsets) { set_error("pbsc_seed_batch: null argument"); return PBSC_ERR_ARG; }
    (void)keep_outcast;
    PBSC_CUDA(cudaSetDevice(idx->device));
    DeviceBatch b;
    SeedBuffers s;
    Workspace w;
    int rc = upload_reads(idx, reads, offsets, n_reads, b, idx->stream);
    if (rc != PBSC_OK) return rc;
    std::vector<uint64_t> h_off(offsets, offsets + n_reads + 1);
    rc = alloc_seed_workspace(p, h_off, b, s, w, idx->stream);
    if (rc != PBSC_OK) return rc;
    rc = run_seed_phase(idx, p, b, s, w, nullptr);
    if (rc != PBSC_OK) return rc;
    PBSC_CUDA(cudaStreamSynchronize(idx->stream));
    std::vector<uint32_t> cnt(n_reads);
    std::vector<uint64_t> region(n_reads + 1);
    if (n_reads)
    {
        PBSC_CUDA(cudaMemcpy(cnt.data(), s.count.p, n_reads * 4, cudaMemcpyDeviceToHost));
        PBSC_CUDA(cudaMemcpy(region.data(), s.region.p, (n_reads + 1) * 8, cudaMemcpyDeviceToHost));
    }
    seed_offsets[0] = 0;
    for (uint64_t r = 0; r < n_reads; r++) seed_offsets[r + 1] = seed_offsets[r] + cnt[r];
    if (seeds_needed) *seeds_needed = seed_offsets[n_reads];
    if (seed_offsets[n_reads] > seeds_cap || (!seeds_out && seed_offsets[n_reads])) { set_error("pbsc_seed_batch: %llu seeds do not fit in seeds_cap %llu", (unsigned long long)seed_offsets[n_reads], (unsigned long long)seeds_cap); return PBSC_ERR_LIMIT; }
    std::vector<pbsc_seed> all(s.total_slots);
    if (s.total_slots) PBSC_CUDA(cudaMemcpy(all.data(), s.seeds.p, s.total_slots * sizeof(pbsc_seed), cudaMemcpyDeviceToHost));
    for (uint64_t r = 0; r < n_reads; r++)
        for (uint32_t i = 0; i < cnt[r]; i++) seeds_out[seed_offsets[r] + i] = all[region[r] + i];
    return PBSC_OK;
}
PBSC_CATCH_ALL("pbsc_seed_batch")
