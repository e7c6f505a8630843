// =====================================================================================
// TEST INFRASTRUCTURE ONLY — CPU oracle for the `stride pbcorrect` hot path.
//
// A from-scratch restatement, in plain sequential C++, of the algorithm the reference
// (ccuchengwei/LongReadSelfCorrect, mounted at /root/reference) runs for
// `stride pbcorrect` (the DP/MSA fallback lives in pbsc_oracle_dp.hpp).  Every function cites the
// reference file:line it follows.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build or execute anything under oracle/; the product
// (longreadselfcorrect_b200/) never includes, links or calls it.
//
// Why C++ and not C: the order in which the reference's IntervalTree returns hits
// depends on libstdc++'s std::sort permutation of equal keys (PacBio/IntervalTree.cpp:18),
// so the oracle calls the same std::sort with the same comparator.
//
// PARITY PINNED: tests/test_oracle_vs_ref.py compares this oracle with the reference
// binary itself (oracle/_ref/stride, built by oracle/build_ref.py from the unmodified
// sources) on correct.fa / discard.fa / threshold-table / seed dumps, and with
// oracle/_ref/fm_dump on raw findInterval intervals; tests/golden/ holds fixtures that
// the reference binary generated (tests/golden/make_golden.py).
// =====================================================================================
#ifndef PBSC_ORACLE_HPP
#define PBSC_ORACLE_HPP

#include <atomic>
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace pbo {

// ---------------------------------------------------------------------------------
// Sequence helpers — Util/Util.cpp:18-56, Util/Util.h:268-287
// ---------------------------------------------------------------------------------
inline char comp(char b)
{
    switch (b) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}
inline std::string reverse(const std::string& s) { return std::string(s.rbegin(), s.rend()); }
inline std::string complement(const std::string& s) { std::string o(s); for (auto& c : o) c = comp(c); return o; }
inline std::string reverseComplement(const std::string& s) { return complement(reverse(s)); }
// rank in the BWT alphabet "$ACGT" (Util/Alphabet.h:39,87-104)
inline int bwtRank(char b) { switch (b) { case 'A': return 1; case 'C': return 2; case 'G': return 3; case 'T': return 4; default: return 0; } }
// index in the DNA alphabet "ACGT" (Util/Alphabet.cpp:15)
inline int dnaIdx(char b) { return bwtRank(b) - 1; }

// ---------------------------------------------------------------------------------
// FM-index: on-disk run-length BWT (SuffixTools/BWTReaderBinary.cpp:55-85,
// RLUnit.h:13-16,118-143) expanded to one symbol per byte with a checkpoint every 64
// symbols.  occ() has the semantics of RLBWT::getOcc (RLBWT.h:121-140): occurrences of
// b in bwt[0..idx] INCLUSIVE, and idx == size_t(-1) yields 0.
// ---------------------------------------------------------------------------------
struct FMIndex
{
    uint64_t numStrings = 0, numSymbols = 0, numRuns = 0;
    std::vector<uint8_t> sym;          // ranks 0..4 ($ACGT)
    std::vector<uint64_t> ckpt;        // 5 counts per 64-symbol block (exclusive prefix)
    uint64_t C[5] = {0, 0, 0, 0, 0};   // RLBWT.cpp:243-247
    uint64_t total[5] = {0, 0, 0, 0, 0};
    std::vector<uint8_t> rl;           // raw run bytes (kept for round-trip tests)

    bool load(const std::string& path, std::string* err = nullptr)
    {
        std::ifstream in(path.c_str(), std::ios::binary);
        if (!in) { if (err) *err = "cannot open " + path; return false; }
        uint16_t magic = 0; int32_t flag = 0;
        in.read((char*)&magic, 2);
        if (magic != 0xCACA) { if (err) *err = "BWT file is not properly formatted, aborting"; return false; }
        in.read((char*)&numStrings, 8); in.read((char*)&numSymbols, 8); in.read((char*)&numRuns, 8); in.read((char*)&flag, 4);
        rl.resize(numRuns);
        in.read((char*)rl.data(), numRuns);
        if ((uint64_t)in.gcount() != numRuns) { if (err) *err = "truncated BWT file " + path; return false; }
        build();
        return true;
    }
    void fromRuns(const std::vector<uint8_t>& runs, uint64_t nStrings)
    {
        rl = runs; numRuns = runs.size(); numStrings = nStrings; numSymbols = 0;
        for (uint8_t u : rl) numSymbols += (u & 0x1F);
        build();
    }
    void build()
    {
        sym.clear(); sym.reserve(numSymbols);
        for (uint8_t u : rl) { uint8_t s = u >> 5, n = u & 0x1F; for (uint8_t i = 0; i < n; i++) sym.push_back(s); }
        assert(sym.size() == numSymbols);
        size_t nb = sym.size() / 64 + 1;
        ckpt.assign(nb * 5, 0);
        uint64_t run[5] = {0, 0, 0, 0, 0};
        for (size_t i = 0; i < sym.size(); i++)
        {
            if ((i & 63) == 0) for (int c = 0; c < 5; c++) ckpt[(i >> 6) * 5 + c] = run[c];
            run[sym[i]]++;
        }
        if ((sym.size() & 63) == 0) for (int c = 0; c < 5; c++) ckpt[(sym.size() >> 6) * 5 + c] = run[c];
        for (int c = 0; c < 5; c++) total[c] = run[c];
        C[0] = 0; C[1] = run[0]; C[2] = C[1] + run[1]; C[3] = C[2] + run[2]; C[4] = C[3] + run[3];
    }
    // RLBWT::getOcc(b, idx) — RLBWT.h:121-140
    inline uint64_t occ(int rank, uint64_t idx) const
    {
        uint64_t p = idx + 1;               // exclusive end; wraps to 0 for idx == -1
        if (p > numSymbols) p = numSymbols;
        uint64_t blk = p >> 6;
        uint64_t r = ckpt[blk * 5 + rank];
        for (uint64_t i = blk << 6; i < p; i++) r += (sym[i] == rank);
        return r;
    }
    inline uint64_t getPC(int rank) const { return C[rank]; }
    inline uint64_t getBWLen() const { return numSymbols; }
    // RLBWT::getChar — RLBWT.h:42-63
    inline char getChar(uint64_t idx) const { return "$ACGT"[sym[idx]]; }
};

// SuffixTools/BWTInterval.h:19-100
struct BWTInterval
{
    int64_t lower = 0, upper = 0;
    BWTInterval() {}
    BWTInterval(int64_t l, int64_t u) : lower(l), upper(u) {}
    inline bool isValid() const { return lower <= upper; }
    inline int64_t size() const { return upper - lower + 1; }
    inline int64_t getFreq() const { return isValid() ? size() : 0; }
};
struct BiBWTInterval
{
    BWTInterval fwdInterval, rvcInterval;
    inline bool isValid() const { return fwdInterval.isValid() && rvcInterval.isValid(); }
    inline int64_t getFreq() const { return fwdInterval.getFreq() + rvcInterval.getFreq(); }
};

struct IndexSet { const FMIndex* pBWT = nullptr; const FMIndex* pRBWT = nullptr; };

// rank-query counter for the roofline numerator
struct OccCounter { static uint64_t& n() { static thread_local uint64_t c = 0; return c; } };
// PBSC_ORACLE_REFINE_STATS=1 (analysis aid): how many refineSAInterval re-searches there are, how many of them look
// up a k-mer that occurs at least twice (both strands together), and how many steps they take; process-wide atomics
struct RefineStats
{
    static bool on() { static const bool v = getenv("PBSC_ORACLE_REFINE_STATS") != nullptr; return v; }
    static std::atomic<unsigned long long>& calls() { static std::atomic<unsigned long long> c{0}; return c; }
    static std::atomic<unsigned long long>& twice() { static std::atomic<unsigned long long> c{0}; return c; }
    static std::atomic<unsigned long long>& bases() { static std::atomic<unsigned long long> c{0}; return c; }
};

// BWTAlgorithms::updateInterval — SuffixTools/BWTAlgorithms.h:66-72
inline void updateInterval(BWTInterval& iv, char b, const FMIndex* bwt, int* count = nullptr)
{
    if (count != nullptr) count[dnaIdx(b)]++;
    int r = bwtRank(b);
    uint64_t pb = bwt->getPC(r);
    OccCounter::n() += 2;
    iv.lower = pb + bwt->occ(r, (uint64_t)(iv.lower - 1));
    iv.upper = pb + bwt->occ(r, (uint64_t)iv.upper) - 1;
}
// BWTAlgorithms::initInterval — BWTAlgorithms.h:136-140
inline void initInterval(BWTInterval& iv, char b, const FMIndex* bwt)
{
    int r = bwtRank(b);
    OccCounter::n() += 1;
    iv.lower = bwt->getPC(r);
    iv.upper = iv.lower + bwt->occ(r, bwt->getBWLen() - 1) - 1;
}
// BWTAlgorithms::findInterval — BWTAlgorithms.cpp:14-31 (breaks on the first invalid interval)
inline BWTInterval findInterval(const FMIndex* bwt, const std::string& w, int* count = nullptr)
{
    int len = w.size();
    int j = len - 1;
    char curr = w[j];
    if (count != nullptr) count[dnaIdx(curr)]++;
    BWTInterval iv;
    initInterval(iv, curr, bwt);
    --j;
    for (; j >= 0; --j)
    {
        curr = w[j];
        updateInterval(iv, curr, bwt, count);
        if (!iv.isValid()) break;
    }
    return iv;
}
// BWTAlgorithms::findBiInterval — BWTAlgorithms.cpp:32-38
inline BiBWTInterval findBiInterval(const IndexSet& idx, const std::string& w, int* count = nullptr)
{
    BiBWTInterval bi;
    bi.fwdInterval = findInterval(idx.pRBWT, reverse(w), count);
    bi.rvcInterval = findInterval(idx.pBWT, reverseComplement(w));
    return bi;
}
// BWTAlgorithms::updateBiInterval — BWTAlgorithms.h:73-77
inline void updateBiInterval(BiBWTInterval& bi, char b, const IndexSet& idx, int* count = nullptr)
{
    updateInterval(bi.fwdInterval, b, idx.pRBWT, count);
    updateInterval(bi.rvcInterval, comp(b), idx.pBWT);
}
// BWTAlgorithms::countSequenceOccurrences(w, pBWT) — BWTAlgorithms.cpp:135-141
inline size_t countSequenceOccurrences(const std::string& w, const FMIndex* bwt)
{
    BiBWTInterval bi;
    bi.fwdInterval = findInterval(bwt, w);
    bi.rvcInterval = findInterval(bwt, reverseComplement(w));
    return bi.getFreq();
}

// ---------------------------------------------------------------------------------
// KmerThreshold — PacBio/KmerThreshold.cpp:11-79 (float arithmetic, source order)
// ---------------------------------------------------------------------------------
struct KmerThreshold
{
    int start = 15, end = 50, cov = 0;
    std::vector<float> table[3];
    void initialize(int s, int e, int c)
    {
        static const float formula[3][6] = {
            {0.0004799107143, -0.008037815126, 0.03673552754, 0.1850695903, -1.572552521, 18.0522088},
            {0.0003348214286, -0.009112394958, 0.04286714686, 0.240519958, -1.8793367350, 21.29319228},
            {0.01714285714, -0.6193907563, 2.266956783, 17.28450630, -100.6983493, 1103.571729}};
        start = std::max(s, 15); end = e; cov = c;
        for (int mode = 0; mode <= 2; mode++)
        {
            table[mode].assign(end + 2, 0.0f);
            float cavity = std::numeric_limits<float>::max();
            for (int ksize = start; ksize <= end; ksize++)
            {
                const float* f = formula[mode];
                int x = cov, y = ksize;
                float v = f[0]*x*x + f[1]*x*y + f[2]*y*y + f[3]*x + f[4]*y + f[5];
                cavity = std::fmin(cavity, std::fmax(v, 2.0f));
                table[mode][ksize] = cavity;
            }
        }
    }
    inline float get(int mode, int ksize) const { return table[mode][ksize]; }
    // KmerThreshold::~KmerThreshold + write — KmerThreshold.cpp:31-41,65-72
    void write(std::ostream& out) const
    {
        out << "Coverage : " << cov << "\n" << "size\tlowcov\tunique\trepeat\n";
        for (int k = start; k <= end; k++)
            out << k << "\t" << table[0][k] << "\t" << table[1][k] << "\t" << table[2][k] << "\n";
    }
};

// ---------------------------------------------------------------------------------
// Parameters — StriDe/PacBioSelfCorrection.cpp:71-101,195-231
// ---------------------------------------------------------------------------------
struct Params
{
    int PBcoverage = 90;
    double ErrorRate = 0.15;
    int startKmerLen = 19;
    int nextTarget = 1;
    int maxLeaves = 32;
    int idmerLen = 9;
    int minKmerLen = 13;
    int genome = 10;
    int mode = 1;
    bool Manual = false, Adjust = false, Split = false, NoDp = false, DebugSeed = false;
    std::array<int, 3> offset = {{0, 0, 0}};
    std::set<int> pool = {5, 9, 19};
    int scanKmerLen = 19, kmerLenUpBound = 50, radius = 100;
    float hhRatio = 0.6;
    std::string directory;
    IndexSet indices;
    KmerThreshold thr;

    // PacBioSelfCorrection.cpp:195-206,231
    void derive()
    {
        static const int size[3] = {17, 19, 21};
        std::map<int, int> order = {{5, 0}, {10, 1}, {100, 2}};
        if (!Adjust)
        {
            startKmerLen = size[order[genome]];
            offset[1] = 2 * std::min(std::max((PBcoverage / 30 - 1), 0), (order[genome] + 1));
            offset[2] = -2 * (order[genome] + 1);
        }
        for (auto& o : offset) pool.insert(startKmerLen + o);
        thr.initialize(-1, 50, PBcoverage);
    }
};

// ---------------------------------------------------------------------------------
// KmerFeature — PacBio/KmerFeature.h:37-136
// ---------------------------------------------------------------------------------
struct KmerFeature
{
    int count[4] = {0, 0, 0, 0};
    std::string word;
    int size = 0;
    BiBWTInterval biInterval;
    bool fake = false;
    int frequency = 0;

    KmerFeature() {}
    KmerFeature(const IndexSet& idx, const std::string& seq, size_t pos, int len, const KmerFeature* base)
    {
        if (base == nullptr)
        {
            word = seq.substr(pos, len);
            size = word.length();
            biInterval = findBiInterval(idx, word, count);
        }
        else
        {
            *this = *base;
            for (size_t i = (pos + this->size); (i < seq.length()) && (this->size < len); i++) expand(seq[i], idx);
        }
        fake = (len != size);
        frequency = biInterval.getFreq();
    }
    inline int getSize() const { return size; }
    inline int getFreq() const { return fake ? -1 : frequency; }
    inline void expand(char b, const IndexSet& idx)
    {
        size++;
        word += b;
        updateBiInterval(biInterval, b, idx, count);
        frequency = biInterval.getFreq();
    }
    inline void shrink(int len)   // update == false on every pbcorrect call site (LongReadProbe.cpp:74,83)
    {
        size -= len;
        for (size_t i = size; i < word.size(); i++) count[dnaIdx(word[i])]--;
        word.erase(size, len);
    }
    inline bool isFake() const { return fake; }
    inline bool isValid() const { return biInterval.isValid(); }
    inline bool isLowComplexity(float m = 0.7, float d = 0.9) const
    {
        int copy[4] = {count[0], count[1], count[2], count[3]};
        std::sort(copy, copy + 4);
        bool isMonmer = (float)copy[3] / this->size >= m;
        bool isDimer = (float)(copy[2] + copy[3]) / this->size >= d;
        return isMonmer || isDimer;
    }
};

// ---------------------------------------------------------------------------------
// SeedFeature — PacBio/SeedFeature.h:22-45, SeedFeature.cpp:22-78
// ---------------------------------------------------------------------------------
struct SeedFeature
{
    std::string seedStr;
    int seedLen = 0, seedStartPos = 0, seedEndPos = 0, maxFixedMerFreq = 0;
    bool isRepeat = false, isHitchhiked = false;
    int startBestKmerSize = 0, endBestKmerSize = 0, startKmerFreq = 0, endKmerFreq = 0;
    int sizeUpperBound = 0, sizeLowerBound = 0, freqUpperBound = 0, freqLowerBound = 0;

    SeedFeature() {}
    SeedFeature(std::string str, int startPos, int frequency, bool repeat, int kmerSize, int PBcoverage)
        : seedStr(str), seedLen(seedStr.length()), seedStartPos(startPos), seedEndPos(startPos + seedLen - 1),
          maxFixedMerFreq(frequency), isRepeat(repeat), isHitchhiked(false), startBestKmerSize(kmerSize),
          endBestKmerSize(kmerSize), sizeUpperBound(seedLen), sizeLowerBound(kmerSize),
          freqUpperBound(PBcoverage >> 1), freqLowerBound(PBcoverage >> 2) {}

    void estimateBestKmerSize(const IndexSet& idx) { modifyKmerSize(idx, true); modifyKmerSize(idx, false); }
    void modifyKmerSize(const IndexSet& idx, bool pole)
    {
        int& kmerSize = pole ? startBestKmerSize : endBestKmerSize;
        int& kmerFreq = pole ? startKmerFreq : endKmerFreq;
        const FMIndex* const pSelBWT = pole ? idx.pRBWT : idx.pBWT;
        std::string seed = pole ? reverse(seedStr) : seedStr;
        kmerFreq = countSequenceOccurrences(seed.substr(seedLen - kmerSize), pSelBWT);
        int bit;
        if (kmerFreq > freqUpperBound) bit = 1;
        else if (kmerFreq < freqLowerBound) bit = -1;
        else return;
        const int freqBound = bit > 0 ? freqUpperBound : freqLowerBound;
        const int corsFreqBound = bit > 0 ? freqLowerBound : freqUpperBound;
        const int sizeBound = bit > 0 ? sizeUpperBound : sizeLowerBound;
        while ((bit ^ kmerFreq) > (bit ^ freqBound) && (bit ^ kmerSize) < (bit ^ sizeBound))
        {
            kmerSize += bit;
            kmerFreq = countSequenceOccurrences(seed.substr(seedLen - kmerSize), pSelBWT);
        }
        if ((bit ^ kmerFreq) < (bit ^ corsFreqBound))
        {
            kmerSize -= bit;
            kmerFreq = countSequenceOccurrences(seed.substr(seedLen - kmerSize), pSelBWT);
        }
    }
    // SeedFeature.h:22-33
    inline void append(std::string extendedStr, const SeedFeature& target)
    {
        seedStr += extendedStr;
        seedLen += extendedStr.length();
        startBestKmerSize = target.startBestKmerSize;
        endBestKmerSize = target.endBestKmerSize;
        isRepeat = target.isRepeat;
        maxFixedMerFreq = target.maxFixedMerFreq;
        seedStartPos = target.seedStartPos;
        seedEndPos = target.seedEndPos;
    }
};
typedef std::vector<SeedFeature> SeedVector;

// per-position record the GPU seed kernel is checked against
struct KmerRecord { int freq; int count[4]; int64_t fwdLower, fwdUpper, rvcLower, rvcUpper; bool fake, valid; };

// ---------------------------------------------------------------------------------
// LongReadProbe — PacBio/LongReadProbe.cpp:34-227
// ---------------------------------------------------------------------------------
struct LongReadProbe
{
    const Params& P;
    std::map<int, std::vector<KmerFeature>> Log;   // KmerFeature::Log()
    std::vector<float> ratioLog;                   // extend/<id>.log values
    SeedVector outcast;                            // seed/error/<id>.seed
    bool outcastWritten = false;                   // the reference only writes it past the size<2 early return
    explicit LongReadProbe(const Params& p) : P(p) {}

    // LongReadProbe.cpp:120-182
    void getSeqAttribute(const std::string& seq, int* const attribute)
    {
        const size_t seqLen = seq.length();
        std::fill_n(attribute, seqLen, 1);
        ratioLog.assign(seqLen, 0.0f);
        int range = 300;
        const int ksize = P.scanKmerLen;
        float repeatValue = P.thr.get(2, ksize);
        int front = 0, fear = -1;
        std::map<int, int> box;
        for (size_t pos = 0; pos < seqLen; pos++)
        {
            int left = pos - (range >> 1);
            int right = pos + (range >> 1);
            left = std::max(left, 0);
            right = std::min(right, (int)(seqLen - 1));
            while (fear < right)
            {
                fear++;
                KmerFeature* prev = nullptr;
                for (auto& iter : P.pool)
                {
                    Log[iter][fear] = KmerFeature(P.indices, seq, fear, iter, prev);
                    prev = &Log[iter][fear];
                }
                const KmerFeature& inKmer = Log[ksize][fear];
                int freq = inKmer.isLowComplexity() ? -1 : inKmer.getFreq();
                int mode;
                if (freq < 0) mode = -1;
                else if (freq >= repeatValue) mode = 2;
                else mode = 1;
                box[mode]++;
            }
            while (front < left)
            {
                const KmerFeature& outKmer = Log[ksize][front];
                front++;
                int freq = outKmer.isLowComplexity() ? -1 : outKmer.getFreq();
                int mode;
                if (freq <= 0) mode = -1;
                else if (freq >= repeatValue) mode = 2;
                else mode = 1;
                box[mode]--;
            }
            int size = (right - left + 1) - box[-1];
            float ratio = (float)box[2] / size + 0.0005;
            ratioLog[pos] = ratio;
            if (ratio >= 0.02) attribute[pos] = 2;
        }
    }

    // LongReadProbe.cpp:34-117
    void searchSeedsWithHybridKmers(const std::string& readSeq, SeedVector& seedVec, std::vector<int>* attrOut = nullptr)
    {
        const size_t readSeqLen = readSeq.length();
        int staticSize = P.startKmerLen;
        if ((int)readSeqLen < staticSize) return;
        for (auto& k : P.pool) Log[k].assign(readSeqLen, KmerFeature());   // PacBioSelfCorrectionProcess.cpp:32-33
        std::vector<int> attribute(readSeqLen);
        getSeqAttribute(readSeq, attribute.data());
        if (P.Manual) std::fill_n(attribute.begin(), readSeqLen, P.mode);
        if (attrOut) *attrOut = attribute;

        for (size_t initPos = 0; initPos < readSeqLen; initPos++)
        {
            int dynamicMode = attribute[initPos];
            staticSize += P.offset[dynamicMode];
            KmerFeature dynamicKmer = Log[staticSize][initPos];
            bool isSeed = false, isRepeat = false;
            int maxFixedMerFreq = dynamicKmer.getFreq();
            size_t seedPos = initPos;
            for (size_t currPos = initPos; currPos < readSeqLen; currPos++)
            {
                int staticMode = attribute[currPos];
                const KmerFeature& staticKmer = Log[staticSize][currPos];
                if (staticKmer.isFake()) break;
                if (isSeed)
                {
                    char b = readSeq[(currPos + staticSize - 1)];
                    dynamicKmer.expand(b, P.indices);
                }
                float dynamicThreshold = P.thr.get(dynamicMode, dynamicKmer.getSize());
                float staticThreshold = P.thr.get(staticMode, staticKmer.getSize());
                float repeatThreshold = (5 - ((staticMode >> 1) << 2)) * staticThreshold;
                if (staticKmer.getFreq() < staticThreshold
                    || dynamicKmer.getFreq() < dynamicThreshold
                    || !dynamicKmer.isValid()
                    || dynamicKmer.getSize() > P.kmerLenUpBound)
                {
                    if (isSeed) dynamicKmer.shrink(1);
                    break;
                }
                float freqDiff = (float)staticKmer.getFreq() / maxFixedMerFreq;
                if (freqDiff < P.hhRatio)
                {
                    initPos++;
                    dynamicKmer.shrink(1);
                    break;
                }
                else if (freqDiff > 1 / P.hhRatio)
                {
                    initPos = currPos - 1;
                    isSeed = false;
                    break;
                }
                initPos = seedPos + dynamicKmer.getSize() - 1;
                isSeed = true;
                isRepeat |= (staticKmer.getFreq() >= repeatThreshold);
                maxFixedMerFreq = std::max(maxFixedMerFreq, staticKmer.getFreq());
            }
            if (isSeed && !dynamicKmer.isLowComplexity())
            {
                seedVec.push_back(SeedFeature(dynamicKmer.word, seedPos, maxFixedMerFreq, isRepeat, staticSize, P.PBcoverage));
                seedVec.back().estimateBestKmerSize(P.indices);
            }
            staticSize -= P.offset[dynamicMode];
        }
        seedVec = removeHitchhikingSeeds(seedVec);
    }

    // LongReadProbe.cpp:187-227
    SeedVector removeHitchhikingSeeds(SeedVector initSeedVec)
    {
        outcast.clear();
        outcastWritten = false;
        if (initSeedVec.size() < 2) return initSeedVec;
        outcastWritten = true;
        for (SeedVector::iterator iterQuery = initSeedVec.begin(); (iterQuery + 1) != initSeedVec.end(); iterQuery++)
        {
            SeedFeature& query = *iterQuery;
            SeedVector::iterator iterSubject = iterQuery + 1;
            for (; iterSubject != initSeedVec.end(); iterSubject++)
            {
                SeedFeature& subject = *iterSubject;
                if ((int)(subject.seedStartPos - query.seedEndPos) > P.radius) break;
                float freqDiff = (float)subject.maxFixedMerFreq / query.maxFixedMerFreq;
                subject.isHitchhiked |= (query.isRepeat && freqDiff < P.hhRatio);
                query.isHitchhiked |= (subject.isRepeat && freqDiff > 1 / P.hhRatio);
            }
        }
        SeedVector finalSeedVec;
        for (const auto& iter : initSeedVec)
        {
            if (iter.isHitchhiked) outcast.push_back(iter);
            else finalSeedVec.push_back(iter);
        }
        return finalSeedVec;
    }
};

// ---------------------------------------------------------------------------------
// IntervalTree<size_t> — PacBio/IntervalTree.h:10-70, IntervalTree.cpp:5-92
// ---------------------------------------------------------------------------------
struct TreeInterval
{
    size_t start, stop, value;
    TreeInterval(size_t s, size_t e, size_t v) : start(s), stop(e), value(v) {}
    friend bool operator<(const TreeInterval& a, const TreeInterval& b) { return a.stop < b.stop; }
    friend bool operator>(const TreeInterval& a, const TreeInterval& b) { return a.start > b.start; }
};
struct IntervalTree
{
    typedef std::vector<TreeInterval> intervalVector;
    intervalVector intervals;
    std::unique_ptr<IntervalTree> left, right;
    size_t center = 0;
    IntervalTree() {}
    IntervalTree(intervalVector& ivals, size_t depth = 16, size_t minbucket = 8, size_t leftextent = 0,
                 size_t rightextent = 0)
    {
        size_t leftp = leftextent, rightp = rightextent, centerp = 0;
        if (leftp == 0 && rightp == 0) std::sort(ivals.begin(), ivals.end(), std::greater<TreeInterval>());
        if (--depth == 0 || ivals.size() < minbucket) intervals = ivals;
        else
        {
            leftp = ivals.back().start;
            rightp = std::max_element(ivals.begin(), ivals.end())->stop;
            centerp = ivals[ivals.size() >> 1].start;
            center = centerp;
            intervalVector lefts, rights;
            for (const auto& iv : ivals)
            {
                if (iv.stop < center) lefts.push_back(iv);
                else if (iv.start > center) rights.push_back(iv);
                else intervals.push_back(iv);
            }
            if (!lefts.empty()) left.reset(new IntervalTree(lefts, depth, minbucket, leftp, centerp));
            if (!rights.empty()) right.reset(new IntervalTree(rights, depth, minbucket, centerp, rightp));
        }
    }
    void findOverlapping(size_t start, size_t stop, intervalVector& overlapping) const
    {
        if (!intervals.empty() && !(stop < intervals.back().start))
            for (const auto& iv : intervals)
                if (iv.start <= start && iv.stop >= stop) overlapping.push_back(iv);
        if (left && start < center) left->findOverlapping(start, stop, overlapping);
        if (right && stop > center) right->findOverlapping(start, stop, overlapping);
    }
};

// ---------------------------------------------------------------------------------
// LongReadSelfCorrectByOverlap — PacBio/LongReadCorrectByOverlap.{h,cpp},
// node state from FMIndexWalk/SAINode.h:301-354 and SAINode.cpp:166-189.
// Each live leaf owns exactly one node, so node and leafInfo are merged in `Leaf`
// and the label chain is kept as one string per leaf (copied on branching).
// ---------------------------------------------------------------------------------
struct Leaf
{
    // SAIOverlapNode3
    std::string str;
    BWTInterval fwdInterval, rvcInterval;
    size_t lastSeedIdx = 0;
    double numRedeemSeed = 0;
    size_t lastOverlapLen = 0, totalSeeds = 0, currOverlapLen = 0, numOfErrors = 0;
    int lastSeedIdxOffset = 0, initSeedIdx = 0;
    size_t queryOverlapLen = 0;
    std::pair<int, int> resultindex = std::make_pair(-1, -1);
    std::vector<double> LocalErrorRateRecord, GlobalErrorRateRecord;
    // leafInfo (LongReadCorrectByOverlap.h:157-208)
    int kmerFrequency = 0;
    char tailLetter = 0;
    size_t tailLetterCount = 0;

    std::string getSuffix(size_t l) const { assert(l <= str.size()); return str.substr(str.size() - l, l); }
};
struct FMidx
{
    std::string SearchLetters;
    BWTInterval fwdInterval, rvcInterval;
    int kmerFrequency;
    FMidx(const std::string& s, const BWTInterval& f, const BWTInterval& r)
        : SearchLetters(s), fwdInterval(f), rvcInterval(r), kmerFrequency(f.size() + r.size()) {}
    void setInterval(const BWTInterval& f, const BWTInterval& r) { fwdInterval = f; rvcInterval = r; kmerFrequency = f.size() + r.size(); }
};
struct WalkResultCand { std::string thread; double errorRate; };

struct WalkTrace   // optional per-level trace for debugging GPU/CPU divergences
{
    std::vector<int> leavesPerLevel;
    uint64_t levels = 0;
};

struct FMExtend
{
    const std::string m_sourceSeed, m_strBetweenSrcTarget, m_targetSeed;
    const int m_disBetweenSrcTarget;
    const size_t m_initkmersize, m_minOverlap, m_maxOverlap;
    const FMIndex* m_pBWT; const FMIndex* m_pRBWT;
    const size_t m_PBcoverage;
    size_t m_min_SA_threshold;
    double m_errorRate;
    const size_t m_maxLeaves, m_seedSize;
    size_t m_localSimilarlykmerSize;
    const double m_PacBioErrorRate;
    size_t m_maxIndelSize;
    double freqsOfKmerSize[101];
    std::string m_query;
    size_t m_maxLength, m_minLength;
    std::vector<BWTInterval> m_fwdTerminatedInterval, m_rvcTerminatedInterval;
    std::list<Leaf> m_leaves;
    size_t m_currentLength, m_currentKmerSize;
    IntervalTree m_fwdIntervalTree, m_rvcIntervalTree, m_fwdIntervalTree2, m_rvcIntervalTree2;
    WalkTrace* trace = nullptr;

    // LongReadCorrectByOverlap.cpp:17-95,106-124
    FMExtend(const std::string& sourceSeed, const std::string& strBetweenSrcTarget, const std::string& targetSeed,
             int disBetweenSrcTarget, size_t initkmersize, size_t maxOverlap, const Params& P, size_t min_SA_threshold,
             double errorRate = 0.25, size_t localSimilarlykmerSize = 100)
        : m_sourceSeed(sourceSeed), m_strBetweenSrcTarget(strBetweenSrcTarget), m_targetSeed(targetSeed),
          m_disBetweenSrcTarget(disBetweenSrcTarget), m_initkmersize(initkmersize), m_minOverlap(P.minKmerLen),
          m_maxOverlap(maxOverlap), m_pBWT(P.indices.pBWT), m_pRBWT(P.indices.pRBWT), m_PBcoverage(P.PBcoverage),
          m_min_SA_threshold(min_SA_threshold), m_errorRate(errorRate), m_maxLeaves(P.maxLeaves),
          m_seedSize(P.idmerLen), m_localSimilarlykmerSize(localSimilarlykmerSize), m_PacBioErrorRate(P.ErrorRate)
    {
        std::string beginningkmer = m_sourceSeed.substr(m_sourceSeed.length() - m_initkmersize);
        if (m_disBetweenSrcTarget > 100) m_maxIndelSize = m_disBetweenSrcTarget * 0.2;
        else m_maxIndelSize = 20;

        Leaf root;
        root.str = beginningkmer;
        root.fwdInterval = findInterval(m_pRBWT, reverse(beginningkmer));
        root.rvcInterval = findInterval(m_pBWT, reverseComplement(beginningkmer));
        root.lastOverlapLen = m_currentLength = root.currOverlapLen = root.queryOverlapLen = m_currentKmerSize = m_initkmersize;
        root.lastSeedIdx = root.initSeedIdx = m_initkmersize - m_seedSize;
        root.totalSeeds = m_initkmersize - m_seedSize + 1;
        root.numRedeemSeed = 0;
        root.LocalErrorRateRecord.push_back(0);
        root.GlobalErrorRateRecord.push_back(0);
        initLeafInfo(root);
        m_leaves.push_back(root);

        for (int i = 0; i <= 100; i++) freqsOfKmerSize[i] = 0;
        for (int i = m_minOverlap; i <= 100; i++) freqsOfKmerSize[i] = pow(1 - m_PacBioErrorRate, i) * m_PBcoverage;

        m_maxLength = (1.2 * (m_disBetweenSrcTarget + 10)) + 2 * m_initkmersize;
        m_minLength = (0.8 * (m_disBetweenSrcTarget - 20)) + 2 * m_initkmersize;

        for (size_t i = 0; i <= m_targetSeed.length() - m_minOverlap; i++)
        {
            std::string endingkmer = m_targetSeed.substr(i, m_minOverlap);
            m_fwdTerminatedInterval.push_back(findInterval(m_pRBWT, reverse(endingkmer)));
            m_rvcTerminatedInterval.push_back(findInterval(m_pBWT, reverseComplement(endingkmer)));
        }
        m_query = beginningkmer + m_strBetweenSrcTarget + m_targetSeed;
        buildOverlapbyFMindex(m_fwdIntervalTree, m_rvcIntervalTree, m_seedSize);
        buildOverlapbyFMindex(m_fwdIntervalTree2, m_rvcIntervalTree2, 5);
    }

    // leafInfo(SAIOverlapNode3*, lastLeafNum) — LongReadCorrectByOverlap.h:160-178
    static void initLeafInfo(Leaf& l)
    {
        l.tailLetterCount = 0;
        for (auto r = l.str.crbegin(); r != l.str.crend(); ++r)
        {
            if (r == l.str.crbegin()) l.tailLetter = *r;
            if (l.tailLetter == *r) l.tailLetterCount++;
            else break;
        }
        l.kmerFrequency = l.fwdInterval.size() + l.rvcInterval.size();
    }

    // LongReadCorrectByOverlap.cpp:127-152
    void buildOverlapbyFMindex(IntervalTree& fwdTree, IntervalTree& rvcTree, const int& overlapSize)
    {
        std::vector<TreeInterval> fwdIntervals, rvcIntervals;
        for (int i = 0; i <= (int)m_query.length() - (int)overlapSize; i++)
        {
            std::string seedStr = m_query.substr(i, overlapSize);
            BWTInterval bi = findInterval(m_pRBWT, reverse(seedStr));
            if (bi.isValid()) fwdIntervals.emplace_back(bi.lower, bi.upper, i);
            bi = findInterval(m_pBWT, reverseComplement(seedStr));
            if (bi.isValid()) rvcIntervals.emplace_back(bi.lower, bi.upper, i);
        }
        fwdTree = IntervalTree(fwdIntervals);
        rvcTree = IntervalTree(rvcIntervals);
    }

    // LongReadCorrectByOverlap.cpp:155-211
    int extendOverlap(std::string& mergedSeq)
    {
        std::vector<WalkResultCand> results;
        while (!m_leaves.empty() && m_leaves.size() <= m_maxLeaves && m_currentLength <= m_maxLength)
        {
            std::list<Leaf> newLeaves;
            extendLeaves(newLeaves);
            PrunedBySeedSupport(newLeaves);
            m_leaves.clear();
            m_leaves = newLeaves;
            if (trace) { trace->leavesPerLevel.push_back((int)m_leaves.size()); trace->levels++; }
            if (m_currentLength >= m_minLength) isTerminated(results);
        }
        if (results.size() > 0) return findTheBestPath(results, mergedSeq);
        if (m_leaves.empty()) return -1;
        else if (m_currentLength > m_maxLength) return -2;
        else if (m_leaves.size() > m_maxLeaves) return -3;
        else return -4;
    }

    // LongReadCorrectByOverlap.cpp:214-236
    int findTheBestPath(const std::vector<WalkResultCand>& results, std::string& mergedSeq)
    {
        double minErrorRate = 1;
        for (size_t i = 0; i < results.size(); i++)
            if (results[i].errorRate < minErrorRate) { minErrorRate = results[i].errorRate; mergedSeq = results[i].thread; }
        if (mergedSeq.length() != 0) return 1;
        return -4;
    }

    // LongReadCorrectByOverlap.cpp:239-278
    void extendLeaves(std::list<Leaf>& newLeaves)
    {
        if (m_currentKmerSize > m_maxOverlap) refineSAInterval(m_leaves, m_maxOverlap);
        attempToExtend(newLeaves);
        if (newLeaves.empty())
        {
            size_t LowerBound = std::max(m_currentKmerSize - 2, m_minOverlap);
            size_t ReduceSize = SelectFreqsOfrange(LowerBound, m_currentKmerSize, m_leaves);
            refineSAInterval(m_leaves, ReduceSize);
            attempToExtend(newLeaves);
            if (newLeaves.empty())
            {
                m_min_SA_threshold--;
                attempToExtend(newLeaves);
                m_min_SA_threshold++;
            }
        }
        if (!newLeaves.empty())
        {
            m_currentLength++;
            m_currentKmerSize++;
            if (isInsufficientFreqs(newLeaves))
            {
                size_t LowerBound = std::max(m_currentKmerSize - 2, m_minOverlap);
                size_t ReduceSize = SelectFreqsOfrange(LowerBound, m_currentKmerSize, newLeaves);
                refineSAInterval(newLeaves, ReduceSize);
            }
        }
    }

    // LongReadCorrectByOverlap.cpp:281-331
    size_t SelectFreqsOfrange(const size_t LowerBound, const size_t UpperBound, std::list<Leaf>& newLeaves)
    {
        std::vector<FMidx> maxKmerArray;
        int tempmaxfmfreqs = 0;
        for (auto& leaf : newLeaves)
        {
            std::string maxKmer = leaf.getSuffix(UpperBound);
            std::string startkmer = maxKmer.substr(UpperBound - LowerBound);
            BWTInterval Fwdinterval = findInterval(m_pBWT, startkmer);
            BWTInterval Rvcinterval = findInterval(m_pRBWT, reverseComplement(reverse(startkmer)));
            maxKmerArray.emplace_back(maxKmer, Fwdinterval, Rvcinterval);
            FMidx& currKmer = maxKmerArray.back();
            if (currKmer.kmerFrequency > tempmaxfmfreqs) tempmaxfmfreqs = currKmer.kmerFrequency;
        }
        if (tempmaxfmfreqs - (int)freqsOfKmerSize[LowerBound] < 5) return LowerBound;
        for (size_t i = 1; i <= UpperBound - LowerBound; i++)
        {
            tempmaxfmfreqs = 0;
            for (size_t j = 0; j < maxKmerArray.size(); j++)
            {
                std::string startkmer = maxKmerArray.at(j).SearchLetters.substr(UpperBound - LowerBound - i);
                BWTInterval Fwdinterval = maxKmerArray.at(j).fwdInterval;
                BWTInterval Rvcinterval = maxKmerArray.at(j).rvcInterval;
                char b = startkmer[0];
                char rcb = comp(b);
                updateInterval(Fwdinterval, b, m_pBWT);
                updateInterval(Rvcinterval, rcb, m_pRBWT);
                maxKmerArray.at(j).setInterval(Fwdinterval, Rvcinterval);
                if (maxKmerArray.at(j).kmerFrequency > tempmaxfmfreqs) tempmaxfmfreqs = maxKmerArray.at(j).kmerFrequency;
            }
            if (tempmaxfmfreqs - (int)freqsOfKmerSize[LowerBound + i] < 5) return LowerBound + i;
        }
        return UpperBound;
    }

    // LongReadCorrectByOverlap.cpp:334-352
    bool isInsufficientFreqs(std::list<Leaf>& newLeaves)
    {
        size_t highfreqscount = 0;
        for (auto& leaf : newLeaves)
        {
            int highfreqThreshold = m_PBcoverage > 60 ? (size_t)(m_PBcoverage / 60) * 3 : 3;
            if (leaf.kmerFrequency > highfreqThreshold) highfreqscount++;
        }
        if (highfreqscount == 0) return true;
        else if (highfreqscount <= 2 && newLeaves.size() >= 5) return true;
        else if (highfreqscount <= 1 && newLeaves.size() >= 3) return true;
        return false;
    }

    // LongReadCorrectByOverlap.cpp:355-369
    void refineSAInterval(std::list<Leaf>& leaves, const size_t newKmerSize)
    {
        for (auto& leaf : leaves)
        {
            std::string reducedKmer = leaf.getSuffix(newKmerSize);
            leaf.fwdInterval = findInterval(m_pRBWT, reverse(reducedKmer));
            leaf.rvcInterval = findInterval(m_pBWT, reverseComplement(reducedKmer));
            if (RefineStats::on())
            {
                const int64_t f = leaf.fwdInterval.isValid() ? leaf.fwdInterval.upper - leaf.fwdInterval.lower + 1 : 0;
                const int64_t r = leaf.rvcInterval.isValid() ? leaf.rvcInterval.upper - leaf.rvcInterval.lower + 1 : 0;
                RefineStats::calls()++; RefineStats::bases() += newKmerSize;
                if (f + r >= 2) RefineStats::twice()++;
            }
        }
        m_currentKmerSize = newKmerSize;
    }

    // LongReadCorrectByOverlap.cpp:373-465
    void attempToExtend(std::list<Leaf>& newLeaves)
    {
        double minimumErrorRate = 1;
        for (auto& leaf : m_leaves)
            if (leaf.LocalErrorRateRecord.back() < minimumErrorRate) minimumErrorRate = leaf.LocalErrorRateRecord.back();
        auto iter = m_leaves.begin();
        while (iter != m_leaves.end())
        {
            double errorRateDiff = (iter->LocalErrorRateRecord.back()) - minimumErrorRate;
            if ((errorRateDiff > 0.05 && m_currentLength > m_localSimilarlykmerSize / 2)
                || (errorRateDiff > 0.1 && m_currentLength > 15))
            {
                iter = m_leaves.erase(iter);
                continue;
            }
            ++iter;
        }
        iter = m_leaves.begin();
        while (iter != m_leaves.end())
        {
            std::vector<FMidx> extensions;
            int count = 0;
            while (count < 2)
            {
                if (count == 1 && !(iter->LocalErrorRateRecord.back() == minimumErrorRate && m_leaves.size() > 1)) break;
                extensions = getFMIndexExtensions(*iter);
                if (extensions.size() > 0)
                {
                    updateLeaves(newLeaves, extensions, *iter);
                    break;
                }
                m_min_SA_threshold--;
                count++;
            }
            m_min_SA_threshold += count;
            ++iter;
        }
    }

    // LongReadCorrectByOverlap.cpp:468-488 + leafInfo ctor (.h:179-208) + createChild (SAINode.cpp:166-189)
    void updateLeaves(std::list<Leaf>& newLeaves, std::vector<FMidx>& extensions, Leaf& leaf)
    {
        for (size_t i = 0; i < extensions.size(); ++i)
        {
            Leaf child = leaf;   // single extension mutates in place, branching copies every counter and both histories
            FMidx& ext = extensions[i];
            child.str += ext.SearchLetters;
            child.kmerFrequency = ext.kmerFrequency;
            child.fwdInterval = ext.fwdInterval;
            child.rvcInterval = ext.rvcInterval;
            child.currOverlapLen++;
            child.queryOverlapLen++;
            if (leaf.tailLetter == ext.SearchLetters[0]) { child.tailLetter = leaf.tailLetter; child.tailLetterCount = leaf.tailLetterCount + 1; }
            else { child.tailLetter = ext.SearchLetters[0]; child.tailLetterCount = 1; }
            newLeaves.push_back(child);
        }
    }

    // LongReadCorrectByOverlap.cpp:491-559
    bool PrunedBySeedSupport(std::list<Leaf>& newLeaves)
    {
        size_t currSeedIdx = m_currentLength - m_seedSize;
        size_t indelOffset = m_seedSize + m_maxIndelSize;
        size_t smallSeedIdx = currSeedIdx <= indelOffset ? 0 : currSeedIdx - indelOffset;
        size_t largeSeedIdx = (currSeedIdx + indelOffset) >= (m_query.length() - m_seedSize) ?
                              (m_query.length() - m_seedSize) : currSeedIdx + indelOffset;
        auto iter = newLeaves.begin();
        while (iter != newLeaves.end())
        {
            bool isNewSeedFound = false;
            Leaf* leaf = &*iter;
            if (m_currentLength - leaf->lastOverlapLen > m_seedSize || m_currentLength - leaf->lastOverlapLen <= 1)
            {
                size_t preSeedIdx = leaf->lastSeedIdx;
                isNewSeedFound = isSupportedByNewSeed(leaf, smallSeedIdx, largeSeedIdx);
                if (isNewSeedFound)
                {
                    if (currSeedIdx + leaf->lastSeedIdxOffset - preSeedIdx > m_seedSize)
                        leaf->numRedeemSeed += (m_seedSize - 1) * m_PacBioErrorRate;
                    leaf->lastSeedIdxOffset = (int)leaf->lastSeedIdx - (int)currSeedIdx;
                }
                else
                {
                    if ((currSeedIdx + leaf->lastSeedIdxOffset - leaf->lastSeedIdx) % m_seedSize == 1)
                        leaf->numOfErrors++;
                    else if ((currSeedIdx + leaf->lastSeedIdxOffset - leaf->lastSeedIdx) > m_seedSize - 1)
                        leaf->numRedeemSeed += 1 - m_PacBioErrorRate;
                }
            }
            else
                leaf->numRedeemSeed += 1 - m_PacBioErrorRate;
            double currErrorRate = computeErrorRate(leaf);
            if (currErrorRate > m_errorRate)
            {
                iter = newLeaves.erase(iter);
                continue;
            }
            iter++;
        }
        return true;
    }

    // LongReadCorrectByOverlap.cpp:562-633
    bool isSupportedByNewSeed(Leaf* currNode, size_t smallSeedIdx, size_t largeSeedIdx)
    {
        size_t seedIdxOffset = currNode->lastOverlapLen < m_currentLength - m_seedSize ?
                               m_seedSize : m_currentLength - currNode->lastOverlapLen;
        size_t startSeedIdx = std::max(smallSeedIdx, currNode->lastSeedIdx + seedIdxOffset);
        bool isNewSeedFound = false;
        BWTInterval currFwdInterval = currNode->fwdInterval;
        BWTInterval currRvcInterval = currNode->rvcInterval;
        std::vector<TreeInterval> resultsFwd, resultsRvc;
        if (currFwdInterval.isValid()) m_fwdIntervalTree.findOverlapping(currFwdInterval.lower, currFwdInterval.upper, resultsFwd);
        if (currRvcInterval.isValid()) m_rvcIntervalTree.findOverlapping(currRvcInterval.lower, currRvcInterval.upper, resultsRvc);
        int minIdxDiff = 10000;
        size_t currSeedIdx = m_currentLength - m_seedSize;
        for (size_t i = 0; i < resultsFwd.size() || i < resultsRvc.size(); i++)
        {
            if (currFwdInterval.isValid() && i < resultsFwd.size() && resultsFwd.at(i).value >= startSeedIdx && resultsFwd.at(i).value <= largeSeedIdx)
            {
                if (std::abs((int)resultsFwd.at(i).value - (int)currSeedIdx) < minIdxDiff)
                {
                    currNode->lastSeedIdx = resultsFwd.at(i).value;
                    currNode->queryOverlapLen = resultsFwd.at(i).value + m_seedSize;
                    minIdxDiff = std::abs((int)resultsFwd.at(i).value - (int)currSeedIdx);
                }
                currNode->lastOverlapLen = m_currentLength;
                currNode->currOverlapLen = m_currentLength;
                isNewSeedFound = true;
            }
            else if (currRvcInterval.isValid() && i < resultsRvc.size() && resultsRvc.at(i).value >= startSeedIdx && resultsRvc.at(i).value <= largeSeedIdx)
            {
                if (std::abs((int)currSeedIdx - (int)resultsRvc.at(i).value) < minIdxDiff)
                {
                    currNode->lastSeedIdx = resultsRvc.at(i).value;
                    currNode->queryOverlapLen = resultsRvc.at(i).value + m_seedSize;
                    minIdxDiff = std::abs((int)resultsRvc.at(i).value - (int)currSeedIdx);
                }
                currNode->lastOverlapLen = m_currentLength;
                currNode->currOverlapLen = m_currentLength;
                isNewSeedFound = true;
            }
        }
        if (isNewSeedFound) currNode->totalSeeds++;
        return isNewSeedFound;
    }

    // LongReadCorrectByOverlap.cpp:636-664
    double computeErrorRate(Leaf* currNode)
    {
        double matchedLen = (double)currNode->totalSeeds + m_seedSize - 1;
        matchedLen += currNode->numRedeemSeed;
        double totalLen = (double)currNode->currOverlapLen;
        double unmatchedLen = totalLen - matchedLen;
        double currErrorRate = unmatchedLen / totalLen;
        currNode->GlobalErrorRateRecord.push_back(currErrorRate);
        if (currNode->GlobalErrorRateRecord.size() >= m_localSimilarlykmerSize)
        {
            size_t totalsize = currNode->GlobalErrorRateRecord.size();
            currErrorRate = (currErrorRate * totalLen - currNode->GlobalErrorRateRecord.at(totalsize - m_localSimilarlykmerSize) * (totalLen - m_localSimilarlykmerSize)) / m_localSimilarlykmerSize;
        }
        currNode->LocalErrorRateRecord.push_back(currErrorRate);
        return currErrorRate;
    }

    // LongReadCorrectByOverlap.cpp:667-784
    std::vector<FMidx> getFMIndexExtensions(const Leaf& currLeaf)
    {
        std::vector<FMidx> output, totalExt;
        size_t IntervalSizeCutoff = m_min_SA_threshold;
        size_t totalcount = 0;
        int maxfreqsofleave = 0;
        static const char RANK[5] = {'$', 'A', 'C', 'G', 'T'};
        for (int i = 1; i < 5; ++i)
        {
            char b = RANK[i];
            BWTInterval fwdProbe = currLeaf.fwdInterval;
            if (fwdProbe.isValid()) updateInterval(fwdProbe, b, m_pRBWT);
            char rcb = RANK[5 - i];
            BWTInterval rvcProbe = currLeaf.rvcInterval;
            if (rvcProbe.isValid()) updateInterval(rvcProbe, rcb, m_pBWT);
            FMidx currExt = FMidx(std::string(1, b), fwdProbe, rvcProbe);
            totalcount += currExt.kmerFrequency;
            if (currExt.kmerFrequency > maxfreqsofleave) maxfreqsofleave = currExt.kmerFrequency;
            totalExt.push_back(currExt);
        }
        for (int i = 1; i < 5; ++i)
        {
            size_t kmerFreq = totalExt.at(i - 1).kmerFrequency;
            BWTInterval fwdInterval = totalExt.at(i - 1).fwdInterval;
            BWTInterval rvcInterval = totalExt.at(i - 1).rvcInterval;
            const double kmerRatioNotPass = 2;
            double kmerRatioCutoff = 0;
            double kmerRatio = (double)kmerFreq / (double)maxfreqsofleave;
            char b = RANK[i];
            bool isHomopolymer = (currLeaf.tailLetterCount >= 3);
            bool isMatchedBy5mer = ismatchedbykmer(fwdInterval, rvcInterval);
            bool isFreqPass = kmerFreq >= IntervalSizeCutoff;
            bool isLowCoverage = totalcount >= IntervalSizeCutoff + 2;
            bool isRepeat = maxfreqsofleave > 100;
            bool isHighlyRepeat = maxfreqsofleave > 150;
            bool isLowlyRepeat = maxfreqsofleave > 50;
            if (isMatchedBy5mer && isHighlyRepeat) kmerRatioCutoff = 0.125;
            else if (isMatchedBy5mer && isLowlyRepeat) kmerRatioCutoff = 0.2;
            else if (isFreqPass) kmerRatioCutoff = 0.25;
            else if (isLowCoverage) kmerRatioCutoff = 0.6;
            else kmerRatioCutoff = kmerRatioNotPass;
            if (isHomopolymer && isRepeat) kmerRatioCutoff = std::max(kmerRatioCutoff, 0.3);
            else if (isHomopolymer) kmerRatioCutoff = std::max(kmerRatioCutoff, 0.6);
            if (kmerRatio >= kmerRatioCutoff) output.emplace_back(std::string(1, b), fwdInterval, rvcInterval);
        }
        return output;
    }

    // LongReadCorrectByOverlap.cpp:787-821
    bool ismatchedbykmer(BWTInterval currFwdInterval, BWTInterval currRvcInterval)
    {
        bool match = false;
        std::vector<TreeInterval> resultsFwd, resultsRvc;
        if (currFwdInterval.isValid()) m_fwdIntervalTree2.findOverlapping(currFwdInterval.lower, currFwdInterval.upper, resultsFwd);
        if (currRvcInterval.isValid()) m_rvcIntervalTree2.findOverlapping(currRvcInterval.lower, currRvcInterval.upper, resultsRvc);
        size_t startSeedIdx = std::max((int)m_currentLength - (int)m_maxIndelSize, 0);
        size_t largeSeedIdx = m_currentLength + m_maxIndelSize;
        for (size_t i = 0; i < resultsFwd.size() || i < resultsRvc.size(); i++)
        {
            if (currFwdInterval.isValid() && i < resultsFwd.size() && resultsFwd.at(i).value >= startSeedIdx && resultsFwd.at(i).value <= largeSeedIdx) { match = true; break; }
            else if (currRvcInterval.isValid() && i < resultsRvc.size() && resultsRvc.at(i).value >= startSeedIdx && resultsRvc.at(i).value <= largeSeedIdx) { match = true; break; }
        }
        return match;
    }

    // LongReadCorrectByOverlap.cpp:825-878
    bool isTerminated(std::vector<WalkResultCand>& results)
    {
        bool found = false;
        for (auto& leafRef : m_leaves)
        {
            Leaf* leaf = &leafRef;
            BWTInterval currfwd = leaf->fwdInterval, currrvc = leaf->rvcInterval;
            for (size_t i = std::max(leaf->resultindex.second, 0); i <= m_targetSeed.length() - (int)m_minOverlap; i++)
            {
                bool isFwdTerminated = currfwd.isValid() && currfwd.lower >= m_fwdTerminatedInterval.at(i).lower && currfwd.upper <= m_fwdTerminatedInterval.at(i).upper;
                bool isRvcTerminated = currrvc.isValid() && currrvc.lower >= m_rvcTerminatedInterval.at(i).lower && currrvc.upper <= m_rvcTerminatedInterval.at(i).upper;
                if (isFwdTerminated || isRvcTerminated)
                {
                    std::string STNodeStr = leaf->str;
                    if (m_targetSeed.length() > m_minOverlap) STNodeStr += m_targetSeed.substr(i + m_minOverlap);
                    WalkResultCand r; r.thread = STNodeStr; r.errorRate = leaf->GlobalErrorRateRecord.back();
                    if (leaf->resultindex.first == -1)
                    {
                        results.push_back(r);
                        leaf->resultindex = std::make_pair((int)results.size(), (int)i);
                    }
                    else
                    {
                        results.at(leaf->resultindex.first - 1) = r;
                        leaf->resultindex = std::make_pair(leaf->resultindex.first, (int)i);
                    }
                    found = true;
                }
            }
        }
        return found;
    }
};

// ---------------------------------------------------------------------------------
// PacBioSelfCorrectionProcess — PacBio/PacBioSelfCorrectionProcess.cpp:23-206
// ---------------------------------------------------------------------------------
}  // namespace pbo
#include "pbsc_oracle_dp.hpp"
namespace pbo {

struct PairRecord   // one FM-extension attempt, for structured parity checks
{
    int srcStart, trgStart, extendKmerSize, dis, status;
    bool fromRtoU;
    std::string src, path, trg, out;
};
struct ReadResult
{
    std::string readid;
    bool merge = false;
    std::vector<std::string> correctedStrs;
    int64_t totalReadsLen = 0, correctedLen = 0, totalSeedNum = 0, totalWalkNum = 0, highErrorNum = 0,
            exceedDepthNum = 0, exceedLeaveNum = 0, FMNum = 0, DPNum = 0, seedDis = 0;
    SeedVector seeds, outcastSeeds;
    bool outcastWritten = false, seedWritten = false;
    uint64_t occSeed = 0, occExtend = 0;   // rank queries issued by each phase (roofline numerator)
    // occExtend split: the LongReadSelfCorrectByOverlap constructor (terminal intervals, query idmer and 5-mer interval trees, root
    // intervals; LongReadCorrectByOverlap.cpp:17-152) and the level loop (extendOverlap, :155-211)
    uint64_t occExtendSetup = 0, occExtendWalk = 0;
    uint64_t dpCells = 0, dpRows = 0, dpAttempts = 0, occDP = 0;   // banded-DP cells / rows kept / fallbacks tried / LF steps of the DP fallback
    std::vector<std::pair<int,int>> dpFailLog;                     // extend/<id>.dp rows
    std::vector<PairRecord> pairs;
    std::vector<std::pair<std::pair<int,int>,int>> extLog;   // extend/<id>.ext rows
    std::vector<float> ratioLog;
};

struct Corrector
{
    const Params& P;
    explicit Corrector(const Params& p) : P(p) {}

    // PacBioSelfCorrectionProcess.cpp:159-206
    int correctByFMExtension(const SeedFeature& source, const SeedFeature& target, const std::string& in, std::string& out, ReadResult& result)
    {
        int interval = target.seedStartPos - source.seedEndPos - 1;
        int extendKmerSize = std::min(source.endBestKmerSize, target.startBestKmerSize) - 2;
        if (source.isRepeat || target.isRepeat)
        {
            extendKmerSize = std::min(source.seedLen, target.seedLen);
            extendKmerSize = std::min(extendKmerSize, P.startKmerLen + 2);
        }
        std::string src, trg, path;
        src = source.seedStr.substr(source.seedLen - extendKmerSize);
        trg = target.seedStr;
        path = in.substr(source.seedEndPos + 1, interval);
        int min_SA_threshold = 3, isFMExtensionSuccess = 0;
        min_SA_threshold = P.PBcoverage > 60 ? ((P.PBcoverage / 60) * 3) : min_SA_threshold;
        bool isFromRtoU = source.isRepeat && !target.isRepeat;
        if (isFromRtoU)
        {
            std::swap(src, trg);
            src = reverseComplement(src);
            trg = reverseComplement(trg);
            path = reverseComplement(path);
        }
        PairRecord rec;
        rec.srcStart = source.seedStartPos; rec.trgStart = target.seedStartPos; rec.extendKmerSize = extendKmerSize;
        rec.dis = interval; rec.fromRtoU = isFromRtoU; rec.src = src; rec.path = path; rec.trg = trg;
        std::string mergedSeq;
        const uint64_t occCtor0 = OccCounter::n();
        FMExtend tree(src, path, trg, interval, extendKmerSize, extendKmerSize + 2, P, min_SA_threshold);
        const uint64_t occCtor1 = OccCounter::n();
        isFMExtensionSuccess = tree.extendOverlap(mergedSeq);
        result.occExtendSetup += occCtor1 - occCtor0;
        result.occExtendWalk += OccCounter::n() - occCtor1;
        rec.status = isFMExtensionSuccess;
        if (isFMExtensionSuccess < 0) { result.pairs.push_back(rec); return isFMExtensionSuccess; }
        rec.out = mergedSeq;
        result.pairs.push_back(rec);
        if (isFromRtoU)
        {
            mergedSeq = reverseComplement(mergedSeq);
            mergedSeq += reverseComplement(src).substr(extendKmerSize);
        }
        out = mergedSeq;
        out.erase(0, extendKmerSize);
        result.correctedLen += out.length();
        result.seedDis += interval;
        result.FMNum++;
        return isFMExtensionSuccess;
    }

    // DP/MSA fallback — PacBioSelfCorrectionProcess.cpp:208-245 (alignment and consensus in pbsc_oracle_dp.hpp)
    bool correctByMSAlignment(const SeedFeature& source, const SeedFeature& target, const std::string& in, std::string& out, ReadResult& result)
    {
        if (P.NoDp) return false;
        int interval = target.seedStartPos - source.seedEndPos - 1;
        int extendKmerSize = std::min(source.endBestKmerSize, target.startBestKmerSize) - 2;
        if (source.isRepeat || target.isRepeat)
        {
            extendKmerSize = std::min(source.seedLen, target.seedLen);
            extendKmerSize = std::min(extendKmerSize, P.startKmerLen + 2);
        }
        std::string path = source.seedStr.substr(source.seedLen - extendKmerSize) + in.substr(source.seedEndPos + 1, interval) + target.seedStr;
        double identity = 0.65;
        size_t totalMaxFixedMerFreq = source.maxFixedMerFreq + target.maxFixedMerFreq, min_call_coverage = 15;
        identity += (totalMaxFixedMerFreq > 50 ? 0.05 : 0);
        identity += (totalMaxFixedMerFreq > 100 ? 0.05 : 0);
        min_call_coverage = totalMaxFixedMerFreq > 50 ? totalMaxFixedMerFreq * 0.4 : min_call_coverage;
        const uint64_t occBefore = OccCounter::n();
        Msa msa = buildMultipleAlignment(path, extendKmerSize, extendKmerSize, path.length() / 10, identity, P.PBcoverage, P.indices, &result.dpCells);
        result.occDP += OccCounter::n() - occBefore;
        result.dpRows += msa.rows.size() - 1;
        result.dpAttempts++;
        if (msa.rows.size() <= 3) return false;
        out = msa.consensus((int)min_call_coverage);
        out.erase(0, extendKmerSize);
        result.correctedLen += out.length();
        result.seedDis += interval;
        result.DPNum++;
        return true;
    }

    // PacBioSelfCorrectionProcess.cpp:56-157
    void initCorrect(std::string& readSeq, const SeedVector& seedVec, SeedVector& pieceVec, ReadResult& result)
    {
        if (seedVec.size() < 2) return;
        pieceVec.push_back(seedVec[0]);
        for (SeedVector::const_iterator iterTarget = seedVec.begin() + 1; iterTarget != seedVec.end(); iterTarget++)
        {
            int isFMExtensionSuccess = 0, firstFMExtensionType = 0;
            SeedFeature& source = pieceVec.back();
            std::string mergedSeq;
            for (int next = 0; next < P.nextTarget && (iterTarget + next) != seedVec.end(); next++)
            {
                const SeedFeature& target = *(iterTarget + next);
                isFMExtensionSuccess = correctByFMExtension(source, target, readSeq, mergedSeq, result);
                firstFMExtensionType = (next == 0 ? isFMExtensionSuccess : firstFMExtensionType);
                if (isFMExtensionSuccess > 0)
                {
                    result.totalWalkNum++;
                    source.append(mergedSeq, target);
                    iterTarget += next;
                    break;
                }
            }
            if (isFMExtensionSuccess <= 0)
            {
                const SeedFeature& target = *iterTarget;
                switch (firstFMExtensionType)
                {
                    case -1: result.highErrorNum++; break;
                    case -2: result.exceedDepthNum++; break;
                    case -3: result.exceedLeaveNum++; break;
                    default: std::cerr << "Does it really happen?\n"; exit(EXIT_FAILURE);
                }
                result.extLog.push_back(std::make_pair(std::make_pair(source.seedStartPos, target.seedStartPos), firstFMExtensionType + 4));
                result.totalWalkNum++;
                bool isMSAlignmentSuccess = correctByMSAlignment(source, target, readSeq, mergedSeq, result);
                if (isMSAlignmentSuccess) source.append(mergedSeq, target);
                else
                {
                    result.dpFailLog.push_back(std::make_pair(source.seedStartPos, target.seedStartPos));
                    if (P.Split) pieceVec.push_back(target);
                    else
                    {
                        mergedSeq = readSeq.substr((source.seedEndPos + 1), (target.seedEndPos - source.seedEndPos));
                        source.append(mergedSeq, target);
                    }
                    result.correctedLen += target.seedStr.length();
                }
            }
        }
    }

    // PacBioSelfCorrectionProcess.cpp:23-54
    ReadResult process(const std::string& id, const std::string& seq)
    {
        ReadResult result;
        result.readid = id;
        std::string readSeq = seq;
        SeedVector seedVec, pieceVec;
        LongReadProbe probe(P);
        const uint64_t occ0 = OccCounter::n();
        probe.searchSeedsWithHybridKmers(readSeq, seedVec);
        const uint64_t occ1 = OccCounter::n();
        result.totalSeedNum = seedVec.size();
        result.seeds = seedVec;
        result.outcastSeeds = probe.outcast;
        result.outcastWritten = probe.outcastWritten;
        result.seedWritten = !((int)readSeq.length() < P.startKmerLen);
        result.ratioLog = probe.ratioLog;
        initCorrect(readSeq, seedVec, pieceVec, result);
        result.occSeed = occ1 - occ0;
        result.occExtend = OccCounter::n() - occ1 - result.occDP;
        result.merge = !pieceVec.empty();
        result.totalReadsLen = readSeq.length();
        for (const auto& iter : pieceVec) result.correctedStrs.push_back(iter.seedStr);
        return result;
    }
};

// ---------------------------------------------------------------------------------
// FASTA reader — Util/SeqReader.cpp:26-135 (FASTA branch only; upper-cases; exits on non-ACGT;
// a last line without trailing '\n' is dropped, as in the reference)
// ---------------------------------------------------------------------------------
struct FastaReader
{
    std::ifstream in;
    explicit FastaReader(const std::string& path) : in(path.c_str()) {}
    bool get(std::string& id, std::string& seq)
    {
        std::string header;
        bool found = false;
        while (in.good())
        {
            getline(in, header);
            if (header.empty()) continue;
            if (header[0] == '>') { found = true; break; }
        }
        if (!found) return false;
        seq.clear();
        std::string temp;
        while (in.good() && in.peek() != '>' && in.peek() != '@')
        {
            getline(in, temp);
            if (in.good() && temp.size() > 0) seq.append(temp);
        }
        if (seq.empty()) return false;
        size_t endPos = std::min(header.find_first_of(' '), header.find_first_of('\t'));
        id = (endPos != std::string::npos) ? header.substr(1, endPos - 1) : header.substr(1);
        std::transform(seq.begin(), seq.end(), seq.begin(), ::toupper);
        if (seq.find_first_not_of("ACGT") != std::string::npos)
        {
            std::cerr << "Error: read " << id << " contains non-ACGT characters.\n";
            std::cerr << "Please run sga preprocess on the data first.\n";
            exit(EXIT_FAILURE);
        }
        return true;
    }
};

}  // namespace pbo
#endif
