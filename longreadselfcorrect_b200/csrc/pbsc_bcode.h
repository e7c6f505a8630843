// pbsc_bcode.h — `--onlyseed -b BARCODE`: the reference's check of seeds against alignment barcodes, host side only
// (PacBio/BCode.cpp:27-165 load / fetch / sum / getPys / validate; PacBioSelfCorrectionProcess.cpp:315-335,372-380 the per-read
// and total summaries).  Plain C++, no CUDA: tests/cpp/test_bcode.cpp compiles it with g++ and compares the summaries with what
// the unmodified reference wrote for the same seeds (tests/golden/tiny.onlyseed.*).
//
// A barcode record describes how a stretch [start, end) of a read aligns to its source: `code` holds two hex digits per base —
// the even places count inserted bases at that position, the odd places encode the bases deleted after it (A = 1, T = 2, C = 4,
// G = 8, or-ed) — and `rvc` says on which strand the alignment was made, which decides from which end of a k-mer the gaps are
// looked at.  A seed is "correct" when its bases carry no gap, or only a gap at the far end that its own bases explain.
#ifndef PBSC_BCODE_H
#define PBSC_BCODE_H

#include <stdio.h>
#include <fstream>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <zlib.h>

namespace pbsc { namespace bcode {

struct Block { int start = 0, end = 0; std::string code; bool rvc = false; };
typedef std::map<std::string, std::vector<Block> > Table;

// nine blank-separated fields per record: qname qstart qend tname tstart tend code rvc sup (BCode.cpp:27-49); plain or gzip
inline bool load(const std::string& path, Table& table, std::string& err)
{
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) { err = "could not open " + path + " for read"; return false; }
    std::string all;
    char buf[1 << 16];
    for (int n; (n = gzread(f, buf, sizeof buf)) > 0;) all.append(buf, (size_t)n);
    gzclose(f);
    std::istringstream in(all);
    std::string qname, tname, code, rvc, sup;
    int qs, qe, ts, te;
    while (in >> qname >> qs >> qe >> tname >> ts >> te >> code >> rvc >> sup)
    {
        Block b;
        b.start = qs; b.end = qe; b.code = code; b.rvc = rvc == "True";
        table[qname].push_back(b);
    }
    return true;
}

inline int digit(char c)
{
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    throw std::out_of_range(std::string("barcode: '") + c + "' is not a hex digit");   // std::map::at in the reference
}
inline int base_bit(char c)
{
    switch (c) { case 'a': case 'A': return 1; case 't': case 'T': return 2; case 'c': case 'C': return 4; case 'g': case 'G': return 8; }
    throw std::out_of_range(std::string("barcode: '") + c + "' is not a base");
}
// a negative position counts from the end, once (BCode::getPys)
inline int from_end(int pos, int len)
{
    if (pos < 0) pos += len;
    if (pos < 0) throw std::out_of_range("barcode: position before the start of its string");   // the reference's assert
    return pos;
}
// in[pos::step] as the reference walks it: from the (wrapped) position while the index stays inside the string
inline std::string every(const std::string& in, int pos, int step)
{
    std::string out;
    for (int i = from_end(pos, (int)in.size()); i >= 0 && i < (int)in.size(); i += step) out += in[(size_t)i];
    return out;
}
inline int digit_sum(const std::string& s) { int t = 0; for (size_t i = 0; i < s.size(); i++) t += digit(s[i]); return t; }
inline char at_or_nul(const std::string& s, long i) { return i >= 0 && i < (long)s.size() ? s[(size_t)i] : '\0'; }

// BCode::validate(pos, ksize, block, seq): is the k-mer seq[pos, pos + k) free of alignment gaps it does not explain itself?
inline bool validate(int pos, int k, const Block& blk, const std::string& seq)
{
    const int base = blk.start;
    const int lo = (pos - base) * 2, hi = (pos + k - base) * 2 - 1;
    const std::string kmer = seq.substr((size_t)pos, (size_t)k);
    const std::string& code = blk.code;
    const std::string info = code.substr((size_t)lo, (size_t)(hi - lo));   // two digits per base of the k-mer, less the last odd one
    // the end of the k-mer the gaps are read from: its last base on the forward strand, its first on the reverse strand
    const int sign = blk.rvc ? -1 : 1, bit = blk.rvc ? 0 : 1;
    const int pole = blk.rvc ? pos : pos + k;
    // inserted bases (even places)
    const int inserted = digit_sum(every(info, 0, 2));
    if (inserted > 0)
    {
        // from the pole inwards: bases without insertion, then single insertions; that run must hold every insertion of the k-mer
        int run = 0, ones = 0;
        const std::string evens = every(info, -bit, -sign * 2);
        for (size_t i = 0; i < evens.size(); i++)
        {
            const int v = digit(evens[i]);
            const bool fits = ones == 0 ? (v == 0 || v == 1) : v == 1;
            if (!fits) break;
            run++;
            ones += v;
        }
        if (inserted != ones) return false;
        if (ones > 0)
        {
            const std::string upper = every(code, 0, 2);   // the even places of the whole block
            // single insertions that continue beyond the pole
            int beyond = 0;
            const std::string outside = every(upper, pole - base + bit - 1, sign);
            for (size_t i = 0; i < outside.size() && digit(outside[i]) == 1; i++) beyond++;
            if (run - ones > 0 && beyond > 0) return false;
            // the `run` bases past the insertions must be gap-free and repeat the k-mer's own end
            for (int i = 0; i < run; i++)
            {
                const long where = (long)pole + sign * (1 - bit + beyond + i) - sign * (run - ones);
                if (at_or_nul(upper, where - base) != '0') return false;
                if (kmer[(size_t)from_end(-sign * (run + bit - 1 - i), k)] != at_or_nul(seq, where)) return false;
            }
        }
    }
    // deleted bases (odd places)
    const int deleted = digit_sum(every(info, 1, 2));
    if (deleted > 0)
    {
        // from the pole inwards up to and including the first deletion mark; the bases passed on the way are collected as bits
        int mark = 0, steps = 0, seen = 0;
        const std::string odds = every(info, -sign * (1 + bit), -sign * 2);
        for (size_t i = 0; i < odds.size() && mark == 0; i++)
        {
            const int v = digit(odds[i]);
            seen |= base_bit(kmer[(size_t)from_end(-sign * (bit + steps), k)]);
            steps++;
            mark += v;
        }
        if (deleted != mark) return false;
        if (mark > 0)
        {
            const int two_bases = ((mark & 1) + ((mark >> 1) & 1) + ((mark >> 2) & 1) + ((mark >> 3) & 1)) == 2;
            if (!(mark == seen || (steps == 1 && (mark & seen) > 0 && two_bases))) return false;
        }
    }
    return true;
}

// one seed against the blocks of its read: 0 = inside a block and valid, 1 = inside a block and not, 2 = in no block
// (PacBioSelfCorrectionProcess.cpp:320-331; the first block that contains the seed decides)
inline int classify(const std::vector<Block>* blocks, int start, int len, const std::string& seq)
{
    if (blocks)
        for (size_t i = 0; i < blocks->size(); i++)
        {
            const Block& b = (*blocks)[i];
            if (start >= b.start && start + len - 1 <= b.end) return validate(start, len, b, seq) ? 0 : 1;
        }
    return 2;
}

// PacBioSelfCorrectionPostProcess::summarize (:372-380): a line only when at least one seed is wrong
inline void summarize(FILE* out, const size_t status[3], const std::string& subject)
{
    const size_t sum = (size_t)(int)(status[0] + status[1] + status[2]);   // std::accumulate with an int seed
    if (status[1] == 0) return;
    fprintf(out, "%s [%ld] %.2lf%% %.2lf%% %.2lf%%\n", subject.c_str(), (long)sum, (double)(100 * status[0]) / sum, (double)(100 * status[1]) / sum,
            (double)(100 * status[2]) / sum);
}

// ---- `kmercheck` (PacBio/KmerCheckProcess.cpp:12-63): frequencies of the correct and of the erroneous k-mers of every barcode block ----
// KmerDistribution (Util/KmerDistribution.h:25, KmerDistribution.cpp:84-152) as far as kmercheck uses it: add, computeKDAttributes'
// five-number summary (integer quartile positions, whiskers at 1.5 x the interquartile range truncated to int) and compare.
struct Histogram
{
    std::map<int, int> bins;
    int total = 0;
    void add(int freq) { bins[freq]++; total++; }
    struct Five { int min = 0, q1 = 0, q2 = 0, q3 = 0, max = 0; };
    Five summary() const
    {
        Five f;
        const int low = total / 4, mid = total * 2 / 4, upp = total * 3 / 4;
        int before = 0, upto = 0;
        for (std::map<int, int>::const_iterator it = bins.begin(); it != bins.end(); ++it)
        {
            before = upto;
            upto += it->second;
            if (low >= before && low <= upto) f.q1 = it->first;   // the last bin whose cumulative range holds the position wins
            if (mid >= before && mid <= upto) f.q2 = it->first;
            if (upp >= before && upp <= upto) f.q3 = it->first;
        }
        const int iqr = f.q3 - f.q1;
        const int small = f.q1 - (int)(iqr * 1.5), large = f.q3 + (int)(iqr * 1.5);
        int prev = 0, cur = 0;
        for (std::map<int, int>::const_iterator it = bins.begin(); it != bins.end(); ++it)
        {
            prev = cur;
            cur = it->first;
            if (f.min == 0 && cur >= small) f.min = cur;
            if (prev <= large && cur > large) f.max = prev;
        }
        if (f.max == 0) f.max = cur;
        return f;
    }
};
// compare (KmerDistribution.cpp:140-152): one line of DIR/total.box ("cov k | wrong | correct") and one of DIR/value.box
inline void compare(std::ostream& total_box, std::ostream& value_box, int cov, int k, const Histogram& correct, const Histogram& wrong)
{
    const Histogram::Five c = correct.summary(), e = wrong.summary();
    total_box << cov << ' ' << k << " | " << e.min << ' ' << e.q1 << ' ' << e.q2 << ' ' << e.q3 << ' ' << e.max << " | " << c.min << ' ' << c.q1 << ' ' << c.q2 << ' '
              << c.q3 << ' ' << c.max << '\n';
    const int value = c.min >= e.max ? c.min : c.q1;   // (the reference's second and third branch both take q1)
    value_box << cov << ' ' << k << ' ' << value << '\n';
}

}}  // namespace pbsc::bcode

#endif
