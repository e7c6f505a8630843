// Host check of the thread-per-alignment banded DP (longreadselfcorrect_b200/csrc/pbsc_dp_thread.cuh): the same fill and
// traceback templates the CUDA kernel instantiates, run here over host accessors that mimic the device storage (16-bit
// circular column, 2-bit packed read with a funnel shift, flag words), against the oracle's Overlapper::extendMatch.
// TEST INFRASTRUCTURE: the oracle is the checker.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include "../../longreadselfcorrect_b200/csrc/pbsc_dp_thread.cuh"
#include "../../oracle/pbsc_oracle.hpp"   // pulls in pbsc_oracle_dp.hpp (pbo::extendMatch)

using namespace pbsc::dpt;

struct HHost
{
    int16_t a[HSLOTS];
    HHost() { for (int i = 0; i < HSLOTS; i++) a[i] = 0; }
    int get(int j) const { return a[j & (HSLOTS - 1)]; }
    bool pairs_ok(int j) const { return ((j >> 1) & (HSLOTS / 2 - 1)) <= HSLOTS / 2 - 5; }
    uint32_t get2(int j, int u) const { const int s = (j & (HSLOTS - 1)) + 2 * u; if ((j & 1) || s + 1 >= HSLOTS) { printf("bad pair access\n"); exit(2); } return (uint32_t)(uint16_t)a[s] | ((uint32_t)(uint16_t)a[s + 1] << 16); }
    void set2(int j, int u, uint32_t w) { const int s = (j & (HSLOTS - 1)) + 2 * u; a[s] = (int16_t)(w & 0xFFFFu); a[s + 1] = (int16_t)(w >> 16); }
    void set(int j, int v) { if (v < -32768 || v > 32767) { printf("16-bit overflow\n"); exit(2); } a[j & (HSLOTS - 1)] = (int16_t)v; }
};
struct SHost
{
    std::vector<uint32_t> w;
    explicit SHost(const std::vector<uint8_t>& s) : w(s.size() / 16 + 2, 0u) { for (size_t x = 0; x < s.size(); x++) w[x >> 4] |= (uint32_t)s[x] << (2 * (x & 15)); }
    int base(int x) const { return (int)((w[x >> 4] >> (2 * (x & 15))) & 3u); }
    uint32_t bits(int x) const
    {
        const uint64_t both = (uint64_t)w[x >> 4] | ((uint64_t)w[(x >> 4) + 1] << 32);
        return (uint32_t)(both >> (2 * (x & 15))) & 0xFFFFFu;
    }
};
struct FHost
{
    std::vector<uint32_t> w; std::vector<char> written;
    explicit FHost(size_t n) : w(n, 0xDEADBEEFu), written(n, 0) {}
    void put(int n, uint32_t v) { if (n < 0 || (size_t)n >= w.size()) { printf("flag word %d out of range\n", n); exit(2); } w[n] = v; written[n] = 1; }
    uint32_t get(int n) const { if (n < 0 || (size_t)n >= w.size() || !written[n]) { printf("flag word %d read before written\n", n); exit(2); } return w[n]; }
};
struct OHost { std::string ops; void put(int, int op) { ops.push_back("MID"[op]); } };

static std::mt19937_64 rng(12345);
static int rnd(int lo, int hi) { return lo + (int)(rng() % (uint64_t)(hi - lo + 1)); }

static std::vector<uint8_t> random_seq(int n, int alphabet, int homop)
{
    std::vector<uint8_t> s;
    while ((int)s.size() < n)
    {
        const uint8_t c = (uint8_t)rnd(0, alphabet - 1);
        const int run = homop ? rnd(1, homop) : 1;
        for (int r = 0; r < run && (int)s.size() < n; r++) s.push_back(c);
    }
    return s;
}
static std::vector<uint8_t> mutate(const std::vector<uint8_t>& s, double err, int alphabet)
{
    std::vector<uint8_t> o;
    for (size_t x = 0; x < s.size(); x++)
    {
        const double u = (double)(rng() % 1000000) / 1e6;
        if (u < err * 0.55) { o.push_back((uint8_t)rnd(0, alphabet - 1)); o.push_back(s[x]); }
        else if (u < err * 0.85) { }
        else if (u < err) o.push_back((uint8_t)rnd(0, alphabet - 1));
        else o.push_back(s[x]);
    }
    return o;
}
static std::string to_str(const std::vector<uint8_t>& s) { std::string o; for (uint8_t c : s) o.push_back("ACGT"[c]); return o; }

// The reference's own answers: tests/golden/dp_units.txt holds "A s1 s2 start_1 start_2" records (and "M" pile-up records,
// skipped here), dp_units.ref.txt what oracle/_ref/dp_dump — linked against the reference's Overlapper — printed for them:
// "A score start0 end0 start1 end1 editDistance totalColumns cigar".  Everything but the score is compared.
static std::string compact_ops(const std::string& ops)
{
    std::string out;
    for (size_t i = 0; i < ops.size();)
    {
        size_t j = i;
        while (j < ops.size() && ops[j] == ops[i]) j++;
        out += std::to_string(j - i) + ops[i];
        i = j;
    }
    return out;
}
static int check_vectors(const char* in_path, const char* ref_path)
{
    std::ifstream in(in_path), ref(ref_path);
    if (!in || !ref) { printf("cannot open the vector files\n"); return 2; }
    std::string tag;
    long tested = 0, failed = 0, skipped = 0;
    while (in >> tag)
    {
        std::string line;
        if (tag == "M")
        {
            std::string q; int minCall, n;
            in >> q >> minCall >> n;
            for (int i = 0; i < n; i++) { std::string s; int a, b; in >> s >> a >> b; }
            std::getline(ref, line);
            continue;
        }
        std::string s1, s2; int a, b;
        in >> s1 >> s2 >> a >> b;
        std::getline(ref, line);
        std::istringstream rs(line);
        std::string rtag, cigar; int score, st0, en0, st1, en1, ed, cols;
        rs >> rtag >> score >> st0 >> en0 >> st1 >> en1 >> ed >> cols >> cigar;
        if (rtag != "A") { printf("answer file out of step\n"); return 2; }
        const int qlen = (int)s1.size(), mlen = (int)s2.size();
        const int origin = (b - a + 1) - (HALF + 1);
        if (!eligible(qlen, mlen, origin, 4000)) { skipped++; continue; }
        std::vector<uint8_t> q(qlen), r(mlen);
        auto code = [](char c) { return (uint8_t)(c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3); };
        for (int x = 0; x < qlen; x++) q[x] = code(s1[x]);
        for (int x = 0; x < mlen; x++) r[x] = code(s2[x]);
        HHost H; SHost S(r); FHost F((size_t)qlen * words_per_col(mlen));
        auto qf = [&](int x) { return (int)q[x]; };
        int bi = 0, bj = 0;
        fill(qlen, mlen, origin, H, S, F, qf, bi, bj);
        tested++;
        if (bi <= 0) { if (cols > 0) { failed++; printf("vector %ld: no traceback start\n", tested); } continue; }
        OHost O; int n = 0, e = 0, i0 = 0, j0 = 0;
        traceback(qlen, mlen, origin, S, F, qf, bi, bj, O, n, e, i0, j0);
        const std::string fwd(O.ops.rbegin(), O.ops.rend());
        if (compact_ops(fwd) != cigar || n != cols || e != ed || i0 != st0 || j0 != st1 || bi - 1 != en0 || bj - 1 != en1)
        {
            failed++;
            if (failed < 10) printf("vector %ld MISMATCH: %s\n  got %d %d %d %d %d %d %s\n", tested, line.c_str(), i0, bi - 1, j0, bj - 1, e, n, compact_ops(fwd).c_str());
        }
    }
    printf("vectors tested %ld skipped %ld failed %ld\n", tested, skipped, failed);
    return failed ? 1 : (tested < 300 ? 3 : 0);
}

int main(int argc, char** argv)
{
    if (argc > 3 && std::string(argv[1]) == "--vectors") return check_vectors(argv[2], argv[3]);
    const int cases = argc > 1 ? atoi(argv[1]) : 6000;
    long tested = 0, skipped = 0, failed = 0, badstart = 0;
    for (int it = 0; it < cases; it++)
    {
        const int qmax = (it % 99 == 13) ? 4000 : (it % 9 == 4) ? 1024 : (it % 3 == 0) ? 384 : 256;   // the kernel runs with 256, 1024 and 4000
        const int alphabet = (it % 5 == 0) ? 2 : 4;
        const int homop = (it % 4 == 0) ? 4 : 0;
        int qlen;
        switch (it % 7) { case 0: qlen = rnd(1, 30); break; case 1: qlen = rnd(qmax - 20, qmax); break; case 2: qlen = rnd(180, 230); break; case 3: qlen = rnd(20, qmax); break; default: qlen = rnd(20, 200); }
        if (qlen > qmax) qlen = qmax;
        std::vector<uint8_t> q = random_seq(qlen, alphabet, homop);
        const int maxLen = (int)((double)qlen * 1.1 + 20.0);
        std::vector<uint8_t> s2;
        const bool isRC = (it & 1) != 0;
        const int mode = it % 11;
        if (mode == 10) s2 = random_seq(rnd(1, maxLen), alphabet, homop);   // unrelated
        else
        {
            std::vector<uint8_t> ext = q;
            const std::vector<uint8_t> flank = random_seq(rnd(0, 60), alphabet, homop);
            if (isRC) ext.insert(ext.begin(), flank.begin(), flank.end()); else ext.insert(ext.end(), flank.begin(), flank.end());
            s2 = mutate(ext, mode < 3 ? 0.02 : (mode < 8 ? 0.15 : 0.35), alphabet);
            if ((int)s2.size() > maxLen) { if (isRC) s2.erase(s2.begin(), s2.begin() + (s2.size() - maxLen)); else s2.resize(maxLen); }
            if (mode == 9 && s2.size() > 8) { const int cut = rnd(1, (int)s2.size() - 1); if (isRC) s2.erase(s2.begin(), s2.begin() + cut); else s2.resize(s2.size() - cut); }
        }
        if (s2.empty()) s2.push_back(0);
        const int mlen = (int)s2.size();
        const int k = std::min(qlen, std::min(mlen, rnd(13, 19)));
        int start_1 = isRC ? qlen - k : 0, start_2 = isRC ? mlen - k : 0;
        // arbitrary band placements (not only the two the caller uses): first column down to one cell at band row 200, last
        // column down to one cell at band row 0
        if (it % 13 == 12) { start_1 = 0; start_2 = rnd(-(BW - 1) - qlen / 2, std::max(mlen - qlen, 0) + qlen / 2) + HALF; }   // also bands that miss the matrix
        const int origin = (start_2 - start_1 + 1) - (HALF + 1);
        if (!eligible(qlen, mlen, origin, qmax)) { skipped++; continue; }
        tested++;
        bool ok = false;
        const pbo::PairOverlap ref = pbo::extendMatch(to_str(q), to_str(s2), start_1, start_2, 200, 1, -1, -8, &ok);
        HHost H; SHost S(s2); FHost F((size_t)qlen * words_per_col(mlen));
        auto qf = [&](int x) { return (int)q[x]; };
        int bi = 0, bj = 0;
        fill(qlen, mlen, origin, H, S, F, qf, bi, bj);
        if (!ok || bi <= 0)
        {
            if (ok != (bi > 0)) { failed++; printf("case %d: traceback start disagrees (oracle ok=%d, bi=%d)\n", it, (int)ok, bi); }
            badstart++;
            continue;
        }
        OHost O; int n = 0, ed = 0, i0 = 0, j0 = 0;
        traceback(qlen, mlen, origin, S, F, qf, bi, bj, O, n, ed, i0, j0);
        const std::string fwd(O.ops.rbegin(), O.ops.rend());
        if (fwd != ref.ops || n != ref.totalColumns || ed != ref.editDistance || i0 != ref.start[0] || j0 != ref.start[1] || bi - 1 != ref.end[0] ||
            bj - 1 != ref.end[1])
        {
            failed++;
            if (failed < 10)
                printf("case %d MISMATCH qlen=%d mlen=%d rc=%d origin=%d: n %d/%d ed %d/%d start (%d,%d)/(%d,%d) end (%d,%d)/(%d,%d)\n", it, qlen, mlen, (int)isRC,
                       origin, n, ref.totalColumns, ed, ref.editDistance, i0, j0, ref.start[0], ref.start[1], bi - 1, bj - 1, ref.end[0], ref.end[1]);
        }
    }
    printf("tested %ld skipped %ld (not eligible) empty-start %ld failed %ld\n", tested, skipped, badstart, failed);
    return failed ? 1 : (tested < cases / 2 ? 3 : 0);
}
