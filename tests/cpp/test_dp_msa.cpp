// Host check of the multiple-alignment column model (longreadselfcorrect_b200/csrc/pbsc_dp_msa.cuh): the same
// msa::consensus that dp_msa_kernel runs one job per thread, compiled here with g++,
//   (a) against the pile-ups answered by the reference's own MultipleAlignment (tests/golden/dp_units.txt "M" records and
//       dp_units.ref.txt, written by oracle/_ref/dp_dump), and
//   (b) against the oracle's padded-row restatement on random pile-ups (insertion-rich, homopolymers, rows starting with an
//       insertion, short rows).
// The alignments fed to it come from the oracle's extendMatch.  TEST INFRASTRUCTURE: the oracle is the checker.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include "../../longreadselfcorrect_b200/csrc/pbsc_dp_msa.cuh"
#include "../../oracle/pbsc_oracle.hpp"   // pulls in pbsc_oracle_dp.hpp (pbo::extendMatch, pbo::Msa)

using namespace pbsc;

struct HostCtx
{
    int minCall; uint8_t* o; uint32_t c;
    int min_call() const { return minCall; }
    uint8_t* out() const { return o; }
    uint32_t cap() const { return c; }
};

static uint8_t code(char c) { return (uint8_t)(c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3); }

struct RowIn { std::string s; int a, b; };

// returns the status of msa::consensus; `cons` receives the consensus on status 0
static int run_model(const std::string& q, const std::vector<RowIn>& rows_in, int minCall, std::string& cons)
{
    const uint32_t qlen = (uint32_t)q.size(), nr = (uint32_t)rows_in.size();
    size_t maxLen = 1;
    for (const auto& r : rows_in) maxLen = std::max(maxLen, r.s.size());
    const uint64_t seqBytes = (maxLen + 15) / 16 * 16, opsBytes = (qlen + maxLen + 15) / 16 * 16, rowBytes = seqBytes + opsBytes;
    std::vector<uint8_t> qc(qlen), rowbuf((size_t)nr * rowBytes + 16, 0);
    for (uint32_t x = 0; x < qlen; x++) qc[x] = code(q[x]);
    std::vector<DpRow> R(nr);
    for (uint32_t r = 0; r < nr; r++)
    {
        const pbo::PairOverlap ov = pbo::extendMatch(q, rows_in[r].s, rows_in[r].a, rows_in[r].b, 200, 1, -1, -8);
        uint8_t* buf = rowbuf.data() + (size_t)r * rowBytes;
        for (size_t x = 0; x < rows_in[r].s.size(); x++) buf[x] = code(rows_in[r].s[x]);
        const size_t n = ov.ops.size();
        if (n > opsBytes) { printf("ops do not fit their slot\n"); exit(2); }
        for (size_t x = 0; x < n; x++) buf[seqBytes + (n - 1 - x)] = (uint8_t)(ov.ops[x] == 'M' ? OP_M : ov.ops[x] == 'I' ? OP_I : OP_D);   // last column first
        R[r].job = 0; R[r].local = r; R[r].len = (uint32_t)rows_in[r].s.size(); R[r].seq_start = 0; R[r].nops = (uint32_t)n;
        R[r].start0 = ov.start[0]; R[r].start1 = ov.start[1]; R[r].pass = 1;
    }
    const uint64_t np = (uint64_t)qlen + 1;
    std::vector<uint16_t> baseCnt(np * 5, 0xABCD), startAt(np, 0xABCD);
    std::vector<uint32_t> head(np, 0xDEADBEEF), tail(np, 0xDEADBEEF);
    std::vector<GapCol> pool(dp_gap_cap(qlen));
    JobView v;
    v.q = qc.data(); v.rows = rowbuf.data(); v.baseCnt = baseCnt.data(); v.startAt = startAt.data(); v.head = head.data(); v.tail = tail.data();
    v.pool = pool.data(); v.rowBytes = rowBytes; v.seqBytes = seqBytes;
    std::vector<uint8_t> out((size_t)qlen * 4 + 512);
    const HostCtx ctx{minCall, out.data(), (uint32_t)out.size()};
    uint32_t n = 0;
    const int rc = msa::consensus(v, qlen, 0u, R.data(), nr, ctx, n);
    cons.clear();
    if (rc == 0) for (uint32_t x = 0; x < n; x++) cons.push_back("ACGT"[out[x]]);
    return rc;
}

static std::string oracle_consensus(const std::string& q, const std::vector<RowIn>& rows_in, int minCall)
{
    pbo::Msa msa;
    msa.addBase(q);
    for (const auto& r : rows_in) msa.addOverlap(r.s, pbo::extendMatch(q, r.s, r.a, r.b, 200, 1, -1, -8));
    return msa.consensus(minCall);
}

static int check_vectors(const char* in_path, const char* ref_path)
{
    std::ifstream in(in_path), ref(ref_path);
    if (!in || !ref) { printf("cannot open the vector files\n"); return 2; }
    std::string tag;
    long tested = 0, failed = 0, few = 0;
    while (in >> tag)
    {
        std::string line;
        if (tag == "A")
        {
            std::string s1, s2; int a, b;
            in >> s1 >> s2 >> a >> b;
            std::getline(ref, line);
            continue;
        }
        std::string q; int minCall, n;
        in >> q >> minCall >> n;
        std::vector<RowIn> rows(n);
        for (int i = 0; i < n; i++) in >> rows[i].s >> rows[i].a >> rows[i].b;
        std::getline(ref, line);
        std::istringstream rs(line);
        std::string rtag, want; int nrows;
        rs >> rtag >> nrows >> want;
        if (rtag != "M" || nrows != n + 1) { printf("answer file out of step\n"); return 2; }
        std::string got;
        const int rc = run_model(q, rows, minCall, got);
        if (rc == 1) { few++; continue; }   // fewer than three rows: the caller never builds a consensus (PacBioSelfCorrectionProcess.cpp:238)
        tested++;
        if (rc != 0 || got != want)
        {
            failed++;
            if (failed < 5) printf("pile-up %ld MISMATCH (status %d)\n  want %s\n  got  %s\n", tested, rc, want.c_str(), got.c_str());
        }
    }
    printf("pile-ups tested %ld, fewer than three rows %ld, failed %ld\n", tested, few, failed);
    return failed ? 1 : (tested < 40 ? 3 : 0);
}

static std::mt19937_64 rng(2024);
static int rnd(int lo, int hi) { return lo + (int)(rng() % (uint64_t)(hi - lo + 1)); }
static std::string random_seq(int n, int alphabet, int homop)
{
    std::string s;
    while ((int)s.size() < n)
    {
        const char c = "ACGT"[rnd(0, alphabet - 1)];
        const int run = homop ? rnd(1, homop) : 1;
        for (int r = 0; r < run && (int)s.size() < n; r++) s.push_back(c);
    }
    return s;
}
static std::string mutate(const std::string& s, double err, int alphabet)
{
    std::string o;
    for (char ch : s)
    {
        const double u = (double)(rng() % 1000000) / 1e6;
        if (u < err * 0.55) { o.push_back("ACGT"[rnd(0, alphabet - 1)]); o.push_back(ch); }
        else if (u < err * 0.85) { }
        else if (u < err) o.push_back("ACGT"[rnd(0, alphabet - 1)]);
        else o.push_back(ch);
    }
    return o;
}

int main(int argc, char** argv)
{
    if (argc > 3 && std::string(argv[1]) == "--vectors") return check_vectors(argv[2], argv[3]);
    const int cases = argc > 1 ? atoi(argv[1]) : 2000;
    long tested = 0, failed = 0, limits = 0;
    for (int it = 0; it < cases; it++)
    {
        const int alphabet = (it % 5 == 0) ? 2 : 4, homop = (it % 4 == 0) ? 4 : 0;
        const int qlen = (it % 6 == 0) ? rnd(20, 60) : rnd(40, 260);
        const std::string q = random_seq(qlen, alphabet, homop);
        const int k = std::min(qlen, rnd(13, 19));
        const int nrows = rnd(3, (it % 3 == 0) ? 60 : 20);
        std::vector<RowIn> rows;
        for (int r = 0; r < nrows; r++)
        {
            const bool isRC = rnd(0, 1) != 0;
            std::string ext = q;
            const std::string flank = random_seq(rnd(0, 30), alphabet, homop);
            if (isRC) ext = flank + ext; else ext += flank;
            std::string s = mutate(ext, (it % 7 == 0) ? 0.3 : 0.14, alphabet);
            const int maxLen = (int)((double)qlen * 1.1 + 20.0);
            if ((int)s.size() > maxLen) s = isRC ? s.substr(s.size() - maxLen) : s.substr(0, maxLen);
            if (rnd(0, 9) == 0 && s.size() > 30) s = isRC ? s.substr(rnd(1, (int)s.size() - 20)) : s.substr(0, rnd(20, (int)s.size() - 1));   // the read ends early
            if (s.empty()) s = "A";
            const int kk = std::min(k, (int)s.size());
            rows.push_back(RowIn{s, isRC ? qlen - kk : 0, isRC ? (int)s.size() - kk : 0});
        }
        const int minCall = (it % 2) ? 15 : rnd(3, 40);
        std::string got;
        const int rc = run_model(q, rows, minCall, got);
        if (rc == 2) { limits++; continue; }   // more inserted columns than this build holds: reported, never truncated
        tested++;
        const std::string want = oracle_consensus(q, rows, minCall);
        if (rc != 0 || got != want)
        {
            failed++;
            if (failed < 5) printf("case %d MISMATCH (status %d, %d rows, query %d)\n  want %s\n  got  %s\n", it, rc, nrows, qlen, want.c_str(), got.c_str());
        }
    }
    printf("tested %ld, outside limits %ld, failed %ld\n", tested, limits, failed);
    return failed ? 1 : (tested < cases / 2 ? 3 : 0);
}
