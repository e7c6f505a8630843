// Host check of lf_step (longreadselfcorrect_b200/csrc/fm_table.cuh: BWT symbol + LF-mapping from one sector, the step of
// dp_retrieve_kernel = LongReadOverlap::retrieveStr, PacBio/LongReadOverlap.cpp:697-747) on the reference-built index:
// LF-walking from each of the first n_strings rows (the suffixes that start with '$') until the next '$' must spell every read
// of tests/golden/tiny.reads.fa exactly once — backwards on tiny.bwt (BWT of the reads), forwards on tiny.rbwt (BWT of the
// reversed reads).
//     test_fm_lf GOLDEN_DIR
#include <algorithm>
#include <cstdio>
#include "host_fm_table.hpp"

int main(int argc, char** argv)
{
    if (argc < 2) { printf("usage: test_fm_lf GOLDEN_DIR\n"); return 2; }
    const std::string dir = argv[1];
    std::vector<std::string> reads;
    { std::ifstream f((dir + "/tiny.reads.fa").c_str()); std::string s; while (std::getline(f, s)) if (!s.empty() && s[0] != '>') reads.push_back(s); }
    std::vector<std::string> want = reads;
    std::sort(want.begin(), want.end());
    long bad = 0, steps = 0;
    for (int rev = 0; rev < 2; rev++)
    {
        HostTable H;
        if (!H.load(dir + (rev ? "/tiny.rbwt" : "/tiny.bwt"))) { printf("cannot load the index\n"); return 2; }
        if (H.n_strings != reads.size()) { printf("index holds %llu strings, the FASTA %zu\n", (unsigned long long)H.n_strings, reads.size()); return 2; }
        std::vector<std::string> got;
        for (uint64_t row = 0; row < H.n_strings; row++)
        {
            uint64_t idx = row;
            std::string s;
            for (;;)
            {
                const int c = lf_step(H.t, idx);
                if (c < 0) break;
                s.push_back("ACGT"[c]);
                steps++;
                if (s.size() > 100000) { printf("LF walk does not end\n"); return 1; }
            }
            if (!rev) std::reverse(s.begin(), s.end());   // the BWT of the reads spells a read from its last base to its first
            got.push_back(s);
        }
        std::sort(got.begin(), got.end());
        if (got != want) { bad++; printf("%s: the LF walks do not spell the read set\n", rev ? "tiny.rbwt" : "tiny.bwt"); }
    }
    printf("%s: %zu reads spelled on both strands, %ld LF steps\n", bad ? "FAILED" : "ok", reads.size(), steps);
    return bad ? 1 : 0;
}
