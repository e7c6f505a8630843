// Host check of the backward search of longreadselfcorrect_b200/csrc/fm_table.cuh (init_interval / update_interval: the code
// every kernel calls, compiled here for the host) against the REFERENCE's own BWTAlgorithms::findInterval
// (SuffixTools/BWTAlgorithms.cpp:14-31): tests/golden/tiny.fm_queries.txt holds 2000 queries, tiny.fm_bwt.txt / tiny.fm_rbwt.txt
// the (lower, upper) pairs that oracle/_ref/fm_dump — linked against the reference's objects — printed for them on the index
// files tiny.bwt / tiny.rbwt written by the reference's `stride index`.  The run-length file (BWTReaderBinary.cpp:55-85,
// RLUnit.h:13-16) is decoded into the block layout the way decode_runs (pbsc_index.cu) does it.
//     test_fm_search GOLDEN_DIR
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "host_fm_table.hpp"

using namespace pbsc;

static int code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }

int main(int argc, char** argv)
{
    if (argc < 2) { printf("usage: test_fm_search GOLDEN_DIR\n"); return 2; }
    const std::string dir = argv[1];
    std::vector<std::string> queries;
    { std::ifstream q((dir + "/tiny.fm_queries.txt").c_str()); std::string s; while (std::getline(q, s)) if (!s.empty()) queries.push_back(s); }
    long checked = 0, bad = 0, empty = 0;
    for (const char* ext : {"bwt", "rbwt"})
    {
        HostTable H;
        if (!H.load(dir + "/tiny." + ext)) { printf("cannot load tiny.%s\n", ext); return 2; }
        std::ifstream ans((dir + "/tiny.fm_" + ext + ".txt").c_str());
        for (const std::string& w : queries)
        {
            long long lower, upper;
            if (!(ans >> lower >> upper)) { printf("answers of tiny.fm_%s.txt end early\n", ext); return 2; }
            int j = (int)w.size() - 1;
            Interval iv = init_interval(H.t, code(w[j]));
            for (--j; j >= 0; --j)
            {
                iv = update_interval(H.t, iv, code(w[j]));
                if (!iv.valid()) break;   // BWTAlgorithms.cpp:27
            }
            checked++;
            if (!iv.valid()) empty++;
            // half-open [lo, hi) against the reference's inclusive (lower, upper): also its raw values at the early break
            if ((long long)iv.lo != lower || (long long)iv.hi - 1 != upper) { if (bad++ < 5) printf("%s %s: got (%llu, %lld), reference (%lld, %lld)\n", ext, w.c_str(), (unsigned long long)iv.lo, (long long)iv.hi - 1, lower, upper); }
        }
    }
    printf("%s: %ld intervals (%ld empty), %ld differ from the reference\n", bad ? "FAILED" : "ok", checked, empty, bad);
    return bad ? 1 : (checked < 4000 ? 3 : 0);
}
