// TEST INFRASTRUCTURE ONLY (oracle/): a tiny driver linked against the reference's own
// objects.  The reference never prints SA intervals, so this exposes
// BWTAlgorithms::findInterval(const BWT*, w) (SuffixTools/BWTAlgorithms.cpp:14-31) on an
// index file the reference itself loads (RLBWT ctor, SuffixTools/RLBWT.cpp:23-32).
//
//   fm_dump FILE.bwt QUERIES            one "lower upper" line per query line
//   fm_dump --time T FILE.bwt QUERIES   time findInterval under `omp parallel for` with T threads;
//                                       prints "queries seconds Mq/s steps checksum"
#include <omp.h>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "BWT.h"
#include "BWTAlgorithms.h"
#include "Timer.h"

int main(int argc, char** argv)
{
    int threads = 0, a = 1;
    if (argc > 2 && std::string(argv[1]) == "--time") { threads = atoi(argv[2]); a = 3; }
    if (argc - a < 2) { fprintf(stderr, "usage: fm_dump [--time T] FILE.bwt QUERIES\n"); return 2; }
    BWT* pBWT = new BWT(argv[a], BWT::DEFAULT_SAMPLE_RATE_SMALL);
    std::vector<std::string> q;
    { std::ifstream in(argv[a + 1]); std::string s; while (std::getline(in, s)) if (!s.empty()) q.push_back(s); }
    if (threads <= 0)
    {
        for (size_t i = 0; i < q.size(); i++)
        {
            BWTInterval iv = BWTAlgorithms::findInterval(pBWT, q[i]);
            printf("%ld %ld\n", (long)iv.lower, (long)iv.upper);
        }
    }
    else
    {
        omp_set_num_threads(threads);
        unsigned long long checksum = 0, steps = 0;
        Timer t("fm", true);
        #pragma omp parallel for schedule(static) reduction(+:checksum,steps)
        for (size_t i = 0; i < q.size(); i++)
        {
            int cnt[4] = {0, 0, 0, 0};
            BWTInterval iv = BWTAlgorithms::findInterval(pBWT, q[i], cnt);
            checksum += (unsigned long long)iv.lower * 31u + (unsigned long long)iv.upper;
            steps += cnt[0] + cnt[1] + cnt[2] + cnt[3] - 1;   // updateInterval calls executed
        }
        double s = t.getElapsedWallTime();
        printf("%zu %.6f %.4f %llu %llu\n", q.size(), s, q.size() / s / 1e6, steps, checksum);
    }
    delete pBWT;
    return 0;
}
