// Host check of the rank primitives of longreadselfcorrect_b200/csrc/fm_table.cuh (occ, occ4, occ_pair, count_dollars,
// update_interval: the code every kernel calls, compiled here for the host) against naive counting over random BWTs with '$'
// symbols, block boundaries at every offset and lengths that are and are not multiples of 64.
//   occ(c, p) = occurrences of base c in bwt[0, p) = RLBWT::getOcc(c, p - 1) (SuffixTools/RLBWT.h:121-140)
//   update_interval = BWTAlgorithms::updateInterval (SuffixTools/BWTAlgorithms.h:66-72) on half-open intervals
// The block layout is built here the way decode_runs (pbsc_index.cu) builds it.
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>
#include "../../longreadselfcorrect_b200/csrc/fm_table.cuh"

using namespace pbsc;

int main()
{
    std::mt19937_64 rng(99);
    long checks = 0, bad = 0;
    for (int it = 0; it < 60; it++)
    {
        const uint64_t n = (it % 5 == 0) ? 64ull * (1 + rng() % 40) : 1 + rng() % 3000;
        const int dollar_per = (it % 3 == 0) ? 7 : 90;   // dense and sparse '$'
        std::vector<int> bwt(n);   // 0 = '$', 1..4 = A C G T
        for (auto& s : bwt) s = (rng() % dollar_per == 0) ? 0 : 1 + (int)(rng() % 4);
        if (it % 7 == 0) for (uint64_t x = 0; x < n; x++) bwt[x] = (x / 50) % 2 ? 1 : bwt[x];   // long runs of A next to '$'
        const uint64_t nb = n / 64 + 1;
        std::vector<FmBlock> blocks(nb, FmBlock{{0, 0, 0, 0}, {0, 0, 0, 0}});
        std::vector<uint64_t> dmask(nb, 0);
        std::vector<uint32_t> dpos;
        uint64_t cnt[5] = {0, 0, 0, 0, 0};
        for (uint64_t pos = 0; pos < n; pos++)
        {
            const uint64_t b = pos >> 6; const uint32_t j = (uint32_t)pos & 63u;
            if (j == 0) { blocks[b].cnt[0] = (uint32_t)cnt[1]; blocks[b].cnt[1] = (uint32_t)cnt[2]; blocks[b].cnt[2] = (uint32_t)cnt[3]; blocks[b].cnt[3] = (uint32_t)cnt[4]; }
            if (bwt[pos] == 0) { dpos.push_back((uint32_t)pos); blocks[b].cnt[0] |= 0x80000000u; dmask[b] |= 1ull << j; }
            else blocks[b].bases[j >> 4] |= (uint32_t)(bwt[pos] - 1) << (2 * (j & 15));
            cnt[bwt[pos]]++;
        }
        if ((n & 63) == 0) { FmBlock& h = blocks[nb - 1]; h.cnt[0] = (uint32_t)cnt[1]; h.cnt[1] = (uint32_t)cnt[2]; h.cnt[2] = (uint32_t)cnt[3]; h.cnt[3] = (uint32_t)cnt[4]; }
        dpos.push_back(0);   // the product allocates one spare entry
        FmTable t;
        t.blocks = blocks.data(); t.dollar_pos = dpos.data(); t.dollar_mask = dmask.data(); t.n = n; t.n_dollar = (uint32_t)cnt[0];
        t.C[0] = cnt[0]; t.C[1] = t.C[0] + cnt[1]; t.C[2] = t.C[1] + cnt[2]; t.C[3] = t.C[2] + cnt[3];
        for (int c = 0; c < 4; c++) t.total[c] = cnt[c + 1];
        // naive prefix counts
        std::vector<uint64_t> pre[5];
        for (int s = 0; s < 5; s++) { pre[s].assign(n + 1, 0); for (uint64_t x = 0; x < n; x++) pre[s][x + 1] = pre[s][x] + (bwt[x] == s); }
        for (uint64_t p = 0; p <= n; p++)
            for (int c = 0; c < 4; c++)
            {
                checks++;
                if (occ(t, c, p) != pre[c + 1][p]) { if (bad++ < 5) printf("occ(%d, %llu) = %llu, naive %llu (n = %llu)\n", c, (unsigned long long)p, (unsigned long long)occ(t, c, p), (unsigned long long)pre[c + 1][p], (unsigned long long)n); }
            }
        for (uint64_t p = 0; p <= n; p++)
        {
            uint64_t r4[4];
            occ4(t, p, r4);   // the walk kernel's probe: all four counts from one sector
            checks++;
            for (int c = 0; c < 4; c++) if (r4[c] != pre[c + 1][p]) { if (bad++ < 5) printf("occ4(%llu)[%d] wrong\n", (unsigned long long)p, c); }
        }
        for (int q = 0; q < 4000; q++)
        {
            uint64_t lo = rng() % (n + 1), hi = (q % 3 == 0) ? lo + rng() % 70 : rng() % (n + 1);
            if (hi > n) hi = n;
            if (hi < lo) std::swap(lo, hi);
            const int c = (int)(rng() % 4);
            uint64_t a, b;
            occ_pair(t, c, lo, hi, a, b);
            checks++;
            if (a != pre[c + 1][lo] || b != pre[c + 1][hi]) { if (bad++ < 5) printf("occ_pair(%d, %llu, %llu) wrong\n", c, (unsigned long long)lo, (unsigned long long)hi); }
            if (count_dollars(t, lo, hi) != pre[0][hi] - pre[0][lo]) { if (bad++ < 5) printf("count_dollars(%llu, %llu) wrong\n", (unsigned long long)lo, (unsigned long long)hi); }
            const Interval r = update_interval(t, Interval{lo, hi}, c);
            if (r.lo != t.C[c] + pre[c + 1][lo] || r.hi != t.C[c] + pre[c + 1][hi]) { if (bad++ < 5) printf("update_interval wrong\n"); }
        }
    }
    printf("%s: %ld checks, %ld wrong\n", bad ? "FAILED" : "ok", checks, bad);
    return bad ? 1 : 0;
}
