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
#include "../../longreadselfcorrect_b200/csrc/fm_table.cuh"

using namespace pbsc;

struct HostTable
{
    std::vector<FmBlock> blocks; std::vector<uint64_t> dmask; std::vector<uint32_t> dpos;
    FmTable t;
    bool load(const std::string& path)
    {
        std::ifstream in(path.c_str(), std::ios::binary);
        if (!in) return false;
        uint16_t magic = 0; uint64_t nstr = 0, nsym = 0, nruns = 0; int32_t flag = 0;
        in.read((char*)&magic, 2); in.read((char*)&nstr, 8); in.read((char*)&nsym, 8); in.read((char*)&nruns, 8); in.read((char*)&flag, 4);
        if (!in || magic != 0xCACA) return false;
        std::vector<uint8_t> runs(nruns);
        in.read((char*)runs.data(), (std::streamsize)nruns);
        if ((uint64_t)in.gcount() != nruns) return false;
        const uint64_t nb = nsym / 64 + 1;
        blocks.assign(nb, FmBlock{{0, 0, 0, 0}, {0, 0, 0, 0}});
        dmask.assign(nb, 0);
        uint64_t cnt[5] = {0, 0, 0, 0, 0}, pos = 0;
        for (uint8_t u : runs)
        {
            const uint32_t sym = u >> 5, len = u & 0x1f;
            for (uint32_t i = 0; i < len; i++, pos++)
            {
                const uint64_t b = pos >> 6; const uint32_t j = (uint32_t)pos & 63u;
                if (j == 0) { blocks[b].cnt[0] = (uint32_t)cnt[1]; blocks[b].cnt[1] = (uint32_t)cnt[2]; blocks[b].cnt[2] = (uint32_t)cnt[3]; blocks[b].cnt[3] = (uint32_t)cnt[4]; }
                if (sym == 0) { dpos.push_back((uint32_t)pos); blocks[b].cnt[0] |= 0x80000000u; dmask[b] |= 1ull << j; }
                else blocks[b].bases[j >> 4] |= (sym - 1) << (2 * (j & 15));
                cnt[sym]++;
            }
        }
        if (pos != nsym) return false;
        if ((nsym & 63) == 0) { FmBlock& h = blocks[nb - 1]; h.cnt[0] = (uint32_t)cnt[1]; h.cnt[1] = (uint32_t)cnt[2]; h.cnt[2] = (uint32_t)cnt[3]; h.cnt[3] = (uint32_t)cnt[4]; }
        dpos.push_back(0);
        t.blocks = blocks.data(); t.dollar_pos = dpos.data(); t.dollar_mask = dmask.data(); t.n = nsym; t.n_dollar = (uint32_t)cnt[0];
        t.C[0] = cnt[0]; t.C[1] = t.C[0] + cnt[1]; t.C[2] = t.C[1] + cnt[2]; t.C[3] = t.C[2] + cnt[3];
        for (int c = 0; c < 4; c++) t.total[c] = cnt[c + 1];
        return true;
    }
};

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
