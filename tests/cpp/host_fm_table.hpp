// host_fm_table.hpp — TEST HELPER: decodes a reference-format run-length BWT file (BWTReaderBinary.cpp:55-85, RLUnit.h:13-16)
// into the block layout of longreadselfcorrect_b200/csrc/fm_table.cuh, in host memory, the way decode_runs (pbsc_index.cu)
// does it for the device.  The primitives under test come from fm_table.cuh itself.
#ifndef HOST_FM_TABLE_HPP
#define HOST_FM_TABLE_HPP
#include <cstdint>
#include <fstream>
#include <string>
#include <vector>
#include "../../longreadselfcorrect_b200/csrc/fm_table.cuh"

using namespace pbsc;

struct HostTable
{
    std::vector<FmBlock> blocks; std::vector<uint64_t> dmask; std::vector<uint32_t> dpos;
    FmTable t; uint64_t n_strings = 0;
    bool load(const std::string& path)
    {
        std::ifstream in(path.c_str(), std::ios::binary);
        if (!in) return false;
        uint16_t magic = 0; uint64_t nstr = 0, nsym = 0, nruns = 0; int32_t flag = 0;
        in.read((char*)&magic, 2); in.read((char*)&nstr, 8); in.read((char*)&nsym, 8); in.read((char*)&nruns, 8); in.read((char*)&flag, 4);
        if (!in || magic != 0xCACA) return false;
        n_strings = nstr;
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

#endif
