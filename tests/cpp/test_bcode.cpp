// Host check of csrc/pbsc_bcode.h (the product's `--onlyseed` arithmetic) against the unmodified reference:
//   test_bcode READS.fa SEEDS.tsv BARCODE OUT_TOTAL_SEED      (TOTAL line on stdout)
// SEEDS.tsv is the reference's own seed dump (tests/golden/tiny.seeds.tsv: "#id" then "seed<TAB>freq<TAB>start<TAB>repeat"),
// so the seeds are exactly the ones the reference validated when it wrote tests/golden/tiny.onlyseed.*.
#include <stdio.h>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "../../longreadselfcorrect_b200/csrc/pbsc_bcode.h"

// kmercheck mode:  test_bcode --kmercheck READS.fa BARCODE FREQS COV LOWER UPPER STEP OUT_TOTAL_BOX OUT_VALUE_BOX
// FREQS holds one line "read<TAB>block<TAB>k<TAB>pos<TAB>frequency" per k-mer of every barcode block (the caller computes the
// frequencies with the oracle's backward search); the barcode arithmetic and the five-number summaries are the product's.
static int kmercheck_mode(int argc, char** argv)
{
    if (argc != 11) { fprintf(stderr, "usage: test_bcode --kmercheck READS.fa BARCODE FREQS COV LOWER UPPER STEP TOTAL_BOX VALUE_BOX\n"); return 2; }
    std::map<std::string, std::string> reads;
    {
        std::ifstream f(argv[2]);
        std::string line, cur;
        while (std::getline(f, line))
        {
            if (line.empty()) continue;
            if (line[0] == '>') { cur = line.substr(1, line.find_first_of(" \t") == std::string::npos ? std::string::npos : line.find_first_of(" \t") - 1); reads[cur]; }
            else reads[cur] += line;
        }
    }
    pbsc::bcode::Table table;
    std::string err;
    if (!pbsc::bcode::load(argv[3], table, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    const int cov = atoi(argv[5]), lower = atoi(argv[6]), upper = atoi(argv[7]), step = atoi(argv[8]);
    std::map<int, pbsc::bcode::Histogram> crt, wrong;
    std::ifstream fr(argv[4]);
    std::string id;
    int block, k, pos;
    long freq;
    while (fr >> id >> block >> k >> pos >> freq)
    {
        if (freq == 1) continue;
        const bool ok = pbsc::bcode::validate(pos, k, table[id][(size_t)block], reads[id]);
        (ok ? crt : wrong)[k].add((int)freq);
    }
    std::ofstream tb(argv[9]), vb(argv[10]);
    for (int kk = lower; kk <= upper; kk += step) pbsc::bcode::compare(tb, vb, cov, kk, crt[kk], wrong[kk]);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc > 1 && std::string(argv[1]) == "--kmercheck") return kmercheck_mode(argc, argv);
    if (argc != 5) { fprintf(stderr, "usage: test_bcode READS.fa SEEDS.tsv BARCODE OUT\n"); return 2; }
    std::vector<std::pair<std::string, std::string> > reads;
    {
        std::ifstream f(argv[1]);
        std::string line;
        while (std::getline(f, line))
        {
            if (line.empty()) continue;
            if (line[0] == '>') reads.push_back(std::make_pair(line.substr(1, line.find_first_of(" \t") == std::string::npos ? std::string::npos : line.find_first_of(" \t") - 1), std::string()));
            else reads.back().second += line;
        }
    }
    std::map<std::string, std::vector<std::pair<int, int> > > seeds;
    {
        std::ifstream f(argv[2]);
        std::string line, cur;
        while (std::getline(f, line))
        {
            if (line.empty()) continue;
            if (line[0] == '#') { cur = line.substr(1); seeds[cur]; continue; }
            std::istringstream in(line);
            std::string s, rep; int freq, start;
            in >> s >> freq >> start >> rep;
            seeds[cur].push_back(std::make_pair(start, (int)s.size()));
        }
    }
    pbsc::bcode::Table table;
    std::string err;
    if (!pbsc::bcode::load(argv[3], table, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    FILE* out = fopen(argv[4], "w");
    if (!out) return 1;
    size_t total[3] = {0, 0, 0};
    for (size_t r = 0; r < reads.size(); r++)
    {
        size_t status[3] = {0, 0, 0};
        pbsc::bcode::Table::const_iterator it = table.find(reads[r].first);
        const std::vector<std::pair<int, int> >& sv = seeds[reads[r].first];
        for (size_t i = 0; i < sv.size(); i++) status[pbsc::bcode::classify(it == table.end() ? nullptr : &it->second, sv[i].first, sv[i].second, reads[r].second)]++;
        pbsc::bcode::summarize(out, status, reads[r].first);
        for (int k = 0; k < 3; k++) total[k] += status[k];
    }
    fclose(out);
    pbsc::bcode::summarize(stdout, total, "TOTAL");
    return 0;
}
