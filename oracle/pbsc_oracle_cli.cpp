// TEST INFRASTRUCTURE ONLY — command-line front end of the CPU oracle (see pbsc_oracle.hpp).
//
//   pbsc_oracle pbcorrect [reference options] [--dump FILE] [--threads T] READS
//       mirrors `stride pbcorrect` (StriDe/PacBioSelfCorrection.cpp:145-434) for --nodp runs:
//       writes DIR/correct.fa, DIR/discard.fa, DIR/threshold-table and, with --debugseed,
//       DIR/seed/<id>.seed, DIR/seed/error/<id>.seed, DIR/extend/<id>.{log,ext,dp};
//       prints the reference's stdout summary.  --dump writes one JSON line per read with
//       seeds / per-pair walk records / pieces for structured parity tests.
//   pbsc_oracle findinterval FILE.bwt QUERIES     "lower upper" per query (cf. oracle/_ref/fm_dump)
//   pbsc_oracle occcount ...pbcorrect args...     like pbcorrect, also prints rank queries issued
//   pbsc_oracle dpcheck FILE                      extendMatch / multiple-alignment records (cf. oracle/_ref/dp_dump)
#include <getopt.h>
#include <omp.h>
#include <sys/stat.h>
#include <chrono>
#include "pbsc_oracle.hpp"

using namespace pbo;

static void writeSeeds(std::ostream& out, const SeedVector& vec)   // SeedFeature.cpp:11-20
{
    for (const auto& s : vec)
        out << s.seedStr << '\t' << s.maxFixedMerFreq << '\t' << s.seedStartPos << '\t' << (s.isRepeat ? "Yes" : "No") << '\n';
}

static int findIntervalMain(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: pbsc_oracle findinterval FILE.bwt QUERIES\n"); return 2; }
    FMIndex fm; std::string err;
    if (!fm.load(argv[2], &err)) { std::cerr << err << "\n"; return 1; }
    std::ifstream in(argv[3]); std::string s;
    while (std::getline(in, s)) if (!s.empty()) { BWTInterval iv = findInterval(&fm, s); printf("%ld %ld\n", (long)iv.lower, (long)iv.upper); }
    return 0;
}

// run-length form of an expanded cigar, as Overlapper::compactCigar prints it
static std::string compactOps(const std::string& ops)
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

// same record format as oracle/refbuild/dp_dump.cpp, answered by the restatement in pbsc_oracle_dp.hpp
static int dpCheckMain(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: pbsc_oracle dpcheck FILE\n"); return 2; }
    std::ifstream in(argv[2]);
    std::string tag;
    while (in >> tag)
    {
        if (tag == "A")
        {
            std::string s1, s2; int a, b;
            in >> s1 >> s2 >> a >> b;
            PairOverlap o = extendMatch(s1, s2, a, b, 200, 1, -1, -8);
            std::cout << "A " << o.score << " " << o.start[0] << " " << o.end[0] << " " << o.start[1] << " " << o.end[1] << " " << o.editDistance << " "
                      << o.totalColumns << " " << compactOps(o.ops) << "\n";
        }
        else if (tag == "M")
        {
            std::string q; int minCall, n;
            in >> q >> minCall >> n;
            Msa msa;
            msa.addBase(q);
            for (int i = 0; i < n; i++)
            {
                std::string s; int a, b;
                in >> s >> a >> b;
                msa.addOverlap(s, extendMatch(q, s, a, b, 200, 1, -1, -8));
            }
            std::cout << "M " << msa.rows.size() << " " << msa.consensus(minCall) << "\n";
        }
    }
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: pbsc_oracle {pbcorrect|findinterval|dpcheck} ...\n"); return 2; }
    std::string cmd = argv[1];
    if (cmd == "findinterval") return findIntervalMain(argc, argv);
    if (cmd == "dpcheck") return dpCheckMain(argc, argv);
    if (cmd != "pbcorrect") { fprintf(stderr, "unknown command %s\n", cmd.c_str()); return 2; }

    Params P;
    std::string prefix, dumpFile;
    int threads = 1;
    enum { OPT_SPLIT = 256, OPT_DEBUGSEED, OPT_NODP, OPT_DUMP, OPT_THREADS, OPT_DEBUGEXTEND };
    static const struct option longopts[] = {
        {"thread", required_argument, nullptr, 't'}, {"prefix", required_argument, nullptr, 'p'},
        {"output", required_argument, nullptr, 'o'}, {"PBcoverage", required_argument, nullptr, 'c'},
        {"error-rate", required_argument, nullptr, 'e'}, {"kmer-size", required_argument, nullptr, 'k'},
        {"unique-offset", required_argument, nullptr, 'u'}, {"repeat-offset", required_argument, nullptr, 'r'},
        {"next-target", required_argument, nullptr, 'n'}, {"max-leaves", required_argument, nullptr, 'l'},
        {"idmer-length", required_argument, nullptr, 'i'}, {"min-kmer-size", required_argument, nullptr, 's'},
        {"genome", required_argument, nullptr, 'g'}, {"mode", required_argument, nullptr, 'm'},
        {"split", no_argument, nullptr, OPT_SPLIT}, {"debugseed", no_argument, nullptr, OPT_DEBUGSEED},
        {"debugextend", no_argument, nullptr, OPT_DEBUGEXTEND},
        {"nodp", no_argument, nullptr, OPT_NODP}, {"dump", required_argument, nullptr, OPT_DUMP},
        {"threads", required_argument, nullptr, OPT_THREADS}, {nullptr, 0, nullptr, 0}};
    optind = 2;
    for (int c; (c = getopt_long(argc, argv, "t:p:o:c:e:k:u:r:n:l:i:s:g:m:v", longopts, nullptr)) != -1;)
    {
        std::istringstream arg(optarg != nullptr ? optarg : "");
        switch (c)
        {
            case 't': break;   // reference thread count: output order here is always the -t 1 order
            case 'p': arg >> prefix; break;
            case 'o': arg >> P.directory; break;
            case 'c': arg >> P.PBcoverage; break;
            case 'e': arg >> P.ErrorRate; break;
            case 'k': arg >> P.startKmerLen; P.Adjust = true; break;
            case 'u': arg >> P.offset[1]; P.Adjust = true; break;
            case 'r': arg >> P.offset[2]; P.Adjust = true; break;
            case 'n': arg >> P.nextTarget; break;
            case 'l': arg >> P.maxLeaves; break;
            case 'i': arg >> P.idmerLen; break;
            case 's': arg >> P.minKmerLen; break;
            case 'g': arg >> P.genome; break;
            case 'm': arg >> P.mode; P.Manual = true; break;
            case OPT_SPLIT: P.Split = true; break;
            case OPT_DEBUGSEED: P.DebugSeed = true; break;
            case OPT_DEBUGEXTEND: break;
            case OPT_NODP: P.NoDp = true; break;
            case OPT_DUMP: dumpFile = optarg; break;
            case OPT_THREADS: threads = atoi(optarg); break;
            default: break;
        }
    }
    if (argc - optind != 1 || prefix.empty() || P.directory.empty()) { fprintf(stderr, "pbsc_oracle pbcorrect: bad arguments\n"); return 1; }
    std::string readsFile = argv[optind];
    P.directory += "/";
    std::vector<std::string> subdir = {""};
    if (P.DebugSeed) subdir = {"extend/", "seed/error/"};
    for (auto& d : subdir) if (system(("mkdir -p " + P.directory + d).c_str()) != 0) return 1;

    FMIndex bwt, rbwt; std::string err;
    if (!bwt.load(prefix + ".bwt", &err) || !rbwt.load(prefix + ".rbwt", &err)) { std::cerr << err << "\n"; return 1; }
    P.indices.pBWT = &bwt; P.indices.pRBWT = &rbwt;
    P.derive();

    std::vector<std::pair<std::string, std::string>> reads;
    { FastaReader rd(readsFile); std::string id, seq; while (rd.get(id, seq)) reads.push_back(std::make_pair(id, seq)); }

    std::vector<ReadResult> results(reads.size());
    auto t0 = std::chrono::steady_clock::now();
    uint64_t occTotal = 0;
    omp_set_num_threads(std::max(1, threads));
    #pragma omp parallel for schedule(dynamic, 1) reduction(+:occTotal)
    for (size_t i = 0; i < reads.size(); i++)
    {
        Corrector c(P);
        uint64_t before = OccCounter::n();
        results[i] = c.process(reads[i].first, reads[i].second);
        occTotal += OccCounter::n() - before;
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // PacBioSelfCorrectionPostProcess — PacBioSelfCorrectionProcess.cpp:250-370
    std::ofstream correct((P.directory + "correct.fa").c_str()), discard((P.directory + "discard.fa").c_str());
    std::ofstream dump; if (!dumpFile.empty()) dump.open(dumpFile.c_str());
    int64_t totalReadsLen = 0, correctedLen = 0, totalSeedNum = 0, totalWalkNum = 0, highErrorNum = 0, exceedDepthNum = 0,
            exceedLeaveNum = 0, FMNum = 0, DPNum = 0, seedDis = 0;
    for (size_t i = 0; i < reads.size(); i++)
    {
        const ReadResult& r = results[i];
        if (P.DebugSeed)
        {
            if (r.seedWritten) { std::ofstream s((P.directory + "seed/" + r.readid + ".seed").c_str()); writeSeeds(s, r.seeds); }
            if (r.outcastWritten) { std::ofstream e((P.directory + "seed/error/" + r.readid + ".seed").c_str()); writeSeeds(e, r.outcastSeeds); }
            if (r.seedWritten)
            {
                std::ofstream lg((P.directory + "extend/" + r.readid + ".log").c_str());
                for (size_t p = 0; p < r.ratioLog.size(); p++) lg << p << '\t' << r.ratioLog[p] << '\n';
            }
            if (r.seeds.size() >= 2)
            {
                std::ofstream x((P.directory + "extend/" + r.readid + ".ext").c_str());
                std::ofstream d((P.directory + "extend/" + r.readid + ".dp").c_str());
                for (auto& row : r.extLog) x << row.first.first << "\t" << row.first.second << "\t" << row.second << "\n";
                for (auto& row : r.dpFailLog) d << row.first << "\t" << row.second << "\n";
            }
        }
        if (r.merge)
        {
            totalReadsLen += r.totalReadsLen; correctedLen += r.correctedLen; totalSeedNum += r.totalSeedNum;
            totalWalkNum += r.totalWalkNum; highErrorNum += r.highErrorNum; exceedDepthNum += r.exceedDepthNum;
            exceedLeaveNum += r.exceedLeaveNum; FMNum += r.FMNum; DPNum += r.DPNum; seedDis += r.seedDis;
            for (size_t k = 0; k < r.correctedStrs.size(); k++)
                correct << ">" << reads[i].first << (P.Split ? ("_" + std::to_string(k)) : "") << "\n" << r.correctedStrs[k] << "\n";
        }
        else discard << ">" << reads[i].first << "\n" << reads[i].second << "\n";
        if (dump.is_open())
        {
            dump << "{\"id\":\"" << r.readid << "\",\"seeds\":[";
            for (size_t k = 0; k < r.seeds.size(); k++)
            {
                const SeedFeature& s = r.seeds[k];
                dump << (k ? "," : "") << "[" << s.seedStartPos << "," << s.seedLen << "," << s.maxFixedMerFreq << "," << (s.isRepeat ? 1 : 0)
                     << "," << s.startBestKmerSize << "," << s.endBestKmerSize << "]";
            }
            dump << "],\"pairs\":[";
            for (size_t k = 0; k < r.pairs.size(); k++)
            {
                const PairRecord& p = r.pairs[k];
                dump << (k ? "," : "") << "{\"s\":" << p.srcStart << ",\"t\":" << p.trgStart << ",\"k\":" << p.extendKmerSize << ",\"dis\":" << p.dis
                     << ",\"rtou\":" << (p.fromRtoU ? 1 : 0) << ",\"status\":" << p.status << ",\"src\":\"" << p.src << "\",\"path\":\"" << p.path
                     << "\",\"trg\":\"" << p.trg << "\",\"out\":\"" << p.out << "\"}";
            }
            dump << "],\"pieces\":[";
            for (size_t k = 0; k < r.correctedStrs.size(); k++) dump << (k ? "," : "") << "\"" << r.correctedStrs[k] << "\"";
            dump << "]}\n";
        }
    }
    { std::ofstream tt((P.directory + "threshold-table").c_str()); P.thr.write(tt); }
    if (totalWalkNum > 0 && totalReadsLen > 0)
    {
        int64_t OutcastNum = totalWalkNum - FMNum - DPNum;
        std::cout << "\n"
                  << "TotalReadsLen: " << totalReadsLen << "\n"
                  << "CorrectedLen: " << correctedLen << ", ratio: " << (float)(correctedLen) / totalReadsLen << "\n"
                  << "TotalSeedNum: " << totalSeedNum << "\n"
                  << "TotalWalkNum: " << totalWalkNum << "\n"
                  << "FMNum: " << FMNum << ", ratio: " << (float)(FMNum * 100) / totalWalkNum << "%\n"
                  << "DPNum: " << DPNum << ", ratio: " << (float)(DPNum * 100) / totalWalkNum << "%\n"
                  << "OutcastNum: " << OutcastNum << ", ratio: " << (float)(OutcastNum * 100) / totalWalkNum << "%\n"
                  << "HighErrorNum: " << highErrorNum << ", ratio: " << (float)(highErrorNum * 100) / (DPNum + OutcastNum) << "%\n"
                  << "ExceedDepthNum: " << exceedDepthNum << ", ratio: " << (float)(exceedDepthNum * 100) / (DPNum + OutcastNum) << "%\n"
                  << "ExceedLeaveNum: " << exceedLeaveNum << ", ratio: " << (float)(exceedLeaveNum * 100) / (DPNum + OutcastNum) << "%\n"
                  << "DisBetweenSeeds: " << seedDis / totalWalkNum << "\n";
    }
    uint64_t occSeed = 0, occExtend = 0, occSetup = 0, occWalk = 0, walkAttempts = 0, dpCells = 0, dpRows = 0, dpAttempts = 0, occDP = 0;
    for (const auto& r : results)
    {
        occSeed += r.occSeed; occExtend += r.occExtend; occSetup += r.occExtendSetup; occWalk += r.occExtendWalk; walkAttempts += r.pairs.size();
        dpCells += r.dpCells; dpRows += r.dpRows; dpAttempts += r.dpAttempts; occDP += r.occDP;
    }
    fprintf(stderr, "[pbsc_oracle] dp fallbacks %llu, rows kept %llu, band cells %llu, LF steps %llu\n", (unsigned long long)dpAttempts,
            (unsigned long long)dpRows, (unsigned long long)dpCells, (unsigned long long)occDP);
    if (RefineStats::on())
        fprintf(stderr, "[pbsc_oracle] refineSAInterval: %llu re-searches, mean %.2f bases, %llu (%.1f %%) of k-mers that occur at least twice\n",
                (unsigned long long)RefineStats::calls(), RefineStats::calls() ? (double)RefineStats::bases() / (double)RefineStats::calls() : 0.0,
                (unsigned long long)RefineStats::twice(), RefineStats::calls() ? 100.0 * (double)RefineStats::twice() / (double)RefineStats::calls() : 0.0);
    fprintf(stderr, "[pbsc_oracle] %zu reads, %.3f s, threads %d, rank queries %llu (seed %llu, extend %llu: setup %llu, walk loop %llu), walks %llu\n", reads.size(), secs,
            threads, (unsigned long long)occTotal, (unsigned long long)occSeed, (unsigned long long)occExtend, (unsigned long long)occSetup, (unsigned long long)occWalk,
            (unsigned long long)walkAttempts);
    return 0;
}
