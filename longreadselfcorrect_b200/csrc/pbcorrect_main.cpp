// pbcorrect — drop-in for `stride pbcorrect` (StriDe/PacBioSelfCorrection.cpp:145-434) on B200 GPUs.
//
// Host side only: option parsing, FASTA/FASTQ input (Util/SeqReader.cpp:26-135), batching over the GPUs of the box,
// and the post-processor that writes DIR/correct.fa, DIR/discard.fa, DIR/threshold-table and the stdout summary
// (PacBio/PacBioSelfCorrectionProcess.cpp:250-370).  All computation goes through the C ABI of libpbsc.so
// (include/pbsc.h); there is no CPU fallback.  Records are written in input order, i.e. the reference's `-t 1` order
// (with -t T > 1 the reference permutes records inside blocks of 500*T reads; contents are identical).
#include <getopt.h>
#include <zlib.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include "pbsc.h"

#define SUBPROGRAM "PacBioSelfCorrection"
#define PACKAGE_NAME "StriDe"
#define PACKAGE_VERSION "0.0.1"
#define PACKAGE_BUGREPORT "ythuang@cs.ccu.edu.tw"

static const char* CORRECT_VERSION_MESSAGE =
    SUBPROGRAM " Version " PACKAGE_VERSION " (B200 hot path)\n"
    "Written by Yao-Ting Huang & Ping-Yeh Chen.\n"
    "\n"
    "Copyright 2015 National Chung Cheng University\n";

static const char* CORRECT_USAGE_MESSAGE =
    "Usage: " PACKAGE_NAME " " SUBPROGRAM " [OPTION] ... READSFILE\n"
    "Correct PacBio reads via FM-index walk\n"
    "\n"
    "      -t, --thread=NUM                 Host threads (I/O only here; the computation runs on the GPUs) (default: 1)\n"
    "      -p, --prefix=PREFIX              Use PREFIX for the names of the index files\n"
    "      -o, --output=DIR                 Output results in the directory\n"
    "      -b, --barcode=FILE               Barcode of raw reads\n"
    "      --gpus=N                         Number of GPUs to use (default: all visible)\n"
    "      --batch-mbp=N                    Read bases per GPU batch, in Mbp (default: 64)\n"
    "      --prefix-k=N                     Length of the short-prefix interval table, 0 = off (default: 13)\n"
    "\nPacBio correction parameters:\n"
    "      -c, --PBcoverage=N               Coverage of PacBio reads (default: 90)\n"
    "      -e, --error-rate=N               The error rate of PacBio reads.(default:0.15)\n"
    "      -k, --kmer-size=N                The start kmer length (default: 19 (PacBioS).)\n"
    "      -n, --next-target                The number of next FMWalk target seed(default: 1)\n"
    "      -l, --max-leaves=N               Number of maximum leaves in the search tree. (default: 32)\n"
    "      -i, --idmer-length=N             The length of the kmer to identify similar reads.(default: 9)\n"
    "      -s, --min-kmer-size=N            The minimum length of the kmer to use. (default: 13.)\n"
    "      -g, --genome=(5/10/100)[m]       Genome size of the species (default: 10m)\n"
    "      -m, --mode=(0/1/2)               Mode in seed-searching (default: 1)\n"
    "      -v, --verbose                    Display verbose output\n"
    "      --help                           Display this help and exit\n"
    "      --version                        Display version and exit\n"
    "      --debugseed                      Output seeds file for each reads (default: false)\n"
    "      --debugextend                    Show extension information (default: false)\n"
    "      --onlyseed                       Only search seeds file for each reads (default: false)\n"
    "      --nodp                           Don't use dp (default: false)\n"
    "      --split                          Split the uncorrected reads (default: false)\n"
    "\nReport bugs to " PACKAGE_BUGREPORT "\n\n";

namespace opt
{
static int thread = 1;
static std::string prefix, directory, barcode, readsFile;
static pbsc_params params;
static bool DebugSeed = false, OnlySeed = false;
static int gpus = 0, prefix_k = 13;
static double batch_mbp = 64;
static int verbose = 0;
}

enum { OPT_HELP = 1, OPT_VERSION, OPT_SPLIT, OPT_DEBUGEXTEND, OPT_DEBUGSEED, OPT_ONLYSEED, OPT_NODP, OPT_GPUS, OPT_BATCH, OPT_PREFIXK };
static const char* shortopts = "t:p:o:b:c:e:k:u:r:n:l:i:s:g:m:v";
static const struct option longopts[] = {
    {"thread", required_argument, nullptr, 't'}, {"prefix", required_argument, nullptr, 'p'},
    {"output", required_argument, nullptr, 'o'}, {"barcode", required_argument, nullptr, 'b'},
    {"PBcoverage", required_argument, nullptr, 'c'}, {"error-rate", required_argument, nullptr, 'e'},
    {"kmer-size", required_argument, nullptr, 'k'}, {"unique-offset", required_argument, nullptr, 'u'},
    {"repeat-offset", required_argument, nullptr, 'r'}, {"next-target", required_argument, nullptr, 'n'},
    {"max-leaves", required_argument, nullptr, 'l'}, {"idmer-length", required_argument, nullptr, 'i'},
    {"min-kmer-size", required_argument, nullptr, 's'}, {"genome", required_argument, nullptr, 'g'},
    {"mode", required_argument, nullptr, 'm'}, {"verbose", no_argument, nullptr, 'v'},
    {"help", no_argument, nullptr, OPT_HELP}, {"version", no_argument, nullptr, OPT_VERSION},
    {"split", no_argument, nullptr, OPT_SPLIT}, {"debugextend", no_argument, nullptr, OPT_DEBUGEXTEND},
    {"debugseed", no_argument, nullptr, OPT_DEBUGSEED}, {"onlyseed", no_argument, nullptr, OPT_ONLYSEED},
    {"nodp", no_argument, nullptr, OPT_NODP}, {"gpus", required_argument, nullptr, OPT_GPUS},
    {"batch-mbp", required_argument, nullptr, OPT_BATCH}, {"prefix-k", required_argument, nullptr, OPT_PREFIXK},
    {nullptr, 0, nullptr, 0}};

// StriDe/PacBioSelfCorrection.cpp:262-434, same messages and exit codes
static void parseOptions(int argc, char** argv)
{
    pbsc_params_default(&opt::params);
    optind = 1;
    bool die = false;
    for (int c; (c = getopt_long(argc, argv, shortopts, longopts, nullptr)) != -1;)
    {
        std::istringstream arg(optarg != nullptr ? optarg : "");
        switch (c)
        {
            case 't': arg >> opt::thread; break;
            case 'p': arg >> opt::prefix; break;
            case 'o': arg >> opt::directory; break;
            case 'b': arg >> opt::barcode; break;
            case 'c': arg >> opt::params.pb_coverage; break;
            case 'e': arg >> opt::params.error_rate; break;
            case 'k': arg >> opt::params.start_kmer; opt::params.adjust = 1; break;
            case 'u': arg >> opt::params.offset[1]; opt::params.adjust = 1; break;
            case 'r': arg >> opt::params.offset[2]; opt::params.adjust = 1; break;
            case 'n': arg >> opt::params.next_target; break;
            case 'l': arg >> opt::params.max_leaves; break;
            case 'i': arg >> opt::params.idmer_len; break;
            case 's': arg >> opt::params.min_kmer; break;
            case 'g': arg >> opt::params.genome; break;
            case 'm': arg >> opt::params.mode; opt::params.manual = 1; break;
            case 'v': opt::verbose++; break;
            case OPT_HELP: std::cerr << CORRECT_USAGE_MESSAGE; exit(EXIT_SUCCESS);
            case OPT_VERSION: std::cerr << CORRECT_VERSION_MESSAGE; exit(EXIT_SUCCESS);
            case OPT_SPLIT: opt::params.split = 1; break;
            case OPT_DEBUGEXTEND: break;
            case OPT_DEBUGSEED: opt::DebugSeed = true; break;
            case OPT_NODP: opt::params.no_dp = 1; break;
            case OPT_ONLYSEED: opt::DebugSeed = true; opt::OnlySeed = true; break;
            case OPT_GPUS: arg >> opt::gpus; break;
            case OPT_BATCH: arg >> opt::batch_mbp; break;
            case OPT_PREFIXK: arg >> opt::prefix_k; break;
            default: die = true; break;
        }
    }
    if (argc - optind < 1) { std::cerr << SUBPROGRAM ": missing arguments\n"; die = true; }
    else if (argc - optind > 1) { std::cerr << SUBPROGRAM ": too many arguments\n"; die = true; }
    if (opt::thread <= 0) { std::cerr << SUBPROGRAM ": invalid number of threads: " << opt::thread << "\n"; die = true; }
    if (opt::prefix.empty()) { std::cerr << SUBPROGRAM << ": no prefix\n"; die = true; }
    if (opt::directory.empty()) { std::cerr << SUBPROGRAM << ": no directory\n"; die = true; }
    else
    {
        opt::directory += "/";
        if (system(("mkdir -p " + opt::directory).c_str()) != 0)
        {
            std::cerr << SUBPROGRAM << ": something wrong making directory: " << opt::directory << "\n";
            die = true;
        }
    }
    const pbsc_params& P = opt::params;
    if (P.pb_coverage <= 0) { std::cerr << SUBPROGRAM ": invalid number of coverage: " << P.pb_coverage << ", must be greater than zero\n"; die = true; }
    if (P.error_rate < 0 || P.error_rate > 1) { std::cerr << SUBPROGRAM ":invalid error rate: " << P.error_rate << ", must be 0 ~ 1\n"; die = true; }
    if (P.start_kmer <= 0) { std::cerr << SUBPROGRAM ": invalid start kmer length: " << P.start_kmer << ", must be greater than zero\n"; die = true; }
    if (P.next_target <= 0) { std::cerr << SUBPROGRAM ": invalid number of next target: " << P.next_target << ", must be greater than zero\n"; die = true; }
    if (P.max_leaves <= 0) { std::cerr << SUBPROGRAM ":invalid number of max leaves:" << P.max_leaves << ", must be greater than zero\n"; die = true; }
    if (P.idmer_len <= 0) { std::cerr << SUBPROGRAM ":invalid kmer length to identify similar reads" << P.idmer_len << ", must be greater than zero\n"; die = true; }
    if (P.min_kmer <= 0) { std::cerr << SUBPROGRAM ":invalid min kmer length:" << P.min_kmer << ", must be greater than zero\n"; die = true; }
    if (P.genome != 5 && P.genome != 10 && P.genome != 100) { std::cerr << SUBPROGRAM ": invalid genome size: " << P.genome << ", must be (5/10/100)[m]\n"; die = true; }
    if (P.mode < 0 || P.mode > 2) { std::cerr << SUBPROGRAM ": invalid mode: " << P.mode << ", must be (0/1/2)\n"; die = true; }
    if (opt::OnlySeed && opt::barcode.empty()) { std::cerr << SUBPROGRAM ": no barcode\n"; die = true; }
    if (die) { std::cerr << "\n" << CORRECT_USAGE_MESSAGE; exit(EXIT_FAILURE); }
    if (opt::OnlySeed || opt::DebugSeed)
    {
        std::cerr << SUBPROGRAM ": --debugseed/--onlyseed diagnostics are not part of this build (hot path only)\n";
        exit(EXIT_FAILURE);
    }
    opt::readsFile = argv[optind++];
}

// ---- sequence input: Util/SeqReader.cpp:26-135 over zlib (plain or .gz, Util/Util.cpp:276-291) ----
class LineReader
{
  public:
    explicit LineReader(const std::string& path) { f_ = gzopen(path.c_str(), "rb"); if (f_) gzbuffer(f_, 1 << 20); }
    ~LineReader() { if (f_) gzclose(f_); }
    bool ok() const { return f_ != nullptr; }
    bool good() const { return good_; }
    bool eof() const { return eof_; }
    int peek() { if (pos_ >= len_ && !fill()) return EOF; return (unsigned char)buf_[pos_]; }
    // std::getline: the line is returned even when EOF ends it, but the stream is no longer good()
    void getline(std::string& out)
    {
        out.clear();
        if (!good_) return;
        bool any = false;
        for (;;)
        {
            if (pos_ >= len_ && !fill()) { eof_ = true; good_ = false; (void)any; return; }
            const char* nl = (const char*)memchr(buf_ + pos_, '\n', len_ - pos_);
            if (nl) { out.append(buf_ + pos_, nl - (buf_ + pos_)); pos_ = (nl - buf_) + 1; return; }
            out.append(buf_ + pos_, len_ - pos_);
            any = true;
            pos_ = len_;
        }
    }
  private:
    bool fill() { if (!f_) return false; int n = gzread(f_, buf_, sizeof buf_); if (n <= 0) return false; len_ = n; pos_ = 0; return true; }
    gzFile f_ = nullptr;
    char buf_[1 << 16];
    int pos_ = 0, len_ = 0;
    bool good_ = true, eof_ = false;
};

static bool readRecord(LineReader& in, std::string& id, std::string& seq)
{
    std::string header;
    int rt = 0;
    while (in.good())
    {
        in.getline(header);
        if (header.empty()) continue;
        if (header[0] == '>') { rt = 1; break; }
        if (header[0] == '@') { rt = 2; break; }
    }
    if (rt == 0) return false;
    bool valid = false;
    seq.clear();
    std::string temp, qual;
    if (rt == 1)
    {
        while (in.good() && in.peek() != '>' && in.peek() != '@')
        {
            in.getline(temp);
            if (in.good() && temp.size() > 0) seq.append(temp);
        }
        valid = seq.size() > 0;
    }
    else
    {
        in.getline(seq); in.getline(temp); in.getline(qual);
        if (seq.empty() || qual.empty()) std::cerr << "Warning, read " << header << " has no sequence or quality values\n";
        valid = !in.eof();
    }
    if (!valid) return false;
    size_t endPos = std::min(header.find_first_of(' '), header.find_first_of('\t'));
    id = endPos != std::string::npos ? header.substr(1, endPos - 1) : header.substr(1);
    for (auto& c : seq) c = (char)toupper((unsigned char)c);
    if (seq.find_first_not_of("ACGT") != std::string::npos)
    {
        std::cerr << "Error: read " << id << " contains non-ACGT characters.\n";
        std::cerr << "Please run sga preprocess on the data first.\n";
        exit(EXIT_FAILURE);
    }
    return true;
}

struct Batch
{
    std::vector<std::string> ids;
    std::string bases;
    std::vector<uint64_t> offsets{0};
    // results
    std::vector<char> pieces;
    std::vector<uint64_t> piece_off, first;
    std::vector<pbsc_read_stats> stats;
    pbsc_timing timing{};
    bool done = false;
    int rc = 0;
    std::string err;
};

int main(int argc, char** argv)
{
    // accept both `pbcorrect [opts] READS` and `pbcorrect pbcorrect [opts] READS` (as `stride pbcorrect`)
    if (argc > 1 && std::string(argv[1]) == "pbcorrect") { argv++; argc--; }
    parseOptions(argc, argv);
    int rc = pbsc_params_derive(&opt::params);
    if (rc != PBSC_OK) { std::cerr << SUBPROGRAM ": " << pbsc_last_error() << "\n"; return EXIT_FAILURE; }
    int ndev = pbsc_device_count();
    if (ndev <= 0) { std::cerr << SUBPROGRAM ": no CUDA device available (this build has no CPU path)\n"; return EXIT_FAILURE; }
    const int ngpu = opt::gpus > 0 ? std::min(opt::gpus, ndev) : ndev;

    auto t_load = std::chrono::steady_clock::now();
    std::cerr << "Loading BWT: " << opt::prefix << ".bwt\n" << "Loading RBWT: " << opt::prefix << ".rbwt\n"
              << "Loading Sampled Suffix Array: " << opt::prefix << ".sai\n";
    std::vector<pbsc_index*> index(ngpu, nullptr);
    {
        std::vector<std::thread> th;
        std::vector<int> rcs(ngpu, 0);
        std::vector<std::string> errs(ngpu);
        for (int g = 0; g < ngpu; g++)
            th.emplace_back([&, g]() {
                rcs[g] = pbsc_index_load(opt::prefix.c_str(), g, 1, &index[g]);
                if (rcs[g] == PBSC_OK && opt::prefix_k > 0) rcs[g] = pbsc_index_build_prefix_table(index[g], opt::prefix_k);
                if (rcs[g] != PBSC_OK) errs[g] = pbsc_last_error();
            });
        for (auto& t : th) t.join();
        for (int g = 0; g < ngpu; g++)
            if (rcs[g] != PBSC_OK) { std::cerr << errs[g] << "\n"; return EXIT_FAILURE; }
    }
    double load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load).count();
    std::cerr << "[timer - index to " << ngpu << " GPU(s)] wall clock: " << load_s << "s (" << pbsc_index_device_bytes(index[0]) / 1e9 << " GB per GPU)\n";

    const pbsc_params& P = opt::params;
    std::cerr << "\nCorrecting PacBio reads for " << opt::readsFile << " using--\n"
              << "number of threads:\t" << opt::thread << "\n"
              << "number of GPUs:\t" << ngpu << "\n"
              << "PB reads coverage:\t" << P.pb_coverage << "\n"
              << "num of next Targets:\t" << P.next_target << "\n"
              << "large kmer size:\t" << P.start_kmer << "\n"
              << "small kmer size:\t" << P.min_kmer << "\n"
              << "max leaves:\t" << P.max_leaves << "\n"
              << "max depth:\t1.2~0.8* (length between two seeds +- 20)" << "\n";

    LineReader in(opt::readsFile);
    if (!in.ok()) { std::cerr << "Error: could not open " << opt::readsFile << " for read\n"; return EXIT_FAILURE; }
    std::ofstream correct((opt::directory + "correct.fa").c_str()), discard((opt::directory + "discard.fa").c_str());
    if (!correct || !discard) { std::cerr << "Error: could not open output files in " << opt::directory << "\n"; return EXIT_FAILURE; }

    auto t0 = std::chrono::steady_clock::now();
    // ---- pipeline: reader (this thread) -> one worker per GPU -> in-order writer (this thread) ----
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::unique_ptr<Batch>> batches;
    size_t next_job = 0;
    bool reading_done = false;
    auto worker = [&](int g) {
        for (;;)
        {
            Batch* b = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return next_job < batches.size() || reading_done; });
                if (next_job >= batches.size()) return;
                b = batches[next_job++].get();
            }
            const uint64_t n = b->ids.size();
            uint64_t cap = (uint64_t)(b->bases.size() * 1.3) + (1 << 16), need = 0;
            b->first.assign(n + 1, 0);
            b->stats.assign(n ? n : 1, pbsc_read_stats{});
            for (;;)
            {
                b->pieces.resize(cap);
                b->piece_off.assign((P.split ? b->bases.size() / 10 + 4 * n : n) + 16, 0);
                b->rc = pbsc_correct_batch(index[g], &P, b->bases.data(), b->offsets.data(), n, b->pieces.data(), cap, b->piece_off.data(),
                                           b->piece_off.size(), b->first.data(), b->stats.data(), &need);
                if (b->rc == PBSC_ERR_LIMIT && need > cap) { cap = need + 64; continue; }
                break;
            }
            if (b->rc != PBSC_OK) b->err = pbsc_last_error();
            pbsc_last_timing(&b->timing);
            { std::lock_guard<std::mutex> lk(mu); b->done = true; }
            cv.notify_all();
        }
    };
    std::vector<std::thread> workers;
    for (int g = 0; g < ngpu; g++) workers.emplace_back(worker, g);

    int64_t totalReadsLen = 0, correctedLen = 0, totalSeedNum = 0, totalWalkNum = 0, highErrorNum = 0, exceedDepthNum = 0, exceedLeaveNum = 0, FMNum = 0,
            DPNum = 0, seedDis = 0;
    double seed_s = 0, fm_s = 0, dp_s = 0;
    size_t written = 0, nreads = 0;
    uint64_t inBases = 0;
    auto flush = [&](bool all) -> bool {
        for (;;)
        {
            Batch* b = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (written >= batches.size()) return true;
                b = batches[written].get();
                if (!b->done) { if (!all) return true; cv.wait(lk, [&] { return b->done; }); }
            }
            if (b->rc != PBSC_OK) { std::cerr << SUBPROGRAM ": " << b->err << "\n"; return false; }
            for (size_t r = 0; r < b->ids.size(); r++)
            {
                const pbsc_read_stats& st = b->stats[r];
                if (st.merge)
                {
                    totalReadsLen += st.total_reads_len; correctedLen += st.corrected_len; totalSeedNum += st.total_seed_num;
                    totalWalkNum += st.total_walk_num; highErrorNum += st.high_error_num; exceedDepthNum += st.exceed_depth_num;
                    exceedLeaveNum += st.exceed_leave_num; FMNum += st.fm_num; DPNum += st.dp_num; seedDis += st.seed_dis;
                    for (uint64_t j = b->first[r]; j < b->first[r + 1]; j++)
                    {
                        correct << ">" << b->ids[r];
                        if (P.split) correct << "_" << (j - b->first[r]);
                        correct << "\n";
                        correct.write(b->pieces.data() + b->piece_off[j], (std::streamsize)(b->piece_off[j + 1] - b->piece_off[j]));
                        correct << "\n";
                    }
                }
                else
                {
                    discard << ">" << b->ids[r] << "\n";
                    discard.write(b->bases.data() + b->offsets[r], (std::streamsize)(b->offsets[r + 1] - b->offsets[r]));
                    discard << "\n";
                }
            }
            seed_s += b->timing.seed_ms / 1e3; fm_s += (b->timing.extend_ms - b->timing.dp_ms) / 1e3; dp_s += b->timing.dp_ms / 1e3;
            { std::lock_guard<std::mutex> lk(mu); batches[written].reset(new Batch()); batches[written]->done = true; }
            written++;
        }
    };

    const uint64_t batch_bases = (uint64_t)(opt::batch_mbp * 1e6);
    std::unique_ptr<Batch> cur(new Batch());
    std::string id, seq;
    bool ok = true;
    while (ok && readRecord(in, id, seq))
    {
        cur->ids.push_back(id);
        cur->bases += seq;
        cur->offsets.push_back(cur->bases.size());
        inBases += seq.size();
        nreads++;
        if (cur->bases.size() >= batch_bases)
        {
            { std::lock_guard<std::mutex> lk(mu); batches.push_back(std::move(cur)); }
            cv.notify_all();
            cur.reset(new Batch());
            ok = flush(false);
        }
    }
    if (!cur->ids.empty()) { std::lock_guard<std::mutex> lk(mu); batches.push_back(std::move(cur)); }
    { std::lock_guard<std::mutex> lk(mu); reading_done = true; }
    cv.notify_all();
    if (ok) ok = flush(true);
    for (auto& t : workers) t.join();
    for (auto* ix : index) pbsc_index_destroy(ix);
    if (!ok) return EXIT_FAILURE;

    double proc = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "Processed %zu sequences in %lfs (%lf sequences/s)\n", nreads, proc, (double)nreads / proc);
    fprintf(stderr, "[timer - " PACKAGE_NAME "::" SUBPROGRAM "] wall clock: %.2fs, %.3f Mbp/s over %d GPU(s)\n", proc, inBases / 1e6 / proc, ngpu);

    // PacBioSelfCorrectionPostProcess::~PacBioSelfCorrectionPostProcess — PacBioSelfCorrectionProcess.cpp:281-311
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
                  << "DisBetweenSeeds: " << seedDis / totalWalkNum << "\n"
                  << "Time of searching Seeds: " << seed_s << "\n"
                  << "Time of searching FM: " << fm_s << "\n"
                  << "Time of searching DP: " << dp_s << "\n";
    }
    // KmerThreshold::~KmerThreshold — PacBio/KmerThreshold.cpp:31-41
    {
        std::vector<char> buf(8192);
        int n = pbsc_threshold_table_text(&P, buf.data(), buf.size());
        if (n > 0) { std::ofstream tt((opt::directory + "threshold-table").c_str()); tt.write(buf.data(), n); }
    }
    return 0;
}
