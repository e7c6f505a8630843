// pbcorrect — drop-in for `stride pbcorrect` (StriDe/PacBioSelfCorrection.cpp:145-434) on B200 GPUs.
//
// Host side only: option parsing, FASTA/FASTQ input (Util/SeqReader.cpp:26-135), batching over the GPUs of the box,
// and the post-processor that writes DIR/correct.fa, DIR/discard.fa, DIR/threshold-table and the stdout summary
// (PacBio/PacBioSelfCorrectionProcess.cpp:250-370).  All computation goes through the C ABI of libpbsc.so
// (include/pbsc.h); there is no CPU fallback.  Records are written in input order, i.e. the reference's `-t 1` order
// (with -t T > 1 the reference permutes records inside blocks of 500*T reads; contents are identical).
#include <getopt.h>
#include <sys/stat.h>
#include <zlib.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
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
#include "pbsc_bcode.h"
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
    "      --batch-mbp=N                    Read bases per GPU batch, in Mbp (default: 256)\n"
    "      --prefix-k=N                     Length of the short-prefix interval table, 0 = off (default: 13)\n"
    "      --lanes=N                        Batches in flight per GPU, 1..4 (default: 1)\n"
    "      --write-fmg                      Leave PREFIX.fmg (the flat index as it sits in GPU memory) for later runs;\n"
    "                                       a PREFIX.fmg that matches PREFIX.bwt/.rbwt is always used when present\n"
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
static int gpus = 0, prefix_k = 13, lanes = 1;
static bool write_fmg = false;
static double batch_mbp = 256;
static int verbose = 0;
}

enum { OPT_HELP = 1, OPT_VERSION, OPT_SPLIT, OPT_DEBUGEXTEND, OPT_DEBUGSEED, OPT_ONLYSEED, OPT_NODP, OPT_GPUS, OPT_BATCH, OPT_PREFIXK, OPT_LANES, OPT_WRITEFMG };
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
    {"lanes", required_argument, nullptr, OPT_LANES}, {"write-fmg", no_argument, nullptr, OPT_WRITEFMG},
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
            case OPT_LANES: arg >> opt::lanes; break;
            case OPT_WRITEFMG: opt::write_fmg = true; break;
            default: die = true; break;
        }
    }
    if (argc - optind < 1) { std::cerr << SUBPROGRAM ": missing arguments\n"; die = true; }
    else if (argc - optind > 1) { std::cerr << SUBPROGRAM ": too many arguments\n"; die = true; }
    if (opt::thread <= 0) { std::cerr << SUBPROGRAM ": invalid number of threads: " << opt::thread << "\n"; die = true; }
    if (opt::lanes < 1 || opt::lanes > 4) { std::cerr << SUBPROGRAM ": invalid number of lanes: " << opt::lanes << ", must be 1..4\n"; die = true; }
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
    if (opt::DebugSeed)
    {
        // StriDe/PacBioSelfCorrection.cpp:351-361
        if (system(("mkdir -p " + opt::directory + "seed/error/").c_str()) != 0 || system(("mkdir -p " + opt::directory + "extend/").c_str()) != 0)
        { std::cerr << SUBPROGRAM << ": something wrong making directory: " << opt::directory << "\n"; exit(EXIT_FAILURE); }
        opt::params.debug_seed = 1;
    }
    opt::readsFile = argv[optind++];
}

// ---- sequence input: Util/SeqReader.cpp:26-135 over zlib (plain or .gz, Util/Util.cpp:276-291) ----
class LineReader
{
  public:
    explicit LineReader(const std::string& path) : buf_(1 << 22) { f_ = gzopen(path.c_str(), "rb"); if (f_) gzbuffer(f_, 1 << 20); }
    ~LineReader() { if (f_) gzclose(f_); }
    bool ok() const { return f_ != nullptr; }
    bool good() const { return good_; }
    bool eof() const { return eof_; }
    int peek() { if (pos_ >= len_ && !fill()) return EOF; return (unsigned char)buf_[pos_]; }
    // std::getline semantics: the line is delivered even when EOF ends it, but the stream is no longer good().  The line's
    // bytes go to sink(ptr, n), possibly in several pieces (no intermediate string: sequence lines are copied once, from the
    // read buffer into the batch's page-locked buffer).
    template <class Sink>
    void getline(Sink sink)
    {
        if (!good_) return;
        for (;;)
        {
            if (pos_ >= len_ && !fill()) { eof_ = true; good_ = false; return; }
            const char* nl = (const char*)memchr(buf_.data() + pos_, '\n', len_ - pos_);
            if (nl) { sink(buf_.data() + pos_, (size_t)(nl - (buf_.data() + pos_))); pos_ = (size_t)(nl - buf_.data()) + 1; return; }
            sink(buf_.data() + pos_, len_ - pos_);
            pos_ = len_;
        }
    }
    void getline(std::string& out) { out.clear(); getline([&](const char* p, size_t n) { out.append(p, n); }); }
  private:
    bool fill() { if (!f_) return false; int n = gzread(f_, buf_.data(), (unsigned)buf_.size()); if (n <= 0) return false; len_ = (size_t)n; pos_ = 0; return true; }
    gzFile f_ = nullptr;
    std::vector<char> buf_;
    size_t pos_ = 0, len_ = 0;
    bool good_ = true, eof_ = false;
};

// ---- page-locked host buffers (pbsc_host_alloc), recycled: allocating them costs milliseconds per 100 MB ----
struct PinBuf { char* p = nullptr; size_t cap = 0; };
class PinPool
{
  public:
    ~PinPool() { for (auto& b : free_) pbsc_host_free(b.p); }
    PinBuf get(size_t need)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t i = 0; i < free_.size(); i++)
                if (free_[i].cap >= need) { PinBuf b = free_[i]; free_[i] = free_.back(); free_.pop_back(); return b; }
            if (!free_.empty()) { pbsc_host_free(free_.back().p); free_.pop_back(); }   // too small for this one: do not hoard it
        }
        PinBuf b;
        b.cap = need + need / 4 + 4096;
        void* p = nullptr;
        if (pbsc_host_alloc(&p, b.cap) != PBSC_OK) { std::cerr << SUBPROGRAM ": " << pbsc_last_error() << "\n"; exit(EXIT_FAILURE); }
        b.p = (char*)p;
        return b;
    }
    void put(PinBuf b) { if (!b.p) return; std::lock_guard<std::mutex> lk(mu_); free_.push_back(b); }
  private:
    std::mutex mu_;
    std::vector<PinBuf> free_;
};
static PinPool g_pins;

struct Batch
{
    size_t seq = 0;
    std::vector<std::string> ids;
    PinBuf bases;                     // raw sequence bytes as read (upper-cased and validated by normalize())
    size_t n_bases = 0;
    std::vector<uint64_t> offsets{0};
    // results
    PinBuf pieces;
    std::vector<uint64_t> piece_off, first;
    std::vector<pbsc_read_stats> stats;
    pbsc_timing timing{};
    // --debugseed
    std::vector<pbsc_seed> dbg_seeds; std::vector<uint64_t> dbg_seed_off, dbg_log_off; std::vector<uint32_t> dbg_surv; std::vector<float> dbg_ratio;
    std::vector<pbsc_walk_log> dbg_log;
    bool normalized = false, taken = false, done = false;
    int rc = 0;
    std::string err;
    void append(const char* p, size_t n)
    {
        if (n_bases + n > bases.cap)
        {
            PinBuf nb = g_pins.get((n_bases + n) * 2);
            if (n_bases) memcpy(nb.p, bases.p, n_bases);
            g_pins.put(bases);
            bases = nb;
        }
        memcpy(bases.p + n_bases, p, n);
        n_bases += n;
    }
    void release() { g_pins.put(bases); g_pins.put(pieces); bases = PinBuf(); pieces = PinBuf(); }
};

// One record, exactly as SeqReader::get delivers it (Util/SeqReader.cpp:26-135): FASTA (multi-line) or FASTQ by the header's
// first character; a FASTA sequence line that EOF ends without a newline is lost (:64-69); ids stop at the first blank or tab.
// The sequence is appended to the batch as read; upper-casing and the ACGT check (:112-125) happen in normalize().
static bool readRecord(LineReader& in, Batch& b)
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
    const size_t start = b.n_bases;
    bool valid = false;
    if (rt == 1)
    {
        while (in.good() && in.peek() != '>' && in.peek() != '@')
        {
            const size_t before = b.n_bases;
            in.getline([&](const char* p, size_t n) { b.append(p, n); });
            if (!in.good()) b.n_bases = before;   // std::getline hit EOF: the reference drops this last, unterminated line
        }
        valid = b.n_bases > start;
    }
    else
    {
        std::string temp, qual;
        in.getline([&](const char* p, size_t n) { b.append(p, n); });
        in.getline(temp); in.getline(qual);
        if (b.n_bases == start || qual.empty()) std::cerr << "Warning, read " << header << " has no sequence or quality values\n";
        valid = !in.eof();
    }
    if (!valid) { b.n_bases = start; return false; }
    size_t endPos = std::min(header.find_first_of(' '), header.find_first_of('\t'));
    b.ids.push_back(endPos != std::string::npos ? header.substr(1, endPos - 1) : header.substr(1));
    b.offsets.push_back(b.n_bases);
    return true;
}

// upper-case in place and insist on ACGT (SeqReader.cpp:112-125): false + the offending read's id on anything else
static bool normalize(Batch& b, std::string& bad_id)
{
    static unsigned char lut[256];
    static std::once_flag once;
    std::call_once(once, [] {
        memset(lut, 0, sizeof lut);
        for (const char* c = "ACGT"; *c; c++) { lut[(unsigned char)*c] = (unsigned char)*c; lut[(unsigned char)tolower(*c)] = (unsigned char)*c; }
    });
    unsigned char* s = (unsigned char*)b.bases.p;
    unsigned char ok = 0xff;
    for (size_t i = 0; i < b.n_bases; i++) { const unsigned char u = lut[s[i]]; ok &= (unsigned char)(u ? 0xff : 0); s[i] = u ? u : s[i]; }
    if (ok) return true;
    for (size_t r = 0; r + 1 < b.offsets.size(); r++)
        for (uint64_t i = b.offsets[r]; i < b.offsets[r + 1]; i++)
            if (!lut[s[i]]) { bad_id = b.ids[r]; return false; }
    return false;
}

// --debugseed: the reference's per-read dump files (same names, same text):
//   seed/<id>.seed        the seeds that survive removeHitchhikingSeeds (operator<< of SeedVector, PacBio/SeedFeature.cpp:11-20)
//   seed/error/<id>.seed  the hitchhiked ones; only when the read had at least two seeds (LongReadProbe.cpp:189,221-226)
//   extend/<id>.log       position <tab> repeat ratio, one line per base (LongReadProbe.cpp:172-173; default ostream float format)
//   extend/<id>.ext       source start <tab> target start <tab> walk outcome + 4 for every failed FM walk (PacBioSelfCorrectionProcess.cpp:130-131)
//   extend/<id>.dp        source start <tab> target start where the DP fallback failed too (:139-140); both only for reads
//                         with at least two seeds (:63,72-76)
static void writeDebugFiles(const Batch& b)
{
    const char* s = b.bases.p;
    std::string text;
    char num[64];
    for (size_t r = 0; r < b.ids.size(); r++)
    {
        const pbsc_seed* sv = b.dbg_seeds.data() + b.dbg_seed_off[r];
        const uint32_t nsurv = b.dbg_surv[r], nall = (uint32_t)(b.dbg_seed_off[r + 1] - b.dbg_seed_off[r]);
        auto seedText = [&](uint32_t from, uint32_t to) {
            text.clear();
            for (uint32_t i = from; i < to; i++)
            {
                text.append(s + b.offsets[r] + sv[i].start, (size_t)sv[i].len);
                snprintf(num, sizeof num, "\t%d\t%d\t%s\n", sv[i].max_fixed_freq, sv[i].start, sv[i].is_repeat ? "Yes" : "No");
                text += num;
            }
        };
        auto put = [&](const std::string& path) { std::ofstream f(path.c_str()); f << text; };
        seedText(0, nsurv);
        put(opt::directory + "seed/" + b.ids[r] + ".seed");
        if (nall >= 2) { seedText(nsurv, nall); put(opt::directory + "seed/error/" + b.ids[r] + ".seed"); }
        {
            std::ofstream f((opt::directory + "extend/" + b.ids[r] + ".log").c_str());
            const uint64_t L = b.offsets[r + 1] - b.offsets[r];
            for (uint64_t p = 0; p < L; p++) f << p << '\t' << b.dbg_ratio[b.offsets[r] + p] << '\n';
        }
        if (nsurv >= 2 && !opt::OnlySeed)   // --onlyseed stops before the walks (PacBioSelfCorrectionProcess.cpp:58-62): no .ext / .dp
        {
            std::ofstream x((opt::directory + "extend/" + b.ids[r] + ".ext").c_str()), d((opt::directory + "extend/" + b.ids[r] + ".dp").c_str());
            for (uint64_t i = b.dbg_log_off[r]; i < b.dbg_log_off[r + 1]; i++)
            {
                const pbsc_walk_log& e = b.dbg_log[i];
                x << e.src_start << "\t" << e.trg_start << "\t" << e.code << "\n";
                if (e.dp_failed) d << e.src_start << "\t" << e.trg_start << "\n";
            }
        }
    }
}

// ---- `pbcorrect index [OPTION] ... READSFILE`: `stride index` (StriDe/index.cpp:34-52,86-214,262-352) with the suffix sorting on
// the GPU (pbsc_build_index_files, pbsc_build.cu).  Same options, same default prefix (the reads file's name without directory
// and extensions, in the current directory), same four files, byte for byte.  -t, -a, -d, -g and -c are accepted and ignored:
// there is one algorithm here and it yields what ropebwt2 yields.
static const char* INDEX_USAGE_MESSAGE =
"Usage: pbcorrect index [OPTION] ... READSFILE\n"
"Index the reads in READSFILE using a suffixarray/bwt\n"
"\n"
"  -v, --verbose                        display verbose output\n"
"      --help                           display this help and exit\n"
"  -a, --algorithm=STR                  accepted for compatibility (sais, ropebwt, ropebwt2): the GPU builder writes the same files\n"
"  -t, --threads=NUM                    accepted for compatibility\n"
"  -p, --prefix=PREFIX                  write index to file using PREFIX instead of prefix of READSFILE\n"
"      --no-reverse                     suppress construction of the reverse BWT\n"
"      --no-forward                     suppress construction of the forward BWT\n"
"      --gpu=NUM                        CUDA device to build on (default: 0)\n\n";

static int indexMain(int argc, char** argv)
{
    enum { IOPT_HELP = 1, IOPT_NO_REVERSE, IOPT_NO_FWD, IOPT_GPU };
    static const struct option iopts[] = {
        {"verbose", no_argument, nullptr, 'v'}, {"check", no_argument, nullptr, 'c'}, {"prefix", required_argument, nullptr, 'p'},
        {"threads", required_argument, nullptr, 't'}, {"disk", required_argument, nullptr, 'd'}, {"gap-array", required_argument, nullptr, 'g'},
        {"algorithm", required_argument, nullptr, 'a'}, {"no-reverse", no_argument, nullptr, IOPT_NO_REVERSE},
        {"no-forward", no_argument, nullptr, IOPT_NO_FWD}, {"help", no_argument, nullptr, IOPT_HELP}, {"gpu", required_argument, nullptr, IOPT_GPU},
        {nullptr, 0, nullptr, 0}};
    const auto t0 = std::chrono::steady_clock::now();
    const clock_t c0 = clock();
    std::string prefix, reads;
    int flags = 0, device = 0;
    bool die = false;
    optind = 1;
    for (int c; (c = getopt_long(argc, argv, "p:a:m:t:d:g:cv", iopts, nullptr)) != -1;)
    {
        std::istringstream arg(optarg != nullptr ? optarg : "");
        switch (c)
        {
            case 'p': arg >> prefix; break;
            case 'a': case 't': case 'd': case 'g': case 'm': case 'c': case 'v': break;
            case IOPT_NO_REVERSE: flags |= PBSC_BUILD_NO_REVERSE; break;
            case IOPT_NO_FWD: flags |= PBSC_BUILD_NO_FORWARD; break;
            case IOPT_GPU: arg >> device; break;
            case IOPT_HELP: std::cout << INDEX_USAGE_MESSAGE; exit(EXIT_SUCCESS);
            default: die = true; break;
        }
    }
    if (argc - optind < 1) { std::cerr << "index: missing arguments\n"; die = true; }
    else if (argc - optind > 1) { std::cerr << "index: too many arguments\n"; die = true; }
    if (die) { std::cout << "\n" << INDEX_USAGE_MESSAGE; exit(EXIT_FAILURE); }
    reads = argv[optind];
    if (prefix.empty())
    {
        // getFilename (Util/Util.cpp:218-225)
        std::string out = reads.substr(reads.find_last_of('/') == std::string::npos ? 0 : reads.find_last_of('/') + 1);
        auto strip = [](const std::string& f) { const size_t dot = f.find_last_of('.'); return dot == std::string::npos ? f : f.substr(0, dot); };
        if (out.size() >= 3 && out.compare(out.size() - 3, 3, ".gz") == 0) out = strip(out);
        prefix = strip(out);
    }
    else
    {
        const size_t slash = prefix.find_last_of('/');
        const std::string dir = (slash == std::string::npos ? std::string(".") : prefix.substr(0, slash)) + "/";
        if (system(("mkdir -p " + dir).c_str()) != 0) { std::cerr << "index: something wrong making directory: " << dir << "\n"; return EXIT_FAILURE; }
    }
    if (pbsc_device_count() <= 0) { std::cerr << "index: no CUDA device available (this build has no CPU path)\n"; return EXIT_FAILURE; }
    LineReader in(reads);
    if (!in.ok()) { std::cerr << "Error: could not open " << reads << " for read\n"; return EXIT_FAILURE; }
    Batch all;
    while (readRecord(in, all)) {}
    if (all.ids.empty()) { std::cerr << "index: input file is empty\n"; return EXIT_FAILURE; }
    std::string bad;
    if (!normalize(all, bad)) { std::cerr << "index: read " << bad << " holds letters other than ACGT\n"; return EXIT_FAILURE; }
    std::cout << "Building index for " << reads << " on GPU " << device << "\n";
    if (pbsc_build_index_files(all.bases.p, all.offsets.data(), all.ids.size(), prefix.c_str(), device, flags) != PBSC_OK)
    { std::cerr << "index: " << pbsc_last_error() << "\n"; return EXIT_FAILURE; }
    if (!(flags & PBSC_BUILD_NO_FORWARD)) std::cout << "\t done bwt construction, generating .sai file\n";
    if (!(flags & PBSC_BUILD_NO_REVERSE)) std::cout << "\t done rbwt construction, generating .rsai file\n";
    all.release();
    char line[160];
    snprintf(line, sizeof line, "[timer - Build FM index] wall clock: %.2fs CPU: %.2fs\n",
             std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), (double)(clock() - c0) / CLOCKS_PER_SEC);
    std::cerr << line;
    return 0;
}

// ---- `pbcorrect kmercheck [OPTION] ... READSFILE`: `stride kmercheck` (StriDe/kmercheck.cpp:25-226, PacBio/KmerCheckProcess.cpp:12-63).
// For every barcode block of every read and every k of the range, the frequency (both strands, KmerFeature.h:37-65) of each k-mer of
// the block; the k-mers the barcode calls correct and the ones it calls wrong go to two histograms per k, summarised into
// DIR/total.box and DIR/value.box (appended to, like the reference).  The frequencies are batched backward searches on the GPU
// (pbsc_findinterval_batch); the barcode arithmetic and the histograms are host code (pbsc_bcode.h).
static const char* KMERCHECK_USAGE_MESSAGE =
"Usage: pbcorrect kmercheck [OPTION] ... READSFILE\n"
"Get sequences kmer frequency\n"
"  -t, --threads=NUM         Use NUM threads for the computation (default: 1)\n"
"  -c, --coverage=NUM        Coverage of PacBio reads (default: 90)\n"
"  -p, --prefix=PREFIX       Use PREFIX for the names of the index files\n"
"  -o, --directory=PATH      Put results in the directory\n"
"  -b, --barcode=FILE        Use the barcode to check kmer \n"
"  -l, --lower=NUM           Kmer size lower bound (default: 15)\n"
"  -u, --upper=NUM           Kmer size upper bound (default: 35)\n"
"  -s, --step=NUM            Kmer size step (default: 1)\n"
"  -v, --verbose             Display verbose output\n"
"      --help                Display this help and exit\n"
"      --version             Display version\n";

static int kmercheckMain(int argc, char** argv)
{
    enum { KOPT_HELP = 1, KOPT_VERSION };
    static const struct option kopts[] = {
        {"threads", required_argument, nullptr, 't'}, {"coverage", required_argument, nullptr, 'c'}, {"prefix", required_argument, nullptr, 'p'},
        {"directory", required_argument, nullptr, 'o'}, {"barcode", required_argument, nullptr, 'b'}, {"lower", required_argument, nullptr, 'l'},
        {"upper", required_argument, nullptr, 'u'}, {"step", required_argument, nullptr, 's'}, {"verbose", no_argument, nullptr, 'v'},
        {"help", no_argument, nullptr, KOPT_HELP}, {"version", no_argument, nullptr, KOPT_VERSION}, {nullptr, 0, nullptr, 0}};
    int thread = 1, coverage = 90, lower = 15, upper = 35, step = 1;
    std::string prefix, directory, barcode;
    bool die = false;
    optind = 1;
    for (int c; (c = getopt_long(argc, argv, "t:c:p:o:b:l:u:s:v", kopts, nullptr)) != -1;)
    {
        std::istringstream arg(optarg != nullptr ? optarg : "");
        switch (c)
        {
            case 't': arg >> thread; break;
            case 'c': arg >> coverage; break;
            case 'p': arg >> prefix; break;
            case 'o': arg >> directory; break;
            case 'b': arg >> barcode; break;
            case 'l': arg >> lower; break;
            case 'u': arg >> upper; break;
            case 's': arg >> step; break;
            case 'v': break;
            case KOPT_HELP: std::cerr << KMERCHECK_USAGE_MESSAGE; exit(EXIT_SUCCESS);
            case KOPT_VERSION: std::cerr << "kmercheck Version " PACKAGE_VERSION "\n\n"; exit(EXIT_SUCCESS);
            default: die = true; break;
        }
    }
    if (argc - optind < 1) { std::cerr << "kmercheck: missing arguments\n"; die = true; }
    else if (argc - optind > 1) { std::cerr << "kmercheck: too many arguments\n"; die = true; }
    if (thread <= 0) { std::cerr << "kmercheck: invalid number of threads: " << thread << "\n"; die = true; }
    if (coverage <= 0) { std::cerr << "kmercheck: invalid coverage: " << coverage << "\n"; die = true; }
    if (prefix.empty()) { std::cerr << "kmercheck: no prefix\n"; die = true; }
    if (directory.empty()) { std::cerr << "kmercheck: no directory\n"; die = true; }
    else
    {
        directory += "/";
        if (system(("mkdir -p " + directory).c_str()) != 0) { std::cerr << "kmercheck: something wrong in directory: " << directory << "\n"; die = true; }
    }
    if (barcode.empty()) { std::cerr << "kmercheck: no barcode\n"; die = true; }
    if (!(lower >= 9 && upper >= lower)) { std::cerr << "kmercheckinvalid range of kmer size:" << lower << " - " << upper << '\n'; die = true; }
    if (step <= 0) { std::cerr << "kmercheckinvalid step size: " << step << '\n'; die = true; }
    if (die) { std::cerr << "\n" << KMERCHECK_USAGE_MESSAGE; exit(EXIT_FAILURE); }
    const std::string readsFile = argv[optind];
    if (pbsc_device_count() <= 0) { std::cerr << "kmercheck: no CUDA device available (this build has no CPU path)\n"; return EXIT_FAILURE; }
    std::cerr << "Loading BWT: " << prefix << ".bwt\n" << "Loading RBWT: " << prefix << ".rbwt\n";
    pbsc_index* idx = nullptr;
    int from_fmg = 0;
    if (pbsc_index_open(prefix.c_str(), 0, 0, 13, 0, &from_fmg, &idx) != PBSC_OK) { std::cerr << pbsc_last_error() << "\n"; return EXIT_FAILURE; }
    std::cerr << "Loading BARCODE: " << barcode << '\n';
    pbsc::bcode::Table table;
    std::string err;
    if (!pbsc::bcode::load(barcode, table, err)) { std::cerr << "Error: " << err << "\n"; return EXIT_FAILURE; }
    std::cerr << "Using kmer size : " << lower << " - " << upper << " (" << step << ")\n";
    const auto t0 = std::chrono::steady_clock::now();
    LineReader in(readsFile);
    if (!in.ok()) { std::cerr << "Error: could not open " << readsFile << " for read\n"; return EXIT_FAILURE; }
    std::map<int, pbsc::bcode::Histogram> crt, wrong;
    // queries of one flush: both strands' strings of every k-mer, and where the answer goes
    struct Query { uint32_t read; int32_t pos; int32_t k; uint32_t block; };
    std::vector<Query> queries;
    std::string fwd_s, rvc_s;                 // reverse(w) for the RBWT, reverse-complement(w) for the BWT, concatenated
    std::vector<uint64_t> q_off(1, 0);
    std::vector<std::string> seqs;            // reads of the current flush
    std::vector<const std::vector<pbsc::bcode::Block>*> seq_blocks;
    size_t nreads = 0;
    int rc_all = 0;
    auto flush = [&]() -> bool {
        const uint64_t n = queries.size();
        if (n)
        {
            std::vector<int64_t> lo(n), hi(n), lo2(n), hi2(n);
            if (pbsc_findinterval_batch(idx, PBSC_RBWT, fwd_s.data(), q_off.data(), n, lo.data(), hi.data(), nullptr) != PBSC_OK ||
                pbsc_findinterval_batch(idx, PBSC_BWT, rvc_s.data(), q_off.data(), n, lo2.data(), hi2.data(), nullptr) != PBSC_OK)
            { std::cerr << "kmercheck: " << pbsc_last_error() << "\n"; return false; }
            for (uint64_t i = 0; i < n; i++)
            {
                const Query& q = queries[i];
                const int64_t freq = (hi[i] >= lo[i] ? hi[i] - lo[i] + 1 : 0) + (hi2[i] >= lo2[i] ? hi2[i] - lo2[i] + 1 : 0);
                if (freq == 0) { std::cerr << "kmercheck: a k-mer of read " << q.read << " does not occur in the index (is " << prefix << " the index of these reads?)\n"; return false; }
                if (freq == 1) continue;
                bool ok;
                try { ok = pbsc::bcode::validate(q.pos, q.k, (*seq_blocks[q.read])[q.block], seqs[q.read]); }
                catch (const std::exception& e) { std::cerr << "kmercheck: " << e.what() << "\n"; return false; }
                (ok ? crt : wrong)[q.k].add((int)freq);
            }
        }
        queries.clear(); fwd_s.clear(); rvc_s.clear(); q_off.assign(1, 0); seqs.clear(); seq_blocks.clear();
        return true;
    };
    for (;;)
    {
        Batch one;
        one.bases = g_pins.get(1 << 16);
        if (!readRecord(in, one)) { one.release(); break; }
        std::string bad;
        if (!normalize(one, bad)) { std::cerr << "kmercheck: read " << bad << " holds letters other than ACGT\n"; one.release(); rc_all = 1; break; }
        nreads++;
        pbsc::bcode::Table::const_iterator it = table.find(one.ids[0]);
        if (it != table.end())
        {
            const uint32_t r = (uint32_t)seqs.size();
            seqs.push_back(std::string(one.bases.p, one.n_bases));
            seq_blocks.push_back(&it->second);
            const std::string& seq = seqs.back();
            for (uint32_t b = 0; b < it->second.size(); b++)
                for (int k = lower; k <= upper; k += step)
                    for (int pos = it->second[b].start; pos <= it->second[b].end - k; pos++)
                    {
                        if (pos < 0 || (size_t)(pos + k) > seq.size()) continue;   // KmerFeature::isFake: the reference asserts on these
                        for (int j = k - 1; j >= 0; j--)
                        {
                            const char ch = seq[(size_t)(pos + j)];
                            fwd_s += ch;
                            rvc_s += ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : 'A';
                        }
                        q_off.push_back(fwd_s.size());
                        queries.push_back(Query{r, pos, k, b});
                    }
        }
        one.release();
        if (fwd_s.size() > (64u << 20) && !flush()) { rc_all = 1; break; }
    }
    if (rc_all == 0 && !flush()) rc_all = 1;
    pbsc_index_destroy(idx);
    if (rc_all) return EXIT_FAILURE;
    {
        // KmerCheckPostProcess (KmerCheckProcess.cpp:42-53): both files are opened for appending
        std::ofstream total_box((directory + "total.box").c_str(), std::ios_base::app), value_box((directory + "value.box").c_str(), std::ios_base::app);
        for (int k = lower; k <= upper; k += step) pbsc::bcode::compare(total_box, value_box, coverage, k, crt[k], wrong[k]);
    }
    const double proc = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "Processed %zu sequences in %lfs (%lf sequences/s)\n", nreads, proc, (double)nreads / proc);
    return 0;
}

int main(int argc, char** argv)
{
    // `pbcorrect kmercheck ...` = `stride kmercheck ...`
    if (argc > 1 && std::string(argv[1]) == "kmercheck") return kmercheckMain(argc - 1, argv + 1);
    // `pbcorrect index ...` = `stride index ...`
    if (argc > 1 && std::string(argv[1]) == "index") return indexMain(argc - 1, argv + 1);
    // accept both `pbcorrect [opts] READS` and `pbcorrect pbcorrect [opts] READS` (as `stride pbcorrect`)
    if (argc > 1 && std::string(argv[1]) == "pbcorrect") { argv++; argc--; }
    parseOptions(argc, argv);
    int rc = pbsc_params_derive(&opt::params);
    if (rc != PBSC_OK) { std::cerr << SUBPROGRAM ": " << pbsc_last_error() << "\n"; return EXIT_FAILURE; }
    int ndev = pbsc_device_count();
    if (ndev <= 0) { std::cerr << SUBPROGRAM ": no CUDA device available (this build has no CPU path)\n"; return EXIT_FAILURE; }
    const int ngpu = opt::gpus > 0 ? std::min(opt::gpus, ndev) : ndev;

    // StriDe/PacBioSelfCorrection.cpp:191 (BCode::load, PacBio/BCode.cpp:27-49)
    pbsc::bcode::Table barcodes;
    if (opt::OnlySeed)
    {
        std::cerr << "Loading BARCODE: " << opt::barcode << '\n';
        std::string err;
        if (!pbsc::bcode::load(opt::barcode, barcodes, err)) { std::cerr << "Error: " << err << "\n"; return EXIT_FAILURE; }
        opt::params.no_dp = 1;   // nothing of the correction is reported in this mode: no reason to run the fallback
    }
    auto t_load = std::chrono::steady_clock::now();
    std::cerr << "Loading BWT: " << opt::prefix << ".bwt\n" << "Loading RBWT: " << opt::prefix << ".rbwt\n"
              << "Loading Sampled Suffix Array: " << opt::prefix << ".sai\n";
    // The index is read and decoded ONCE (PREFIX.fmg, the persisted flat tables, when it matches PREFIX.bwt/.rbwt; else the
    // run-length files), on GPU 0; the other GPUs receive it by peer copies over NVLink, all at the same time.
    std::vector<pbsc_index*> index(ngpu, nullptr);
    int from_fmg = 0;
    if (pbsc_index_open(opt::prefix.c_str(), 0, 1, opt::prefix_k, opt::write_fmg ? 1 : 0, &from_fmg, &index[0]) != PBSC_OK)
    { std::cerr << pbsc_last_error() << "\n"; return EXIT_FAILURE; }
    {
        std::vector<std::thread> th;
        std::vector<int> rcs(ngpu, 0);
        std::vector<std::string> errs(ngpu);
        for (int g = 1; g < ngpu; g++)
            th.emplace_back([&, g]() {
                rcs[g] = pbsc_index_clone(index[0], g, &index[g]);
                if (rcs[g] != PBSC_OK) errs[g] = pbsc_last_error();
            });
        for (auto& t : th) t.join();
        for (int g = 0; g < ngpu; g++)
        {
            if (rcs[g] != PBSC_OK) { std::cerr << errs[g] << "\n"; return EXIT_FAILURE; }
            if (pbsc_index_set_lanes(index[g], opt::lanes) != PBSC_OK) { std::cerr << pbsc_last_error() << "\n"; return EXIT_FAILURE; }
        }
    }
    double load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load).count();
    std::cerr << "[timer - index to " << ngpu << " GPU(s)] wall clock: " << load_s << "s (" << pbsc_index_device_bytes(index[0]) / 1e9 << " GB per GPU, from "
              << (from_fmg ? "PREFIX.fmg" : "PREFIX.bwt/.rbwt") << (ngpu > 1 ? ", peer copies to the other GPUs" : "") << ")\n";

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
    // --onlyseed (PacBioSelfCorrectionProcess.cpp:265-287): DIR/total.seed instead of correct.fa / discard.fa
    FILE* correct = opt::OnlySeed ? nullptr : fopen((opt::directory + "correct.fa").c_str(), "wb");
    FILE* discard = opt::OnlySeed ? nullptr : fopen((opt::directory + "discard.fa").c_str(), "wb");
    FILE* seed_status = opt::OnlySeed ? fopen((opt::directory + "total.seed").c_str(), "w") : nullptr;
    if (opt::OnlySeed ? !seed_status : (!correct || !discard)) { std::cerr << "Error: could not open output files in " << opt::directory << "\n"; return EXIT_FAILURE; }
    std::vector<char> wbuf_c(8 << 20), wbuf_d(1 << 20);
    if (correct) setvbuf(correct, wbuf_c.data(), _IOFBF, wbuf_c.size());
    if (discard) setvbuf(discard, wbuf_d.data(), _IOFBF, wbuf_d.size());
    size_t seed_total[3] = {0, 0, 0};

    auto t0 = std::chrono::steady_clock::now();
    // ---- pipeline (replaces Concurrency/SequenceProcessFramework.h:91-230):
    //   reader (this thread: record boundaries, one copy of the sequence bytes into page-locked memory)
    //   -> normalizers (-t threads: upper-case + ACGT check)
    //   -> GPU workers (lanes + 1 host threads per GPU: pbsc_correct_batch = upload on a copy stream, kernels on a lane of the
    //      index, fetch on the copy stream; one thread's copies overlap the other threads' kernels)
    //   -> writer (one thread, batches in input order: the reference's `-t 1` record order)
    // The reader stops while too many batches are in flight: memory is bounded by the GPU count, not by the file size. ----
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::unique_ptr<Batch>> inflight;   // batches [written, produced), in input order
    size_t produced = 0, written = 0;
    bool reading_done = false, failed = false;
    std::string fail_msg;
    auto fail = [&](const std::string& m) { std::lock_guard<std::mutex> lk(mu); if (!failed) { failed = true; fail_msg = m; } cv.notify_all(); };
    auto at = [&](size_t seq) -> Batch* { return inflight[seq - written].get(); };   // call with mu held
    const int n_workers = ngpu * (opt::lanes + 1);
    const size_t max_inflight = (size_t)2 * n_workers + 2;

    auto normalizer = [&]() {
        for (;;)
        {
            Batch* b = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] {
                    if (failed) return true;
                    for (size_t s = written; s < produced; s++) { Batch* x = at(s); if (!x->normalized && !x->taken) { b = x; return true; } }
                    return reading_done;
                });
                if (failed || !b) return;
                b->taken = true;
            }
            std::string bad;
            const bool ok = normalize(*b, bad);
            if (!ok) { fail("Error: read " + bad + " contains non-ACGT characters.\nPlease run sga preprocess on the data first."); return; }
            { std::lock_guard<std::mutex> lk(mu); b->normalized = true; b->taken = false; }
            cv.notify_all();
        }
    };
    auto worker = [&](int g) {
        for (;;)
        {
            Batch* b = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] {
                    if (failed) return true;
                    bool pending = false;
                    for (size_t s = written; s < produced; s++)
                    {
                        Batch* x = at(s);
                        if (x->done) continue;
                        if (x->normalized && !x->taken) { b = x; return true; }
                        pending = true;
                    }
                    return reading_done && !pending;
                });
                if (failed || !b) return;
                b->taken = true;
            }
            const uint64_t n = b->ids.size();
            uint64_t cap = (uint64_t)(b->n_bases * 1.3) + (1 << 16), need = 0;
            b->first.assign(n + 1, 0);
            b->stats.assign(n ? n : 1, pbsc_read_stats{});
            if (P.debug_seed)
            {
                // the staged interface, so that the batch is still there for pbsc_batch_fetch_debug
                pbsc_batch* h = nullptr;
                b->rc = pbsc_batch_upload(index[g], &P, b->bases.p, b->offsets.data(), n, &h);
                if (b->rc == PBSC_OK) b->rc = pbsc_batch_run(h, nullptr);
                uint64_t nb = 0, np = 0, ns = 0, nl = 0;
                if (b->rc == PBSC_OK) b->rc = pbsc_batch_result_size(h, &nb, &np);
                if (b->rc == PBSC_OK)
                {
                    b->pieces = g_pins.get(nb + 64);
                    b->piece_off.assign(np + 16, 0);
                    b->rc = pbsc_batch_fetch(h, b->pieces.p, b->pieces.cap, b->piece_off.data(), b->piece_off.size(), b->first.data(), b->stats.data());
                }
                if (b->rc == PBSC_OK) b->rc = pbsc_batch_debug_size(h, &ns, &nl);
                if (b->rc == PBSC_OK)
                {
                    b->dbg_seeds.resize(ns + 1); b->dbg_log.resize(nl + 1); b->dbg_seed_off.assign(n + 1, 0); b->dbg_log_off.assign(n + 1, 0);
                    b->dbg_surv.assign(n + 1, 0); b->dbg_ratio.resize(b->n_bases + 1);
                    b->rc = pbsc_batch_fetch_debug(h, b->dbg_seeds.data(), b->dbg_seeds.size(), b->dbg_seed_off.data(), b->dbg_surv.data(), b->dbg_ratio.data(),
                                                   b->dbg_ratio.size(), b->dbg_log.data(), b->dbg_log.size(), b->dbg_log_off.data());
                }
                if (b->rc != PBSC_OK) { fail(pbsc_last_error()); pbsc_batch_destroy(h); return; }
                pbsc_batch_destroy(h);
            }
            else
            for (;;)
            {
                if (b->pieces.cap < cap) { g_pins.put(b->pieces); b->pieces = g_pins.get(cap); }
                b->piece_off.assign((P.split ? b->n_bases / 10 + 4 * n : n) + 16, 0);
                b->rc = pbsc_correct_batch(index[g], &P, b->bases.p, b->offsets.data(), n, b->pieces.p, b->pieces.cap, b->piece_off.data(),
                                           b->piece_off.size(), b->first.data(), b->stats.data(), &need);
                if (b->rc == PBSC_ERR_LIMIT && need > b->pieces.cap) { cap = need + 64; continue; }
                break;
            }
            if (b->rc != PBSC_OK) { fail(pbsc_last_error()); return; }
            pbsc_last_timing(&b->timing);
            { std::lock_guard<std::mutex> lk(mu); b->done = true; }
            cv.notify_all();
        }
    };

    int64_t totalReadsLen = 0, correctedLen = 0, totalSeedNum = 0, totalWalkNum = 0, highErrorNum = 0, exceedDepthNum = 0, exceedLeaveNum = 0, FMNum = 0,
            DPNum = 0, seedDis = 0;
    double seed_s = 0, fm_s = 0, dp_s = 0;
    auto writer = [&]() {
        std::string line;
        for (;;)
        {
            Batch* b = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || (written < produced && at(written)->done) || (reading_done && written == produced); });
                if (failed || written == produced) return;
                b = at(written);
            }
            for (size_t r = 0; r < b->ids.size() && opt::OnlySeed; r++)
            {
                // PacBioSelfCorrectionPostProcess::process, OnlySeed branch (:315-335): every surviving seed against the read's barcode blocks
                size_t status[3] = {0, 0, 0};
                const std::string seq(b->bases.p + b->offsets[r], (size_t)(b->offsets[r + 1] - b->offsets[r]));
                pbsc::bcode::Table::const_iterator it = barcodes.find(b->ids[r]);
                const pbsc_seed* sv = b->dbg_seeds.data() + b->dbg_seed_off[r];
                try
                {
                    for (uint32_t i = 0; i < b->dbg_surv[r]; i++)
                        status[pbsc::bcode::classify(it == barcodes.end() ? nullptr : &it->second, sv[i].start, sv[i].len, seq)]++;
                }
                catch (const std::exception& e) { fail(std::string(SUBPROGRAM ": read ") + b->ids[r] + ": " + e.what()); return; }
                pbsc::bcode::summarize(seed_status, status, b->ids[r]);
                for (int k = 0; k < 3; k++) seed_total[k] += status[k];
            }
            for (size_t r = 0; r < b->ids.size() && !opt::OnlySeed; r++)
            {
                const pbsc_read_stats& st = b->stats[r];
                if (st.merge)
                {
                    totalReadsLen += st.total_reads_len; correctedLen += st.corrected_len; totalSeedNum += st.total_seed_num;
                    totalWalkNum += st.total_walk_num; highErrorNum += st.high_error_num; exceedDepthNum += st.exceed_depth_num;
                    exceedLeaveNum += st.exceed_leave_num; FMNum += st.fm_num; DPNum += st.dp_num; seedDis += st.seed_dis;
                    for (uint64_t j = b->first[r]; j < b->first[r + 1]; j++)
                    {
                        line.assign(">"); line += b->ids[r];
                        if (P.split) { line += "_"; line += std::to_string(j - b->first[r]); }
                        line += "\n";
                        fwrite(line.data(), 1, line.size(), correct);
                        fwrite(b->pieces.p + b->piece_off[j], 1, (size_t)(b->piece_off[j + 1] - b->piece_off[j]), correct);
                        fputc('\n', correct);
                    }
                }
                else
                {
                    line.assign(">"); line += b->ids[r]; line += "\n";
                    fwrite(line.data(), 1, line.size(), discard);
                    fwrite(b->bases.p + b->offsets[r], 1, (size_t)(b->offsets[r + 1] - b->offsets[r]), discard);
                    fputc('\n', discard);
                }
            }
            if (P.debug_seed) writeDebugFiles(*b);
            seed_s += b->timing.seed_ms / 1e3; fm_s += (b->timing.extend_ms - b->timing.dp_ms) / 1e3; dp_s += b->timing.dp_ms / 1e3;
            b->release();
            { std::lock_guard<std::mutex> lk(mu); inflight.pop_front(); written++; }
            cv.notify_all();
        }
    };

    std::vector<std::thread> threads;
    for (int i = 0; i < std::max(1, opt::thread - 1); i++) threads.emplace_back(normalizer);
    for (int w = 0; w < n_workers; w++) threads.emplace_back(worker, w % ngpu);
    threads.emplace_back(writer);

    uint64_t batch_bases = (uint64_t)(opt::batch_mbp * 1e6);
    {
        // several GPUs: no batch larger than an even share of the input, so that every GPU gets work (a plain FASTA file is
        // about one byte per base; a .gz or FASTQ input only makes the batches smaller than they had to be)
        struct stat sb;
        if (ngpu > 1 && stat(opt::readsFile.c_str(), &sb) == 0 && sb.st_size > 0)
            batch_bases = std::min<uint64_t>(batch_bases, std::max<uint64_t>(16000000ull, ((uint64_t)sb.st_size + ngpu - 1) / ngpu));
    }
    size_t nreads = 0;
    uint64_t inBases = 0;
    auto push = [&](std::unique_ptr<Batch>& cur) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return failed || produced - written < max_inflight; });   // backpressure
        if (failed) return;
        cur->seq = produced++;
        inflight.push_back(std::move(cur));
        lk.unlock();
        cv.notify_all();
    };
    std::unique_ptr<Batch> cur(new Batch());
    cur->bases = g_pins.get(batch_bases + (1 << 20));
    for (;;)
    {
        { std::lock_guard<std::mutex> lk(mu); if (failed) break; }
        const size_t before = cur->n_bases;
        if (!readRecord(in, *cur)) break;
        inBases += cur->n_bases - before;
        nreads++;
        if (cur->n_bases >= batch_bases)
        {
            push(cur);
            cur.reset(new Batch());
            cur->bases = g_pins.get(batch_bases + (1 << 20));
        }
    }
    if (!cur->ids.empty()) push(cur); else if (cur) cur->release();
    { std::lock_guard<std::mutex> lk(mu); reading_done = true; }
    cv.notify_all();
    for (auto& t : threads) t.join();
    if (correct) fclose(correct);
    if (discard) fclose(discard);
    for (auto* ix : index) pbsc_index_destroy(ix);
    if (failed) { std::cerr << fail_msg << "\n"; return EXIT_FAILURE; }

    double proc = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "Processed %zu sequences in %lfs (%lf sequences/s)\n", nreads, proc, (double)nreads / proc);
    fprintf(stderr, "[timer - " PACKAGE_NAME "::" SUBPROGRAM "] wall clock: %.2fs, %.3f Mbp/s over %d GPU(s)\n", proc, inBases / 1e6 / proc, ngpu);

    // PacBioSelfCorrectionPostProcess::~PacBioSelfCorrectionPostProcess — PacBioSelfCorrectionProcess.cpp:281-311
    if (opt::OnlySeed)
    {
        pbsc::bcode::summarize(stdout, seed_total, "TOTAL");
        fclose(seed_status);
    }
    else if (totalWalkNum > 0 && totalReadsLen > 0)
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
