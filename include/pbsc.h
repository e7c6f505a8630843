/*
 * pbsc.h — C ABI of the B200-native `stride pbcorrect` hot path (libpbsc.so).
 *
 * The reference (ccuchengwei/LongReadSelfCorrect) has no plugin/FFI interface for this
 * path: its boundary is the template policy pair Processor/PostProcessor of
 * Concurrency/SequenceProcessFramework.h:359-386 around PacBioSelfCorrectionProcess::process
 * (PacBio/PacBioSelfCorrectionProcess.cpp:23-54).  Each entry point below names the
 * reference interface it replaces.  Conventions: plain pointers and sizes only, int return
 * code (0 = ok, negative = error, text via pbsc_last_error()), no exceptions cross the
 * boundary, the caller owns every host buffer, the library owns device memory behind
 * opaque handles.  There is NO CPU fallback: every compute entry point fails with
 * PBSC_ERR_CUDA when no CUDA device is usable.
 */
#ifndef PBSC_H
#define PBSC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBSC_OK 0
#define PBSC_ERR_ARG (-1)
#define PBSC_ERR_IO (-2)
#define PBSC_ERR_FORMAT (-3)
#define PBSC_ERR_CUDA (-4)
#define PBSC_ERR_LIMIT (-5)
#define PBSC_ERR_INTERNAL (-6)

#define PBSC_BWT 0  /* PREFIX.bwt  : BWT of the reads            (BWTIndexSet::pBWT)  */
#define PBSC_RBWT 1 /* PREFIX.rbwt : BWT of the reversed reads   (BWTIndexSet::pRBWT) */

/* Walk outcome codes of LongReadSelfCorrectByOverlap::extendOverlap
 * (PacBio/LongReadCorrectByOverlap.cpp:155-211). */
#define PBSC_WALK_OK 1
#define PBSC_WALK_HIGH_ERROR (-1)
#define PBSC_WALK_EXCEED_DEPTH (-2)
#define PBSC_WALK_EXCEED_LEAVES (-3)
#define PBSC_WALK_NO_PATH (-4)

typedef struct pbsc_index pbsc_index; /* both rank tables + prefix table, resident on one GPU */

/* Option block of `stride pbcorrect` (StriDe/PacBioSelfCorrection.cpp:71-101) plus the
 * values PacBioSelfCorrectionMain derives from it (:195-231).  Fill with
 * pbsc_params_default(), set the option fields, then call pbsc_params_derive(). */
typedef struct pbsc_params
{
    /* command-line options */
    int32_t pb_coverage;  /* -c, default 90  */
    double error_rate;    /* -e, default 0.15 */
    int32_t start_kmer;   /* -k, default 19 (sets `adjust`) */
    int32_t next_target;  /* -n, default 1   */
    int32_t max_leaves;   /* -l, default 32  */
    int32_t idmer_len;    /* -i, default 9   */
    int32_t min_kmer;     /* -s, default 13  */
    int32_t genome;       /* -g, 5/10/100, default 10 */
    int32_t mode;         /* -m, default 1 (sets `manual`) */
    int32_t manual;       /* -m given */
    int32_t adjust;       /* -k/-u/-r given */
    int32_t split;        /* --split */
    int32_t no_dp;        /* --nodp  */
    int32_t offset[3];    /* offset[1] = -u, offset[2] = -r */
    /* derived by pbsc_params_derive() */
    int32_t pool[8];      /* ascending k-mer sizes, KmerFeature::Log() keys */
    int32_t n_pool;
    int32_t scan_kmer;    /* 19 */
    int32_t kmer_up_bound;/* 50 */
    int32_t radius;       /* 100 */
    float hh_ratio;       /* 0.6f */
    float threshold[3][52]; /* KmerThreshold table[mode][k] (PacBio/KmerThreshold.cpp:43-79) */
    double freqs_of_kmer[101]; /* pow(1-e,i)*cov, i>=min_kmer (LongReadCorrectByOverlap.cpp:68-70) */
    int32_t debug_seed;   /* --debugseed: batches keep what pbsc_batch_fetch_debug returns */
    int32_t reserved;
} pbsc_params;

/* One seed of a read: SeedFeature (PacBio/SeedFeature.h:22-45) after
 * LongReadProbe::searchSeedsWithHybridKmers (PacBio/LongReadProbe.cpp:34-117). */
typedef struct pbsc_seed
{
    int32_t start;            /* seedStartPos */
    int32_t len;              /* seedLen */
    int32_t max_fixed_freq;   /* maxFixedMerFreq */
    int32_t is_repeat;        /* isRepeat */
    int32_t start_best_k;     /* startBestKmerSize */
    int32_t end_best_k;       /* endBestKmerSize */
    int32_t hitchhiked;       /* isHitchhiked (such seeds are reported only by pbsc_seed_batch with keep_outcast) */
    int32_t static_k;         /* static k-mer size the seed was grown from */
} pbsc_seed;

/* Per-read counters of PacBioSelfCorrectionResult (PacBio/PacBioSelfCorrectionProcess.h:59-95). */
typedef struct pbsc_read_stats
{
    int64_t total_reads_len, corrected_len, total_seed_num, total_walk_num, high_error_num,
        exceed_depth_num, exceed_leave_num, fm_num, dp_num, seed_dis;
    int32_t merge;      /* result.merge: 1 -> correct.fa, 0 -> discard.fa */
    int32_t n_pieces;   /* >1 only with --split */
} pbsc_read_stats;

const char* pbsc_last_error(void);
int pbsc_device_count(void);
/* Page-locked host memory for the caller's read and result buffers (optional: every entry point also accepts pageable
 * memory).  From pinned buffers the H2D/D2H copies of pbsc_correct_batch / pbsc_batch_upload / pbsc_batch_fetch run at
 * PCIe speed.  The reference has no counterpart (its reads live in std::string, SequenceWorkItem.h). */
int pbsc_host_alloc(void** p, size_t bytes);
void pbsc_host_free(void* p);
/* page-lock memory the caller owns (cudaHostRegister), e.g. a shared mapping that several processes fill with results */
int pbsc_host_register(void* p, size_t bytes);
void pbsc_host_unregister(void* p);
/* give the library's cache of per-batch device blocks on `device` back to the driver */
void pbsc_trim(int device);
/* measured peak (GB/s) of independent random 32-byte sector reads over `bytes` bytes of HBM: the roofline denominator for
 * kernels whose every rank query is one aligned sector (no reference counterpart; measurement aid) */
int pbsc_random_sector_bench(int device, uint64_t bytes, float* gbps);

/* ---- parameters: PacBioSelfCorrectionMain, StriDe/PacBioSelfCorrection.cpp:195-231 ---- */
void pbsc_params_default(pbsc_params* p);
int pbsc_params_derive(pbsc_params* p);
/* text of DIR/threshold-table (KmerThreshold::~KmerThreshold/write, PacBio/KmerThreshold.cpp:31-41,65-72);
 * returns the length written (excluding NUL) or a negative error */
int pbsc_threshold_table_text(const pbsc_params* p, char* buf, size_t cap);

/* ---- index: replaces RLBWT (SuffixTools/RLBWT.{h,cpp}), BWTReaderBinary::read
 *      (SuffixTools/BWTReaderBinary.cpp:27-85) and BWTIndexSet for this path ---- */
/* run bytes as stored on disk after the 30-byte header: high 3 bits symbol rank ($ACGT), low 5 bits run length */
int pbsc_index_create(const uint8_t* bwt_runs, uint64_t bwt_n_runs, uint64_t bwt_n_symbols, uint64_t bwt_n_strings,
                      const uint8_t* rbwt_runs, uint64_t rbwt_n_runs, uint64_t rbwt_n_symbols, uint64_t rbwt_n_strings,
                      int device, pbsc_index** out);
/* reads PREFIX.bwt and PREFIX.rbwt; PREFIX.sai must exist (content unused, as in the reference) unless require_sai==0 */
int pbsc_index_load(const char* prefix, int device, int require_sai, pbsc_index** out);
/* synthetic index for the FM microbenchmark: i.i.d. uniform ACGT symbols with n_strings '$' at random
 * positions, generated on the device from `seed` (both strands get independent streams) */
int pbsc_index_create_synthetic(uint64_t n_symbols, uint64_t n_strings, uint64_t seed, int device, pbsc_index** out);
/* build the short-prefix interval table for all k0-mers (k0 in 1..15); 0 disables it */
int pbsc_index_build_prefix_table(pbsc_index* idx, int k0);
void pbsc_index_destroy(pbsc_index* idx);

/* ---- the index as one relocatable blob (512-byte header + the tables exactly as they sit in HBM).  No counterpart in the
 *      reference, which re-reads PREFIX.bwt/.rbwt and rebuilds its rank directory on every start
 *      (SuffixTools/BWTReaderBinary.cpp:55-85, SuffixTools/RLBWT.cpp:109-248) once per process
 *      (StriDe/PacBioSelfCorrection.cpp:155-172). ---- */
/* PREFIX.fmg, the persisted flat index: written once, later runs skip the run-length decode and the prefix-table build */
int pbsc_index_save(const pbsc_index* idx, const char* path);
int pbsc_index_load_fmg(const char* path, int device, pbsc_index** out);
/* `pbcorrect -p PREFIX`: PREFIX.fmg if present, well-formed and describing the same PREFIX.bwt/.rbwt (string, symbol and run
 * counts of both headers) with a k0 prefix table; else the run-length files (+ prefix table), and PREFIX.fmg is written for
 * the next run when write_fmg != 0.  *from_fmg (may be NULL) reports which way it went. */
int pbsc_index_open(const char* prefix, int device, int require_sai, int k0, int write_fmg, int* from_fmg, pbsc_index** out);
/* a copy of `src` on another GPU of the box: one peer copy per table over NVLink (cudaMemcpyPeerAsync), no host decode */
int pbsc_index_clone(const pbsc_index* src, int device, pbsc_index** out);
/* the same bytes through caller-owned memory, for callers that move the index themselves (one process per GPU + NCCL
 * broadcast): export writes the blob to a device buffer on the index's GPU; import copies a blob that sits on GPU
 * src_device (or in host memory when src_device < 0) into a new index on `device`. */
int pbsc_index_blob_size(const pbsc_index* idx, uint64_t* bytes);
int pbsc_index_export_blob(const pbsc_index* idx, void* d_dst, uint64_t cap);
int pbsc_index_import_blob(const void* src, uint64_t bytes, int src_device, int device, pbsc_index** out);

/* ---- lanes: how many batches of this index may RUN at the same time (1..4, default 1), each on its own stream with its own
 *      scratch arena, so that the tail of one batch's rounds overlaps the dense kernels of another.  Together with the
 *      per-batch copy streams of pbsc_batch_upload / pbsc_batch_fetch this replaces the worker threads of
 *      Concurrency/SequenceProcessFramework.h:91-230: callers drive it from `lanes` (or more) host threads. ---- */
int pbsc_index_set_lanes(pbsc_index* idx, int lanes);
int pbsc_index_lanes(const pbsc_index* idx);
uint64_t pbsc_index_num_symbols(const pbsc_index* idx, int which);
uint64_t pbsc_index_num_strings(const pbsc_index* idx, int which);
uint64_t pbsc_index_device_bytes(const pbsc_index* idx);
/* decoded BWT symbol at position i (RLBWT::getChar, SuffixTools/RLBWT.h:42-63); test hook */
int pbsc_index_get_symbols(const pbsc_index* idx, int which, uint64_t first, uint64_t count, char* out);

/* ---- backward search: BWTAlgorithms::findInterval(const BWT*, w) (SuffixTools/BWTAlgorithms.cpp:14-31) ----
 * kmers: n strings over ACGT concatenated, offsets[n+1].  lower/upper follow the reference convention
 * (inclusive; lower > upper when w is absent).  steps[i] (optional) = updateInterval calls executed
 * before the early break.  Host-buffer entry: copies are part of the call. */
int pbsc_findinterval_batch(pbsc_index* idx, int which, const char* kmers, const uint64_t* offsets, uint64_t n,
                            int64_t* lower, int64_t* upper, uint8_t* steps);
/* fixed-k device-resident variant for timing: kmers2bit is a device pointer to n k-mers packed 2 bits/base
 * in one uint64 each (base j of the k-mer at bits 2j..2j+1, A=0,C=1,G=2,T=3; k<=32), outputs are device pointers.
 * Returns kernel milliseconds in *ms (CUDA events on the launch stream). */
int pbsc_findinterval_device(pbsc_index* idx, int which, const uint64_t* d_kmers2bit, int k, uint64_t n,
                             int64_t* d_lower, int64_t* d_upper, uint8_t* d_steps, float* ms);

/* ---- seed phase: LongReadProbe::searchSeedsWithHybridKmers (PacBio/LongReadProbe.cpp:34-227) ----
 * reads: n_reads strings over ACGT concatenated, offsets[n_reads+1].  seeds_out receives the surviving seeds of
 * read r at seed_offsets[r]..seed_offsets[r+1]; seed_offsets has n_reads+1 entries.  If seeds_cap is too small the
 * call fails with PBSC_ERR_LIMIT and *seeds_needed says how many are required. */
int pbsc_seed_batch(pbsc_index* idx, const pbsc_params* p, const char* reads, const uint64_t* offsets, uint64_t n_reads,
                    pbsc_seed* seeds_out, uint64_t seeds_cap, uint64_t* seed_offsets, uint64_t* seeds_needed,
                    int keep_outcast);

/* ---- FM extension of explicit seed pairs: LongReadSelfCorrectByOverlap ctor + extendOverlap
 *      (PacBio/LongReadCorrectByOverlap.cpp:17-95,155-211), as called from correctByFMExtension
 *      (PacBio/PacBioSelfCorrectionProcess.cpp:186-192).  For pair i: src/path/trg strings (already swapped and
 *      reverse-complemented by the caller when isFromRtoU), dis = disBetweenSrcTarget, k = initkmersize
 *      (maxOverlap = k+2), min_sa = min_SA_threshold.  status[i] = PBSC_WALK_*; on success the merged
 *      sequence is out[out_offsets[i]..out_offsets[i+1]). ---- */
int pbsc_extend_batch(pbsc_index* idx, const pbsc_params* p, uint64_t n_pairs,
                      const char* src, const uint64_t* src_off, const char* path, const uint64_t* path_off,
                      const char* trg, const uint64_t* trg_off, const int32_t* dis, const int32_t* k,
                      const int32_t* min_sa, int32_t* status, char* out, uint64_t out_cap, uint64_t* out_offsets);

/* ---- whole hot path for a batch of reads: PacBioSelfCorrectionProcess::process
 *      (PacBio/PacBioSelfCorrectionProcess.cpp:23-245): seeds, FM extension and, unless p->no_dp, the DP /
 *      multiple-alignment fallback for failed walks (correctByMSAlignment, LongReadOverlap::buildMultipleAlignment,
 *      Overlapper::extendMatch, MultipleAlignment::calculateBaseConsensus).
 *      pieces_out holds the corrected pieces of read r, concatenated, at piece_offsets[...]; piece p of read r
 *      is pieces_out[piece_offsets[first_piece[r]+p] .. piece_offsets[first_piece[r]+p+1]).  Reads with
 *      stats[r].merge==0 have no pieces (they go to discard.fa).  first_piece has n_reads+1 entries. ---- */
int pbsc_correct_batch(pbsc_index* idx, const pbsc_params* p, const char* reads, const uint64_t* offsets,
                       uint64_t n_reads, char* pieces_out, uint64_t pieces_cap, uint64_t* piece_offsets,
                       uint64_t piece_offsets_cap, uint64_t* first_piece, pbsc_read_stats* stats,
                       uint64_t* bytes_needed);

/* ---- the same path with the batch resident on the device, for callers that pipeline copies themselves
 *      (Concurrency/SequenceProcessFramework.h:91-230 is replaced by this staged interface):
 *      upload (pinned async H2D) -> run (seed + extend kernels only, nothing crosses PCIe) -> fetch (D2H). ---- */
typedef struct pbsc_batch pbsc_batch;
int pbsc_batch_upload(pbsc_index* idx, const pbsc_params* p, const char* reads, const uint64_t* offsets,
                      uint64_t n_reads, pbsc_batch** out);
/* runs the kernels; *ms = device time between CUDA events on the launch stream (may be NULL) */
int pbsc_batch_run(pbsc_batch* b, float* ms);
/* sizes of the packed result: total piece bytes and number of pieces */
int pbsc_batch_result_size(pbsc_batch* b, uint64_t* piece_bytes, uint64_t* n_pieces);
int pbsc_batch_fetch(pbsc_batch* b, char* pieces_out, uint64_t pieces_cap, uint64_t* piece_offsets,
                     uint64_t piece_offsets_cap, uint64_t* first_piece, pbsc_read_stats* stats);
void pbsc_batch_destroy(pbsc_batch* b);

/* ---- --debugseed (PacBio/LongReadProbe.cpp:109-114,172-173,221-226; PacBio/PacBioSelfCorrectionProcess.cpp:72-76,130-131,
 *      139-140): what the reference's per-read dump files hold, for a batch that ran with params.debug_seed != 0.
 *      seeds: the surviving seeds of read r at seed_offsets[r] .. +n_surviving[r], then its hitchhiked ones up to
 *      seed_offsets[r+1] (seed/<id>.seed, seed/error/<id>.seed); ratio: the repeat ratio of every read position
 *      (extend/<id>.log), n_bases floats; log: one record per seed pair whose FM walk failed, in walk order
 *      (extend/<id>.ext: src_start, trg_start, code = walk outcome + 4; dp_failed != 0 also goes to extend/<id>.dp). ---- */
typedef struct pbsc_walk_log { int32_t src_start, trg_start, code, dp_failed; } pbsc_walk_log;
int pbsc_batch_debug_size(pbsc_batch* b, uint64_t* n_seeds, uint64_t* n_log);
int pbsc_batch_fetch_debug(pbsc_batch* b, pbsc_seed* seeds, uint64_t seeds_cap, uint64_t* seed_offsets, uint32_t* n_surviving, float* ratio,
                           uint64_t ratio_cap, pbsc_walk_log* log, uint64_t log_cap, uint64_t* log_offsets);

/* timing/launch counters of the last pbsc_correct_batch / pbsc_batch_run on this thread (ms from CUDA events
 * on the launch stream) */
typedef struct pbsc_timing
{
    float h2d_ms, seed_ms, extend_ms, d2h_ms, total_ms;
    uint64_t kernel_launches;
    uint64_t seed_pairs;     /* FM walks attempted */
    uint64_t rank_queries;   /* occ() lookups issued by the kernels (0 unless built with PBSC_COUNT_OCC) */
    float dp_ms;             /* part of extend_ms spent in the DP / multiple-alignment fallback */
    uint64_t dp_jobs;        /* failed walks that went through correctByMSAlignment */
    uint64_t dp_rows;        /* overlapping reads retrieved and aligned for them */
    float walk_ms;           /* walk_levels_kernel alone: CUDA events around each of its launches, summed */
    uint64_t walk_launches;
    uint64_t dp_thread_rows; /* of dp_rows: aligned by the thread-per-alignment kernel (the rest by the warp-per-row kernel) */
} pbsc_timing;
int pbsc_last_timing(pbsc_timing* t);
/* measurement build only (libpbsc_count.so, compiled with -DPBSC_COUNT_OCC): distinct 32-byte index sectors the kernels asked for
 * since the last reset, per kernel family: out[0] seed phase, out[1] walk setup, out[2] level loop, out[3] DP fallback, out[4]
 * other.  Returns 1 in the measurement build, 0 (and zeros) in the product build. */
int pbsc_occ_counts(uint64_t* out, int reset);

/* ---- index construction: `stride index READS` (StriDe/index.cpp:86-214) --------------------------------------------------------------
 * Replaces BWTCA::runRopebwt2 (SuffixTools/BWTCARopebwt.cpp:160-247, ropebwt2 with MR_SO_IO: one sentinel per read, ordered by read
 * index), BWTWriterBinary (SuffixTools/BWTWriterBinary.cpp:28-94, run-length units of at most 31 symbols) and
 * SampledSuffixArray::buildLexicoIndex / writeLexicoIndex (SuffixTools/SampledSuffixArray.cpp:158-190,248-258).  The suffixes are
 * sorted on the GPU (pbsc_build.cu); PREFIX.bwt, PREFIX.rbwt, PREFIX.sai and PREFIX.rsai come out byte-identical to the reference's.
 * `reads`: the bases of all reads, one ASCII letter each (ACGT, either case; anything else is PBSC_ERR_ARG: the reference's index
 * holds $ACGT only), read r at [offsets[r], offsets[r + 1]).  Limits: fewer than 2^32-1 symbols (bases + reads). */
#define PBSC_BUILD_NO_FORWARD 1   /* --no-forward: skip PREFIX.bwt / .sai */
#define PBSC_BUILD_NO_REVERSE 2   /* --no-reverse: skip PREFIX.rbwt / .rsai */
int pbsc_build_index_files(const char* reads, const uint64_t* offsets, uint64_t n_reads, const char* prefix, int device, int flags);
/* One strand in memory: *runs = the RLUnit bytes of the .bwt (reverse = 0) or .rbwt (reverse = 1) body, *n_symbols = bases + reads,
 * *lex_order (optional) = the read index of the r-th '$' of the BWT (the body of the .sai / .rsai).  Both arrays are malloc'd here;
 * release them with pbsc_free.  pbsc_index_from_runs takes the bytes as they are. */
int pbsc_build_bwt(const char* reads, const uint64_t* offsets, uint64_t n_reads, int reverse, int device, uint8_t** runs, uint64_t* n_runs,
                   uint64_t* n_symbols, uint32_t** lex_order);
void pbsc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* PBSC_H */
