// pbsc_batch.cuh — device-side batch layout shared by the seed and extend phases.
#ifndef PBSC_BATCH_CUH
#define PBSC_BATCH_CUH

#include "pbsc_internal.h"

namespace pbsc {

// Per-position record of one static k-mer size: what LongReadProbe reads back from
// KmerFeature::Log()[staticSize][pos] (PacBio/KmerFeature.h:37-136).
struct __align__(32) StaticFeat
{
    uint64_t fwd_lo, rvc_lo;
    uint32_t fwd_size, rvc_size;
    int32_t freq;        // KmerFeature::frequency (not yet masked by `fake`)
    uint32_t count_fake; // count[A..T] in the low 4x7 bits (<=51 each), bit 31 = fake
};

struct SeedParamsDev
{
    int32_t pool[8];
    int32_t n_pool;
    int32_t start_kmer, scan_kmer, kmer_up_bound, radius, pb_coverage, manual, mode;
    int32_t offset[3];
    int32_t slot_of_mode[3];   // mode -> index into the distinct static sizes
    int32_t static_size[3];    // distinct static sizes
    int32_t n_static;
    float hh_ratio;
    float threshold[3][52];
};

// A batch of reads resident on the device.
struct DeviceBatch
{
    DevBuf<uint8_t> codes;      // 2-bit codes, one byte per base, reads concatenated
    DevBuf<uint64_t> offsets;   // n_reads + 1
    uint64_t n_reads = 0, n_bases = 0;
};

// Output of the seed phase, resident on the device.
struct SeedBuffers
{
    DevBuf<pbsc_seed> seeds;       // per-read regions [region[r], region[r+1])
    DevBuf<uint64_t> region;       // n_reads + 1
    DevBuf<uint32_t> count;        // surviving seeds of read r (packed at the start of its region)
    DevBuf<uint32_t> outcast;      // hitchhiked seeds of read r (packed after the surviving ones)
    uint64_t total_slots = 0;
};

// Everything one batch needs on the device, allocated once by pbsc_batch_upload so that pbsc_batch_run only
// launches kernels.
struct Workspace
{
    // seed phase
    DevBuf<StaticFeat> feats;
    DevBuf<uint8_t> cls, attr, ntriv;
    DevBuf<ulonglong2> prefix;   // uint4 running counts per base
    DevBuf<uint64_t> cand;       // SeedCand per base (16 bytes)
    DevBuf<pbsc_seed> seed_tmp;
    // extend phase
    DevBuf<uint8_t> pieces, scratch;
    DevBuf<uint64_t> piece_region, bounds_region;
    DevBuf<uint32_t> bounds, order;
    DevBuf<pbsc_read_stats> stats;
    DevBuf<int32_t> status;
    DevBuf<unsigned long long> counters;   // [0] work queue, [1] walks attempted
    DevBuf<unsigned int> maxima;           // [0] largest seed gap, [1] largest target seed
    // --debugseed only: repeat ratio per read position; failed walks per read (same per-read regions as the seeds) and their counts
    DevBuf<float> dbg_ratio;
    DevBuf<pbsc_walk_log> dbg_log;
    DevBuf<uint32_t> dbg_log_n;
    std::vector<uint64_t> h_piece_region, h_bounds_region;
    uint32_t q_cap = 0, node_cap = 0, merged_cap = 0;
    int blocks = 0;
    size_t scratch_stride = 0;
    float piece_factor = 1.5f;
    uint32_t pool_nodes = 256;   // label-tree pool: nodes per task (successful light walks copy their tree there)
    bool thread_engine = true;
};

int upload_reads(pbsc_index* idx, const char* reads, const uint64_t* offsets, uint64_t n_reads, DeviceBatch& b, cudaStream_t st);
int alloc_seed_workspace(const pbsc_params* p, const std::vector<uint64_t>& h_offsets, DeviceBatch& b, SeedBuffers& s, Workspace& w, cudaStream_t st);
int run_seed_phase(pbsc_index* idx, const pbsc_params* p, DeviceBatch& b, SeedBuffers& s, Workspace& w, uint64_t* launches);
int alloc_extend_workspace(pbsc_index* idx, const pbsc_params* p, const std::vector<uint64_t>& h_offsets, DeviceBatch& b, SeedBuffers& s, Workspace& w, cudaStream_t st);
// launches the chain kernel; returns PBSC_ERR_LIMIT when some read overflowed its scratch/piece capacity
int run_extend_chain(pbsc_index* idx, const pbsc_params* p, DeviceBatch& b, SeedBuffers& s, Workspace& w, uint64_t* launches);
// thread-per-walk engine with speculative pair scheduling (pbsc_extend_thread.cu); same outputs in the same buffers
int run_extend_threads(pbsc_index* idx, const pbsc_params* p, DeviceBatch& b, SeedBuffers& s, Workspace& w, uint64_t* launches);
// PBSC_ENGINE=warp selects the warp-per-read chain kernel, anything else the thread engine
bool use_thread_engine();

}  // namespace pbsc
#endif
