// pbsc_pipeline.cu — the whole hot path for one batch of reads
// (PacBioSelfCorrectionProcess::process, PacBio/PacBioSelfCorrectionProcess.cpp:23-54) as a staged,
// stream-ordered pipeline: upload (H2D + workspace) -> run (seed kernels, extend-chain kernel) -> fetch (D2H).
// pbsc_correct_batch is the three stages back to back on host buffers.
#include <string.h>
#include <stdlib.h>
#include <chrono>
#include <memory>
#include <new>
#include "pbsc_batch.cuh"
#include "pbsc_dp.cuh"

using namespace pbsc;

struct pbsc_batch
{
    pbsc_index* idx = nullptr;            // the index (lane 0); the lane that runs the batch is picked by pbsc_batch_run
    pbsc_params params;
    DeviceBatch b;
    SeedBuffers s;
    Workspace w;
    std::vector<uint64_t> h_offsets;
    // fetched results
    std::vector<uint64_t> h_packed_off;   // start of read r's bytes in the packed output
    std::vector<uint32_t> h_bounds;
    std::vector<pbsc_read_stats> h_stats;
    bool ran = false, fetched = false;
    float h2d_ms = 0, seed_ms = 0, extend_ms = 0, d2h_ms = 0;
    uint64_t launches = 0, walks = 0;
    // copies of this batch run on its own stream: the upload of batch i+1 and the fetch of batch i-1 overlap the kernels of
    // batch i, which run on a lane's stream (north_star (4); Concurrency/SequenceProcessFramework.h:91-230)
    cudaStream_t cs = nullptr;
    cudaEvent_t ev_up = nullptr;
    ~pbsc_batch() { if (ev_up) cudaEventDestroy(ev_up); if (cs) cudaStreamDestroy(cs); }
};

namespace {
// events that are destroyed on every return path
struct Events
{
    std::vector<cudaEvent_t> e;
    cudaError_t create(int n) { e.assign(n, nullptr); for (auto& x : e) { cudaError_t rc = cudaEventCreate(&x); if (rc != cudaSuccess) return rc; } return cudaSuccess; }
    ~Events() { for (auto x : e) if (x) cudaEventDestroy(x); }
    cudaEvent_t operator[](int i) const { return e[i]; }
};
struct LaneGuard
{
    pbsc_index* primary; pbsc_index* lane;
    ~LaneGuard() { if (lane) lane_release(primary, lane); }
};
// no C++ exception may cross the C ABI (std::bad_alloc from the host-side vectors, mostly)
template <class F>
int guarded(const char* who, F f)
{
    try { return f(); }
    catch (const std::bad_alloc&) { set_error("%s: out of host memory", who); return PBSC_ERR_LIMIT; }
    catch (const std::exception& e) { set_error("%s: %s", who, e.what()); return PBSC_ERR_INTERNAL; }
    catch (...) { set_error("%s: unknown exception", who); return PBSC_ERR_INTERNAL; }
}
}  // namespace

static int batch_upload_impl(pbsc_index* idx, const pbsc_params* p, const char* reads, const uint64_t* offsets, uint64_t n_reads, pbsc_batch** out)
{
    if (!p->no_dp && !use_thread_engine())
    { set_error("the DP/MSA fallback (PacBioSelfCorrectionProcess.cpp:208-245) runs on the thread engine only; unset PBSC_ENGINE=warp or pass --nodp"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(idx->device));
    std::unique_ptr<pbsc_batch> bt(new pbsc_batch());
    bt->idx = idx;
    bt->params = *p;
    bt->h_offsets.assign(offsets, offsets + n_reads + 1);
    PBSC_CUDA(cudaStreamCreateWithFlags(&bt->cs, cudaStreamNonBlocking));
    PBSC_CUDA(cudaEventCreateWithFlags(&bt->ev_up, cudaEventDisableTiming));
    {
        std::lock_guard<std::mutex> lk(idx->run_mu);
        if (idx->learned_node_cap) bt->w.node_cap = idx->learned_node_cap;
        if (idx->learned_piece_factor > 0) bt->w.piece_factor = idx->learned_piece_factor;
        if (idx->learned_pool_nodes) bt->w.pool_nodes = idx->learned_pool_nodes;
    }
    Events ev;
    PBSC_CUDA(ev.create(2));
    cudaEventRecord(ev[0], bt->cs);
    // only the reads travel now (one byte per base on the device): the ~130 bytes per base of seed and extend workspace are
    // taken when the batch gets a lane and given back when its kernels are done, so a batch that waits for its turn, or for its
    // results to be fetched, holds almost no device memory
    int rc = upload_reads(idx, reads, offsets, n_reads, bt->b, bt->cs);
    cudaEventRecord(ev[1], bt->cs);
    if (rc == PBSC_OK) { cudaEventSynchronize(ev[1]); cudaEventElapsedTime(&bt->h2d_ms, ev[0], ev[1]); PBSC_CUDA(cudaEventRecord(bt->ev_up, bt->cs)); }
    if (rc != PBSC_OK) return rc;
    *out = bt.release();
    return PBSC_OK;
}

static int batch_run_impl(pbsc_batch* bt, float* ms)
{
    pbsc_index* primary = bt->idx;
    PBSC_CUDA(cudaSetDevice(primary->device));
    LaneGuard g{primary, nullptr};
    int rc = lane_acquire(primary, bt->params.idmer_len, &g.lane);
    if (rc != PBSC_OK) return rc;
    pbsc_index* idx = g.lane;
    cudaStream_t st = idx->stream;
    Events e;
    PBSC_CUDA(e.create(4));
    PBSC_CUDA(cudaStreamWaitEvent(st, bt->ev_up, 0));
    bt->launches = 0; bt->walks = 0;
    // capacities other batches of this index had to grow to since this one was uploaded
    bt->w.node_cap = std::max(bt->w.node_cap, idx->learned_node_cap);
    bt->w.pool_nodes = std::max(bt->w.pool_nodes, idx->learned_pool_nodes);
    bt->w.piece_factor = std::max(bt->w.piece_factor, idx->learned_piece_factor);
    rc = alloc_seed_workspace(&bt->params, bt->h_offsets, bt->b, bt->s, bt->w, st);
    if (rc == PBSC_OK) rc = alloc_extend_workspace(idx, &bt->params, bt->h_offsets, bt->b, bt->s, bt->w, st);
    if (rc != PBSC_OK) return rc;
    if (bt->params.debug_seed)
    {
        // --debugseed: the repeat ratio of every position and the failed walks of every read are kept for pbsc_batch_fetch_debug
        PBSC_CUDA(bt->w.dbg_ratio.alloc(bt->b.n_bases)); PBSC_CUDA(bt->w.dbg_log.alloc(bt->s.total_slots)); PBSC_CUDA(bt->w.dbg_log_n.alloc(bt->b.n_reads));
        PBSC_CUDA(cudaMemsetAsync(bt->w.dbg_log_n.p, 0, (bt->b.n_reads ? bt->b.n_reads : 1) * 4, st));
    }
    // what only the kernels need goes back to the block cache on every way out; the results (pieces, bounds, counters) stay
    struct Scratch
    {
        pbsc_batch* bt;
        ~Scratch()
        {
            Workspace& w = bt->w; SeedBuffers& s = bt->s;
            w.feats.release(); w.cls.release(); w.attr.release(); w.ntriv.release(); w.prefix.release(); w.cand.release(); w.seed_tmp.release();
            w.scratch.release(); w.order.release();
            if (!bt->params.debug_seed) { s.seeds.release(); s.region.release(); s.count.release(); s.outcast.release(); }
        }
    } scratch_guard{bt};
    bool grew = false;
    float extend_total = 0;
    for (int attempt = 0;; attempt++)
    {
        last_dp_stats() = DpStats();   // counters of the attempt that succeeds, not a sum over retries
        bt->launches = 0;
        if (attempt == 0)
        {
            cudaEventRecord(e[0], st);
            rc = run_seed_phase(idx, &bt->params, bt->b, bt->s, bt->w, &bt->launches);   // seeds do not depend on the capacities: once
            cudaEventRecord(e[1], st);
            if (rc != PBSC_OK) break;
        }
        cudaEventRecord(e[2], st);
        rc = bt->w.thread_engine ? run_extend_threads(idx, &bt->params, bt->b, bt->s, bt->w, &bt->launches)
                                 : run_extend_chain(idx, &bt->params, bt->b, bt->s, bt->w, &bt->launches);
        cudaEventRecord(e[3], st);
        if (rc != PBSC_OK) break;
        cudaError_t ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) { rc = cuda_fail(ce, "hot-path kernels", __FILE__, __LINE__); break; }
        { float t = 0; cudaEventElapsedTime(&t, e[2], e[3]); extend_total += t; }
        // any read that ran out of scratch or piece capacity? (4 bytes per read)
        const uint64_t n = bt->b.n_reads;
        std::vector<int32_t> status(n);
        if (n) PBSC_CUDA(cudaMemcpy(status.data(), bt->w.status.p, n * 4, cudaMemcpyDeviceToHost));
        uint64_t ovf_pieces = 0, ovf_tree = 0, ovf_pool = 0, ovf_dp = 0, ovf_other = 0;
        for (uint64_t r = 0; r < n && rc == PBSC_OK; r++)
        {
            switch (status[r])
            {
                case PBSC_OVF_PIECES: ovf_pieces++; break;
                case PBSC_OVF_TREE: ovf_tree++; break;
                case PBSC_OVF_POOL: ovf_pool++; break;
                case PBSC_OVF_DP: ovf_dp++; break;
                case PBSC_WALK_OVERFLOW: ovf_other++; break;
                case PBSC_WALK_UNSUPPORTED:
                    set_error("read %llu: a seed pair is outside this build's limits (walk k-mer > 61 or target shorter than -s)", (unsigned long long)r); rc = PBSC_ERR_LIMIT; break;
                case PBSC_WALK_NO_PATH: set_error("Does it really happen?"); rc = PBSC_ERR_INTERNAL; break;   // PacBioSelfCorrectionProcess.cpp:125-127
                default: break;
            }
        }
        if (rc != PBSC_OK) break;
        if (ovf_dp)
        {
            // not a capacity: larger walk scratch cannot help
            set_error("DP fallback: %llu read(s) need an alignment or consensus outside this build's limits (consensus longer than the walk's output slot, "
                      "or more than 3 x query + 128 inserted columns)", (unsigned long long)ovf_dp);
            rc = PBSC_ERR_LIMIT; break;
        }
        if (!(ovf_pieces || ovf_tree || ovf_pool || ovf_other)) break;
        if (attempt == 3)
        {
            set_error("capacity exceeded after %d retries (piece regions %llu, label trees %llu, label-tree pool %llu, other %llu reads)", attempt,
                      (unsigned long long)ovf_pieces, (unsigned long long)ovf_tree, (unsigned long long)ovf_pool, (unsigned long long)ovf_other);
            rc = PBSC_ERR_LIMIT; break;
        }
        // grow only what ran out, then run the extend phase again (the seeds stay)
        grew = true;
        if (ovf_pieces || ovf_other) bt->w.piece_factor *= 2.0f;
        if (ovf_tree || ovf_other) bt->w.node_cap *= 8;
        if (ovf_pool) bt->w.pool_nodes *= 4;
        if (ovf_pieces || ovf_other)
        {
            rc = alloc_extend_workspace(idx, &bt->params, bt->h_offsets, bt->b, bt->s, bt->w, st);
            if (rc != PBSC_OK) break;
        }
    }
    if (rc == PBSC_OK)
    {
        cudaEventElapsedTime(&bt->seed_ms, e[0], e[1]);
        bt->extend_ms = extend_total;
        unsigned long long hw = 0;
        cudaMemcpy(&hw, bt->w.counters.p + 1, 8, cudaMemcpyDeviceToHost);
        bt->walks = hw;
        bt->ran = true;
        bt->fetched = false;
        if (ms) *ms = bt->seed_ms + bt->extend_ms;
        Timing& T = last_timing();
        T.h2d_ms = bt->h2d_ms; T.seed_ms = bt->seed_ms; T.extend_ms = bt->extend_ms; T.d2h_ms = 0;
        T.total_ms = bt->h2d_ms + bt->seed_ms + bt->extend_ms;
        T.kernel_launches = bt->launches; T.seed_pairs = bt->walks;
        const DpStats& D = last_dp_stats();
        T.dp_ms = D.ms; T.dp_jobs = D.jobs; T.dp_rows = D.rows; T.dp_thread_rows = D.thread_rows;
        if (D.bad) { set_error("DP fallback: %llu alignments or consensus buffers outside this build's limits", (unsigned long long)D.bad); rc = PBSC_ERR_LIMIT; }
        if (grew && rc == PBSC_OK)
        {
            // later batches start from what this one needed -- recorded only now that it is known to work
            std::lock_guard<std::mutex> lk(primary->run_mu);
            primary->learned_node_cap = std::max(primary->learned_node_cap, bt->w.node_cap);
            primary->learned_piece_factor = std::max(primary->learned_piece_factor, bt->w.piece_factor);
            primary->learned_pool_nodes = std::max(primary->learned_pool_nodes, bt->w.pool_nodes);
        }
    }
    return rc;
}

extern "C" {

int pbsc_batch_upload(pbsc_index* idx, const pbsc_params* p, const char* reads, const uint64_t* offsets, uint64_t n_reads, pbsc_batch** out)
{
    if (!idx || !p || !reads || !offsets || !out) { set_error("pbsc_batch_upload: null argument"); return PBSC_ERR_ARG; }
    *out = nullptr;
    return guarded("pbsc_batch_upload", [&] { return batch_upload_impl(idx, p, reads, offsets, n_reads, out); });
}

int pbsc_batch_run(pbsc_batch* bt, float* ms)
{
    if (!bt) { set_error("pbsc_batch_run: null"); return PBSC_ERR_ARG; }
    return guarded("pbsc_batch_run", [&] { return batch_run_impl(bt, ms); });
}

}  // extern "C"

// the small per-read results: counters and piece bounds
static int fetch_device_results(pbsc_batch* bt)
{
    if (bt->fetched) return PBSC_OK;
    cudaStream_t st = bt->cs;
    const uint64_t n = bt->b.n_reads;
    bt->h_bounds.resize(bt->w.h_bounds_region[n]);
    bt->h_stats.resize(n);
    if (n) PBSC_CUDA(cudaMemcpyAsync(bt->h_stats.data(), bt->w.stats.p, n * sizeof(pbsc_read_stats), cudaMemcpyDeviceToHost, st));
    if (bt->h_bounds.size()) PBSC_CUDA(cudaMemcpyAsync(bt->h_bounds.data(), bt->w.bounds.p, bt->h_bounds.size() * 4, cudaMemcpyDeviceToHost, st));
    PBSC_CUDA(cudaStreamSynchronize(st));
    bt->h_packed_off.assign(n + 1, 0);
    for (uint64_t r = 0; r < n; r++)
    {
        uint64_t used = 0;
        if (bt->h_stats[r].merge) used = (bt->h_bounds.data() + bt->w.h_bounds_region[r])[bt->h_stats[r].n_pieces];
        bt->h_packed_off[r + 1] = bt->h_packed_off[r] + used;
    }
    bt->fetched = true;
    return PBSC_OK;
}

static int result_size_impl(pbsc_batch* bt, uint64_t* piece_bytes, uint64_t* n_pieces)
{
    if (!bt || !bt->ran) { set_error("pbsc_batch_result_size: batch has not run"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(bt->idx->device));
    int rc = fetch_device_results(bt);
    if (rc != PBSC_OK) return rc;
    uint64_t np = 0;
    for (uint64_t r = 0; r < bt->b.n_reads; r++) if (bt->h_stats[r].merge) np += (uint64_t)bt->h_stats[r].n_pieces;
    if (piece_bytes) *piece_bytes = bt->h_packed_off[bt->b.n_reads];
    if (n_pieces) *n_pieces = np;
    return PBSC_OK;
}

// warp per read: gather the bytes each read produced into one packed ASCII buffer
__global__ void pack_pieces_kernel(uint64_t n_reads, const uint8_t* __restrict__ pieces, const uint64_t* __restrict__ piece_region,
                                   const uint64_t* __restrict__ packed_off, char* __restrict__ out)
{
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_reads) return;
    const uint64_t a = packed_off[warp], e = packed_off[warp + 1];
    const uint8_t* src = pieces + piece_region[warp];
    for (uint64_t x = lane; x < e - a; x += 32) out[a + x] = "ACGT"[src[x] & 3];
}

static int fetch_impl(pbsc_batch* bt, char* pieces_out, uint64_t pieces_cap, uint64_t* piece_offsets, uint64_t piece_offsets_cap,
                      uint64_t* first_piece, pbsc_read_stats* stats)
{
    if (!bt || !bt->ran || !piece_offsets || !first_piece || !stats) { set_error("pbsc_batch_fetch: bad argument or batch has not run"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(bt->idx->device));
    cudaStream_t st = bt->cs;
    Events ev;
    PBSC_CUDA(ev.create(2));
    cudaEvent_t e0 = ev[0], e1 = ev[1];
    cudaEventRecord(e0, st);
    uint64_t nb = 0, np = 0;
    bt->fetched = false;
    int rc = result_size_impl(bt, &nb, &np);
    if (rc == PBSC_OK && (np + 1 > piece_offsets_cap || nb > pieces_cap || (!pieces_out && nb)))
    {
        set_error("pbsc_batch_fetch: output needs %llu bytes and %llu piece offsets", (unsigned long long)nb, (unsigned long long)(np + 1));
        rc = PBSC_ERR_LIMIT;
    }
    const uint64_t n = bt->b.n_reads;
    if (rc == PBSC_OK && nb)
    {
        DevBuf<uint64_t> d_off; DevBuf<char> d_out;
        cudaError_t ce = d_off.alloc(n + 1);
        if (ce == cudaSuccess) ce = d_out.alloc(nb);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_off.p, bt->h_packed_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st);
        if (ce == cudaSuccess)
        {
            pack_pieces_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(n, bt->w.pieces.p, bt->w.piece_region.p, d_off.p, d_out.p);
            ce = cudaMemcpyAsync(pieces_out, d_out.p, nb, cudaMemcpyDeviceToHost, st);
        }
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) rc = cuda_fail(ce, "pack/fetch pieces", __FILE__, __LINE__);
    }
    if (rc == PBSC_OK)
    {
        uint64_t pi = 0;
        piece_offsets[0] = 0;
        first_piece[0] = 0;
        for (uint64_t r = 0; r < n; r++)
        {
            stats[r] = bt->h_stats[r];
            const uint32_t k = bt->h_stats[r].merge ? (uint32_t)bt->h_stats[r].n_pieces : 0;
            const uint32_t* bounds = bt->h_bounds.data() + bt->w.h_bounds_region[r];
            for (uint32_t j = 0; j < k; j++) piece_offsets[++pi] = bt->h_packed_off[r] + bounds[j + 1];
            first_piece[r + 1] = pi;
        }
    }
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&bt->d2h_ms, e0, e1);
    Timing& T = last_timing();
    T.d2h_ms = bt->d2h_ms;
    T.total_ms = bt->h2d_ms + bt->seed_ms + bt->extend_ms + bt->d2h_ms;
    T.kernel_launches = bt->launches + 1;
    return rc;
}

extern "C" {

int pbsc_batch_result_size(pbsc_batch* bt, uint64_t* piece_bytes, uint64_t* n_pieces)
{
    return guarded("pbsc_batch_result_size", [&] { return result_size_impl(bt, piece_bytes, n_pieces); });
}

int pbsc_batch_fetch(pbsc_batch* bt, char* pieces_out, uint64_t pieces_cap, uint64_t* piece_offsets, uint64_t piece_offsets_cap,
                     uint64_t* first_piece, pbsc_read_stats* stats)
{
    return guarded("pbsc_batch_fetch", [&] { return fetch_impl(bt, pieces_out, pieces_cap, piece_offsets, piece_offsets_cap, first_piece, stats); });
}

int pbsc_batch_debug_size(pbsc_batch* bt, uint64_t* n_seeds, uint64_t* n_log)
{
    if (!bt || !bt->ran || !bt->params.debug_seed) { set_error("pbsc_batch_debug_size: the batch has not run with params.debug_seed set"); return PBSC_ERR_ARG; }
    return guarded("pbsc_batch_debug_size", [&] {
        PBSC_CUDA(cudaSetDevice(bt->idx->device));
        const uint64_t n = bt->b.n_reads;
        std::vector<uint32_t> a(n), b(n), c(n);
        if (n)
        {
            PBSC_CUDA(cudaMemcpy(a.data(), bt->s.count.p, n * 4, cudaMemcpyDeviceToHost));
            PBSC_CUDA(cudaMemcpy(b.data(), bt->s.outcast.p, n * 4, cudaMemcpyDeviceToHost));
            PBSC_CUDA(cudaMemcpy(c.data(), bt->w.dbg_log_n.p, n * 4, cudaMemcpyDeviceToHost));
        }
        uint64_t ns = 0, nl = 0;
        for (uint64_t r = 0; r < n; r++) { ns += (uint64_t)a[r] + b[r]; nl += c[r]; }
        if (n_seeds) *n_seeds = ns;
        if (n_log) *n_log = nl;
        return PBSC_OK;
    });
}

int pbsc_batch_fetch_debug(pbsc_batch* bt, pbsc_seed* seeds, uint64_t seeds_cap, uint64_t* seed_offsets, uint32_t* n_surviving, float* ratio,
                           uint64_t ratio_cap, pbsc_walk_log* log, uint64_t log_cap, uint64_t* log_offsets)
{
    if (!bt || !bt->ran || !bt->params.debug_seed || !seed_offsets || !n_surviving || !log_offsets)
    { set_error("pbsc_batch_fetch_debug: bad argument, or the batch has not run with params.debug_seed set"); return PBSC_ERR_ARG; }
    return guarded("pbsc_batch_fetch_debug", [&] {
        PBSC_CUDA(cudaSetDevice(bt->idx->device));
        const uint64_t n = bt->b.n_reads;
        std::vector<uint32_t> cnt(n), outc(n), nlog(n);
        std::vector<uint64_t> region(n + 1, 0);
        if (n)
        {
            PBSC_CUDA(cudaMemcpy(cnt.data(), bt->s.count.p, n * 4, cudaMemcpyDeviceToHost));
            PBSC_CUDA(cudaMemcpy(outc.data(), bt->s.outcast.p, n * 4, cudaMemcpyDeviceToHost));
            PBSC_CUDA(cudaMemcpy(nlog.data(), bt->w.dbg_log_n.p, n * 4, cudaMemcpyDeviceToHost));
            PBSC_CUDA(cudaMemcpy(region.data(), bt->s.region.p, (n + 1) * 8, cudaMemcpyDeviceToHost));
        }
        seed_offsets[0] = 0; log_offsets[0] = 0;
        for (uint64_t r = 0; r < n; r++)
        {
            n_surviving[r] = cnt[r];
            seed_offsets[r + 1] = seed_offsets[r] + cnt[r] + outc[r];
            log_offsets[r + 1] = log_offsets[r] + nlog[r];
        }
        if (seed_offsets[n] > seeds_cap || log_offsets[n] > log_cap || bt->b.n_bases > ratio_cap || (!seeds && seed_offsets[n]) || (!log && log_offsets[n]) || (!ratio && bt->b.n_bases))
        { set_error("pbsc_batch_fetch_debug: needs %llu seeds, %llu log records, %llu ratios", (unsigned long long)seed_offsets[n], (unsigned long long)log_offsets[n], (unsigned long long)bt->b.n_bases); return PBSC_ERR_LIMIT; }
        std::vector<pbsc_seed> all(bt->s.total_slots);
        std::vector<pbsc_walk_log> lg(bt->s.total_slots);
        if (bt->s.total_slots)
        {
            PBSC_CUDA(cudaMemcpy(all.data(), bt->s.seeds.p, bt->s.total_slots * sizeof(pbsc_seed), cudaMemcpyDeviceToHost));
            PBSC_CUDA(cudaMemcpy(lg.data(), bt->w.dbg_log.p, bt->s.total_slots * sizeof(pbsc_walk_log), cudaMemcpyDeviceToHost));
        }
        for (uint64_t r = 0; r < n; r++)
        {
            for (uint32_t i = 0; i < cnt[r] + outc[r]; i++) seeds[seed_offsets[r] + i] = all[region[r] + i];
            for (uint32_t i = 0; i < nlog[r]; i++) log[log_offsets[r] + i] = lg[region[r] + i];
        }
        if (bt->b.n_bases) PBSC_CUDA(cudaMemcpy(ratio, bt->w.dbg_ratio.p, bt->b.n_bases * sizeof(float), cudaMemcpyDeviceToHost));
        return PBSC_OK;
    });
}

void pbsc_batch_destroy(pbsc_batch* bt)
{
    if (!bt) return;
    cudaSetDevice(bt->idx->device);
    delete bt;
}

int pbsc_correct_batch(pbsc_index* idx, const pbsc_params* p, const char* reads, const uint64_t* offsets, uint64_t n_reads,
                       char* pieces_out, uint64_t pieces_cap, uint64_t* piece_offsets, uint64_t piece_offsets_cap,
                       uint64_t* first_piece, pbsc_read_stats* stats, uint64_t* bytes_needed)
{
    if (!idx || !p || !reads || !offsets || !piece_offsets || !first_piece || !stats) { set_error("pbsc_correct_batch: null argument"); return PBSC_ERR_ARG; }
    pbsc_batch* bt = nullptr;
    const bool trace = getenv("PBSC_TRACE") != nullptr;
    auto now = []() { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t0 = now();
    int rc = pbsc_batch_upload(idx, p, reads, offsets, n_reads, &bt);
    if (rc != PBSC_OK) return rc;
    const auto t1 = now();
    rc = pbsc_batch_run(bt, nullptr);
    const auto t2 = now();
    if (rc == PBSC_OK)
    {
        uint64_t nb = 0, np = 0;
        rc = pbsc_batch_result_size(bt, &nb, &np);
        if (bytes_needed) *bytes_needed = nb;
        if (rc == PBSC_OK) rc = pbsc_batch_fetch(bt, pieces_out, pieces_cap, piece_offsets, piece_offsets_cap, first_piece, stats);
    }
    const auto t3 = now();
    pbsc_batch_destroy(bt);
    if (trace) fprintf(stderr, "[pbsc] correct_batch host wall: upload %.1f ms, run %.1f ms, fetch %.1f ms, destroy %.1f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, now()));
    return rc;
}

}  // extern "C"
