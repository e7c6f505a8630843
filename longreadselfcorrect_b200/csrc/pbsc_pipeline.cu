// pbsc_pipeline.cu — the whole hot path for one batch of reads
// (PacBioSelfCorrectionProcess::process, PacBio/PacBioSelfCorrectionProcess.cpp:23-54) as a staged,
// stream-ordered pipeline: upload (H2D + workspace) -> run (seed kernels, extend-chain kernel) -> fetch (D2H).
// pbsc_correct_batch is the three stages back to back on host buffers.
#include <string.h>
#include <stdlib.h>
#include <chrono>
#include "pbsc_batch.cuh"
#include "pbsc_dp.cuh"

using namespace pbsc;

struct pbsc_batch
{
    pbsc_index* idx = nullptr;
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
};

extern "C" {

int pbsc_batch_upload(pbsc_index* idx, const pbsc_params* p, const char* reads, const uint64_t* offsets, uint64_t n_reads, pbsc_batch** out)
{
    if (!idx || !p || !reads || !offsets || !out) { set_error("pbsc_batch_upload: null argument"); return PBSC_ERR_ARG; }
    *out = nullptr;
    if (!p->no_dp && !use_thread_engine())
    { set_error("the DP/MSA fallback (PacBioSelfCorrectionProcess.cpp:208-245) runs on the thread engine only; unset PBSC_ENGINE=warp or pass --nodp"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(idx->device));
    pbsc_batch* bt = new pbsc_batch();
    bt->idx = idx;
    bt->params = *p;
    bt->h_offsets.assign(offsets, offsets + n_reads + 1);
    if (idx->learned_node_cap) bt->w.node_cap = idx->learned_node_cap;
    if (idx->learned_piece_factor > 0) bt->w.piece_factor = idx->learned_piece_factor;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, idx->stream);
    int rc = upload_reads(idx, reads, offsets, n_reads, bt->b);
    cudaEventRecord(e1, idx->stream);
    if (rc == PBSC_OK) rc = alloc_seed_workspace(p, bt->h_offsets, bt->b, bt->s, bt->w, idx->stream);
    if (rc == PBSC_OK) rc = alloc_extend_workspace(idx, p, bt->h_offsets, bt->b, bt->s, bt->w);
    if (rc == PBSC_OK) { cudaEventSynchronize(e1); cudaEventElapsedTime(&bt->h2d_ms, e0, e1); }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc != PBSC_OK) { delete bt; return rc; }
    *out = bt;
    return PBSC_OK;
}

int pbsc_batch_run(pbsc_batch* bt, float* ms)
{
    if (!bt) { set_error("pbsc_batch_run: null"); return PBSC_ERR_ARG; }
    pbsc_index* idx = bt->idx;
    PBSC_CUDA(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    cudaEvent_t e[3];
    for (auto& x : e) PBSC_CUDA(cudaEventCreate(&x));
    auto done = [&]() { for (auto& x : e) cudaEventDestroy(x); };
    int rc = PBSC_OK;
    bt->launches = 0; bt->walks = 0;
    last_dp_stats() = DpStats();
    for (int attempt = 0;; attempt++)
    {
        cudaEventRecord(e[0], st);
        rc = run_seed_phase(idx, &bt->params, bt->b, bt->s, bt->w, &bt->launches);
        cudaEventRecord(e[1], st);
        if (rc == PBSC_OK) rc = bt->w.thread_engine ? run_extend_threads(idx, &bt->params, bt->b, bt->s, bt->w, &bt->launches)
                                                    : run_extend_chain(idx, &bt->params, bt->b, bt->s, bt->w, &bt->launches);
        cudaEventRecord(e[2], st);
        if (rc != PBSC_OK) break;
        cudaError_t ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) { rc = cuda_fail(ce, "hot-path kernels", __FILE__, __LINE__); break; }
        // any read that ran out of scratch or piece capacity? (4 bytes per read)
        const uint64_t n = bt->b.n_reads;
        std::vector<int32_t> status(n);
        if (n) PBSC_CUDA(cudaMemcpy(status.data(), bt->w.status.p, n * 4, cudaMemcpyDeviceToHost));
        bool overflow = false;
        for (uint64_t r = 0; r < n && rc == PBSC_OK; r++)
        {
            if (status[r] == -100) overflow = true;
            else if (status[r] == -101) { set_error("read %llu: a seed pair is outside this build's limits (walk k-mer > 61 or target shorter than -s)", (unsigned long long)r); rc = PBSC_ERR_LIMIT; }
            else if (status[r] == PBSC_WALK_NO_PATH) { set_error("Does it really happen?"); rc = PBSC_ERR_INTERNAL; }   // PacBioSelfCorrectionProcess.cpp:125-127
        }
        if (rc != PBSC_OK || !overflow) break;
        if (attempt == 3) { set_error("walk scratch capacity exceeded after %d retries", attempt); rc = PBSC_ERR_LIMIT; break; }
        // grow the capacities and run the batch again
        bt->w.piece_factor *= 2.0f;
        bt->w.node_cap *= 8;
        idx->learned_node_cap = bt->w.node_cap;
        idx->learned_piece_factor = bt->w.piece_factor;
        rc = alloc_extend_workspace(idx, &bt->params, bt->h_offsets, bt->b, bt->s, bt->w);
        if (rc != PBSC_OK) break;
    }
    if (rc == PBSC_OK)
    {
        cudaEventElapsedTime(&bt->seed_ms, e[0], e[1]);
        cudaEventElapsedTime(&bt->extend_ms, e[1], e[2]);
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
    }
    done();
    return rc;
}

// the small per-read results: counters and piece bounds
static int fetch_device_results(pbsc_batch* bt)
{
    if (bt->fetched) return PBSC_OK;
    pbsc_index* idx = bt->idx;
    cudaStream_t st = idx->stream;
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

int pbsc_batch_result_size(pbsc_batch* bt, uint64_t* piece_bytes, uint64_t* n_pieces)
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

int pbsc_batch_fetch(pbsc_batch* bt, char* pieces_out, uint64_t pieces_cap, uint64_t* piece_offsets, uint64_t piece_offsets_cap,
                     uint64_t* first_piece, pbsc_read_stats* stats)
{
    if (!bt || !bt->ran || !piece_offsets || !first_piece || !stats) { set_error("pbsc_batch_fetch: bad argument or batch has not run"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(bt->idx->device));
    cudaStream_t st = bt->idx->stream;
    cudaEvent_t e0, e1;
    PBSC_CUDA(cudaEventCreate(&e0)); PBSC_CUDA(cudaEventCreate(&e1));
    cudaEventRecord(e0, st);
    uint64_t nb = 0, np = 0;
    bt->fetched = false;
    int rc = pbsc_batch_result_size(bt, &nb, &np);
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
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    Timing& T = last_timing();
    T.d2h_ms = bt->d2h_ms;
    T.total_ms = bt->h2d_ms + bt->seed_ms + bt->extend_ms + bt->d2h_ms;
    T.kernel_launches = bt->launches + 1;
    return rc;
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
