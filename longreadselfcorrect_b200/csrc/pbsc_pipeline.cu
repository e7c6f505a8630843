// pbsc_pipeline.cu — pbsc_correct_batch: the whole hot path for one batch of reads
// (PacBioSelfCorrectionProcess::process, PacBio/PacBioSelfCorrectionProcess.cpp:23-54):
// H2D of the reads, seed phase, FM-extend chain, D2H of the corrected pieces and counters.
#include <chrono>
#include "pbsc_batch.cuh"

namespace pbsc {
int run_extend_chain(pbsc_index* idx, const pbsc_params* p, DeviceBatch& b, SeedBuffers& s, const std::vector<uint64_t>& h_offsets,
                     std::vector<uint8_t>& h_pieces, std::vector<uint64_t>& h_piece_region, std::vector<uint32_t>& h_bounds,
                     std::vector<uint64_t>& h_bounds_region, std::vector<pbsc_read_stats>& h_stats, uint64_t* launches, uint64_t* walks,
                     float piece_factor, uint32_t node_cap);
}

using namespace pbsc;

extern "C" int pbsc_correct_batch(pbsc_index* idx, const pbsc_params* p, const char* reads, const uint64_t* offsets,
                                  uint64_t n_reads, char* pieces_out, uint64_t pieces_cap, uint64_t* piece_offsets,
                                  uint64_t piece_offsets_cap, uint64_t* first_piece, pbsc_read_stats* stats,
                                  uint64_t* bytes_needed)
{
    if (!idx || !p || !reads || !offsets || !piece_offsets || !first_piece || !stats) { set_error("pbsc_correct_batch: null argument"); return PBSC_ERR_ARG; }
    if (!p->no_dp) { set_error("pbsc_correct_batch: the DP/MSA fallback is not available in this build; pass --nodp (no_dp=1)"); return PBSC_ERR_ARG; }
    PBSC_CUDA(cudaSetDevice(idx->device));
    Timing& T = last_timing();
    T = Timing();
    cudaEvent_t ev[5];
    for (auto& e : ev) PBSC_CUDA(cudaEventCreate(&e));
    auto cleanup = [&]() { for (auto& e : ev) cudaEventDestroy(e); };
    cudaStream_t st = idx->stream;
    DeviceBatch b;
    SeedBuffers s;
    cudaEventRecord(ev[0], st);
    int rc = upload_reads(idx, reads, offsets, n_reads, b);
    if (rc != PBSC_OK) { cleanup(); return rc; }
    cudaEventRecord(ev[1], st);
    uint64_t launches = 1, walks = 0;
    rc = run_seed_phase(idx, p, b, s, &launches);
    if (rc != PBSC_OK) { cleanup(); return rc; }
    cudaEventRecord(ev[2], st);
    std::vector<uint64_t> h_offsets(offsets, offsets + n_reads + 1);
    std::vector<uint8_t> h_pieces;
    std::vector<uint64_t> h_piece_region, h_bounds_region;
    std::vector<uint32_t> h_bounds;
    std::vector<pbsc_read_stats> h_stats;
    float factor = 1.5f;
    uint32_t node_cap = 1u << 15;
    for (int attempt = 0;; attempt++)
    {
        rc = run_extend_chain(idx, p, b, s, h_offsets, h_pieces, h_piece_region, h_bounds, h_bounds_region, h_stats, &launches, &walks, factor, node_cap);
        if (rc != PBSC_ERR_LIMIT || attempt == 3) break;
        factor *= 2.0f;      // scratch or piece capacity exceeded somewhere in the batch: re-run larger
        node_cap *= 8;
        walks = 0;
    }
    if (rc == PBSC_ERR_LIMIT) set_error("pbsc_correct_batch: walk scratch capacity exceeded after retries");
    if (rc != PBSC_OK) { cleanup(); return rc; }
    cudaEventRecord(ev[3], st);
    // ---- pack results for the caller ----
    uint64_t np = 0, nbytes = 0;
    first_piece[0] = 0;
    for (uint64_t r = 0; r < n_reads; r++)
    {
        stats[r] = h_stats[r];
        const uint32_t k = h_stats[r].merge ? (uint32_t)h_stats[r].n_pieces : 0;
        const uint32_t* bounds = h_bounds.data() + h_bounds_region[r];
        for (uint32_t j = 0; j < k; j++) nbytes += bounds[j + 1] - bounds[j];
        np += k;
        first_piece[r + 1] = np;
    }
    if (bytes_needed) *bytes_needed = nbytes;
    if (np + 1 > piece_offsets_cap || nbytes > pieces_cap || (!pieces_out && nbytes))
    {
        set_error("pbsc_correct_batch: output needs %llu bytes and %llu piece offsets", (unsigned long long)nbytes, (unsigned long long)(np + 1));
        cleanup();
        return PBSC_ERR_LIMIT;
    }
    uint64_t w = 0, pi = 0;
    piece_offsets[0] = 0;
    for (uint64_t r = 0; r < n_reads; r++)
    {
        const uint32_t k = h_stats[r].merge ? (uint32_t)h_stats[r].n_pieces : 0;
        const uint32_t* bounds = h_bounds.data() + h_bounds_region[r];
        const uint8_t* src = h_pieces.data() + h_piece_region[r];
        for (uint32_t j = 0; j < k; j++)
        {
            for (uint32_t x = bounds[j]; x < bounds[j + 1]; x++) pieces_out[w++] = "ACGT"[src[x] & 3];
            piece_offsets[++pi] = w;
        }
    }
    cudaEventRecord(ev[4], st);
    cudaEventSynchronize(ev[4]);
    cudaEventElapsedTime(&T.h2d_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&T.seed_ms, ev[1], ev[2]);
    cudaEventElapsedTime(&T.extend_ms, ev[2], ev[3]);
    cudaEventElapsedTime(&T.d2h_ms, ev[3], ev[4]);
    cudaEventElapsedTime(&T.total_ms, ev[0], ev[4]);
    T.kernel_launches = launches;
    T.seed_pairs = walks;
    cleanup();
    return PBSC_OK;
}
